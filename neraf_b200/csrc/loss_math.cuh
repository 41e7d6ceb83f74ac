// The two weighted loss scalars from the (all-reduced) partial sums -- shared by loss.cu and the heads' backward
// kernel (elementwise.cu), which forms them when the loss gradient is fused into it.
#pragma once
#include <stdint.h>

#include "../../include/neraf_b200.h"

namespace neraf {

__device__ __forceinline__ void finalize(const double* sums, int64_t n_total, int criterion, float w_sc, float w_mag,
                                         float* losses) {
  const double n = (double)n_total;
  if (criterion == NERAF_CRIT_MSE) {
    losses[0] = 0.f;
    losses[1] = (float)(w_mag * (sums[2] / n));
  } else {
    losses[0] = (float)(w_sc * (sqrt(sums[0]) / sqrt(sums[1])));          // NeRAF_evaluator.py:26 (no epsilon)
    losses[1] = (float)(w_mag * ((criterion == NERAF_CRIT_SC_SLMSE ? sums[2] : sums[3]) / n));
  }
}

}  // namespace neraf
