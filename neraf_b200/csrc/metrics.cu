// Acoustic yard-sticks of rendered impulse responses, batched on the device: T60, EDT
// and C50 as /root/reference/NeRAF/NeRAF_helper.py computes them one RIR at a time in numpy after a device -> host
// copy (compute_t60 :48-64, measure_rt60_advance :66-77, measure_clarity :104-107, measure_edt :124-146;
// pyroomacoustics.experimental.measure_rt60 for the Schroeder fit).
//
// One THREAD per impulse response.  The reference's arithmetic is sequential by construction -- numpy's cumsum of the
// reversed float32 power is a running float32 sum from the tail, and the decay times are "first index where the
// curve is below a threshold" -- so each thread walks its response from the tail exactly like numpy does (bit-equal
// running sums) and the batch supplies the parallelism: 2 000-16 000 responses per render.  A warp touches 32 rows at
// a time and re-uses every 128-byte line for 32 steps, so the walks run out of L1 (transposing the batch so that a warp
// reads ONE line per step was measured 1.5x slower: every step then waits for L2).
#include "common.cuh"
#include "kernels.h"
#include "metrics_core.h"

namespace neraf {

struct MetricArgs {
  const float* wave; long long S; int L; double fs;
  int highpass; float decay_db; metrics::Biquad biquad;
  int t50; float* filtered;
  double* t60; double* edt; double* c50;
};

// the per-response arithmetic lives in metrics_core.h (host/device-portable: also driven on the CPU by the tests)
__global__ void __launch_bounds__(32) acoustic_metrics_kernel(MetricArgs a) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.S) return;
  metrics::measure(a.wave + s * (long long)a.L, a.L, a.fs, a.highpass != 0, a.biquad, a.decay_db, a.t50,
                   a.highpass ? a.filtered + s * (long long)a.L : nullptr, a.t60 ? a.t60 + s : nullptr,
                   a.edt ? a.edt + s : nullptr, a.c50 ? a.c50 + s : nullptr);
}

}  // namespace neraf

using namespace neraf;

extern "C" int neraf_acoustic_metrics(const neraf_metric_params* p, const float* wave, int64_t n_signals,
                                      void* workspace, size_t workspace_bytes, double* t60, double* edt, double* c50,
                                      neraf_stream_t stream) {
  NERAF_REQUIRE(p && p->n_samples >= 2 && p->fs > 0.0 && p->t60_decay_db > 0.f, "acoustic_metrics: bad parameters");
  NERAF_REQUIRE(n_signals >= 0, "acoustic_metrics: n_signals < 0");
  if (n_signals == 0) return NERAF_OK;
  NERAF_REQUIRE(wave && (t60 || edt || c50), "acoustic_metrics: null pointer");
  MetricArgs a = {};
  a.wave = wave; a.S = n_signals; a.L = p->n_samples; a.fs = p->fs;
  a.highpass = (t60 && p->t60_highpass_hz > 0.0) ? 1 : 0; a.decay_db = p->t60_decay_db;
  a.t50 = (int)((50.0 / 1000.0) * p->fs + 1.0);
  a.t60 = t60; a.edt = edt; a.c50 = c50;
  if (a.highpass) {
    const size_t need = (size_t)n_signals * p->n_samples * sizeof(float);
    if (!workspace || workspace_bytes < need)
      return set_error(NERAF_ERR_WORKSPACE, "acoustic_metrics: the high-passed T60 needs %zu workspace bytes, got %zu",
                       need, workspace_bytes);
    a.filtered = reinterpret_cast<float*>(workspace);
    a.biquad = metrics::highpass_coeffs(p->fs, p->t60_highpass_hz);
  }
  acoustic_metrics_kernel<<<(unsigned)ceil_div(n_signals, 32), 32, 0, (cudaStream_t)stream>>>(a);      // one warp per block: spread over the SMs
  NERAF_CHECK_LAUNCH("acoustic_metrics_kernel");
  return NERAF_OK;
}
