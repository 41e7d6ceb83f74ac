// Drives neraf_b200/csrc/metrics_core.h -- the per-response code the acoustic-metrics kernel executes -- on the CPU.
//   metrics_host_check <n_signals> <n_samples> <fs> <highpass_hz> <decay_db>   < float32 waveforms   > float64 (S, 3)
// Output rows: t60, edt, c50.  Test infrastructure only (tests/test_metrics.py).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../neraf_b200/csrc/metrics_core.h"

int main(int argc, char** argv) {
  if (argc != 6) return 2;
  const long S = atol(argv[1]);
  const int L = atoi(argv[2]);
  const double fs = atof(argv[3]), hz = atof(argv[4]);
  const float decay_db = (float)atof(argv[5]);
  std::vector<float> wave((size_t)S * L), scratch(L);
  if (fread(wave.data(), sizeof(float), wave.size(), stdin) != wave.size()) return 3;
  const bool hp = hz > 0.0;
  const neraf::metrics::Biquad c = hp ? neraf::metrics::highpass_coeffs(fs, hz) : neraf::metrics::Biquad{};
  const int t50 = (int)((50.0 / 1000.0) * fs + 1.0);
  std::vector<double> out((size_t)S * 3);
  for (long s = 0; s < S; ++s)
    neraf::metrics::measure(wave.data() + (size_t)s * L, L, fs, hp, c, decay_db, t50, scratch.data(), &out[3 * s],
                            &out[3 * s + 1], &out[3 * s + 2]);
  fwrite(out.data(), sizeof(double), out.size(), stdout);
  return 0;
}
