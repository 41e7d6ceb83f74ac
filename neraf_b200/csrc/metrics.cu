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

namespace neraf {

struct MetricArgs {
  const float* wave; long long S; int L; double fs;
  int highpass; float decay_db; double b0, b1, b2, a1, a2;
  int t50; float* filtered;
  double* t60; double* edt; double* c50;
};

// energy_db[n] = 10 log10(E[n]) - 10 log10(E[0]) over n < i_nz, float32 like numpy on a float32 response.
// Returns through the references: EDT index (first n with -10 - e_db > 0) and the rt60 pair (i_5db, i_decay, decay).
struct Decay { int i_edt, i_5, i_dec; float decay; bool ok; };

// c50 (optional): the first walk also accumulates measure_clarity's two energy sums (fp64), split at sample t50.
__device__ __forceinline__ Decay decay_walk(const float* __restrict__ h, int L, float decay_db, bool want_edt,
                                            double* c50 = nullptr, int t50 = 0) {
  Decay d; d.i_edt = -1; d.i_5 = -1; d.i_dec = -1; d.decay = decay_db; d.ok = false;
  float e = 0.f;
  int i_nz = -1;
  double early = 0.0, late = 0.0;
  for (int n = L - 1; n >= 0; --n) {                 // np.cumsum(power[::-1])[::-1]: running float32 sum from the tail
    const float v = h[n];
    const float p = __fmul_rn(v, v);
    e = __fadd_rn(e, p);
    if (i_nz < 0 && e > 0.f) i_nz = n;                // np.max(np.where(energy > 0))
    if (c50) { if (n < t50) early += (double)p; else late += (double)p; }
  }
  if (c50) *c50 = 10.0 * log10(early / late);         // measure_clarity: 10 log10(sum h^2[:t] / sum h^2[t:])
  if (i_nz <= 0) return d;                            // all-zero response, or nothing left after energy[:i_nz]
  const float l0 = __fmul_rn(10.f, log10f(e));        // energy_db[0] before the shift
  float thr_dec = 0.f;
  e = 0.f;
  for (int n = L - 1; n >= 0; --n) {
    const float v = h[n];
    e = __fadd_rn(e, __fmul_rn(v, v));
    if (n >= i_nz) continue;                          // energy[:i_nz]
    const float db = __fsub_rn(__fmul_rn(10.f, log10f(e)), l0);
    if (n == i_nz - 1) {                              // the curve is non-increasing: its minimum is its last element
      const float min_db = -db;
      if (min_db - 5.f < decay_db) d.decay = min_db;  // measure_rt60: not enough dynamic range for decay_db
      thr_dec = -5.f - d.decay;
    }
    if (db < -5.f) d.i_5 = n;                         // descending walk: the last hit is np.min(np.where(...))
    if (db < thr_dec) d.i_dec = n;
    if (want_edt && (-10.f - db) > 0.f) d.i_edt = n;  // measure_edt, decay_db = 10
  }
  d.ok = true;
  return d;
}

__global__ void __launch_bounds__(32) acoustic_metrics_kernel(MetricArgs a) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.S) return;
  const float* h = a.wave + s * (long long)a.L;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);

  // raw response: EDT, (SoundSpaces) T60, and C50 (t = int(0.05 fs + 1)) ride on the same two walks
  double c50 = 0.0;
  Decay raw = decay_walk(h, a.L, a.highpass ? 10.f : a.decay_db, a.edt != nullptr, a.c50 ? &c50 : nullptr, a.t50);
  if (a.c50) a.c50[s] = c50;
  if (a.edt) a.edt[s] = (raw.ok && raw.i_edt >= 0) ? (60.0 / 10.0) * ((double)raw.i_edt / a.fs) : nan;

  if (a.t60) {
    Decay d = raw;
    if (a.highpass) {
      // torchaudio.functional.highpass_biquad (RBJ high-pass, direct form I, clamp to [-1, 1]) in float64
      float* y = a.filtered + s * (long long)a.L;
      // y[n] = (b0 x[n] + b1 x[n-1] + b2 x[n-2] - a2 y[n-2]) - a1 y[n-1]: everything but the last product is off the
      // loop-carried chain, which is ONE fp64 FMA per sample
      double x1 = 0.0, x2 = 0.0, y1 = 0.0, y2 = 0.0;
      for (int n = 0; n < a.L; ++n) {
        const double xn = (double)h[n];
        const double u = fma(a.b0, xn, fma(a.b1, x1, fma(a.b2, x2, -a.a2 * y2)));
        const double yn = fma(-a.a1, y1, u);
        x2 = x1; x1 = xn; y2 = y1; y1 = yn;
        y[n] = (float)fmin(fmax(yn, -1.0), 1.0);
      }
      d = decay_walk(y, a.L, a.decay_db, false);
    }
    // compute_t60's try/except: any failure of the fit reads -1
    a.t60[s] = (d.ok && d.i_5 >= 0 && d.i_dec >= 0)
                   ? (double)(60.f / d.decay) * ((double)d.i_dec / a.fs - (double)d.i_5 / a.fs) : -1.0;
  }
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
    const double w0 = 2.0 * 3.14159265358979323846 * p->t60_highpass_hz / p->fs, q = 0.707;
    const double alpha = sin(w0) / 2.0 / q, cw = cos(w0), a0 = 1.0 + alpha;
    a.b0 = (1.0 + cw) / 2.0 / a0; a.b1 = (-1.0 - cw) / a0; a.b2 = a.b0;
    a.a1 = -2.0 * cw / a0; a.a2 = (1.0 - alpha) / a0;
  }
  acoustic_metrics_kernel<<<(unsigned)ceil_div(n_signals, 32), 32, 0, (cudaStream_t)stream>>>(a);      // one warp per block: spread over the SMs
  NERAF_CHECK_LAUNCH("acoustic_metrics_kernel");
  return NERAF_OK;
}
