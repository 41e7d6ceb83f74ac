// Per-response arithmetic of the acoustic-metrics kernel (metrics.cu), written as host/device-portable code so that the
// exact walks the GPU executes can also be driven on the CPU (tests/csrc/metrics_host_check.cpp) and compared with the
// oracle without a device.  References: /root/reference/NeRAF/NeRAF_helper.py -- compute_t60 :48-64,
// measure_rt60_advance :66-77 (pyroomacoustics measure_rt60), measure_clarity :104-107, measure_edt :124-146.
#pragma once

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#else
#define MC_HD inline
#endif
#include <math.h>

namespace neraf {
namespace metrics {

// float32 operations with one rounding each (no contraction): numpy's float32 arithmetic
#if defined(__CUDA_ARCH__)
MC_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
MC_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
MC_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
#else
MC_HD float f_mul(float a, float b) { volatile float r = a * b; return r; }
MC_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
MC_HD float f_sub(float a, float b) { volatile float r = a - b; return r; }
#endif

// energy_db[n] = 10 log10(E[n]) - 10 log10(E[0]) over n < i_nz, float32 like numpy on a float32 response.
// EDT index (first n with -10 - e_db > 0) and the rt60 pair (i_5db, i_decay, decay).
struct Decay { int i_edt, i_5, i_dec; float decay; bool ok; };

// c50 (optional): the first walk also accumulates measure_clarity's two energy sums (fp64), split at sample t50.
MC_HD Decay decay_walk(const float* h, int L, float decay_db, bool want_edt, double* c50 = nullptr, int t50 = 0) {
  Decay d; d.i_edt = -1; d.i_5 = -1; d.i_dec = -1; d.decay = decay_db; d.ok = false;
  float e = 0.f;
  int i_nz = -1;
  double early = 0.0, late = 0.0;
  for (int n = L - 1; n >= 0; --n) {                 // np.cumsum(power[::-1])[::-1]: running float32 sum from the tail
    const float v = h[n];
    const float p = f_mul(v, v);
    e = f_add(e, p);
    if (i_nz < 0 && e > 0.f) i_nz = n;                // np.max(np.where(energy > 0))
    if (c50) { if (n < t50) early += (double)p; else late += (double)p; }
  }
  if (c50) *c50 = 10.0 * log10(early / late);         // measure_clarity: 10 log10(sum h^2[:t] / sum h^2[t:])
  if (i_nz <= 0) return d;                            // all-zero response, or nothing left after energy[:i_nz]
  const float l0 = f_mul(10.f, log10f(e));            // energy_db[0] before the shift
  float thr_dec = 0.f;
  e = 0.f;
  for (int n = L - 1; n >= 0; --n) {
    const float v = h[n];
    e = f_add(e, f_mul(v, v));
    if (n >= i_nz) continue;                          // energy[:i_nz]
    const float db = f_sub(f_mul(10.f, log10f(e)), l0);
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

// torchaudio.functional.highpass_biquad coefficients (RBJ high-pass, Q = 0.707), normalised by a0
struct Biquad { double b0, b1, b2, a1, a2; };

inline Biquad highpass_coeffs(double fs, double cutoff_hz) {
  const double w0 = 2.0 * 3.14159265358979323846 * cutoff_hz / fs, q = 0.707;
  const double alpha = sin(w0) / 2.0 / q, cw = cos(w0), a0 = 1.0 + alpha;
  Biquad c;
  c.b0 = (1.0 + cw) / 2.0 / a0; c.b1 = (-1.0 - cw) / a0; c.b2 = c.b0;
  c.a1 = -2.0 * cw / a0; c.a2 = (1.0 - alpha) / a0;
  return c;
}

// Direct form I in float64, output clamped to [-1, 1] and rounded to float32 (torchaudio's lfilter clamp).
// y[n] = (b0 x[n] + b1 x[n-1] + b2 x[n-2] - a2 y[n-2]) - a1 y[n-1]: everything but the last product is off the
// loop-carried chain, which is ONE fp64 FMA per sample.
MC_HD void highpass(const float* h, int L, const Biquad& c, float* y) {
  double x1 = 0.0, x2 = 0.0, y1 = 0.0, y2 = 0.0;
  for (int n = 0; n < L; ++n) {
    const double xn = (double)h[n];
    const double u = fma(c.b0, xn, fma(c.b1, x1, fma(c.b2, x2, -c.a2 * y2)));
    const double yn = fma(-c.a1, y1, u);
    x2 = x1; x1 = xn; y2 = y1; y1 = yn;
    y[n] = (float)fmin(fmax(yn, -1.0), 1.0);
  }
}

// One impulse response -> T60 / EDT / C50 (any of the three outputs may be null).  `filtered`: L floats of scratch
// when do_highpass.  A failed T60 fit reads -1 (compute_t60's try / except), a curve that never drops 10 dB gives NaN.
MC_HD void measure(const float* h, int L, double fs, bool do_highpass, const Biquad& c, float decay_db, int t50,
                   float* filtered, double* t60, double* edt, double* c50) {
  // raw response: EDT, (SoundSpaces) T60, and C50 (t = int(0.05 fs + 1)) ride on the same two walks
  double c50v = 0.0;
  const Decay raw = decay_walk(h, L, do_highpass ? 10.f : decay_db, edt != nullptr, c50 ? &c50v : nullptr, t50);
  if (c50) *c50 = c50v;
  if (edt) *edt = (raw.ok && raw.i_edt >= 0) ? (60.0 / 10.0) * ((double)raw.i_edt / fs) : (double)NAN;
  if (t60) {
    Decay d = raw;
    if (do_highpass) {
      highpass(h, L, c, filtered);
      d = decay_walk(filtered, L, decay_db, false);
    }
    *t60 = (d.ok && d.i_5 >= 0 && d.i_dec >= 0) ? (double)(60.f / d.decay) * ((double)d.i_dec / fs - (double)d.i_5 / fs) : -1.0;
  }
}

}  // namespace metrics
}  // namespace neraf
