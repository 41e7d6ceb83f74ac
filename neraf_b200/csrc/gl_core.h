// Per-lane building blocks of the fused Griffin-Lim kernel (K3), written as host/device-portable
// code so that the exact index arithmetic the GPU executes can also be driven lane-by-lane on the CPU
// (tests/csrc/gl_host_check.cpp) and compared with a reference FFT before any GPU time is spent.
//
// A warp owns one STFT frame at a time.  The n_fft-point real transform is done as an H = n_fft/2
// point complex FFT (Stockham autosort, radix-8 passes with a radix-4/2 tail) on a per-warp
// shared-memory buffer (split re/im, bank-swizzled so that the strided stores of the early passes are
// conflict free) plus the usual even/odd post-/pre-processing.  Twiddles: the even/odd split uses one
// table tw[m] = exp(-2 pi i m / n_fft), m < n_fft; every FFT pass has its own small table laid out
// [r - 1][k] so that consecutive lanes read consecutive entries (indexing the big table with k*r*step
// made every twiddle load an 8-16 way bank conflict: 43 % of all shared-memory wavefronts, ncu).
//
// Every pass is split in a "load+butterfly" half that only reads the buffer and a "store" half that
// only writes it; the caller separates the halves with a warp barrier (the values live in registers
// in between).
#pragma once

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#include <cmath>
#endif

namespace neraf {
namespace gl {

struct C2 {
  float x, y;
};

GL_HD C2 cadd(C2 a, C2 b) { return C2{a.x + b.x, a.y + b.y}; }
GL_HD C2 csub(C2 a, C2 b) { return C2{a.x - b.x, a.y - b.y}; }
GL_HD C2 cmul(C2 a, C2 b) { return C2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
GL_HD C2 cconj(C2 a) { return C2{a.x, -a.y}; }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
GL_HD C2 rot90(C2 a) {
  return INV ? C2{-a.y, a.x} : C2{a.y, -a.x};
}

// Bank swizzle of the per-warp FFT buffers (4-byte elements, split re/im): low index bits are XOR-ed with higher
// ones so that the three access patterns of the Stockham passes are conflict free --
//   contiguous (loads, last-pass stores), stride 8 (first-pass stores: i = 8 lane + r) and
//   8-element groups 64 apart (second-pass stores: i = 64 g + 8 r + m) --
// without padding (a bijection on [0, H)).  Index-padding (i + i/32) left the second pattern 2-4 way conflicted.
GL_HD int padi(int i) { return i ^ ((i >> 5) & 7) ^ (((i >> 6) & 3) << 3); }

constexpr int padded_size(int h) { return h; }

// ------------------------------------------------------------------------------------------------
// Small DFTs, natural-order output.  Forward: exp(-2 pi i nk/R); INV: conjugate kernel, unnormalised.
// ------------------------------------------------------------------------------------------------
template <bool INV>
GL_HD void dft2(C2* v) {
  const C2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}

template <bool INV>
GL_HD void dft4(C2* v) {
  const C2 a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
  const C2 a2 = cadd(v[1], v[3]), a3 = rot90<INV>(csub(v[1], v[3]));
  v[0] = cadd(a0, a2);
  v[1] = cadd(a1, a3);
  v[2] = csub(a0, a2);
  v[3] = csub(a1, a3);
}

template <bool INV>
GL_HD void dft8(C2* v) {
  C2 e[4] = {v[0], v[2], v[4], v[6]};
  C2 o[4] = {v[1], v[3], v[5], v[7]};
  dft4<INV>(e);
  dft4<INV>(o);
  const float h = 0.70710678118654752440f;
  // o[k] *= exp(-+ 2 pi i k / 8)
  const C2 w1 = INV ? C2{h, h} : C2{h, -h};
  const C2 w3 = INV ? C2{-h, h} : C2{-h, -h};
  o[1] = cmul(o[1], w1);
  o[2] = rot90<INV>(o[2]);
  o[3] = cmul(o[3], w3);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    v[k] = cadd(e[k], o[k]);
    v[k + 4] = csub(e[k], o[k]);
  }
}

template <int R, bool INV>
GL_HD void dftR(C2* v) {
  if (R == 8) dft8<INV>(v);
  else if (R == 4) dft4<INV>(v);
  else dft2<INV>(v);
}

// Number of butterflies each lane owns in a radix-R pass over H points.
template <int R, int H>
struct PassShape {
  static constexpr int BUTTERFLIES = H / R;
  static constexpr int PER_LANE = (BUTTERFLIES + 31) / 32;
};

// One Stockham pass, first half: gather R inputs per butterfly, apply the inter-stage twiddles and
// the radix-R DFT; results stay in `v` (registers).  Ns = product of the radices of earlier passes.
// twp: this pass's twiddle table, twp[(r - 1) * Ns + k] = exp(-2 pi i k r / (Ns R)) (unused when Ns == 1).
template <int R, int H, bool INV>
GL_HD void pass_load(int lane, int Ns, const float* re, const float* im, const C2* twp, C2 (*v)[R]) {
  constexpr int NB = PassShape<R, H>::BUTTERFLIES;
  constexpr int PL = PassShape<R, H>::PER_LANE;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < PL; ++b) {
    const int j = lane + 32 * b;
    if (j < NB) {
      const int k = j % Ns;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int r = 0; r < R; ++r) {
        const int idx = padi(j + r * NB);
        C2 x{re[idx], im[idx]};
        if (r > 0 && Ns > 1) {
          C2 w = twp[(r - 1) * Ns + k];
          if (INV) w.y = -w.y;
          x = cmul(x, w);
        }
        v[b][r] = x;
      }
      dftR<R, INV>(v[b]);
    }
  }
}

// Second half: scatter the R outputs of every butterfly to their autosorted positions.
template <int R, int H>
GL_HD void pass_store(int lane, int Ns, float* re, float* im, const C2 (*v)[R]) {
  constexpr int NB = PassShape<R, H>::BUTTERFLIES;
  constexpr int PL = PassShape<R, H>::PER_LANE;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int b = 0; b < PL; ++b) {
    const int j = lane + 32 * b;
    if (j < NB) {
      const int k = j % Ns;
      const int base = (j / Ns) * Ns * R + k;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int r = 0; r < R; ++r) {
        const int idx = padi(base + r * Ns);
        re[idx] = v[b][r].x;
        im[idx] = v[b][r].y;
      }
    }
  }
}

// Radix schedule for H = 2^p, 32 <= H <= 1024: radix 8 while possible, then one radix 4 or 2.
template <int H>
struct Schedule {
  static constexpr int LOG2 = H == 32 ? 5 : H == 64 ? 6 : H == 128 ? 7 : H == 256 ? 8 : H == 512 ? 9 : H == 1024 ? 10 : -1;
  static_assert(LOG2 > 0, "unsupported FFT size");
  static constexpr int N8 = LOG2 / 3;
  static constexpr int TAIL = LOG2 % 3 == 0 ? 1 : (LOG2 % 3 == 1 ? 2 : 4);
};

// Per-pass twiddle tables, concatenated in pass order (passes with Ns == 1 have none).
template <int H>
struct PassTables {
  static constexpr int N8 = Schedule<H>::N8, TAIL = Schedule<H>::TAIL;
  // offset of the table of radix-8 pass p (p >= 1): 7 * (8 + 64 + ... + 8^(p-1))
  static constexpr int offset8(int p) { return p <= 1 ? 0 : offset8(p - 1) + 7 * ipow8(p - 1); }
  static constexpr int ipow8(int p) { return p == 0 ? 1 : 8 * ipow8(p - 1); }
  static constexpr int TAIL_OFFSET = offset8(N8);
  static constexpr int TOTAL = TAIL_OFFSET + (TAIL > 1 ? (TAIL - 1) * ipow8(N8) : 0);
};

// Entry e of the concatenated tables as (numerator, denominator): exp(-2 pi i num / den).
template <int H>
GL_HD void pass_table_angle(int e, int* num, int* den) {
  int Ns = 8, off = 0;
  for (int p = 1; p < Schedule<H>::N8; ++p) {
    if (e < off + 7 * Ns) {
      const int r = (e - off) / Ns + 1, k = (e - off) % Ns;
      *num = k * r; *den = Ns * 8;
      return;
    }
    off += 7 * Ns;
    Ns *= 8;
  }
  // tail pass (radix 4 or 2) after N8 radix-8 passes: Ns = 8^N8
  const int NsT = PassTables<H>::ipow8(Schedule<H>::N8);
  const int r = (e - off) / NsT + 1, k = (e - off) % NsT;
  *num = k * r; *den = NsT * Schedule<H>::TAIL;
}

// ------------------------------------------------------------------------------------------------
// Frame load: z[j] = x[2j] + i x[2j+1] with x[n] = D[reflect(t*hop + n - H)] * win[n]
// (torch.stft center=True, pad_mode="reflect"; window zero-padded and centred to n_fft).
// ------------------------------------------------------------------------------------------------
GL_HD int reflect_index(int idx, int L) {
  if (idx < 0) idx = -idx;
  if (idx >= L) idx = 2 * (L - 1) - idx;
  return idx;
}

// [j_lo, j_hi): complex samples j = n/2 that touch the window support [left, left + win_length) (zero elsewhere).
template <int H>
GL_HD void load_frame(int lane, int t, int hop, int L, int j_lo, int j_hi, const float* D, const float* win, float* re,
                      float* im) {
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int j = lane; j < H; j += 32) {
    float x0 = 0.f, x1 = 0.f;
    if (j >= j_lo && j < j_hi) {
      const int n0 = 2 * j;
      const int base = t * hop + n0 - H;
      x0 = D[reflect_index(base, L)] * win[n0];
      x1 = D[reflect_index(base + 1, L)] * win[n0 + 1];
    }
    const int p = padi(j);
    re[p] = x0;
    im[p] = x1;
  }
}

// ------------------------------------------------------------------------------------------------
// Spectrum step.  Input: Z = FFT_H(z) in (re, im).  For every bin k of the n_fft-point real transform
//   X[k]   = rebuilt spectrum (torch.stft) of the frame,
//   ang    = X / (|X| + 1e-16)                         (torchaudio griffinlim phase normalisation)
//   Y[k]   = mag[k] * ang, imaginary parts of DC / Nyquist dropped (C2R semantics of torch.istft)
// and the buffer is overwritten with Z' such that IFFT_H(Z') = y[2n] + i y[2n+1], y = n_fft * irfft(Y).
// Pairs (k, H-k) are handled by one lane, so the pass is in place.
// ------------------------------------------------------------------------------------------------
// x / (|x| + 1e-16) (torchaudio functional.py:345).  Evaluated as x * rsqrt(max(|x|^2, 1e-32)): identical to rounding
// for |x| >> 1e-16, 0 for x == 0 like the reference, and differs only for 0 < |x| < ~1e-15 where the reference shrinks
// the unit phasor (never reached by STFTs of fp32 audio).  One MUFU.RSQ instead of an IEEE sqrt + division.
GL_HD C2 unit_phase(C2 x) {
  const float s = x.x * x.x + x.y * x.y;
#if defined(__CUDA_ARCH__)
  const float inv = rsqrtf(fmaxf(s, 1e-32f));
#else
  const float inv = 1.f / std::sqrt(s > 1e-32f ? s : 1e-32f);
#endif
  return C2{x.x * inv, x.y * inv};
}

// Shared tail: given Y[k], Y[H-k] write Z'[k], Z'[H-k].
template <int H>
GL_HD void write_inverse_pair(int k, C2 yk, C2 ykk, const C2* tw, float* re, float* im) {
  const int kk = H - k;
  const C2 e = cadd(yk, cconj(ykk));
  const C2 o = cmul(csub(yk, cconj(ykk)), cconj(tw[k]));     // * exp(+2 pi i k / n_fft)
  // Z'[k] = E + i O ;  Z'[H-k] = conj(E) + i conj(O)
  const int pk = padi(k), pkk = padi(kk);
  re[pk] = e.x - o.y;
  im[pk] = e.y + o.x;
  re[pkk] = e.x + o.y;
  im[pkk] = -e.y + o.x;
}

// The magnitudes a lane needs in spectrum_step (bins k = lane + 32 i and H - k), fetched into registers ahead of the
// forward FFT so that their global-memory latency hides behind it.
template <int H>
struct MagRegs {
  static constexpr int NK = (H / 2) / 32 + 1;
  float k[NK], kk[NK];
};

template <int H>
GL_HD void load_mag(int lane, const float* mag_row, MagRegs<H>& m) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < MagRegs<H>::NK; ++i) {
    const int k = lane + 32 * i;
    const bool ok = k <= H / 2;
    m.k[i] = ok ? mag_row[k] : 0.f;
    m.kk[i] = ok ? mag_row[H - k] : 0.f;
  }
}

template <int H>
GL_HD void spectrum_step(int lane, const C2* tw, const MagRegs<H>& m, float* re, float* im) {
  // k = 0 (DC / Nyquist pair) is done by the lane that owns index 0.
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < MagRegs<H>::NK; ++i) {
    const int k = lane + 32 * i;
    if (k > H / 2) continue;
    if (k == 0) {
      const float zr = re[0], zi = im[0];
      const float x0 = zr + zi, xh = zr - zi;
#if defined(__CUDA_ARCH__)
      const float a0 = x0 / (fabsf(x0) + 1e-16f), ah = xh / (fabsf(xh) + 1e-16f);
#else
      const float a0 = x0 / (std::fabs(x0) + 1e-16f), ah = xh / (std::fabs(xh) + 1e-16f);
#endif
      const float y0 = m.k[i] * a0, yh = m.kk[i] * ah;
      re[0] = y0 + yh;
      im[0] = y0 - yh;
    } else {
      const int kk = H - k;
      const int pk = padi(k), pkk = padi(kk);
      const C2 zk{re[pk], im[pk]}, zkk{re[pkk], im[pkk]};
      const C2 zc = cconj(zkk);
      const C2 xe{0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y)};
      const C2 d = csub(zk, zc);
      const C2 xo{0.5f * d.y, -0.5f * d.x};                      // -i (Zk - conj Zkk) / 2
      const C2 tt = cmul(tw[k], xo);
      const C2 xk = cadd(xe, tt);
      const C2 xkk = cconj(csub(xe, tt));
      const C2 ak = unit_phase(xk), akk = unit_phase(xkk);
      const float mk = m.k[i], mkk = m.kk[i];
      write_inverse_pair<H>(k, C2{mk * ak.x, mk * ak.y}, C2{mkk * akk.x, mkk * akk.y}, tw, re, im);
    }
  }
}

// Iteration 0: Y[k] = mag[k] * init[k] with an arbitrary complex start "phase" (torch.rand cfloat or ones).
// init_row: interleaved complex, element k at init_row[2*k*stride], or nullptr for all-ones.
template <int H>
GL_HD void init_step(int lane, const C2* tw, const float* mag_row, const float* init_row, long long init_stride,
                     float* re, float* im) {
  for (int k = lane; k <= H / 2; k += 32) {
    const int kk = H - k;
    C2 ik{1.f, 0.f}, ikk{1.f, 0.f};
    if (init_row) {
      ik = C2{init_row[2 * k * init_stride], init_row[2 * k * init_stride + 1]};
      ikk = C2{init_row[2 * kk * init_stride], init_row[2 * kk * init_stride + 1]};
    }
    const float mk = mag_row[k], mkk = mag_row[kk];
    if (k == 0) {
      const float y0 = mk * ik.x, yh = mkk * ikk.x;              // imaginary parts of DC / Nyquist are ignored
      re[0] = y0 + yh;
      im[0] = y0 - yh;
    } else {
      write_inverse_pair<H>(k, C2{mk * ik.x, mk * ik.y}, C2{mkk * ikk.x, mkk * ikk.y}, tw, re, im);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Overlap-add of the windowed inverse frame: ACC[t*hop + n - H] += scale * win[n] * y[n]
// (torch.istft: irfft * window, fold, trimmed by n_fft/2 on both sides; the window-envelope division
// happens once per iteration on the whole waveform).
// ------------------------------------------------------------------------------------------------
template <int H>
GL_HD void ola_frame(int lane, int t, int hop, int L, int j_lo, int j_hi, const float* win, const float* re, const float* im,
                     float scale, float* ACC) {
#if defined(__CUDA_ARCH__)
#pragma unroll 4
#endif
  for (int j = j_lo + lane; j < j_hi; j += 32) {               // outside the window support nothing is added
    const int n0 = 2 * j;
    const int p = padi(j);
    const int idx = t * hop + n0 - H;
    const float w0 = win[n0], w1 = win[n0 + 1];
    if (idx >= 0 && idx < L) ACC[idx] += re[p] * (w0 * scale);
    if (idx + 1 >= 0 && idx + 1 < L) ACC[idx + 1] += im[p] * (w1 * scale);
  }
}

}  // namespace gl
}  // namespace neraf
