// Per-element code of the grid-feature producer's data-movement and normalisation kernels (gridnet.cu), written so
// that the SAME functions compile for the device and for the host: tests/csrc/gridnet_host.cpp loops them over the
// whole index space on the CPU and the not-gpu tests compare the result -- and the whole ResNet3D assembled from
// them by neraf_b200/gridnet.py -- with the reference's NeRAF_resnet3d.py run by torch.
//
// Layout: an activation is a matrix (V, C): row v = (d * H + h) * W + w is a voxel, the C channels of a voxel are
// contiguous ("channels last"), row stride ld >= C.  Elements are fp32 or bf16 (bf16_t: the same 16 bits as
// __nv_bfloat16, round-to-nearest-even).  A convolution is im2col (gather) + GEMM; column j of the gathered matrix
// is ((kd * k + kh) * k + kw) * C + c, so a voxel's channel vector is copied as a unit.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GN_HD __host__ __device__ __forceinline__
#else
#define GN_HD inline
#endif

namespace neraf {
namespace gridnet {

struct bf16_t { uint16_t bits; };

GN_HD float bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
GN_HD uint32_t float_to_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
GN_HD float to_float(float v) { return v; }
GN_HD float to_float(bf16_t v) { return bits_to_float((uint32_t)v.bits << 16); }
GN_HD void from_float(float v, float* out) { *out = v; }
GN_HD void from_float(float v, bf16_t* out) {                 // round to nearest even, NaN stays NaN
  uint32_t u = float_to_bits(v);
  if ((u & 0x7fffffffu) > 0x7f800000u) { out->bits = 0x7fc0; return; }
  u += 0x7fffu + ((u >> 16) & 1u);
  out->bits = (uint16_t)(u >> 16);
}

// Cubic-window geometry shared by the convolutions and the max pooling (NeRAF_resnet3d.py:119,122,88-89,83).
struct Window {
  int in_d, in_h, in_w;      // input extent
  int C;                     // channels
  int k, stride, pad;
  int out_d, out_h, out_w;   // (in + 2 pad - k) / stride + 1
};

GN_HD int out_extent(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

GN_HD Window make_window(int in_d, int in_h, int in_w, int C, int k, int stride, int pad) {
  Window w;
  w.in_d = in_d; w.in_h = in_h; w.in_w = in_w; w.C = C; w.k = k; w.stride = stride; w.pad = pad;
  w.out_d = out_extent(in_d, k, stride, pad);
  w.out_h = out_extent(in_h, k, stride, pad);
  w.out_w = out_extent(in_w, k, stride, pad);
  return w;
}

// idx = q * d + r.  Voxel counts and -- for every tensor of the reference's networks -- element counts fit 32 bits, and a
// 64-bit division costs the GPU several times a 32-bit one (these kernels are otherwise a handful of instructions per
// element), so the 32-bit form is taken whenever both operands allow it.
GN_HD void split_index(long long idx, long long d, long long* q, int* r) {
  if ((((unsigned long long)idx | (unsigned long long)d) >> 32) == 0) {
    const unsigned a = (unsigned)idx, b = (unsigned)d, qq = a / b;
    *q = (long long)qq;
    *r = (int)(a - qq * b);
  } else {
    const long long qq = idx / d;
    *q = qq;
    *r = (int)(idx - qq * d);
  }
}
// voxel index (< 2^31, checked by the callers) -> (d, h, w) for an extent (., H, W)
GN_HD void split_voxel(long long v, int H, int W, int* d, int* h, int* w) {
  const unsigned u = (unsigned)v, t = u / (unsigned)W;
  *w = (int)(u - t * (unsigned)W);
  const unsigned dd = t / (unsigned)H;
  *h = (int)(t - dd * (unsigned)H);
  *d = (int)dd;
}

GN_HD long long in_voxels(const Window& w) { return (long long)w.in_d * w.in_h * w.in_w; }
GN_HD long long out_voxels(const Window& w) { return (long long)w.out_d * w.out_h * w.out_w; }

// ---- im2col: element (v_out, j) of the gathered matrix, j in [0, ld): pad columns (j >= k^3 C) read as zero ----------
// The input is addressed as in[v * voxel_stride + c * channel_stride]: (V, C) row-major activations use (ld, 1), the
// reference's channels-first grid (1, C, D, H, W) uses (1, V).
template <class Tin, class Tout>
GN_HD void im2col_element(const Window& w, const Tin* in, long long voxel_stride, long long channel_stride, Tout* col,
                          long long ld, long long idx) {
  long long v;
  int j;
  split_index(idx, ld, &v, &j);
  float val = 0.f;
  const int K = w.k * w.k * w.k * w.C;
  if (j < K) {
    const int kidx = j / w.C, c = j - kidx * w.C;
    const int kw = kidx % w.k, kh = (kidx / w.k) % w.k, kd = kidx / (w.k * w.k);
    int ow, oh, od;
    split_voxel(v, w.out_h, w.out_w, &od, &oh, &ow);
    const int id = od * w.stride - w.pad + kd, ih = oh * w.stride - w.pad + kh, iw = ow * w.stride - w.pad + kw;
    if (id >= 0 && id < w.in_d && ih >= 0 && ih < w.in_h && iw >= 0 && iw < w.in_w) {
      const long long vin = ((long long)id * w.in_h + ih) * w.in_w + iw;
      val = to_float(in[vin * voxel_stride + c * channel_stride]);
    }
  }
  from_float(val, col + idx);
}

// ---- col2im (the data gradient of a convolution, gather form): dx[v_in, c] = sum over the window positions that
// read voxel v_in of dcol[v_out, kidx * C + c].  fp32 accumulation, fixed order (kd, kh, kw ascending): deterministic.
template <class T>
GN_HD void col2im_element(const Window& w, const T* dcol, long long ld_col, T* dx, long long ld_dx, long long idx) {
  long long v;
  int c, iw, ih, id;
  split_index(idx, w.C, &v, &c);
  split_voxel(v, w.in_h, w.in_w, &id, &ih, &iw);
  float acc = 0.f;
  for (int kd = 0; kd < w.k; ++kd) {
    const int td = id + w.pad - kd;
    if (td < 0 || td % w.stride) continue;
    const int od = td / w.stride;
    if (od >= w.out_d) continue;
    for (int kh = 0; kh < w.k; ++kh) {
      const int th = ih + w.pad - kh;
      if (th < 0 || th % w.stride) continue;
      const int oh = th / w.stride;
      if (oh >= w.out_h) continue;
      for (int kw = 0; kw < w.k; ++kw) {
        const int tw = iw + w.pad - kw;
        if (tw < 0 || tw % w.stride) continue;
        const int ow = tw / w.stride;
        if (ow >= w.out_w) continue;
        const long long vout = ((long long)od * w.out_h + oh) * w.out_w + ow;
        const int kidx = (kd * w.k + kh) * w.k + kw;
        acc += to_float(dcol[vout * ld_col + (long long)kidx * w.C + c]);
      }
    }
  }
  from_float(acc, dx + v * ld_dx + c);
}

// ---- 16-byte forms of the two gathers for bf16 activations whose channel count is a multiple of 8: eight channels of
// one voxel move as one 128-bit word (the scalar forms issue 2-byte accesses; these are the HBM-bound passes that move
// the most bytes: 0.46 GB for the stem's gathered matrix of a 128^3 grid) -------------------------------------------
struct alignas(16) vec8_t { uint32_t w[4]; };

GN_HD bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

// usable when input and output are bf16, rows are whole 16-byte words and the channels are contiguous
GN_HD bool gather_can_vec8(const Window& w, bool all_bf16, long long in_row_stride, long long channel_stride,
                           long long ld_col, const void* a, const void* b) {
  return all_bf16 && channel_stride == 1 && w.C % 8 == 0 && in_row_stride % 8 == 0 && ld_col % 8 == 0 && aligned16(a) &&
         aligned16(b);
}

// unit idx = (v_out, j8): columns [8 j8, 8 j8 + 8) of row v_out
GN_HD void im2col_vec8_element(const Window& w, const bf16_t* in, long long voxel_stride, bf16_t* col, long long ld,
                               long long idx) {
  long long v;
  int j;
  split_index(idx, ld / 8, &v, &j);
  j *= 8;
  vec8_t val = {{0u, 0u, 0u, 0u}};
  const int K = w.k * w.k * w.k * w.C;
  if (j < K) {
    const int kidx = j / w.C, c = j - kidx * w.C;
    const int kw = kidx % w.k, kh = (kidx / w.k) % w.k, kd = kidx / (w.k * w.k);
    int ow, oh, od;
    split_voxel(v, w.out_h, w.out_w, &od, &oh, &ow);
    const int id = od * w.stride - w.pad + kd, ih = oh * w.stride - w.pad + kh, iw = ow * w.stride - w.pad + kw;
    if (id >= 0 && id < w.in_d && ih >= 0 && ih < w.in_h && iw >= 0 && iw < w.in_w) {
      const long long vin = ((long long)id * w.in_h + ih) * w.in_w + iw;
      val = *reinterpret_cast<const vec8_t*>(in + vin * voxel_stride + c);
    }
  }
  *reinterpret_cast<vec8_t*>(col + v * ld + j) = val;
}

// unit idx = (v_in, c8): channels [8 c8, 8 c8 + 8) of input voxel v_in; same summation order as col2im_element
GN_HD void col2im_vec8_element(const Window& w, const bf16_t* dcol, long long ld_col, bf16_t* dx, long long ld_dx,
                               long long idx) {
  long long v;
  int c, iw, ih, id;
  split_index(idx, w.C / 8, &v, &c);
  c *= 8;
  split_voxel(v, w.in_h, w.in_w, &id, &ih, &iw);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int kd = 0; kd < w.k; ++kd) {
    const int td = id + w.pad - kd;
    if (td < 0 || td % w.stride) continue;
    const int od = td / w.stride;
    if (od >= w.out_d) continue;
    for (int kh = 0; kh < w.k; ++kh) {
      const int th = ih + w.pad - kh;
      if (th < 0 || th % w.stride) continue;
      const int oh = th / w.stride;
      if (oh >= w.out_h) continue;
      for (int kw = 0; kw < w.k; ++kw) {
        const int tw = iw + w.pad - kw;
        if (tw < 0 || tw % w.stride) continue;
        const int ow = tw / w.stride;
        if (ow >= w.out_w) continue;
        const long long vout = ((long long)od * w.out_h + oh) * w.out_w + ow;
        const int kidx = (kd * w.k + kh) * w.k + kw;
        const vec8_t q = *reinterpret_cast<const vec8_t*>(dcol + vout * ld_col + (long long)kidx * w.C + c);
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += bits_to_float(q.w[i] << 16);
          acc[2 * i + 1] += bits_to_float(q.w[i] & 0xffff0000u);
        }
      }
    }
  }
  vec8_t out;
  for (int i = 0; i < 4; ++i) {
    bf16_t lo, hi;
    from_float(acc[2 * i], &lo);
    from_float(acc[2 * i + 1], &hi);
    out.w[i] = (uint32_t)lo.bits | ((uint32_t)hi.bits << 16);
  }
  *reinterpret_cast<vec8_t*>(dx + v * ld_dx + c) = out;
}

// ---- max pooling (nn.MaxPool3d(3, 2, 1), NeRAF_resnet3d.py:122): first maximum in (d, h, w) scan order, like
// torch's CPU kernel (max_pool3d_with_indices: `val > maxval`), so that gradients of tied maxima -- frequent after a
// ReLU -- take the same route as in the reference.  argmax holds the input voxel index.
template <class T>
GN_HD void maxpool_element(const Window& w, const T* x, long long ld_x, T* y, long long ld_y, int32_t* argmax,
                           long long idx) {
  long long v;
  int c, ow, oh, od;
  split_index(idx, w.C, &v, &c);
  split_voxel(v, w.out_h, w.out_w, &od, &oh, &ow);
  float best = 0.f;
  long long arg = -1;
  for (int kd = 0; kd < w.k; ++kd) {
    const int id = od * w.stride - w.pad + kd;
    if (id < 0 || id >= w.in_d) continue;
    for (int kh = 0; kh < w.k; ++kh) {
      const int ih = oh * w.stride - w.pad + kh;
      if (ih < 0 || ih >= w.in_h) continue;
      for (int kw = 0; kw < w.k; ++kw) {
        const int iw = ow * w.stride - w.pad + kw;
        if (iw < 0 || iw >= w.in_w) continue;
        const long long vin = ((long long)id * w.in_h + ih) * w.in_w + iw;
        const float val = to_float(x[vin * ld_x + c]);
        if (arg < 0 || val > best || val != val) { best = val; arg = vin; }
      }
    }
  }
  from_float(best, y + v * ld_y + c);
  argmax[v * w.C + c] = (int32_t)arg;
}

// gradient of the pooling, gather form: every input voxel sums the gradients of the windows it won; dy2 (optional) is
// a second gradient of the pooled tensor (its two consumers: the first block's convolution and its shortcut)
template <class T>
GN_HD void maxpool_backward_element(const Window& w, const T* dy, const T* dy2, long long ld_dy, const int32_t* argmax,
                                    T* dx, long long ld_dx, long long idx) {
  long long v;
  int c, iw, ih, id;
  split_index(idx, w.C, &v, &c);
  split_voxel(v, w.in_h, w.in_w, &id, &ih, &iw);
  // the windows that contain this voxel: o in [ceil((i + pad - k + 1) / s), floor((i + pad) / s)] per axis (at most
  // ceil(k / s)^3 of them: 8 for the 3/2 pooling), walked from the highest o down -- the order in which the plain k^3
  // loop with its stride test meets them, so the sum is bit-identical to that form, at 8 instead of 27 iterations and
  // without a division per iteration (the kernel was 540 us of a 7 ms producer step)
  float acc = 0.f;
  const int s_ = w.stride;
  const int td = id + w.pad - w.k + 1, th = ih + w.pad - w.k + 1, tw = iw + w.pad - w.k + 1;
  const int od0 = td <= 0 ? 0 : (td + s_ - 1) / s_, oh0 = th <= 0 ? 0 : (th + s_ - 1) / s_, ow0 = tw <= 0 ? 0 : (tw + s_ - 1) / s_;
  int od1 = (id + w.pad) / s_, oh1 = (ih + w.pad) / s_, ow1 = (iw + w.pad) / s_;
  if (od1 > w.out_d - 1) od1 = w.out_d - 1;
  if (oh1 > w.out_h - 1) oh1 = w.out_h - 1;
  if (ow1 > w.out_w - 1) ow1 = w.out_w - 1;
  for (int od = od1; od >= od0; --od)
    for (int oh = oh1; oh >= oh0; --oh)
      for (int ow = ow1; ow >= ow0; --ow) {
        const long long vout = ((long long)od * w.out_h + oh) * w.out_w + ow;
        if (argmax[vout * w.C + c] == (int32_t)v) {
          acc += to_float(dy[vout * ld_dy + c]);
          if (dy2) acc += to_float(dy2[vout * ld_dy + c]);
        }
      }
  from_float(acc, dx + v * ld_dx + c);
}

// ---- convolution weights: the parameter (c_out, c_in, k, k, k) <-> the GEMM operand (c_out, k^3 c_in), element
// (co, j) with j in [0, ld) (pad columns zero) ---------------------------------------------------------------------
template <class Tout>
GN_HD void pack_weight_element(const float* w, long long c_in, long long k3, Tout* out, long long ld, long long idx) {
  long long co;
  int jj;
  split_index(idx, ld, &co, &jj);
  const long long j = jj;
  float val = 0.f;
  if (j < k3 * c_in) {
    const long long kidx = j / c_in, ci = j - kidx * c_in;
    val = w[(co * c_in + ci) * k3 + kidx];
  }
  from_float(val, out + idx);
}
// the weight gradient back in the parameter's layout: element idx of dw (c_out, c_in, k^3).  The GEMM that formed it may
// have been split over the voxels (split-K: the contraction runs over up to 262 144 voxels while c_out x k^3 c_in is a
// handful of tiles): dw_mat then holds n_partials matrices partial_stride floats apart, summed here in a fixed order.
GN_HD void unpack_wgrad_element(const float* dw_mat, long long ld, long long c_in, long long k3, int n_partials,
                                long long partial_stride, float* dw, long long idx) {
  long long t, co;
  int kidx, ci;
  split_index(idx, k3, &t, &kidx);
  split_index(t, c_in, &co, &ci);
  const float* src = dw_mat + co * ld + (long long)kidx * c_in + ci;
  float acc = src[0];
  for (int s = 1; s < n_partials; ++s) acc += src[s * partial_stride];
  dw[idx] = acc;
}

// ---- batch normalisation (nn.BatchNorm3d, NeRAF_resnet3d.py:120 and the blocks) ------------------------------------
// column sums over rows [r0, r1) stepping by `step`: fp32 inside a strip of <= 64 visited rows, fp64 across strips
template <class T>
GN_HD void column_sums_partial(const T* x, long long ld, long long c, long long r0, long long r1, long long step,
                               double* s, double* ss) {
  double S = 0.0, SS = 0.0;
  long long r = r0;
  while (r < r1) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 64 && r < r1; ++i, r += step) {
      const float v = to_float(x[r * ld + c]);
      a += v; b = fmaf(v, v, b);
    }
    S += (double)a; SS += (double)b;
  }
  *s = S; *ss = SS;
}

// One channel of the statistics: training -> batch mean / biased variance (what normalises, torch batch_norm) and the
// running estimates updated with the UNBIASED variance and `momentum` (nn.BatchNorm3d default 0.1; momentum 0 leaves
// them alone); evaluation -> the running estimates themselves.
GN_HD void bn_finalize_channel(const double* sums, long long V, long long C, float eps, float momentum, int training,
                               float* running_mean, float* running_var, float* mean, float* invstd, long long c) {
  if (training) {
    const double m = sums[c] / (double)V;
    double var = sums[C + c] / (double)V - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    if (invstd) invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (momentum > 0.f && running_mean && running_var) {
      const double unbiased = V > 1 ? var * (double)V / (double)(V - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean[c] = running_mean[c];
    if (invstd) invstd[c] = 1.f / sqrtf(running_var[c] + eps);
  }
}

// y = [relu]( gamma (x - mean) invstd + beta [+ residual] )  -- the tail of every conv unit (NeRAF_resnet3d.py:98-113)
GN_HD float bn_apply_value(float x, float mean, float invstd, float gamma, float beta, bool has_res, float res, int relu) {
  float v = (x - mean) * invstd * gamma + beta;
  if (has_res) v += res;
  if (relu && !(v > 0.f)) v = 0.f;
  return v;
}

template <class T>
GN_HD void bn_apply_element(const T* x, long long ld_x, long long C, const float* mean, const float* invstd,
                            const float* gamma, const float* beta, const T* residual, long long ld_res, int relu, T* y,
                            long long ld_y, long long idx) {
  long long r;
  int c;
  split_index(idx, C, &r, &c);
  const float res = residual ? to_float(residual[r * ld_res + c]) : 0.f;
  from_float(bn_apply_value(to_float(x[r * ld_x + c]), mean[c], invstd[c], gamma[c], beta[c], residual != nullptr, res, relu),
             y + r * ld_y + c);
}

// 16-byte forms for bf16 matrices whose rows are whole 16-byte words (all of the network's activations): unit idx =
// (r, c8), eight channels per thread, the same per-channel arithmetic as the scalar forms (same bits).
GN_HD bool rows_can_vec8(bool is_bf16, long long C, long long ld_a, long long ld_b, long long ld_c, const void* a,
                         const void* b, const void* c, const void* p0, const void* p1, const void* p2, const void* p3) {
  return is_bf16 && C % 8 == 0 && ld_a % 8 == 0 && ld_b % 8 == 0 && ld_c % 8 == 0 && aligned16(a) && aligned16(b) &&
         aligned16(c) && aligned16(p0) && aligned16(p1) && aligned16(p2) && aligned16(p3);      // p*: per-channel vectors
}
struct alignas(16) f4_t { float v[4]; };
struct alignas(16) d2_t { double v[2]; };
// eight consecutive per-channel values as two / four 16-byte loads (p is 16-byte aligned: checked by the dispatch)
GN_HD void load8(const float* p, float* out) {
  const f4_t a = *reinterpret_cast<const f4_t*>(p), b = *reinterpret_cast<const f4_t*>(p + 4);
  for (int i = 0; i < 4; ++i) { out[i] = a.v[i]; out[4 + i] = b.v[i]; }
}
GN_HD void load8(const double* p, double* out) {
  for (int i = 0; i < 4; ++i) {
    const d2_t a = *reinterpret_cast<const d2_t*>(p + 2 * i);
    out[2 * i] = a.v[0]; out[2 * i + 1] = a.v[1];
  }
}
GN_HD void load8(const bf16_t* p, float* out) {
  const vec8_t q = *reinterpret_cast<const vec8_t*>(p);          // by value: one 128-bit load
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = bits_to_float(q.w[i] << 16);
    out[2 * i + 1] = bits_to_float(q.w[i] & 0xffff0000u);
  }
}
GN_HD void unpack8(const vec8_t& q, float* f) {
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = bits_to_float(q.w[i] << 16);
    f[2 * i + 1] = bits_to_float(q.w[i] & 0xffff0000u);
  }
}
GN_HD vec8_t pack8(const float* f) {
  vec8_t out;
  for (int i = 0; i < 4; ++i) {
    bf16_t lo, hi;
    from_float(f[2 * i], &lo);
    from_float(f[2 * i + 1], &hi);
    out.w[i] = (uint32_t)lo.bits | ((uint32_t)hi.bits << 16);
  }
  return out;
}

// Pooling gradient, eight channels of one input voxel per thread: unit idx = (v_in, c8).  Per channel the same windows in
// the same order as maxpool_backward_element (same bits); the 32-byte argmax row segment and the 16-byte gradient
// segments of a window are read once for all eight channels (the scalar form was 270 us for the (64^3, 64) tensor).
struct alignas(16) i4_t { int32_t v[4]; };
GN_HD void maxpool_backward_vec8_element(const Window& w, const bf16_t* dy, const bf16_t* dy2, long long ld_dy,
                                         const int32_t* argmax, bf16_t* dx, long long ld_dx, long long idx) {
  long long v;
  int c, iw, ih, id;
  split_index(idx, w.C / 8, &v, &c);
  c *= 8;
  split_voxel(v, w.in_h, w.in_w, &id, &ih, &iw);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int s_ = w.stride;
  const int td = id + w.pad - w.k + 1, th = ih + w.pad - w.k + 1, tw = iw + w.pad - w.k + 1;
  const int od0 = td <= 0 ? 0 : (td + s_ - 1) / s_, oh0 = th <= 0 ? 0 : (th + s_ - 1) / s_, ow0 = tw <= 0 ? 0 : (tw + s_ - 1) / s_;
  int od1 = (id + w.pad) / s_, oh1 = (ih + w.pad) / s_, ow1 = (iw + w.pad) / s_;
  if (od1 > w.out_d - 1) od1 = w.out_d - 1;
  if (oh1 > w.out_h - 1) oh1 = w.out_h - 1;
  if (ow1 > w.out_w - 1) ow1 = w.out_w - 1;
  for (int od = od1; od >= od0; --od)
    for (int oh = oh1; oh >= oh0; --oh)
      for (int ow = ow1; ow >= ow0; --ow) {
        const long long vout = ((long long)od * w.out_h + oh) * w.out_w + ow;
        const i4_t a0 = *reinterpret_cast<const i4_t*>(argmax + vout * w.C + c);
        const i4_t a1 = *reinterpret_cast<const i4_t*>(argmax + vout * w.C + c + 4);
        bool won[8];
        bool any = false;
        for (int i = 0; i < 4; ++i) {
          won[i] = a0.v[i] == (int32_t)v;
          won[4 + i] = a1.v[i] == (int32_t)v;
          any = any || won[i] || won[4 + i];
        }
        if (!any) continue;
        float g[8], g2[8];
        load8(dy + vout * ld_dy + c, g);
        if (dy2) load8(dy2 + vout * ld_dy + c, g2);
        for (int i = 0; i < 8; ++i)
          if (won[i]) {
            acc[i] += g[i];
            if (dy2) acc[i] += g2[i];
          }
      }
  *reinterpret_cast<vec8_t*>(dx + v * ld_dx + c) = pack8(acc);
}

// ---- 128-bit forms of the two batch-norm reductions: eight consecutive channels [c, c + 8) of rows r0, r0 + step, ...
// per call, PER CHANNEL the arithmetic of column_sums_partial / bn_backward_partial (fp32 strips of 64 rows into fp64
// sums: the same bits for the same (r0, r1, step)).  One 16-byte load per row and operand instead of eight 2-byte ones.
GN_HD void column_sums_partial8(const bf16_t* x, long long ld, long long c, long long r0, long long r1, long long step,
                                double* s, double* ss) {
  double S[8], SS[8];
  for (int k = 0; k < 8; ++k) { S[k] = 0.0; SS[k] = 0.0; }
  long long r = r0;
  while (r < r1) {
    float a[8], b[8];
    for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; }
    for (int i = 0; i < 64 && r < r1; ++i, r += step) {
      float v[8];
      load8(x + r * ld + c, v);
      for (int k = 0; k < 8; ++k) { a[k] += v[k]; b[k] = fmaf(v[k], v[k], b[k]); }
    }
    for (int k = 0; k < 8; ++k) { S[k] += (double)a[k]; SS[k] += (double)b[k]; }
  }
  for (int k = 0; k < 8; ++k) { s[k] = S[k]; ss[k] = SS[k]; }
}

GN_HD void bn_backward_partial8(const bf16_t* dy, const bf16_t* dy2, const bf16_t* y, const bf16_t* x, long long ld,
                                const float* mean, const float* invstd, bf16_t* g_out, long long c, long long r0,
                                long long r1, long long step, double* s_g, double* s_gx) {
  float m[8], is[8];
  load8(mean + c, m);
  load8(invstd + c, is);
  double S[8], SX[8];
  for (int k = 0; k < 8; ++k) { S[k] = 0.0; SX[k] = 0.0; }
  long long r = r0;
  while (r < r1) {
    float a[8], b[8];
    for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; }
    for (int i = 0; i < 64 && r < r1; ++i, r += step) {
      float g[8], xv[8];
      load8(dy + r * ld + c, g);
      if (dy2) {
        float g2[8];
        load8(dy2 + r * ld + c, g2);
        for (int k = 0; k < 8; ++k) g[k] += g2[k];
      }
      if (y) {
        float yv[8];
        load8(y + r * ld + c, yv);
        for (int k = 0; k < 8; ++k) if (!(yv[k] > 0.f)) g[k] = 0.f;
      }
      const vec8_t stored = pack8(g);
      *reinterpret_cast<vec8_t*>(g_out + r * ld + c) = stored;
      unpack8(stored, g);                         // the sums see what the second pass will read
      load8(x + r * ld + c, xv);
      for (int k = 0; k < 8; ++k) {
        const float xh = (xv[k] - m[k]) * is[k];
        a[k] += g[k]; b[k] = fmaf(g[k], xh, b[k]);
      }
    }
    for (int k = 0; k < 8; ++k) { S[k] += (double)a[k]; SX[k] += (double)b[k]; }
  }
  for (int k = 0; k < 8; ++k) { s_g[k] = S[k]; s_gx[k] = SX[k]; }
}

GN_HD void bn_apply_vec8_element(const bf16_t* x, long long ld_x, long long C, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, const bf16_t* residual, long long ld_res,
                                 int relu, bf16_t* y, long long ld_y, long long idx) {
  long long r;
  int c;
  split_index(idx, C / 8, &r, &c);
  c *= 8;
  float xv[8], rv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, out[8], m[8], is[8], ga[8], be[8];
  load8(x + r * ld_x + c, xv);
  if (residual) load8(residual + r * ld_res + c, rv);
  load8(mean + c, m); load8(invstd + c, is); load8(gamma + c, ga); load8(beta + c, be);
  for (int i = 0; i < 8; ++i) out[i] = bn_apply_value(xv[i], m[i], is[i], ga[i], be[i], residual != nullptr, rv[i], relu);
  *reinterpret_cast<vec8_t*>(y + r * ld_y + c) = pack8(out);
}

// Backward, first pass over rows [r0, r1) of column c: the gradient that reaches the normalisation is
// g = (dy [+ dy2]) * [y > 0] (dy2: the second consumer of this unit's output -- the next block's shortcut; y: the
// unit's output when it ends in a ReLU, else null).  g is stored (the shortcut branch of a residual unit receives
// exactly g) and the two sums of the batch-norm gradient are formed: sum g, sum g xhat.
template <class T>
GN_HD void bn_backward_partial(const T* dy, const T* dy2, const T* y, const T* x, long long ld, const float* mean,
                               const float* invstd, T* g_out, long long c, long long r0, long long r1, long long step,
                               double* s_g, double* s_gx) {
  const float m = mean[c], is = invstd[c];
  double S = 0.0, SX = 0.0;
  long long r = r0;
  while (r < r1) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < 64 && r < r1; ++i, r += step) {
      float g = to_float(dy[r * ld + c]);
      if (dy2) g += to_float(dy2[r * ld + c]);
      if (y && !(to_float(y[r * ld + c]) > 0.f)) g = 0.f;
      T stored;
      from_float(g, &stored);
      g_out[r * ld + c] = stored;
      g = to_float(stored);                       // the sums see what the second pass will read
      const float xh = (to_float(x[r * ld + c]) - m) * is;
      a += g; b = fmaf(g, xh, b);
    }
    S += (double)a; SX += (double)b;
  }
  *s_g = S; *s_gx = SX;
}

// Second pass: dx = gamma invstd (g - (sum g + xhat sum g xhat) / V) with batch statistics, gamma invstd g with the
// running ones (they are constants then).
GN_HD float bn_backward_value(float g, float x, float mean, float invstd, float gamma, double s_g, double s_gx,
                              double inv_v, int training) {
  float v = g;
  if (training) {
    const float xh = (x - mean) * invstd;
    v -= (float)((s_g + (double)xh * s_gx) * inv_v);
  }
  return v * gamma * invstd;
}

template <class T>
GN_HD void bn_backward_element(const T* g, const T* x, long long ld, long long C, const float* mean, const float* invstd,
                               const float* gamma, const double* sums, long long V, int training, T* dx, long long idx) {
  long long r;
  int c;
  split_index(idx, C, &r, &c);
  const double inv_v = 1.0 / (double)V;            // loop-invariant: one fp64 division per thread, not per element
  from_float(bn_backward_value(to_float(g[r * ld + c]), to_float(x[r * ld + c]), mean[c], invstd[c], gamma[c], sums[c],
                               sums[C + c], inv_v, training),
             dx + r * ld + c);
}

GN_HD void bn_backward_vec8_element(const bf16_t* g, const bf16_t* x, long long ld, long long C, const float* mean,
                                    const float* invstd, const float* gamma, const double* sums, long long V,
                                    int training, bf16_t* dx, long long idx) {
  long long r;
  int c;
  split_index(idx, C / 8, &r, &c);
  c *= 8;
  const double inv_v = 1.0 / (double)V;
  float gv[8], xv[8], out[8], m[8], is[8], ga[8];
  double sg[8], sgx[8];
  load8(g + r * ld + c, gv);
  load8(x + r * ld + c, xv);
  load8(mean + c, m); load8(invstd + c, is); load8(gamma + c, ga);
  load8(sums + c, sg); load8(sums + C + c, sgx);
  for (int i = 0; i < 8; ++i) out[i] = bn_backward_value(gv[i], xv[i], m[i], is[i], ga[i], sg[i], sgx[i], inv_v, training);
  *reinterpret_cast<vec8_t*>(dx + r * ld + c) = pack8(out);
}

// out[r, c] = v[c] * scale: the gradient of the global average pooling (nn.AvgPool3d over the whole extent,
// NeRAF_resnet3d.py:141-157) spread back over the voxels
template <class T>
GN_HD void broadcast_rows_element(const float* v, float scale, long long C, T* out, long long ld, long long idx) {
  long long r;
  int c;
  split_index(idx, C, &r, &c);
  from_float(v[c] * scale, out + r * ld + c);
}

}  // namespace gridnet
}  // namespace neraf
