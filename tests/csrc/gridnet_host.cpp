// Drives neraf_b200/csrc/gridnet_core.h -- the per-element code the grid-feature producer's kernels execute -- on the
// CPU, behind entry points with the argument lists of the neraf_grid_* functions of include/neraf_b200.h (host
// pointers instead of device pointers, the stream ignored).  Built as build/libgridnet_host.so by
// __graft_entry__.build(); loaded only by tests/test_gridnet.py, which runs neraf_b200/gridnet.py's assembly of the
// ResNet3D on it and compares with the reference's network.  Test infrastructure only.
#include <cstdint>
#include <cstring>

#include "../../include/neraf_b200.h"
#include "../../neraf_b200/csrc/gridnet_core.h"

using namespace neraf::gridnet;

namespace {
Window window_of(const neraf_window3d* w) {
  return make_window(w->in_d, w->in_h, w->in_w, w->channels, w->k, w->stride, w->pad);
}
constexpr long long kLanes = 8;   // the strip reductions are walked with the kernels' 8 row lanes
}  // namespace

#define HOST_API extern "C" __attribute__((visibility("default")))

HOST_API int gridhost_im2col(const neraf_window3d* wd, const void* in, int32_t in_dtype, int64_t vs, int64_t cs, void* col,
                             int32_t col_dtype, int64_t ld, void*) {
  const Window w = window_of(wd);
  const long long n = out_voxels(w) * ld;
  if (gather_can_vec8(w, in_dtype == NERAF_DT_BF16 && col_dtype == NERAF_DT_BF16, vs, cs, ld, in, col)) {   // as gridnet.cu
    for (long long i = 0; i < n / 8; ++i) im2col_vec8_element(w, (const bf16_t*)in, vs, (bf16_t*)col, ld, i);
    return 0;
  }
  for (long long i = 0; i < n; ++i) {
    if (in_dtype == NERAF_DT_F32 && col_dtype == NERAF_DT_F32) im2col_element(w, (const float*)in, vs, cs, (float*)col, ld, i);
    else if (in_dtype == NERAF_DT_F32) im2col_element(w, (const float*)in, vs, cs, (bf16_t*)col, ld, i);
    else if (col_dtype == NERAF_DT_BF16) im2col_element(w, (const bf16_t*)in, vs, cs, (bf16_t*)col, ld, i);
    else im2col_element(w, (const bf16_t*)in, vs, cs, (float*)col, ld, i);
  }
  return 0;
}

HOST_API int gridhost_col2im(const neraf_window3d* wd, const void* dcol, int32_t dtype, int64_t ld_col, void* dx,
                             int64_t ld_dx, void*) {
  const Window w = window_of(wd);
  const long long n = in_voxels(w) * w.C;
  if (gather_can_vec8(w, dtype == NERAF_DT_BF16, ld_dx, 1, ld_col, dcol, dx)) {                               // as gridnet.cu
    for (long long i = 0; i < n / 8; ++i) col2im_vec8_element(w, (const bf16_t*)dcol, ld_col, (bf16_t*)dx, ld_dx, i);
    return 0;
  }
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32) col2im_element(w, (const float*)dcol, ld_col, (float*)dx, ld_dx, i);
    else col2im_element(w, (const bf16_t*)dcol, ld_col, (bf16_t*)dx, ld_dx, i);
  }
  return 0;
}

HOST_API int gridhost_pack_weight(const float* weight, int64_t c_out, int64_t c_in, int64_t k3, void* out, int32_t dtype,
                                  int64_t ld, void*) {
  const long long n = c_out * ld;
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32) pack_weight_element(weight, c_in, k3, (float*)out, ld, i);
    else pack_weight_element(weight, c_in, k3, (bf16_t*)out, ld, i);
  }
  return 0;
}

HOST_API int gridhost_unpack_wgrad(const float* dw_mat, int64_t ld, int64_t c_out, int64_t c_in, int64_t k3,
                                   int32_t n_partials, int64_t partial_stride, float* dw, void*) {
  const long long n = c_out * c_in * k3;
  for (long long i = 0; i < n; ++i) unpack_wgrad_element(dw_mat, ld, c_in, k3, n_partials, partial_stride, dw, i);
  return 0;
}

HOST_API int gridhost_bn_stats(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld, double* sums, void*) {
  memset(sums, 0, sizeof(double) * 2 * C);
  if (rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, 8, 8, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)) {      // as gridnet.cu
    for (long long c = 0; c < C; c += 8)
      for (long long lane = 0; lane < kLanes; ++lane) {
        double s[8], ss[8];
        column_sums_partial8((const bf16_t*)x, ld, c, lane, V, kLanes, s, ss);
        for (int k = 0; k < 8; ++k) { sums[c + k] += s[k]; sums[C + c + k] += ss[k]; }
      }
    return 0;
  }
  for (long long c = 0; c < C; ++c)
    for (long long lane = 0; lane < kLanes; ++lane) {
      double s = 0.0, ss = 0.0;
      if (dtype == NERAF_DT_F32) column_sums_partial((const float*)x, ld, c, lane, V, kLanes, &s, &ss);
      else column_sums_partial((const bf16_t*)x, ld, c, lane, V, kLanes, &s, &ss);
      sums[c] += s; sums[C + c] += ss;
    }
  return 0;
}

HOST_API int gridhost_bn_finalize(const double* sums, int64_t V, int64_t C, float eps, float momentum, int32_t training,
                                  float* running_mean, float* running_var, float* mean, float* invstd, void*) {
  for (long long c = 0; c < C; ++c)
    bn_finalize_channel(sums, V, C, eps, momentum, training, running_mean, running_var, mean, invstd, c);
  return 0;
}

HOST_API int gridhost_bn_apply(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld_x, const float* mean,
                               const float* invstd, const float* gamma, const float* beta, const void* residual,
                               int64_t ld_res, int32_t relu, void* y, int64_t ld_y, void*) {
  const long long n = V * C;
  if (rows_can_vec8(dtype == NERAF_DT_BF16, C, ld_x, residual ? ld_res : 0, ld_y, x, residual, y, mean, invstd, gamma, beta)) {           // as gridnet.cu
    for (long long i = 0; i < n / 8; ++i)
      bn_apply_vec8_element((const bf16_t*)x, ld_x, C, mean, invstd, gamma, beta, (const bf16_t*)residual, ld_res, relu,
                            (bf16_t*)y, ld_y, i);
    return 0;
  }
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32)
      bn_apply_element((const float*)x, ld_x, C, mean, invstd, gamma, beta, (const float*)residual, ld_res, relu, (float*)y,
                       ld_y, i);
    else
      bn_apply_element((const bf16_t*)x, ld_x, C, mean, invstd, gamma, beta, (const bf16_t*)residual, ld_res, relu,
                       (bf16_t*)y, ld_y, i);
  }
  return 0;
}

HOST_API int gridhost_bn_backward_reduce(const void* dy, const void* dy2, const void* y, const void* x, int32_t dtype,
                                         int64_t V, int64_t C, int64_t ld, const float* mean, const float* invstd,
                                         void* g_out, double* sums, void*) {
  memset(sums, 0, sizeof(double) * 2 * C);
  if (rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, 8, 8, dy, x, g_out, dy2, y, mean, invstd)) {                              // as gridnet.cu
    for (long long c = 0; c < C; c += 8)
      for (long long lane = 0; lane < kLanes; ++lane) {
        double s[8], ss[8];
        bn_backward_partial8((const bf16_t*)dy, (const bf16_t*)dy2, (const bf16_t*)y, (const bf16_t*)x, ld, mean, invstd,
                             (bf16_t*)g_out, c, lane, V, kLanes, s, ss);
        for (int k = 0; k < 8; ++k) { sums[c + k] += s[k]; sums[C + c + k] += ss[k]; }
      }
    return 0;
  }
  for (long long c = 0; c < C; ++c)
    for (long long lane = 0; lane < kLanes; ++lane) {
      double s = 0.0, ss = 0.0;
      if (dtype == NERAF_DT_F32)
        bn_backward_partial((const float*)dy, (const float*)dy2, (const float*)y, (const float*)x, ld, mean, invstd,
                            (float*)g_out, c, lane, V, kLanes, &s, &ss);
      else
        bn_backward_partial((const bf16_t*)dy, (const bf16_t*)dy2, (const bf16_t*)y, (const bf16_t*)x, ld, mean, invstd,
                            (bf16_t*)g_out, c, lane, V, kLanes, &s, &ss);
      sums[c] += s; sums[C + c] += ss;
    }
  return 0;
}

HOST_API int gridhost_bn_backward_apply(const void* g, const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld,
                                        const float* mean, const float* invstd, const float* gamma, const double* sums,
                                        int32_t training, void* dx, float* dgamma, float* dbeta, void*) {
  const long long n = V * C;
  const bool vec8 = rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, ld, ld, g, x, dx, mean, invstd, gamma, sums);                              // as gridnet.cu
  for (long long i = 0; vec8 && i < n / 8; ++i)
    bn_backward_vec8_element((const bf16_t*)g, (const bf16_t*)x, ld, C, mean, invstd, gamma, sums, V, training, (bf16_t*)dx, i);
  for (long long i = 0; !vec8 && i < n; ++i) {
    if (dtype == NERAF_DT_F32)
      bn_backward_element((const float*)g, (const float*)x, ld, C, mean, invstd, gamma, sums, V, training, (float*)dx, i);
    else
      bn_backward_element((const bf16_t*)g, (const bf16_t*)x, ld, C, mean, invstd, gamma, sums, V, training, (bf16_t*)dx, i);
  }
  for (long long c = 0; c < C; ++c) {
    if (dbeta) dbeta[c] = (float)sums[c];
    if (dgamma) dgamma[c] = (float)sums[C + c];
  }
  return 0;
}

HOST_API int gridhost_maxpool(const neraf_window3d* wd, const void* x, int32_t dtype, int64_t ld_x, void* y, int64_t ld_y,
                              int32_t* argmax, void*) {
  const Window w = window_of(wd);
  const long long n = out_voxels(w) * w.C;
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32) maxpool_element(w, (const float*)x, ld_x, (float*)y, ld_y, argmax, i);
    else maxpool_element(w, (const bf16_t*)x, ld_x, (bf16_t*)y, ld_y, argmax, i);
  }
  return 0;
}

HOST_API int gridhost_maxpool_backward(const neraf_window3d* wd, const void* dy, const void* dy2, int32_t dtype, int64_t ld_dy,
                                       const int32_t* argmax, void* dx, int64_t ld_dx, void*) {
  const Window w = window_of(wd);
  const long long n = in_voxels(w) * w.C;
  if (rows_can_vec8(dtype == NERAF_DT_BF16, w.C, ld_dy, ld_dx, 8, dy, dx, argmax, dy2, nullptr, nullptr, nullptr)) {   // as gridnet.cu
    for (long long i = 0; i < n / 8; ++i)
      maxpool_backward_vec8_element(w, (const bf16_t*)dy, (const bf16_t*)dy2, ld_dy, argmax, (bf16_t*)dx, ld_dx, i);
    return 0;
  }
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32) maxpool_backward_element(w, (const float*)dy, (const float*)dy2, ld_dy, argmax, (float*)dx, ld_dx, i);
    else maxpool_backward_element(w, (const bf16_t*)dy, (const bf16_t*)dy2, ld_dy, argmax, (bf16_t*)dx, ld_dx, i);
  }
  return 0;
}

HOST_API int gridhost_broadcast_rows(const float* v, float scale, int64_t V, int64_t C, void* out, int32_t dtype,
                                     int64_t ld, void*) {
  const long long n = V * C;
  for (long long i = 0; i < n; ++i) {
    if (dtype == NERAF_DT_F32) broadcast_rows_element(v, scale, C, (float*)out, ld, i);
    else broadcast_rows_element(v, scale, C, (bf16_t*)out, ld, i);
  }
  return 0;
}
