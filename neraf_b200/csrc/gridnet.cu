// Grid-feature producer: the non-GEMM operators of the reference's ResNet3D
// (NeRAF_resnet3d.py:116-201; one forward + backward per training step at NeRAF_model.py:554-556).  Every kernel here is
// a grid-stride loop (or a strip reduction) over the per-element functions of gridnet_core.h, which are also compiled
// for the host and checked there against torch (tests/test_gridnet.py).  All of them are HBM-bound gathers / streams
// over channels-last (V, C) matrices: consecutive threads walk consecutive channels of one voxel, so reads and writes
// are coalesced; grids are sized to a few waves of the 148 SMs.  The contractions themselves run on the tcgen05 job-list
// kernel (gemm_mega.cu) or the fp32 GEMM (gemm_simt.cu) -- see neraf_b200/gridnet.py for the assembly.
#include <cstdlib>

#include "common.cuh"
#include "gridnet_core.h"
#include "kernels.h"

namespace neraf {
namespace gridnet {

constexpr int kThreads = 256;

// NERAF_GRID_SCALAR=1 (environment, read once) keeps every kernel on its scalar form: for A/B timing of the 128-bit forms
// and for bisecting a parity failure.
static bool allow_vec8() {
  static const bool scalar_only = getenv("NERAF_GRID_SCALAR") != nullptr;
  return !scalar_only;
}

static unsigned grid_for(long long n) {
  const long long want = ceil_div(n, kThreads);
  const long long cap = (long long)sm_count() * 16;              // 16 blocks of 256 threads per SM = 2 full waves
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

#define GN_LOOP(n)                                                                               \
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (n);             \
       idx += (long long)gridDim.x * blockDim.x)

template <class Tin, class Tout>
__global__ void __launch_bounds__(kThreads) im2col_kernel(Window w, const Tin* in, long long vs, long long cs, Tout* col,
                                                          long long ld, long long n) {
  GN_LOOP(n) im2col_element(w, in, vs, cs, col, ld, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) col2im_kernel(Window w, const T* dcol, long long ld_col, T* dx, long long ld_dx,
                                                          long long n) {
  GN_LOOP(n) col2im_element(w, dcol, ld_col, dx, ld_dx, idx);
}
__global__ void __launch_bounds__(kThreads) im2col_vec8_kernel(Window w, const bf16_t* in, long long vs, bf16_t* col,
                                                               long long ld, long long n) {
  GN_LOOP(n) im2col_vec8_element(w, in, vs, col, ld, idx);
}
__global__ void __launch_bounds__(kThreads) col2im_vec8_kernel(Window w, const bf16_t* dcol, long long ld_col, bf16_t* dx,
                                                               long long ld_dx, long long n) {
  GN_LOOP(n) col2im_vec8_element(w, dcol, ld_col, dx, ld_dx, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) maxpool_kernel(Window w, const T* x, long long ld_x, T* y, long long ld_y,
                                                           int32_t* argmax, long long n) {
  GN_LOOP(n) maxpool_element(w, x, ld_x, y, ld_y, argmax, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) maxpool_backward_kernel(Window w, const T* dy, const T* dy2,
                                                                    long long ld_dy, const int32_t* argmax, T* dx,
                                                                    long long ld_dx, long long n) {
  GN_LOOP(n) maxpool_backward_element(w, dy, dy2, ld_dy, argmax, dx, ld_dx, idx);
}
__global__ void __launch_bounds__(kThreads) maxpool_backward_vec8_kernel(Window w, const bf16_t* dy, const bf16_t* dy2,
                                                                         long long ld_dy, const int32_t* argmax,
                                                                         bf16_t* dx, long long ld_dx, long long n) {
  GN_LOOP(n) maxpool_backward_vec8_element(w, dy, dy2, ld_dy, argmax, dx, ld_dx, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) pack_weight_kernel(const float* wt, long long c_in, long long k3, T* out,
                                                               long long ld, long long n) {
  GN_LOOP(n) pack_weight_element(wt, c_in, k3, out, ld, idx);
}
__global__ void __launch_bounds__(kThreads) unpack_wgrad_kernel(const float* dw_mat, long long ld, long long c_in,
                                                                long long k3, int n_partials, long long partial_stride,
                                                                float* dw, long long n) {
  GN_LOOP(n) unpack_wgrad_element(dw_mat, ld, c_in, k3, n_partials, partial_stride, dw, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const T* x, long long ld_x, long long C, const float* mean,
                                                            const float* invstd, const float* gamma, const float* beta,
                                                            const T* res, long long ld_res, int relu, T* y, long long ld_y,
                                                            long long n) {
  GN_LOOP(n) bn_apply_element(x, ld_x, C, mean, invstd, gamma, beta, res, ld_res, relu, y, ld_y, idx);
}
template <class T>
__global__ void __launch_bounds__(kThreads) bn_backward_kernel(const T* g, const T* x, long long ld, long long C,
                                                               const float* mean, const float* invstd, const float* gamma,
                                                               const double* sums, long long V, int training, T* dx,
                                                               float* dgamma, float* dbeta, long long n) {
  GN_LOOP(n) bn_backward_element(g, x, ld, C, mean, invstd, gamma, sums, V, training, dx, idx);
  if (blockIdx.x == 0)
    for (long long c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dgamma) dgamma[c] = (float)sums[C + c];
    }
}
__global__ void __launch_bounds__(kThreads) bn_apply_vec8_kernel(const bf16_t* x, long long ld_x, long long C,
                                                                 const float* mean, const float* invstd,
                                                                 const float* gamma, const float* beta, const bf16_t* res,
                                                                 long long ld_res, int relu, bf16_t* y, long long ld_y,
                                                                 long long n) {
  GN_LOOP(n) bn_apply_vec8_element(x, ld_x, C, mean, invstd, gamma, beta, res, ld_res, relu, y, ld_y, idx);
}
__global__ void __launch_bounds__(kThreads) bn_backward_vec8_kernel(const bf16_t* g, const bf16_t* x, long long ld,
                                                                    long long C, const float* mean, const float* invstd,
                                                                    const float* gamma, const double* sums, long long V,
                                                                    int training, bf16_t* dx, float* dgamma, float* dbeta,
                                                                    long long n) {
  GN_LOOP(n) bn_backward_vec8_element(g, x, ld, C, mean, invstd, gamma, sums, V, training, dx, idx);
  if (blockIdx.x == 0)
    for (long long c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)sums[c];
      if (dgamma) dgamma[c] = (float)sums[C + c];
    }
}
template <class T>
__global__ void __launch_bounds__(kThreads) broadcast_rows_kernel(const float* v, float scale, long long C, T* out,
                                                                  long long ld, long long n) {
  GN_LOOP(n) broadcast_rows_element(v, scale, C, out, ld, idx);
}
__global__ void bn_finalize_kernel(const double* sums, long long V, long long C, float eps, float momentum, int training,
                                   float* running_mean, float* running_var, float* mean, float* invstd) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) bn_finalize_channel(sums, V, C, eps, momentum, training, running_mean, running_var, mean, invstd, c);
}

// Strip reductions over the rows of a (V, C) matrix.  Block = 32 columns x 8 row lanes; block (bx, by) owns columns
// [32 bx, 32 bx + 32) and rows [by * rows_per_block, ...): lane ty visits rows r0 + ty, r0 + ty + 8, ...  A warp reads
// 32 consecutive channels of one row.  The 8 lanes are combined in shared memory, one fp64 atomic per column and sum.
constexpr int kRedCols = 32, kRedLanes = 8;

template <class T>
__global__ void __launch_bounds__(kRedCols * kRedLanes) column_sums_kernel(const T* x, long long ld, long long V,
                                                                           long long C, long long rows_per_block,
                                                                           double* sums) {
  __shared__ double sh[2][kRedLanes][kRedCols];
  const int tx = threadIdx.x % kRedCols, ty = threadIdx.x / kRedCols;
  const long long c = (long long)blockIdx.x * kRedCols + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < V ? r0 + rows_per_block : V;
  double s = 0.0, ss = 0.0;
  if (c < C) column_sums_partial(x, ld, c, r0 + ty, r1, (long long)kRedLanes, &s, &ss);
  sh[0][ty][tx] = s; sh[1][ty][tx] = ss;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 1; i < kRedLanes; ++i) { s += sh[0][i][tx]; ss += sh[1][i][tx]; }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, ss);
  }
}

template <class T>
__global__ void __launch_bounds__(kRedCols * kRedLanes) bn_backward_reduce_kernel(const T* dy, const T* dy2, const T* y,
                                                                                  const T* x, long long ld, long long V,
                                                                                  long long C, const float* mean,
                                                                                  const float* invstd, T* g_out,
                                                                                  long long rows_per_block, double* sums) {
  __shared__ double sh[2][kRedLanes][kRedCols];
  const int tx = threadIdx.x % kRedCols, ty = threadIdx.x / kRedCols;
  const long long c = (long long)blockIdx.x * kRedCols + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < V ? r0 + rows_per_block : V;
  double s = 0.0, ss = 0.0;
  if (c < C) bn_backward_partial(dy, dy2, y, x, ld, mean, invstd, g_out, c, r0 + ty, r1, (long long)kRedLanes, &s, &ss);
  sh[0][ty][tx] = s; sh[1][ty][tx] = ss;
  __syncthreads();
  if (ty == 0 && c < C) {
    for (int i = 1; i < kRedLanes; ++i) { s += sh[0][i][tx]; ss += sh[1][i][tx]; }
    atomicAdd(sums + c, s);
    atomicAdd(sums + C + c, ss);
  }
}

// 128-bit forms: a thread owns 8 consecutive channels (one 16-byte load per row and operand).  Thread t of the 256:
// column group t % groups, row lane t / groups, where groups = min(C / 8, 32) -- for narrow matrices (C = 64: 8 groups) a
// warp covers four whole rows per load instead of a quarter of its lanes one row; blockIdx.x walks blocks of 256 channels.
// (The scalar forms read 2 bytes per thread and 64-byte row segments per warp: 19 % of a producer step.)
constexpr int kRed8Threads = 256;
struct Red8Shape { int groups, lanes; };
__device__ __forceinline__ void red8_finish(double (&s)[8], double (&ss)[8], int group, int lane, int lanes, int groups,
                                            long long c, long long C, double* sums) {
  __shared__ double sh[16][kRed8Threads];
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 8; ++k) { sh[k][t] = s[k]; sh[8 + k][t] = ss[k]; }
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int i = 1; i < lanes; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += sh[k][i * groups + group]; ss[k] += sh[8 + k][i * groups + group]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(sums + c + k, s[k]);
      atomicAdd(sums + C + c + k, ss[k]);
    }
  }
}
__global__ void __launch_bounds__(kRed8Threads) column_sums_vec8_kernel(const bf16_t* x, long long ld, long long V, long long C,
                                                                        long long rows_per_block, int groups, double* sums) {
  const int lanes = kRed8Threads / groups;
  const int group = threadIdx.x % groups, lane = threadIdx.x / groups;
  const long long c = ((long long)blockIdx.x * groups + group) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < V ? r0 + rows_per_block : V;
  double s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.0; ss[k] = 0.0; }
  if (c < C) column_sums_partial8(x, ld, c, r0 + lane, r1, (long long)lanes, s, ss);
  red8_finish(s, ss, group, lane, lanes, groups, c, C, sums);
}
__global__ void __launch_bounds__(kRed8Threads) bn_backward_reduce_vec8_kernel(const bf16_t* dy, const bf16_t* dy2, const bf16_t* y,
                                                                               const bf16_t* x, long long ld, long long V,
                                                                               long long C, const float* mean,
                                                                               const float* invstd, bf16_t* g_out,
                                                                               long long rows_per_block, int groups,
                                                                               double* sums) {
  const int lanes = kRed8Threads / groups;
  const int group = threadIdx.x % groups, lane = threadIdx.x / groups;
  const long long c = ((long long)blockIdx.x * groups + group) * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = r0 + rows_per_block < V ? r0 + rows_per_block : V;
  double s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.0; ss[k] = 0.0; }
  if (c < C) bn_backward_partial8(dy, dy2, y, x, ld, mean, invstd, g_out, c, r0 + lane, r1, (long long)lanes, s, ss);
  red8_finish(s, ss, group, lane, lanes, groups, c, C, sums);
}
// column groups per block (a divisor of 256, or 0: use the scalar form) and the grid: about four blocks per SM
static int red8_shape(long long V, long long C, dim3* grid, long long* rows_per_block) {
  const long long cg = C / 8;
  int groups = 32;
  if (cg < 32) {
    if (cg <= 0 || (cg & (cg - 1)) != 0) return 0;
    groups = (int)cg;
  }
  const int lanes = kRed8Threads / groups;
  const long long col_blocks = ceil_div(cg, groups);
  long long row_blocks = ceil_div((long long)sm_count() * 4, col_blocks);
  if (row_blocks < 1) row_blocks = 1;
  long long rows = round_up(ceil_div(V, row_blocks), lanes);
  if (rows < 4 * lanes) rows = 4 * lanes;
  row_blocks = ceil_div(V, rows);
  *grid = dim3((unsigned)col_blocks, (unsigned)row_blocks);
  *rows_per_block = rows;
  return groups;
}

// rows per block so that the grid is about four blocks per SM, in multiples of the 8 row lanes
static long long strip_rows(long long V, long long C, dim3* grid) {
  const long long col_blocks = ceil_div(C, kRedCols);
  long long row_blocks = ceil_div((long long)sm_count() * 4, col_blocks);
  if (row_blocks < 1) row_blocks = 1;
  long long rows = round_up(ceil_div(V, row_blocks), kRedLanes);
  if (rows < 8 * kRedLanes) rows = 8 * kRedLanes;
  row_blocks = ceil_div(V, rows);
  *grid = dim3((unsigned)col_blocks, (unsigned)row_blocks);
  return rows;
}

static int check_window(const neraf_window3d* w, Window* out) {
  NERAF_REQUIRE(w, "grid op: null window");
  NERAF_REQUIRE(w->in_d > 0 && w->in_h > 0 && w->in_w > 0 && w->channels > 0, "grid op: empty input extent");
  NERAF_REQUIRE(w->k >= 1 && w->k <= 7 && w->stride >= 1 && w->stride <= 4 && w->pad >= 0 && w->pad < w->k,
                "grid op: unsupported window (k %d, stride %d, pad %d)", w->k, w->stride, w->pad);
  NERAF_REQUIRE(w->in_d + 2 * w->pad >= w->k && w->in_h + 2 * w->pad >= w->k && w->in_w + 2 * w->pad >= w->k,
                "grid op: window larger than the padded input");
  *out = make_window(w->in_d, w->in_h, w->in_w, w->channels, w->k, w->stride, w->pad);
  NERAF_REQUIRE(in_voxels(*out) < 0x7fffffffLL && out_voxels(*out) < 0x7fffffffLL, "grid op: more than 2^31 voxels");
  return NERAF_OK;
}

static int check_dtype(int32_t dt) {
  NERAF_REQUIRE(dt == NERAF_DT_F32 || dt == NERAF_DT_BF16, "grid op: dtype must be NERAF_DT_F32 or NERAF_DT_BF16");
  return NERAF_OK;
}

}  // namespace gridnet
}  // namespace neraf

using namespace neraf;
using namespace neraf::gridnet;

extern "C" int neraf_grid_im2col(const neraf_window3d* wd, const void* in, int32_t in_dtype, int64_t voxel_stride,
                                 int64_t channel_stride, void* col, int32_t col_dtype, int64_t ld_col,
                                 neraf_stream_t stream) {
  Window w;
  NERAF_TRY(check_window(wd, &w));
  NERAF_TRY(check_dtype(in_dtype));
  NERAF_TRY(check_dtype(col_dtype));
  NERAF_REQUIRE(in && col, "im2col: null pointer");
  NERAF_REQUIRE(ld_col >= (int64_t)w.k * w.k * w.k * w.C, "im2col: ld_col < k^3 C");
  NERAF_REQUIRE(voxel_stride >= 1 && channel_stride >= 1, "im2col: bad input strides");
  const long long n = out_voxels(w) * ld_col;
  cudaStream_t s = (cudaStream_t)stream;
  if (allow_vec8() && gather_can_vec8(w, in_dtype == NERAF_DT_BF16 && col_dtype == NERAF_DT_BF16, voxel_stride,
                                      channel_stride, ld_col, in, col)) {
    im2col_vec8_kernel<<<grid_for(n / 8), kThreads, 0, s>>>(w, (const bf16_t*)in, voxel_stride, (bf16_t*)col, ld_col, n / 8);
    NERAF_CHECK_LAUNCH("im2col_vec8_kernel");
    return NERAF_OK;
  }
  const unsigned g = grid_for(n);
  if (in_dtype == NERAF_DT_F32 && col_dtype == NERAF_DT_F32)
    im2col_kernel<<<g, kThreads, 0, s>>>(w, (const float*)in, voxel_stride, channel_stride, (float*)col, ld_col, n);
  else if (in_dtype == NERAF_DT_F32)
    im2col_kernel<<<g, kThreads, 0, s>>>(w, (const float*)in, voxel_stride, channel_stride, (bf16_t*)col, ld_col, n);
  else if (col_dtype == NERAF_DT_BF16)
    im2col_kernel<<<g, kThreads, 0, s>>>(w, (const bf16_t*)in, voxel_stride, channel_stride, (bf16_t*)col, ld_col, n);
  else
    im2col_kernel<<<g, kThreads, 0, s>>>(w, (const bf16_t*)in, voxel_stride, channel_stride, (float*)col, ld_col, n);
  NERAF_CHECK_LAUNCH("im2col_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_col2im(const neraf_window3d* wd, const void* dcol, int32_t dtype, int64_t ld_col, void* dx,
                                 int64_t ld_dx, neraf_stream_t stream) {
  Window w;
  NERAF_TRY(check_window(wd, &w));
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(dcol && dx, "col2im: null pointer");
  NERAF_REQUIRE(ld_col >= (int64_t)w.k * w.k * w.k * w.C && ld_dx >= w.C, "col2im: row stride too small");
  const long long n = in_voxels(w) * w.C;
  cudaStream_t s = (cudaStream_t)stream;
  if (allow_vec8() && gather_can_vec8(w, dtype == NERAF_DT_BF16, ld_dx, 1, ld_col, dcol, dx)) {
    col2im_vec8_kernel<<<grid_for(n / 8), kThreads, 0, s>>>(w, (const bf16_t*)dcol, ld_col, (bf16_t*)dx, ld_dx, n / 8);
    NERAF_CHECK_LAUNCH("col2im_vec8_kernel");
    return NERAF_OK;
  }
  if (dtype == NERAF_DT_F32)
    col2im_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const float*)dcol, ld_col, (float*)dx, ld_dx, n);
  else
    col2im_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const bf16_t*)dcol, ld_col, (bf16_t*)dx, ld_dx, n);
  NERAF_CHECK_LAUNCH("col2im_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_pack_weight(const float* weight, int64_t c_out, int64_t c_in, int64_t k3, void* out,
                                      int32_t out_dtype, int64_t ld_out, neraf_stream_t stream) {
  NERAF_TRY(check_dtype(out_dtype));
  NERAF_REQUIRE(weight && out && c_out > 0 && c_in > 0 && k3 > 0 && ld_out >= c_in * k3, "pack_weight: bad arguments");
  const long long n = c_out * ld_out;
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == NERAF_DT_F32)
    pack_weight_kernel<<<grid_for(n), kThreads, 0, s>>>(weight, c_in, k3, (float*)out, ld_out, n);
  else
    pack_weight_kernel<<<grid_for(n), kThreads, 0, s>>>(weight, c_in, k3, (bf16_t*)out, ld_out, n);
  NERAF_CHECK_LAUNCH("pack_weight_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_unpack_wgrad(const float* dw_mat, int64_t ld, int64_t c_out, int64_t c_in, int64_t k3,
                                       int32_t n_partials, int64_t partial_stride, float* dweight,
                                       neraf_stream_t stream) {
  NERAF_REQUIRE(dw_mat && dweight && c_out > 0 && c_in > 0 && k3 > 0 && ld >= c_in * k3, "unpack_wgrad: bad arguments");
  NERAF_REQUIRE(n_partials >= 1 && n_partials <= 64 && (n_partials == 1 || partial_stride >= c_out * ld),
                "unpack_wgrad: 1..64 partial matrices, at least c_out * ld floats apart");
  const long long n = c_out * c_in * k3;
  unpack_wgrad_kernel<<<grid_for(n), kThreads, 0, (cudaStream_t)stream>>>(dw_mat, ld, c_in, k3, n_partials, partial_stride,
                                                                          dweight, n);
  NERAF_CHECK_LAUNCH("unpack_wgrad_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_bn_stats(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld, double* sums,
                                   neraf_stream_t stream) {
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(x && sums && V > 0 && C > 0 && ld >= C, "bn_stats: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  NERAF_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)(2 * C) * sizeof(double), s));
  dim3 grid;
  if (allow_vec8() && rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, 8, 8, x, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr)) {
    long long rows8 = 0;
    const int groups = red8_shape(V, C, &grid, &rows8);
    if (groups > 0) {
      column_sums_vec8_kernel<<<grid, kRed8Threads, 0, s>>>((const bf16_t*)x, ld, V, C, rows8, groups, sums);
      NERAF_CHECK_LAUNCH("column_sums_vec8_kernel");
      return NERAF_OK;
    }
  }
  const long long rows = strip_rows(V, C, &grid);
  if (dtype == NERAF_DT_F32)
    column_sums_kernel<<<grid, kRedCols * kRedLanes, 0, s>>>((const float*)x, ld, V, C, rows, sums);
  else
    column_sums_kernel<<<grid, kRedCols * kRedLanes, 0, s>>>((const bf16_t*)x, ld, V, C, rows, sums);
  NERAF_CHECK_LAUNCH("column_sums_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_bn_finalize(const double* sums, int64_t V, int64_t C, float eps, float momentum,
                                      int32_t training, float* running_mean, float* running_var, float* mean,
                                      float* invstd, neraf_stream_t stream) {
  NERAF_REQUIRE(mean && V > 0 && C > 0, "bn_finalize: bad arguments");
  NERAF_REQUIRE(training ? sums != nullptr : (running_mean && running_var), "bn_finalize: missing statistics");
  NERAF_REQUIRE(!(training && momentum > 0.f) || (running_mean && running_var), "bn_finalize: momentum without running buffers");
  bn_finalize_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, V, C, eps, momentum, training,
                                                                                  running_mean, running_var, mean, invstd);
  NERAF_CHECK_LAUNCH("bn_finalize_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_bn_apply(const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld_x, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta, const void* residual,
                                   int64_t ld_res, int32_t relu, void* y, int64_t ld_y, neraf_stream_t stream) {
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(x && y && mean && invstd && gamma && beta && V > 0 && C > 0 && ld_x >= C && ld_y >= C &&
                (!residual || ld_res >= C), "bn_apply: bad arguments");
  const long long n = V * C;
  cudaStream_t s = (cudaStream_t)stream;
  if (allow_vec8() &&
      rows_can_vec8(dtype == NERAF_DT_BF16, C, ld_x, residual ? ld_res : 0, ld_y, x, residual, y, mean, invstd, gamma, beta)) {
    bn_apply_vec8_kernel<<<grid_for(n / 8), kThreads, 0, s>>>((const bf16_t*)x, ld_x, C, mean, invstd, gamma, beta,
                                                              (const bf16_t*)residual, ld_res, relu, (bf16_t*)y, ld_y, n / 8);
    NERAF_CHECK_LAUNCH("bn_apply_vec8_kernel");
    return NERAF_OK;
  }
  if (dtype == NERAF_DT_F32)
    bn_apply_kernel<<<grid_for(n), kThreads, 0, s>>>((const float*)x, ld_x, C, mean, invstd, gamma, beta,
                                                     (const float*)residual, ld_res, relu, (float*)y, ld_y, n);
  else
    bn_apply_kernel<<<grid_for(n), kThreads, 0, s>>>((const bf16_t*)x, ld_x, C, mean, invstd, gamma, beta,
                                                     (const bf16_t*)residual, ld_res, relu, (bf16_t*)y, ld_y, n);
  NERAF_CHECK_LAUNCH("bn_apply_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_bn_backward_reduce(const void* dy, const void* dy2, const void* y, const void* x, int32_t dtype,
                                             int64_t V, int64_t C, int64_t ld, const float* mean, const float* invstd,
                                             void* g_out, double* sums, neraf_stream_t stream) {
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(dy && x && mean && invstd && g_out && sums && V > 0 && C > 0 && ld >= C, "bn_backward_reduce: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  NERAF_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)(2 * C) * sizeof(double), s));
  dim3 grid;
  if (allow_vec8() && rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, 8, 8, dy, x, g_out, dy2, y, mean, invstd)) {
    long long rows8 = 0;
    const int groups = red8_shape(V, C, &grid, &rows8);
    if (groups > 0) {
      bn_backward_reduce_vec8_kernel<<<grid, kRed8Threads, 0, s>>>((const bf16_t*)dy, (const bf16_t*)dy2, (const bf16_t*)y,
                                                                   (const bf16_t*)x, ld, V, C, mean, invstd, (bf16_t*)g_out,
                                                                   rows8, groups, sums);
      NERAF_CHECK_LAUNCH("bn_backward_reduce_vec8_kernel");
      return NERAF_OK;
    }
  }
  const long long rows = strip_rows(V, C, &grid);
  if (dtype == NERAF_DT_F32)
    bn_backward_reduce_kernel<<<grid, kRedCols * kRedLanes, 0, s>>>((const float*)dy, (const float*)dy2, (const float*)y,
                                                                    (const float*)x, ld, V, C, mean, invstd,
                                                                    (float*)g_out, rows, sums);
  else
    bn_backward_reduce_kernel<<<grid, kRedCols * kRedLanes, 0, s>>>((const bf16_t*)dy, (const bf16_t*)dy2,
                                                                    (const bf16_t*)y, (const bf16_t*)x, ld, V, C, mean,
                                                                    invstd, (bf16_t*)g_out, rows, sums);
  NERAF_CHECK_LAUNCH("bn_backward_reduce_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_bn_backward_apply(const void* g, const void* x, int32_t dtype, int64_t V, int64_t C, int64_t ld,
                                            const float* mean, const float* invstd, const float* gamma,
                                            const double* sums, int32_t training, void* dx, float* dgamma, float* dbeta,
                                            neraf_stream_t stream) {
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(g && x && mean && invstd && gamma && sums && dx && V > 0 && C > 0 && ld >= C,
                "bn_backward_apply: bad arguments");
  const long long n = V * C;
  cudaStream_t s = (cudaStream_t)stream;
  if (allow_vec8() && rows_can_vec8(dtype == NERAF_DT_BF16, C, ld, ld, ld, g, x, dx, mean, invstd, gamma, sums)) {
    bn_backward_vec8_kernel<<<grid_for(n / 8), kThreads, 0, s>>>((const bf16_t*)g, (const bf16_t*)x, ld, C, mean, invstd,
                                                                 gamma, sums, V, training, (bf16_t*)dx, dgamma, dbeta, n / 8);
    NERAF_CHECK_LAUNCH("bn_backward_vec8_kernel");
    return NERAF_OK;
  }
  if (dtype == NERAF_DT_F32)
    bn_backward_kernel<<<grid_for(n), kThreads, 0, s>>>((const float*)g, (const float*)x, ld, C, mean, invstd, gamma, sums,
                                                        V, training, (float*)dx, dgamma, dbeta, n);
  else
    bn_backward_kernel<<<grid_for(n), kThreads, 0, s>>>((const bf16_t*)g, (const bf16_t*)x, ld, C, mean, invstd, gamma,
                                                        sums, V, training, (bf16_t*)dx, dgamma, dbeta, n);
  NERAF_CHECK_LAUNCH("bn_backward_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_maxpool(const neraf_window3d* wd, const void* x, int32_t dtype, int64_t ld_x, void* y,
                                  int64_t ld_y, int32_t* argmax, neraf_stream_t stream) {
  Window w;
  NERAF_TRY(check_window(wd, &w));
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(x && y && argmax && ld_x >= w.C && ld_y >= w.C, "maxpool: bad arguments");
  const long long n = out_voxels(w) * w.C;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == NERAF_DT_F32)
    maxpool_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const float*)x, ld_x, (float*)y, ld_y, argmax, n);
  else
    maxpool_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const bf16_t*)x, ld_x, (bf16_t*)y, ld_y, argmax, n);
  NERAF_CHECK_LAUNCH("maxpool_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_maxpool_backward(const neraf_window3d* wd, const void* dy, const void* dy2, int32_t dtype,
                                           int64_t ld_dy,
                                           const int32_t* argmax, void* dx, int64_t ld_dx, neraf_stream_t stream) {
  Window w;
  NERAF_TRY(check_window(wd, &w));
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(dy && dx && argmax && ld_dy >= w.C && ld_dx >= w.C, "maxpool_backward: bad arguments");
  const long long n = in_voxels(w) * w.C;
  cudaStream_t s = (cudaStream_t)stream;
  if (allow_vec8() && rows_can_vec8(dtype == NERAF_DT_BF16, w.C, ld_dy, ld_dx, 8, dy, dx, argmax, dy2, nullptr, nullptr, nullptr)) {
    maxpool_backward_vec8_kernel<<<grid_for(n / 8), kThreads, 0, s>>>(w, (const bf16_t*)dy, (const bf16_t*)dy2, ld_dy, argmax,
                                                                      (bf16_t*)dx, ld_dx, n / 8);
    NERAF_CHECK_LAUNCH("maxpool_backward_vec8_kernel");
    return NERAF_OK;
  }
  if (dtype == NERAF_DT_F32)
    maxpool_backward_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const float*)dy, (const float*)dy2, ld_dy, argmax,
                                                             (float*)dx, ld_dx, n);
  else
    maxpool_backward_kernel<<<grid_for(n), kThreads, 0, s>>>(w, (const bf16_t*)dy, (const bf16_t*)dy2, ld_dy, argmax,
                                                             (bf16_t*)dx, ld_dx, n);
  NERAF_CHECK_LAUNCH("maxpool_backward_kernel");
  return NERAF_OK;
}

extern "C" int neraf_grid_broadcast_rows(const float* v, float scale, int64_t V, int64_t C, void* out, int32_t dtype,
                                         int64_t ld, neraf_stream_t stream) {
  NERAF_TRY(check_dtype(dtype));
  NERAF_REQUIRE(v && out && V > 0 && C > 0 && ld >= C, "broadcast_rows: bad arguments");
  const long long n = V * C;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == NERAF_DT_F32)
    broadcast_rows_kernel<<<grid_for(n), kThreads, 0, s>>>(v, scale, C, (float*)out, ld, n);
  else
    broadcast_rows_kernel<<<grid_for(n), kThreads, 0, s>>>(v, scale, C, (bf16_t*)out, ld, n);
  NERAF_CHECK_LAUNCH("broadcast_rows_kernel");
  return NERAF_OK;
}
