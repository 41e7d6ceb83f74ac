// Fused spectral loss of the acoustic field (K2).
//
// Replaces STFTLoss.forward (/root/reference/NeRAF/NeRAF_evaluator.py:88-108: exp x2, sub x3,
// Frobenius norm x2, div, mse/l1 -- ~10 kernels each re-reading the (B,C,F) tensors) and its autograd
// backward by
//   1. one reduction kernel producing the four global sums
//        S_num = sum (e^y - e^x)^2, S_den = sum (e^y - 1e-3)^2, S_sq = sum (y-x)^2, S_abs = sum |y-x|
//      (per-element math in fp32, accumulation in fp64, warp-shuffle -> shared -> one atomic per block),
//   2. a one-thread finalize -- a separate launch only under data parallelism, where the ranks all-reduce the sums
//      first (the Frobenius ratio is a global quantity); on one GPU the last block of the
//      reduction forms the losses itself (neraf_spectral_loss_forward),
//   3. one elementwise kernel for d loss / d pred.
// HBM-bound: 8 B/element forward, 12 B/element backward.
#include "common.cuh"
#include "loss_math.cuh"

namespace neraf {

constexpr float kEpsMag = 1e-3f;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void accumulate(float x, float y, double& s_num, double& s_den, double& s_sq, double& s_abs) {
  const float ex = expf(x), ey = expf(y);
  const float dm = ey - ex;                 // (e^y - 1e-3) - (e^x - 1e-3)
  const float ym = ey - kEpsMag;
  const float d = y - x;
  s_num += (double)(dm * dm);
  s_den += (double)(ym * ym);
  s_sq += (double)(d * d);
  s_abs += (double)fabsf(d);
}

// FUSED != 0: the last block to finish (ticket in sums[4], reinterpreted as an integer) also forms the two losses,
// saving the separate finalize launch of the single-GPU path.
struct FinalizeArgs { int fused; int64_t n_total; int criterion; float w_sc, w_mag; float* losses; };

__global__ void __launch_bounds__(256) loss_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                        int64_t n, double* __restrict__ sums, const FinalizeArgs fin) {
  double s_num = 0, s_den = 0, s_sq = 0, s_abs = 0;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* g4 = reinterpret_cast<const float4*>(gt);
  for (int64_t i = tid; i < n4; i += nthreads) {
    const float4 x = __ldg(p4 + i), y = __ldg(g4 + i);
    accumulate(x.x, y.x, s_num, s_den, s_sq, s_abs);
    accumulate(x.y, y.y, s_num, s_den, s_sq, s_abs);
    accumulate(x.z, y.z, s_num, s_den, s_sq, s_abs);
    accumulate(x.w, y.w, s_num, s_den, s_sq, s_abs);
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nthreads) accumulate(__ldg(pred + i), __ldg(gt + i), s_num, s_den, s_sq, s_abs);

  __shared__ double red[4][8];
  s_num = warp_sum(s_num); s_den = warp_sum(s_den); s_sq = warp_sum(s_sq); s_abs = warp_sum(s_abs);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (lane == 0) { red[0][warp] = s_num; red[1][warp] = s_den; red[2][warp] = s_sq; red[3][warp] = s_abs; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(sums + threadIdx.x, t);
  }
  if (!fin.fused) return;
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long ticket = atomicAdd(reinterpret_cast<unsigned long long*>(sums + 4), 1ull);
    is_last = ticket == (unsigned long long)gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    double v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldcg(sums + i);
    finalize(v, fin.n_total, fin.criterion, fin.w_sc, fin.w_mag, fin.losses);
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ sums, int64_t n_total, int criterion, float w_sc,
                                     float w_mag, float* __restrict__ losses) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  finalize(sums, n_total, criterion, w_sc, w_mag, losses);
}

__global__ void __launch_bounds__(256) loss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                            int64_t n, int64_t n_total, int criterion,
                                                            const double* __restrict__ sums,
                                                            const float* __restrict__ up_sc,
                                                            const float* __restrict__ up_mag, float w_sc, float w_mag,
                                                            float* __restrict__ dpred) {
  // d sc/dx = (e^x - e^y) e^x / (sqrt(S_num) sqrt(S_den)); d mse/dx = 2 (x-y)/N; d l1/dx = sign(x-y)/N
  const float g_sc = up_sc ? up_sc[0] : 1.f, g_mag = up_mag ? up_mag[0] : 1.f;
  float a = 0.f;
  if (criterion != NERAF_CRIT_MSE) a = (float)((double)(g_sc * w_sc) / (sqrt(sums[0]) * sqrt(sums[1])));
  const float b = (float)((double)(g_mag * w_mag) / (double)n_total);
  const bool l1 = criterion == NERAF_CRIT_SC_SLL1;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += nthreads) {
    const float x = __ldg(pred + i), y = __ldg(gt + i);
    const float d = x - y;
    float g = l1 ? b * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * b * d;
    if (criterion != NERAF_CRIT_MSE) {
      const float ex = expf(x), ey = expf(y);
      g += a * (ex - ey) * ex;
    }
    dpred[i] = g;
  }
}

static unsigned reduce_grid(int64_t n) {
  const int64_t want = ceil_div(ceil_div(n, 4), 256);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace neraf

using namespace neraf;

extern "C" int neraf_spectral_loss_sums(const float* pred, const float* gt, int64_t n, double* sums, int accumulate,
                                        neraf_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NERAF_REQUIRE(sums && n >= 0 && (n == 0 || (pred && gt)), "spectral_loss_sums: null pointer");
  if (!accumulate) NERAF_CHECK_CUDA(cudaMemsetAsync(sums, 0, 4 * sizeof(double), stream));
  if (n == 0) return NERAF_OK;
  loss_sums_kernel<<<reduce_grid(n), 256, 0, stream>>>(pred, gt, n, sums, FinalizeArgs{});
  NERAF_CHECK_LAUNCH("loss_sums_kernel");
  return NERAF_OK;
}

extern "C" int neraf_spectral_loss_forward(const float* pred, const float* gt, int64_t n, int criterion, float w_sc,
                                           float w_mag, double* scratch, float* losses, neraf_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  NERAF_REQUIRE(scratch && losses && n > 0 && pred && gt, "spectral_loss_forward: bad arguments");
  NERAF_REQUIRE(criterion >= 0 && criterion <= 2, "spectral_loss_forward: unknown criterion %d", criterion);
  NERAF_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 5 * sizeof(double), stream));
  const FinalizeArgs fin{1, n, criterion, w_sc, w_mag, losses};
  loss_sums_kernel<<<reduce_grid(n), 256, 0, stream>>>(pred, gt, n, scratch, fin);
  NERAF_CHECK_LAUNCH("loss_sums_kernel");
  return NERAF_OK;
}

extern "C" int neraf_spectral_loss_finalize(const double* sums, int64_t n_total, int criterion, float w_sc, float w_mag,
                                            float* losses, neraf_stream_t stream) {
  NERAF_REQUIRE(sums && losses && n_total > 0, "spectral_loss_finalize: bad arguments");
  NERAF_REQUIRE(criterion >= 0 && criterion <= 2, "spectral_loss_finalize: unknown criterion %d", criterion);
  loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, n_total, criterion, w_sc, w_mag, losses);
  NERAF_CHECK_LAUNCH("loss_finalize_kernel");
  return NERAF_OK;
}

extern "C" int neraf_spectral_loss_backward(const float* pred, const float* gt, int64_t n, int64_t n_total,
                                            int criterion, const double* sums, const float* upstream_sc,
                                            const float* upstream_mag, float w_sc, float w_mag, float* dpred,
                                            neraf_stream_t stream) {
  NERAF_REQUIRE(sums && dpred && n >= 0 && n_total > 0 && (n == 0 || (pred && gt)), "spectral_loss_backward: bad arguments");
  NERAF_REQUIRE(criterion >= 0 && criterion <= 2, "spectral_loss_backward: unknown criterion %d", criterion);
  if (n == 0) return NERAF_OK;
  const int64_t want = ceil_div(n, 256 * 4);
  const int64_t cap = (int64_t)sm_count() * 16;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
  loss_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, gt, n, n_total, criterion, sums, upstream_sc,
                                                               upstream_mag, w_sc, w_mag, dpred);
  NERAF_CHECK_LAUNCH("loss_backward_kernel");
  return NERAF_OK;
}
