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
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "loss_math.cuh"

namespace neraf {

constexpr float kEpsMag = 1e-3f;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-element terms in fp32.  Strips of up to 8 elements are summed in fp32 (pairwise) and enter the fp64 running
// sums once per strip: per element that is half a float -> double conversion instead of four (the conversions and the
// fp64 adds, not HBM, bounded the first version at 0.60 of the HBM roofline: profiles/r02a_bench.json), and the
// rounding of an 8-term fp32 sum of non-negative terms (< 5e-7 relative) is far inside the 1e-5 gate.
struct Terms { float num, den, sq, ab; };
__device__ __forceinline__ Terms terms(float x, float y) {
  const float ex = exp_fma(x), ey = exp_fma(y);
  const float dm = ey - ex;                 // (e^y - 1e-3) - (e^x - 1e-3)
  const float ym = ey - kEpsMag;
  const float d = y - x;
  return Terms{dm * dm, ym * ym, d * d, fabsf(d)};
}
__device__ __forceinline__ Terms operator+(const Terms& a, const Terms& b) {
  return Terms{a.num + b.num, a.den + b.den, a.sq + b.sq, a.ab + b.ab};
}
__device__ __forceinline__ Terms terms4(const float4& x, const float4& y) {
  return (terms(x.x, y.x) + terms(x.y, y.y)) + (terms(x.z, y.z) + terms(x.w, y.w));
}
__device__ __forceinline__ void add_strip(const Terms& t, double& s_num, double& s_den, double& s_sq, double& s_abs) {
  s_num += (double)t.num; s_den += (double)t.den; s_sq += (double)t.sq; s_abs += (double)t.ab;
}
__device__ __forceinline__ void accumulate(float x, float y, double& s_num, double& s_den, double& s_sq, double& s_abs) {
  add_strip(terms(x, y), s_num, s_den, s_sq, s_abs);
}

// Partial sums of one thread over elements [0, n) taken with stride (tid, nthreads): two 16-byte loads per array in
// flight per iteration.
__device__ __forceinline__ void thread_sums(const float* __restrict__ pred, const float* __restrict__ gt, int64_t n,
                                            int64_t tid, int64_t nthreads, double& s_num, double& s_den, double& s_sq,
                                            double& s_abs) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* g4 = reinterpret_cast<const float4*>(gt);
  int64_t i = tid;
  for (; i + nthreads < n4; i += 2 * nthreads) {
    const float4 x0 = __ldg(p4 + i), y0 = __ldg(g4 + i);
    const float4 x1 = __ldg(p4 + i + nthreads), y1 = __ldg(g4 + i + nthreads);
    add_strip(terms4(x0, y0) + terms4(x1, y1), s_num, s_den, s_sq, s_abs);
  }
  if (i < n4) add_strip(terms4(__ldg(p4 + i), __ldg(g4 + i)), s_num, s_den, s_sq, s_abs);
  for (int64_t k = n4 * 4 + tid; k < n; k += nthreads) accumulate(__ldg(pred + k), __ldg(gt + k), s_num, s_den, s_sq, s_abs);
}

// FUSED != 0: the last block to finish (ticket in sums[4], reinterpreted as an integer) also forms the two losses,
// saving the separate finalize launch of the single-GPU path.
struct FinalizeArgs { int fused; int64_t n_total; int criterion; float w_sc, w_mag; float* losses; };

__global__ void __launch_bounds__(256) loss_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                        int64_t n, double* __restrict__ sums, const FinalizeArgs fin) {
  double s_num = 0, s_den = 0, s_sq = 0, s_abs = 0;
  thread_sums(pred, gt, n, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x, s_num, s_den,
              s_sq, s_abs);

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
  auto grad = [&](float x, float y) -> float {
    const float d = x - y;
    float g = l1 ? b * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * b * d;
    if (criterion != NERAF_CRIT_MSE) {
      const float ex = exp_fma(x), ey = exp_fma(y);
      g += a * (ex - ey) * ex;
    }
    return g;
  };
  // 16-byte loads and stores (8 B read + 4 B written per element); scalar tail / unaligned buffers
  const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(dpred)) & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* g4 = reinterpret_cast<const float4*>(gt);
  float4* d4 = reinterpret_cast<float4*>(dpred);
  for (int64_t i = tid; i < n4; i += nthreads) {
    const float4 x = __ldg(p4 + i), y = __ldg(g4 + i);
    __stcs(d4 + i, make_float4(grad(x.x, y.x), grad(x.y, y.y), grad(x.z, y.z), grad(x.w, y.w)));
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += nthreads) dpred[i] = grad(__ldg(pred + i), __ldg(gt + i));
}

// ---------------------------------------------------------------------------------------------
// The whole loss in ONE launch (the graphed training step): partial sums -> grid barrier (data parallel: the last
// block to arrive trades the four sums with the other ranks through peer-mapped memory, neraf_rank_exchange) ->
// loss values, d loss / d pred, the heads' 10 tanh' factor, the bf16 head gradient and the heads' bias gradients.
// Replaces loss_sums_kernel + (all-reduce) + head_backward_kernel: one launch, no dependent-launch gap between the two
// halves, and no host-issued collective between forward and backward.  The bias-gradient buffers that the second half
// and the backward GEMMs accumulate into are cleared by the first half (the barrier orders the two), which removes
// the memset node in front of the backward.  Every block must be resident (the grid never exceeds what fits).
// ---------------------------------------------------------------------------------------------
struct LossHeadArgs {
  const float* pred; const float* gt; int64_t M, N;
  float* dz_f32; int64_t ld_f32; __nv_bfloat16* dz_bf16; int64_t ld_bf16;
  HeadColsum cs;
  double* sums; int64_t n_total; int criterion; float w_sc, w_mag; float* losses; float* total;
  unsigned int* sync;                // [0] arrivals  [1] generation  [2] step number (data parallel)
  float4* zero; int64_t n_zero4;     // cleared before the barrier
  int world, rank; void* peers[NERAF_MAX_RANKS];   // world == 0: no exchange (single process)
};

constexpr int kXchgFlagOffset = 2 * NERAF_MAX_RANKS * 4 * 8;     // slots: double [parity][rank][4]; flags: u32 [parity][rank]
static_assert(kXchgFlagOffset + 2 * NERAF_MAX_RANKS * 4 <= NERAF_EXCHANGE_BYTES, "exchange buffer layout");
constexpr long long kSpinLimit = 4000000000LL;                    // ~2 s: a lost peer traps instead of hanging the GPU

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __maxnreg__(40) loss_head_kernel(const LossHeadArgs A) {
  asm volatile("griddepcontrol.wait;" ::: "memory");        // programmatic dependent launch: the forward has completed
  __shared__ double red[4][8];
  __shared__ float csum[8][32];
  __shared__ unsigned int s_gen;
  __shared__ bool s_last;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;

  // ---- first half: clear the bias gradients, partial sums of this rank's elements
  for (int64_t i = gtid; i < A.n_zero4; i += nthreads) A.zero[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  double s_num = 0, s_den = 0, s_sq = 0, s_abs = 0;
  thread_sums(A.pred, A.gt, A.M * A.N, gtid, nthreads, s_num, s_den, s_sq, s_abs);
  s_num = warp_sum(s_num); s_den = warp_sum(s_den); s_sq = warp_sum(s_sq); s_abs = warp_sum(s_abs);
  if (lane == 0) { red[0][warp] = s_num; red[1][warp] = s_den; red[2][warp] = s_sq; red[3][warp] = s_abs; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(A.sums + threadIdx.x, t);
  }
  __threadfence();
  __syncthreads();

  // ---- grid barrier (sense reversal on sync[1]; the last arriver resets the count, so the words are reusable)
  if (threadIdx.x == 0) {
    s_gen = ld_acquire_gpu(A.sync + 1);                      // read BEFORE arriving: it cannot have advanced yet
    __threadfence();
    s_last = atomicAdd(A.sync, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();                                         // every block's atomics and clears are visible here
    if (A.world > 0) {
      // Trade the four sums with the other ranks: store mine into slot [parity][rank] of EVERY rank's buffer, raise
      // the flag next to it, wait for the flags of all ranks in my own buffer, add the slots in rank order (the same
      // order on every rank: bit-identical global sums everywhere).  Two slot sets alternate with the step number: a
      // rank can only overwrite set p again two steps later, which needs every peer's flag of the step in between,
      // which a peer raises after it has read set p.
      const unsigned int seq = A.sync[2] + 1u;
      const int par = (int)(seq & 1u);
      if (threadIdx.x < A.world) {
        uint8_t* peer = reinterpret_cast<uint8_t*>(A.peers[threadIdx.x]);
        double* slot = reinterpret_cast<double*>(peer) + (par * NERAF_MAX_RANKS + A.rank) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) st_relaxed_sys_f64(slot + k, __ldcg(A.sums + k));
        __threadfence_system();                             // the slot is visible before the flag (one fence, plain store)
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<unsigned int*>(peer + kXchgFlagOffset) +
                                                                 par * NERAF_MAX_RANKS + A.rank), "r"(seq) : "memory");
        const unsigned int* flag = reinterpret_cast<const unsigned int*>(reinterpret_cast<uint8_t*>(A.peers[A.rank]) + kXchgFlagOffset) +
                                   par * NERAF_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(flag) != seq) {
          __nanosleep(100);
          if (clock64() - t0 > kSpinLimit) __trap();
        }
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        const double* mine = reinterpret_cast<const double*>(A.peers[A.rank]) + par * NERAF_MAX_RANKS * 4;
        double t = 0;
        for (int q = 0; q < A.world; ++q) t += ld_relaxed_sys_f64(mine + q * 4 + threadIdx.x);
        A.sums[threadIdx.x] = t;
      }
      __syncthreads();
      if (threadIdx.x == 0) A.sync[2] = seq;
    }
    if (threadIdx.x == 0) {
      __threadfence();
      if (A.losses) {
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldcg(A.sums + i);
        finalize(v, A.n_total, A.criterion, A.w_sc, A.w_mag, A.losses);
        if (A.total) A.total[0] = A.losses[0] + A.losses[1];
      }
      A.sync[0] = 0u;
      __threadfence();
      atomicAdd(A.sync + 1, 1u);                             // open the barrier
    }
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire_gpu(A.sync + 1) == s_gen) {
      __nanosleep(64);
      if (clock64() - t0 > kSpinLimit) __trap();
    }
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // every block is resident: the backward may move in

  // ---- second half: gradient of the loss through 10 tanh, bf16 / fp32 head gradient, bias gradients (column sums)
  float a = 0.f;
  if (A.criterion != NERAF_CRIT_MSE) a = (float)((double)A.w_sc / (sqrt(__ldcg(A.sums)) * sqrt(__ldcg(A.sums + 1))));
  const float b = (float)((double)A.w_mag / (double)A.n_total);
  const bool l1 = A.criterion == NERAF_CRIT_SC_SLL1;
  const int tx = lane, ty = warp;
  const int64_t tiles_n = (A.N + 31) / 32, tiles_m = (A.M + 31) / 32;
  for (int64_t item = blockIdx.x; item < tiles_n * tiles_m; item += gridDim.x) {
    const int64_t c = (item % tiles_n) * 32 + tx;
    const int64_t r0 = (item / tiles_n) * 32;
    float sum = 0.f;
    if (c < A.N) {
      float yy[4], dd[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        yy[i] = r < A.M ? __ldg(A.pred + r * A.N + c) : 0.f;
        dd[i] = r < A.M ? __ldg(A.gt + r * A.N + c) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        const float x = yy[i], t = dd[i], d = x - t;
        float g = l1 ? b * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * b * d;
        if (A.criterion != NERAF_CRIT_MSE) {
          const float ex = exp_fma(x), et = exp_fma(t);
          g += a * (ex - et) * ex;
        }
        const float v = g * (10.f - x * x * 0.1f);
        if (r < A.M) {
          if (A.dz_f32) A.dz_f32[r * A.ld_f32 + c] = v;
          if (A.dz_bf16) A.dz_bf16[r * A.ld_bf16 + c] = __float2bfloat16_rn(v);
          sum += v;
        }
      }
    }
    if (A.cs.width != 0) {
      csum[ty][tx] = sum;
      __syncthreads();
      if (ty == 0 && c < A.N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += csum[i][tx];
        atomicAdd(A.cs.ptr[c / A.cs.width] + (c % A.cs.width), t);
      }
      __syncthreads();
    }
  }
}

int loss_head_fused(const float* y, int64_t M, int64_t N, float* dz_f32, int64_t ld_f32, void* dz_bf16, int64_t ld_bf16,
                    const HeadColsum& cs, const neraf_loss_grad* loss, float* zero, int64_t n_zero, cudaStream_t stream) {
  NERAF_REQUIRE(loss && loss->fuse_sums && loss->sync && loss->gt && loss->sums && loss->n_total > 0,
                "loss_head_fused: bad neraf_loss_grad");
  NERAF_REQUIRE(loss->criterion >= 0 && loss->criterion <= 2, "loss_head_fused: unknown criterion %d", loss->criterion);
  NERAF_REQUIRE(n_zero == 0 || (zero && ((uintptr_t)zero & 15) == 0 && n_zero % 4 == 0),
                "loss_head_fused: the region to clear must be 16-byte aligned and a multiple of 4 floats");
  LossHeadArgs A = {};
  A.pred = y; A.gt = loss->gt; A.M = M; A.N = N;
  A.dz_f32 = dz_f32; A.ld_f32 = ld_f32; A.dz_bf16 = (__nv_bfloat16*)dz_bf16; A.ld_bf16 = ld_bf16;
  A.cs = cs;
  A.sums = const_cast<double*>(loss->sums); A.n_total = loss->n_total; A.criterion = loss->criterion;
  A.w_sc = loss->w_sc; A.w_mag = loss->w_mag; A.losses = loss->losses; A.total = loss->losses ? loss->total : nullptr;
  A.sync = loss->sync;
  A.zero = reinterpret_cast<float4*>(zero); A.n_zero4 = n_zero / 4;
  A.world = 0; A.rank = 0;
  if (loss->exchange) {             // a one-rank exchange runs the same protocol against its own buffer
    const neraf_rank_exchange* x = loss->exchange;
    NERAF_REQUIRE(x->world >= 1 && x->world <= NERAF_MAX_RANKS && x->rank >= 0 && x->rank < x->world,
                  "loss_head_fused: bad rank exchange");
    A.world = x->world; A.rank = x->rank;
    for (int r = 0; r < x->world; ++r) {
      NERAF_REQUIRE(x->peers[r], "loss_head_fused: exchange buffer of rank %d is null", r);
      A.peers[r] = x->peers[r];
    }
  }
  // all blocks wait at the barrier: never launch more than fit at once
  static int per_sm[64] = {0};
  int dev = 0;
  NERAF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && per_sm[dev] == 0) {
    int n = 0;
    NERAF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, loss_head_kernel, 256, 0));
    NERAF_REQUIRE(n > 0, "loss_head_fused: the kernel does not fit on this device");
    per_sm[dev] = n;
  }
  int blocks_per_sm = 4;
  if (const char* e = getenv("NERAF_LOSS_BLOCKS")) blocks_per_sm = atoi(e) > 0 ? atoi(e) : 4;      // tuning
  if (dev >= 0 && dev < 64 && blocks_per_sm > per_sm[dev]) blocks_per_sm = per_sm[dev];
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  const int64_t want = ceil_div(M, 32) * ceil_div(N, 32);
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, loss_head_kernel, A));
  NERAF_CHECK_LAUNCH("loss_head_kernel");
  return NERAF_OK;
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
