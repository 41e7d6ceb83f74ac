// Small HBM-bound helpers around the GEMMs of the acoustic field:
//  * fp32 -> bf16 operand copies (row-major and transposed) for the tcgen05 path,
//  * the gradients of the batch-invariant grid-feature block of layer 1 hoisted out of the batch
//    (NeRAF_model.py:557-560 expands the same 1024 values to every row): the rank-1 weight
//    gradient and the gradient w.r.t. the grid feature (the forward mat-vec lives in encode.cu's prep kernel),
//  * bias gradients (column sums), and the gradient through the 10*tanh heads (NeRAF_field.py:57-58).
#include <algorithm>

#include "common.cuh"
#include <cstdlib>

#include "kernels.h"
#include "loss_math.cuh"

namespace neraf {

// ---------------------------------------------------------------------------------------------
// fp32 (rows, cols) -> bf16 (rows, cols) and/or bf16 (cols, rows)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) convert_bf16_kernel(const float* __restrict__ in, int64_t rows, int64_t cols,
                                                           int64_t ld_in, __nv_bfloat16* __restrict__ out,
                                                           int64_t ld_out, __nv_bfloat16* __restrict__ out_t,
                                                           int64_t ld_t) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + i * 8, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = __ldg(in + r * ld_in + c);
      if (out) out[r * ld_out + c] = __float2bfloat16_rn(v);
    }
    tile[ty + i * 8][tx] = v;
  }
  if (!out_t) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t c = c0 + ty + i * 8, r = r0 + tx;
    if (r < rows && c < cols) out_t[c * ld_t + r] = __float2bfloat16_rn(tile[tx][ty + i * 8]);
  }
}

int convert_bf16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out, void* out_t,
                 int64_t ld_t, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return NERAF_OK;
  NERAF_REQUIRE(in && (out || out_t), "convert_bf16: null pointer");
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  NERAF_REQUIRE(grid.y <= 65535, "convert_bf16: too many rows for one launch (%lld)", (long long)rows);
  convert_bf16_kernel<<<grid, 256, 0, stream>>>(in, rows, cols, ld_in, (__nv_bfloat16*)out, ld_out,
                                                (__nv_bfloat16*)out_t, ld_t);
  NERAF_CHECK_LAUNCH("convert_bf16_kernel");
  return NERAF_OK;
}

// ---------------------------------------------------------------------------------------------
// bf16 operand copies of SEVERAL fp32 matrices in one launch (the per-step re-pack of the MLP parameters):
// matrix i is (rows, cols) fp32 with row stride ld_in -> bf16 row stride ld_out (pad columns zero-filled).
// Work unit = 8 consecutive columns of one row (two 16-byte loads, one 16-byte store when aligned).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_list_kernel(const PackList L) {
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t u = tid; u < L.total_units; u += nthreads) {
    int i = 0;
    while (i + 1 < L.n && u >= L.m[i + 1].unit_start) ++i;
    const PackMatrix& M = L.m[i];
    const int64_t local = u - M.unit_start;
    const int64_t upr = M.ld_out / 8;                       // units per row (ld_out is a multiple of 8)
    const int64_t r = local / upr, c = (local - r * upr) * 8;
    const float* src = M.in + r * M.ld_in + c;
    __nv_bfloat16* dst = M.out + r * M.ld_out + c;
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
    if (M.vec && c + 8 <= M.cols) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
      h[2] = __floats2bfloat162_rn(b.x, b.y); h[3] = __floats2bfloat162_rn(b.z, b.w);
    } else {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = c + e < M.cols ? __ldg(src + e) : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    }
    *reinterpret_cast<uint4*>(dst) = pk;
  }
  for (int64_t j = tid; j < L.n_copy; j += nthreads) {     // fp32 vector gathered from up to 8 segments (head biases)
    const int seg = (int)(j / L.copy_width);
    L.copy_dst[j] = __ldg(L.copy_src[seg] + (j - (int64_t)seg * L.copy_width));
  }
}

int pack_list(PackList& L, cudaStream_t stream) {
  int64_t units = 0;
  for (int i = 0; i < L.n; ++i) {
    PackMatrix& M = L.m[i];
    NERAF_REQUIRE(M.in && M.out && M.ld_out % 8 == 0 && M.ld_out >= M.cols && ((uintptr_t)M.out % 16) == 0,
                  "pack_list: matrix %d: output must be 16-byte aligned with a row stride that is a multiple of 8", i);
    M.vec = (M.ld_in % 4 == 0) && (((uintptr_t)M.in % 16) == 0);
    M.unit_start = units;
    units += M.rows * (M.ld_out / 8);
  }
  L.total_units = units;
  if (units == 0 && L.n_copy == 0) return NERAF_OK;
  const int64_t want = ceil_div(units, 256 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
  pack_list_kernel<<<grid, 256, 0, stream>>>(L);
  NERAF_CHECK_LAUNCH("pack_list_kernel");
  return NERAF_OK;
}

// ---------------------------------------------------------------------------------------------
// Both gradients of the hoisted grid block of layer 1 in ONE pass over the rows (they share their input db1 = s):
//   dW1[n, :K] = s[n] * g        (rank 1: 4 K bytes written per row)
//   dg[k]     += sum_n W1[n, k] s[n]                (4 K bytes read per row; dg pre-zeroed)
// HBM-bound (20.9 MB written + 20.9 MB read at n1 = 5096, K = 1024).  A block owns GG_ROWS consecutive rows and ALL K
// columns: thread t holds columns t, t + 256, ... (coalesced 1 KB runs; W1's row stride 1187 rules out 16-byte
// vectors), every load of the block's rows is issued before the first use (GG_ROWS x K/256 independent 4-byte loads
// per thread in flight), the column sums stay in registers and leave as one fire-and-forget atomic per column and block.
// compact (optional): the per-query block of dW1, (N, E) with row stride ld_c, copied into dW[:, K:K+E] by the same
// rows (data parallel: that block was all-reduced in a compact buffer, see neraf_field_backward_dp).
// ---------------------------------------------------------------------------------------------
constexpr int GG_ROWS = 8;
constexpr int GG_TPB = 256;

// dg is a sum over all N rows of every column: the row blocks add their partial sums into an fp64 scratch vector
// (atomics: their order varies from run to run, but at fp64 the variation is ~1e-16 and vanishes in the rounding to
// fp32 -- with fp32 atomics the last bits of dg changed between runs, and the bf16 producer behind it (whose backward
// rounds dg to bf16) turned that into 3e-3 jumps of its parameter gradients and, under data parallelism, into replicas
// that drift apart).  The last block to finish (ticket) rounds the sums into dg and leaves scratch and ticket zero.
template <int KCH>      // column chunks of 256 per thread: K <= 256 * KCH
__global__ void __launch_bounds__(GG_TPB) grid_grads_kernel(const float* __restrict__ s, const float* __restrict__ g,
                                                            const float* __restrict__ W, int64_t ldw, int64_t N, int64_t K,
                                                            float* __restrict__ dW, float* __restrict__ dg,
                                                            double* __restrict__ dg64, unsigned int* __restrict__ ticket,
                                                            const void* __restrict__ compact, int compact_bf16, int64_t E,
                                                            int64_t ld_c, int gg_blocks,
                                                            const __nv_bfloat16* __restrict__ widen_src,
                                                            float* __restrict__ widen_dst, int64_t widen_n) {
  asm volatile("griddepcontrol.wait;" ::: "memory");        // programmatic dependent launch: the backward GEMMs are complete
  const int t = threadIdx.x;
  const int widen_blocks = (int)gridDim.x - gg_blocks;      // they come FIRST: their grid-stride loops want the first wave
  if ((int)blockIdx.x < widen_blocks) {
    // extra blocks (data parallel): the exchanged bf16 gradient sums -> the fp32 .grad buffer, 16 bytes in, 32 out,
    // four loads in flight per thread
    const int64_t n8 = widen_n / 8;
    const int64_t stride = (int64_t)widen_blocks * GG_TPB;
    auto widen8 = [&](int64_t i, const uint4& raw) {
      const unsigned int w[4] = {raw.x, raw.y, raw.z, raw.w};
      float4 lo, hi;
      lo.x = __uint_as_float(w[0] << 16); lo.y = __uint_as_float(w[0] & 0xffff0000u);
      lo.z = __uint_as_float(w[1] << 16); lo.w = __uint_as_float(w[1] & 0xffff0000u);
      hi.x = __uint_as_float(w[2] << 16); hi.y = __uint_as_float(w[2] & 0xffff0000u);
      hi.z = __uint_as_float(w[3] << 16); hi.w = __uint_as_float(w[3] & 0xffff0000u);
      __stcs(reinterpret_cast<float4*>(widen_dst) + 2 * i, lo);
      __stcs(reinterpret_cast<float4*>(widen_dst) + 2 * i + 1, hi);
    };
    int64_t i = (int64_t)blockIdx.x * GG_TPB + t;
    for (; i + 3 * stride < n8; i += 4 * stride) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = __ldcs(reinterpret_cast<const uint4*>(widen_src) + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) widen8(i + u * stride, raw[u]);
    }
    for (; i < n8; i += stride) widen8(i, __ldcs(reinterpret_cast<const uint4*>(widen_src) + i));
    if (blockIdx.x == 0)
      for (int64_t k = n8 * 8 + t; k < widen_n; k += GG_TPB) widen_dst[k] = __bfloat162float(widen_src[k]);
    return;
  }
  // Row groups of GG_ROWS rows, several per block: the column sums stay in registers across them (one fp64 atomic per
  // column and BLOCK instead of per group: 218 k instead of 652 k on 1024 addresses) and the grid is resident at once
  // (one block per group was 637 blocks at 3 per SM, i.e. two waves): 19.7 -> 15.9 us cold (ncu, one GPU)
  const int b = (int)blockIdx.x - widen_blocks;
  const int64_t groups = (N + GG_ROWS - 1) / GG_ROWS;
  float gk[KCH], acc[KCH];
#pragma unroll
  for (int j = 0; j < KCH; ++j) {
    const int64_t k = t + 256 * j;
    gk[j] = k < K ? __ldg(g + k) : 0.f;
    acc[j] = 0.f;
  }
  for (int64_t grp = b; grp < groups; grp += gg_blocks) {
    const int64_t n0 = grp * GG_ROWS;
    const int rows = (int)(N - n0 < GG_ROWS ? N - n0 : GG_ROWS);
    float w[GG_ROWS][KCH], sn[GG_ROWS];
#pragma unroll
    for (int r = 0; r < GG_ROWS; ++r) sn[r] = r < rows ? __ldg(s + n0 + r) : 0.f;
    if (dg) {
#pragma unroll
      for (int r = 0; r < GG_ROWS; ++r)
#pragma unroll
        for (int j = 0; j < KCH; ++j) {
          const int64_t k = t + 256 * j;
          w[r][j] = (r < rows && k < K) ? __ldg(W + (n0 + r) * ldw + k) : 0.f;
        }
    }
    if (dW) {
#pragma unroll
      for (int r = 0; r < GG_ROWS; ++r) {
        if (r < rows) {
          float* row = dW + (n0 + r) * ldw;
#pragma unroll
          for (int j = 0; j < KCH; ++j) {
            const int64_t k = t + 256 * j;
            if (k < K) row[k] = sn[r] * gk[j];
          }
          if (compact) {
            if (compact_bf16)
              for (int64_t e = t; e < E; e += GG_TPB)
                row[K + e] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(compact)[(n0 + r) * ld_c + e]);
            else
              for (int64_t e = t; e < E; e += GG_TPB) row[K + e] = __ldg(reinterpret_cast<const float*>(compact) + (n0 + r) * ld_c + e);
          }
        }
      }
    }
    if (dg) {
#pragma unroll
      for (int r = 0; r < GG_ROWS; ++r)
#pragma unroll
        for (int j = 0; j < KCH; ++j) acc[j] = fmaf(w[r][j], sn[r], acc[j]);
    }
  }
  if (dg) {
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
      const int64_t k = t + 256 * j;
      if (k < K) atomicAdd(dg64 + k, (double)acc[j]);
    }
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (t == 0) last = atomicAdd(ticket, 1u) == (unsigned int)gg_blocks - 1;
    __syncthreads();
    if (last) {
      __threadfence();
      for (int64_t k = t; k < K; k += GG_TPB) {
        dg[k] = (float)__ldcg(dg64 + k);
        dg64[k] = 0.0;
      }
      if (t == 0) *ticket = 0u;
    }
  }
}

// bf16 -> fp32 widening of the exchanged gradient sums as a launch of its own (data parallel): 16 bytes in, 32 out, four
// loads in flight per thread, 32 registers -> full occupancy.  neraf_field_grid_grads runs it on the helper stream BESIDE the
// grid-gradient kernel: fused into that kernel's grid (extra blocks) the two roles shared its 80-register, 3-blocks-per-SM
// footprint and the launch was latency-bound at 37 us for 130 MB (ncu: 31 % occupancy, DRAM 26 % busy).
__global__ void __launch_bounds__(256) widen_bf16_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t n8 = n / 8;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  auto widen8 = [&](int64_t i, const uint4& raw) {
    const unsigned int w[4] = {raw.x, raw.y, raw.z, raw.w};
    float4 lo, hi;
    lo.x = __uint_as_float(w[0] << 16); lo.y = __uint_as_float(w[0] & 0xffff0000u);
    lo.z = __uint_as_float(w[1] << 16); lo.w = __uint_as_float(w[1] & 0xffff0000u);
    hi.x = __uint_as_float(w[2] << 16); hi.y = __uint_as_float(w[2] & 0xffff0000u);
    hi.z = __uint_as_float(w[3] << 16); hi.w = __uint_as_float(w[3] & 0xffff0000u);
    __stcs(reinterpret_cast<float4*>(dst) + 2 * i, lo);
    __stcs(reinterpret_cast<float4*>(dst) + 2 * i + 1, hi);
  };
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n8; i += 4 * stride) {
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) raw[u] = __ldcs(reinterpret_cast<const uint4*>(src) + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) widen8(i + u * stride, raw[u]);
  }
  for (; i < n8; i += stride) widen8(i, __ldcs(reinterpret_cast<const uint4*>(src) + i));
  if (blockIdx.x == 0)
    for (int64_t k = n8 * 8 + threadIdx.x; k < n; k += blockDim.x) dst[k] = __bfloat162float(src[k]);
}

int widen_bf16(const void* src, float* dst, int64_t n, cudaStream_t stream) {
  if (n <= 0) return NERAF_OK;
  NERAF_REQUIRE(src && dst && ((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 15) == 0, "widen_bf16: buffers must be 16-byte aligned");
  const int64_t want = ceil_div(n, 8 * 256 * 4);
  const int64_t cap = 8 * (int64_t)sm_count();
  widen_bf16_kernel<<<(unsigned)std::max<int64_t>(1, std::min(want, cap)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), dst, n);
  NERAF_CHECK_LAUNCH("widen_bf16_kernel");
  return NERAF_OK;
}

// scratch: G doubles + one u32 ticket (zero before the first use, left zero)
int grid_grads(const float* s, const float* g, const float* W, int64_t ldw, int64_t N, int64_t K, float* dW, float* dg,
               void* scratch, cudaStream_t stream, const void* compact, int64_t E, int64_t ld_c, bool compact_bf16,
               const void* widen_src, float* widen_dst, int64_t widen_n) {
  if (N <= 0 || K <= 0 || (!dW && !dg)) return NERAF_OK;
  NERAF_REQUIRE(widen_n == 0 || (widen_src && widen_dst && ((uintptr_t)widen_src & 15) == 0 && ((uintptr_t)widen_dst & 15) == 0),
                "grid_grads: the buffers to widen must be 16-byte aligned");
  NERAF_REQUIRE(K <= 256 * 8, "grid_grads: at most 2048 grid-feature columns (got %lld)", (long long)K);
  NERAF_REQUIRE(!dg || (scratch && ((uintptr_t)scratch & 7) == 0), "grid_grads: dg needs an 8-byte aligned scratch buffer");
  double* dg64 = reinterpret_cast<double*>(scratch);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(dg64 + K);
  const int64_t groups = ceil_div(N, GG_ROWS);
  const int64_t per_block = std::max<int64_t>(1, ceil_div(groups, 3 * (int64_t)sm_count()));   // 80 registers: 3 blocks per SM resident
  const int gg_blocks = (int)ceil_div(groups, per_block);
  const int widen_blocks = widen_n > 0 ? (int)std::min<int64_t>(ceil_div(widen_n, 8 * GG_TPB * 4), 4 * sm_count()) : 0;
  const unsigned grid = (unsigned)(gg_blocks + widen_blocks);
  const void* cp = dW ? compact : nullptr;
  const int cbf = compact_bf16 ? 1 : 0;
  const __nv_bfloat16* wsrc = reinterpret_cast<const __nv_bfloat16*>(widen_src);
  // no memset in front of it: the kernel follows the backward's job-list launch directly and may be set up under its tail
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(GG_TPB); cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  if (K <= 256 * 4)
    NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, grid_grads_kernel<4>, s, g, W, ldw, N, K, dW, dg, dg64, ticket, cp, cbf, E, ld_c,
                                        gg_blocks, wsrc, widen_dst, widen_n));
  else
    NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, grid_grads_kernel<8>, s, g, W, ldw, N, K, dW, dg, dg64, ticket, cp, cbf, E, ld_c,
                                        gg_blocks, wsrc, widen_dst, widen_n));
  NERAF_CHECK_LAUNCH("grid_grads_kernel");
  return NERAF_OK;
}

// ---------------------------------------------------------------------------------------------
// Column sums (bias gradients)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) colsum_f32_kernel(const float* __restrict__ X, int64_t M, int64_t N, int64_t ld,
                                                          float* __restrict__ out) {
  __shared__ float red[32][33];
  const int nx = threadIdx.x % 32, my = threadIdx.x / 32;
  const int64_t n = (int64_t)blockIdx.x * 32 + nx;
  float acc = 0.f;
  if (n < N)
    for (int64_t m = my; m < M; m += 32) acc += __ldg(X + m * ld + n);
  red[my][nx] = acc;
  __syncthreads();
  if (my == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][nx];
    out[n] = t;
  }
}

int colsum_f32(const float* X, int64_t M, int64_t N, int64_t ld, float* out, cudaStream_t stream) {
  if (N <= 0) return NERAF_OK;
  colsum_f32_kernel<<<(unsigned)ceil_div(N, 32), 1024, 0, stream>>>(X, M, N, ld, out);
  NERAF_CHECK_LAUNCH("colsum_f32_kernel");
  return NERAF_OK;
}

// ---------------------------------------------------------------------------------------------
// Gradient through y = 10*tanh(z):  dz = dout * (10 - y*y/10)
// ---------------------------------------------------------------------------------------------
constexpr int kHeadRows = 32;        // rows per block of head_backward_kernel (4 per thread, all loads in flight at once)

// Block = 32 columns x 8 row lanes over kHeadRows rows: coalesced 128-byte row segments, the bias gradient
// (column sums) is accumulated in registers and leaves the block as one atomic per column.
// LOSS: `dout` holds the targets and the upstream gradient is the spectral loss's, evaluated here with exactly the
// arithmetic of loss_backward_kernel (loss.cu) -- fused, the step saves one launch and the dpred round trip.
template <bool LOSS>
__global__ void __launch_bounds__(256) head_backward_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                                            int64_t M, int64_t N, float* __restrict__ dz_f32,
                                                            int64_t ld_f32, __nv_bfloat16* __restrict__ dz_bf16,
                                                            int64_t ld_bf16, HeadColsum cs, neraf_loss_grad lg) {
  __shared__ float red[8][32];
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  if (LOSS && lg.losses && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {    // the loss values themselves
    finalize(lg.sums, lg.n_total, lg.criterion, lg.w_sc, lg.w_mag, lg.losses);
    if (lg.total) lg.total[0] = lg.losses[0] + lg.losses[1];
  }
  const int64_t c = (int64_t)blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * kHeadRows;
  const int64_t r1 = r0 + kHeadRows < M ? r0 + kHeadRows : M;
  float sum = 0.f;
  if (c < N) {
    float yy[kHeadRows / 8], dd[kHeadRows / 8];
#pragma unroll
    for (int i = 0; i < kHeadRows / 8; ++i) {
      const int64_t r = r0 + ty + 8 * i;
      yy[i] = r < r1 ? __ldg(y + r * N + c) : 0.f;
      dd[i] = r < r1 ? __ldg(dout + r * N + c) : 0.f;
    }
    if (LOSS) {
      float a = 0.f;
      if (lg.criterion != NERAF_CRIT_MSE) a = (float)((double)lg.w_sc / (sqrt(lg.sums[0]) * sqrt(lg.sums[1])));
      const float b = (float)((double)lg.w_mag / (double)lg.n_total);
      const bool l1 = lg.criterion == NERAF_CRIT_SC_SLL1;
#pragma unroll
      for (int i = 0; i < kHeadRows / 8; ++i) {
        const float x = yy[i], t = dd[i], d = x - t;
        float g = l1 ? b * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.f * b * d;
        if (lg.criterion != NERAF_CRIT_MSE) {
          const float ex = exp_fma(x), et = exp_fma(t);
          g += a * (ex - et) * ex;
        }
        dd[i] = g;
      }
    }
#pragma unroll
    for (int i = 0; i < kHeadRows / 8; ++i) {
      const int64_t r = r0 + ty + 8 * i;
      const float v = dd[i] * (10.f - yy[i] * yy[i] * 0.1f);
      if (r < r1) {
        if (dz_f32) dz_f32[r * ld_f32 + c] = v;
        if (dz_bf16) dz_bf16[r * ld_bf16 + c] = __float2bfloat16_rn(v);
      }
      sum += v;
    }
  }
  if (cs.width == 0) return;
  red[ty][tx] = sum;
  __syncthreads();
  if (ty == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    atomicAdd(cs.ptr[c / cs.width] + (c % cs.width), t);
  }
}

int head_backward(const float* dout, const float* y, int64_t M, int64_t N, float* dz_f32, int64_t ld_f32, void* dz_bf16,
                  int64_t ld_bf16, float* const* colsum, int64_t head_width, cudaStream_t stream,
                  const neraf_loss_grad* loss, float* zero, int64_t n_zero) {
  if (M <= 0 || N <= 0) return NERAF_OK;
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, kHeadRows));
  NERAF_REQUIRE(grid.y <= 65535, "head_backward: batch too large for one launch (%lld)", (long long)M);
  HeadColsum cs = {};
  if (colsum) {
    const int64_t heads = ceil_div(N, head_width);
    NERAF_REQUIRE(heads <= 8, "head_backward: at most 8 heads");
    for (int64_t c = 0; c < heads; ++c) cs.ptr[c] = colsum[c];
    cs.width = head_width;
  }
  if (loss && loss->fuse_sums) return loss_head_fused(y, M, N, dz_f32, ld_f32, dz_bf16, ld_bf16, cs, loss, zero, n_zero, stream);
  NERAF_REQUIRE(n_zero == 0, "head_backward: only the fused loss kernel clears a region");
  if (loss) {
    NERAF_REQUIRE(loss->gt && loss->sums && loss->n_total > 0 && loss->criterion >= 0 && loss->criterion <= 2,
                  "head_backward: bad neraf_loss_grad");
    head_backward_kernel<true><<<grid, 256, 0, stream>>>(loss->gt, y, M, N, dz_f32, ld_f32, (__nv_bfloat16*)dz_bf16,
                                                         ld_bf16, cs, *loss);
  } else {
    head_backward_kernel<false><<<grid, 256, 0, stream>>>(dout, y, M, N, dz_f32, ld_f32, (__nv_bfloat16*)dz_bf16,
                                                          ld_bf16, cs, neraf_loss_grad{});
  }
  NERAF_CHECK_LAUNCH("head_backward_kernel");
  return NERAF_OK;
}

}  // namespace neraf

extern "C" int neraf_convert_bf16(const float* in, int64_t rows, int64_t cols, int64_t ld_in, void* out, int64_t ld_out,
                                  void* out_t, int64_t ld_t, neraf_stream_t stream) {
  return neraf::convert_bf16(in, rows, cols, ld_in, out, ld_out, out_t, ld_t, (cudaStream_t)stream);
}
