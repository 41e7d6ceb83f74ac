// Persistent "job list" tcgen05 kernel: ONE launch executes a whole list of dependent bf16 GEMMs
// (all forward layers of the acoustic MLP, or all dgrad / wgrad GEMMs of its backward).
//
// Why: at the reference batch (2048 STFT columns) a train step is 17 GEMMs of 1-43 GFLOP.  Launched one by
// one (gemm_umma.cu) each fills every SM with a 1-CTA/SM footprint, so nothing overlaps: the small layers are
// latency-bound (few tiles, long K loops), every kernel boundary costs a drain + ramp, and weight-gradient
// GEMMs cannot hide behind the serial dgrad chain (profiles/r01_launches.md).  Here the scheduling unit is the
// TILE: all tiles of all jobs form one topologically ordered list, CTA pairs walk it with a static stride, and
// dependencies are tracked per 256-row block with global counters:
//   * forward / dgrad tile (row block rb, column tile nt) of job j waits until ALL column tiles of row block rb
//     of the producing job have been stored -> different row blocks run in different layers at the same time;
//   * a weight-gradient tile waits for every row block of the dZ it contracts over, and fills the SMs the
//     dgrad chain leaves idle.
// Deadlock freedom: tiles are processed in increasing index order by every unit and every dependency has a
// smaller index, so the smallest unfinished tile is always runnable.
//
// Tile = 256 x bn (bn in {64,128,256}, per job) on a CTA PAIR (tcgen05 cta_group::2, see gemm_umma.cu): each
// CTA stages its 128 rows of A and bn/2 rows of B per 64-deep k-block through a 6-stage TMA/mbarrier ring, the
// leader issues the MMAs, two 256-column TMEM accumulator stages overlap the epilogue with the next tile.
//
// Operand layouts.  K-major operands (row-major with the contraction innermost) are the forward / dgrad case.
// The weight gradient dW = dZ^T X contracts over the BATCH, which is the OUTER dimension of the row-major
// activations: both operands are consumed MN-major (tcgen05 descriptor major bit, TMA boxes of 64 contiguous
// MN elements x 64 batch rows), so no transposed copy of any activation or gradient is ever written.
//
// Epilogue.  TMEM -> registers (lane = row) -> bias / LeakyReLU / 10 tanh / LeakyReLU' gate -> bf16 (or fp32)
// packed into a 128-byte-swizzled 32-row staging tile per warp -> one TMA store per 64 (32) columns; M/N tails
// are clipped by the tensor map.  Bias gradients are column sums taken from the fp32 values with a
// recursive-halving shuffle reduction (31 shuffles per 32 x 32 block) and one atomic per column.
// Cross-SM visibility: stores complete (bulk wait_group) -> __threadfence -> atomicAdd(counter); consumer:
// ld.acquire spin -> fence.proxy.async -> TMA loads.
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "umma_common.cuh"

namespace neraf {
namespace umma {

constexpr int MEGA_THREADS = 192;
constexpr int MEGA_STAGES = 6;
constexpr int MEGA_A_BYTES = BLOCK_M * BLOCK_K * 2;            // 16 KB
constexpr int MEGA_B_BYTES = 128 * BLOCK_K * 2;                // up to bn/2 = 128 rows: 16 KB
constexpr int MEGA_OUT_BYTES = 32 * 128;                       // per-warp staging tile: 32 rows x 128 B
constexpr int MEGA_SMEM_BYTES = 1024 + MEGA_STAGES * (MEGA_A_BYTES + MEGA_B_BYTES) + 4 * MEGA_OUT_BYTES + (2 * MEGA_STAGES + 4) * 8 + 16;
constexpr int MEGA_TMEM_COLS = 512;
constexpr int MEGA_ACC_COLS = 256;
constexpr int MN_BOX_BYTES = 64 * 128;                         // one MN-major box: 64 k-rows x 64 MN elements

struct alignas(64) DeviceJob {
  CUtensorMap tmA, tmB, tmOutB, tmOutF;
  int M, N, K, bn;
  int tile_start, num_m, num_n, cnt_off;
  int wait_job, wait_all, wait_target, wait_nrb, wait_cnt_off;
  int act, a_mn, b_mn, f32_tma;
  const float* bias;
  const __nv_bfloat16* gate; long long ldg;
  int has_out_bf16, pad0;
  float* out_f32; long long ld_f32;
  float* colsum;
};

struct MegaParams {
  int num_jobs, num_tiles;
  unsigned int* counters;
  DeviceJob jobs[NERAF_MEGA_MAX_JOBS];
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (ld_acquire(p) < target) {
    __nanosleep(64);
    if ((++spins & 255u) == 0 && clock64() - t0 > WATCHDOG_CYCLES) __trap();
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// MN-major, 128-byte-swizzled operand: 64 contiguous MN elements per k row (128 B), 8-row groups 1024 B apart
// (SBO), the next block of 64 MN elements one TMA box (8 KB) further (LBO).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(MN_BOX_BYTES >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Column sums of a 32 (rows = lanes) x 32 (registers) block: lane i ends up with sum over lanes of v[i].
__device__ __forceinline__ float column_sums_32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? v[i + s] : v[i];
      const float send = upper ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) umma_mega_kernel(const __grid_constant__ MegaParams P) {
  constexpr int CG = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + MEGA_STAGES * MEGA_A_BYTES;
  uint8_t* out_buf = smem + MEGA_STAGES * (MEGA_A_BYTES + MEGA_B_BYTES);       // 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_buf + 4 * MEGA_OUT_BYTES);
  uint64_t* empty_bar = full_bar + MEGA_STAGES;
  uint64_t* tmem_full = empty_bar + MEGA_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MEGA_STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4 * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc<CG>(tmem_slot, MEGA_TMEM_COLS); tmem_relinquish<CG>(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int j = 0;
      for (int tile = unit; tile < P.num_tiles; tile += num_units) {
        while (j + 1 < P.num_jobs && tile >= P.jobs[j + 1].tile_start) ++j;
        const DeviceJob& J = P.jobs[j];
        const int local = tile - J.tile_start;
        const int mt = local / J.num_n, nt = local % J.num_n;          // row-block major: a row block completes early
        // ---- dependencies: operands written by earlier tiles of this launch (other SMs)
        if (J.wait_job >= 0) {
          if (J.wait_all) {
            for (int rb = 0; rb < J.wait_nrb; ++rb) spin_until(P.counters + J.wait_cnt_off + rb, (unsigned)J.wait_target);
          } else {
            spin_until(P.counters + J.wait_cnt_off + mt, (unsigned)J.wait_target);
          }
          fence_proxy_async_all();
        }
        const int num_kb = (J.K + BLOCK_K - 1) / BLOCK_K;
        const int b_rows = J.bn / CG;
        const uint32_t stage_bytes = (uint32_t)(MEGA_A_BYTES + b_rows * BLOCK_K * 2) * CG;
        const int a_row0 = mt * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M;
        const int b_row0 = nt * J.bn + (int)cta_rank * b_rows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          if (is_leader) mbar_expect_tx(full_bar + stage, stage_bytes);
          uint8_t* sa = smem_a + stage * MEGA_A_BYTES;
          uint8_t* sb = smem_b + stage * MEGA_B_BYTES;
          if (!J.a_mn) {
            tma_load_2d<CG>(sa, &J.tmA, full_bar + stage, kb * BLOCK_K, a_row0);
          } else {                                        // matrix is (K rows, MN cols): coordinates {mn, k}
            tma_load_2d<CG>(sa, &J.tmA, full_bar + stage, a_row0, kb * BLOCK_K);
            tma_load_2d<CG>(sa + MN_BOX_BYTES, &J.tmA, full_bar + stage, a_row0 + 64, kb * BLOCK_K);
          }
          if (!J.b_mn) {
            tma_load_2d<CG>(sb, &J.tmB, full_bar + stage, kb * BLOCK_K, b_row0);
          } else {
            tma_load_2d<CG>(sb, &J.tmB, full_bar + stage, b_row0, kb * BLOCK_K);
            if (b_rows > 64) tma_load_2d<CG>(sb + MN_BOX_BYTES, &J.tmB, full_bar + stage, b_row0 + 64, kb * BLOCK_K);
          }
          if (++stage == MEGA_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && is_leader) {
      int stage = 0; uint32_t phase = 0;
      int it = 0, j = 0;
      for (int tile = unit; tile < P.num_tiles; tile += num_units, ++it) {
        while (j + 1 < P.num_jobs && tile >= P.jobs[j + 1].tile_start) ++j;
        const DeviceJob& J = P.jobs[j];
        const int num_kb = (J.K + BLOCK_K - 1) / BLOCK_K;
        const bool a_mn = J.a_mn != 0, b_mn = J.b_mn != 0;
        const uint32_t idesc = make_idesc(BLOCK_M * CG, J.bn) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
        // per UMMA_K = 16 step: K-major +32 B inside the swizzle row; MN-major +16 k-rows x 128 B
        const uint64_t a_step = a_mn ? (2048 >> 4) : (32 >> 4);
        const uint64_t b_step = b_mn ? (2048 >> 4) : (32 >> 4);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(tmem_empty + as, aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * MEGA_ACC_COLS);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem_a + stage * MEGA_A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * MEGA_B_BYTES);
          const uint64_t adesc = a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
          const uint64_t bdesc = b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16<CG>(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0);
          umma_commit<CG>(empty_bar + stage);
          if (++stage == MEGA_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(tmem_full + as);
      }
    }
    __syncwarp();
  } else {
    const int quarter = warp & 3;
    uint8_t* wbuf = out_buf + (warp - 2) * MEGA_OUT_BYTES;       // this warp's 32 x 128 B staging tile
    float* wbuf_f = reinterpret_cast<float*>(wbuf);
    int it = 0, j = 0;
    for (int tile = unit; tile < P.num_tiles; tile += num_units, ++it) {
      while (j + 1 < P.num_jobs && tile >= P.jobs[j + 1].tile_start) ++j;
      const DeviceJob& J = P.jobs[j];
      const int local = tile - J.tile_start;
      const int mt = local / J.num_n, nt = local % J.num_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int M = J.M, N = J.N, bn = J.bn;
      const float* bias = J.bias;
      const int act = J.act;
      const __nv_bfloat16* gate = J.gate;
      const long long ldg = J.ldg;
      const bool out_b = J.has_out_bf16 != 0;
      float* out_f32 = J.out_f32;
      const long long ld_f32 = J.ld_f32;
      const bool f32_tma = J.f32_tma != 0;
      float* colsum = J.colsum;
      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      const int m_base = mt * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M + quarter * 32;
      const int m = m_base + lane;
      const bool row_ok = m < M;
      const int nchunks = bn / 32;
#pragma unroll 1
      for (int c = 0; c < nchunks; ++c) {
        const int n0 = nt * bn + c * 32;
        if (n0 >= N) break;
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * MEGA_ACC_COLS + c * 32), v);
        if (c == nchunks - 1 || n0 + 32 >= N) {           // accumulator fully read: hand the TMEM stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(tmem_empty + as);
            else mbar_arrive_remote(tmem_empty + as, 0);
          }
        }
        const bool full_chunk = n0 + 32 <= N;
        if (bias) {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += (full_chunk || n0 + q < N) ? __ldg(bias + n0 + q) : 0.f;
        }
        if (act != NERAF_ACT_NONE) {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = apply_act(v[q], act);
        }
        if (gate && row_ok) {
          const __nv_bfloat16* g = gate + (long long)m * ldg + n0;
          if (full_chunk && (ldg % 8 == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 raw = __ldg(reinterpret_cast<const uint4*>(g) + q);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                v[q * 8 + e * 2] *= f.x > 0.f ? 1.f : kLeakySlope;
                v[q * 8 + e * 2 + 1] *= f.y > 0.f ? 1.f : kLeakySlope;
              }
            }
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (n0 + q < N) v[q] *= __bfloat162float(g[q]) > 0.f ? 1.f : kLeakySlope;
          }
        }
        // ---- bf16 row-major output: two 32-column chunks share one 64-column (128 B) staging tile / TMA store
        if (out_b) {
          const int half = c & 1;
          if (half == 0) {                                // staging tile is about to be rewritten
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[i * 8 + 2 * e], v[i * 8 + 2 * e + 1]);
            const int chunk16 = half * 4 + i;             // 16-byte chunk inside the 128-byte row
            *reinterpret_cast<uint4*>(wbuf + lane * 128 + ((chunk16 ^ (lane & 7)) << 4)) = pk;
          }
          const bool last_half = half == 1 || c == nchunks - 1 || n0 + 32 >= N;
          if (last_half) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&J.tmOutB, wbuf, n0 - half * 32, m_base);
              bulk_commit();
            }
          }
        }
        // ---- fp32 row-major output
        if (out_f32) {
          if (f32_tma) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 pk = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
              *reinterpret_cast<float4*>(wbuf + lane * 128 + ((i ^ (lane & 7)) << 4)) = pk;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&J.tmOutF, wbuf, n0, m_base);
              bulk_commit();
            }
          } else {
            // row stride not 16-byte aligned (e.g. (B, 513) outputs): transpose through the staging tile
            // (XOR-swizzled 32 x 32 floats, conflict-free both ways) so that lanes write consecutive columns
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 32; ++q) wbuf_f[lane * 32 + (q ^ lane)] = v[q];
            __syncwarp();
            const int n = n0 + lane;
            if (n < N) {
#pragma unroll 4
              for (int r = 0; r < 32; ++r) {
                const int mm = m_base + r;
                if (mm >= M) break;
                out_f32[(long long)mm * ld_f32 + n] = wbuf_f[r * 32 + (lane ^ r)];
              }
            }
            __syncwarp();
          }
        }
        // ---- bias gradient: column sums of the fp32 values (destroys v)
        if (colsum) {
          if (!row_ok) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] = 0.f;
          }
          const float s = column_sums_32(v, lane);
          if (n0 + lane < N) atomicAdd(colsum + n0 + lane, s);
        }
      }
      if (lane == 0) bulk_wait_all0();               // this warp's TMA stores have been performed
      __threadfence();                               // ... and every store of the tile is visible device-wide
      __syncwarp();
      if (lane == 0) atomicAdd(P.counters + J.cnt_off + mt, 1u);   // row block progress (8 arrivals per tile)
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, MEGA_TMEM_COLS);
  }
}

}  // namespace umma

int get_tensor_map_2d(const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                      CUtensorMap* out);

int mega_run(const MegaJob* jobs, int n_jobs, void* counters, size_t counters_bytes, cudaStream_t stream) {
  using namespace umma;
  NERAF_REQUIRE(jobs && n_jobs > 0 && n_jobs <= NERAF_MEGA_MAX_JOBS, "mega_run: 1..%d jobs", NERAF_MEGA_MAX_JOBS);
  static MegaParams P;          // large: build in static storage (single-threaded driver, see header conventions)
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  P.num_jobs = n_jobs;
  int tile = 0, cnt = 0;
  int cnt_off[NERAF_MEGA_MAX_JOBS], nrb[NERAF_MEGA_MAX_JOBS], num_n[NERAF_MEGA_MAX_JOBS];
  for (int i = 0; i < n_jobs; ++i) {
    const MegaJob& s = jobs[i];
    NERAF_REQUIRE(s.M > 0 && s.N > 0 && s.K > 0, "mega_run: job %d has an empty dimension", i);
    NERAF_REQUIRE(s.bn == 64 || s.bn == 128 || s.bn == 256, "mega_run: job %d tile width %d", i, s.bn);
    NERAF_REQUIRE(!s.b_mn || s.bn >= 128, "mega_run: job %d: an MN-major B operand needs a tile width >= 128", i);
    NERAF_REQUIRE(s.wait_job < i, "mega_run: job %d depends on a later job", i);
    NERAF_REQUIRE(!s.epi.out_bf16_t && !s.epi.accumulate_f32, "mega_run: job %d: transposed / accumulating outputs are not supported", i);
    NERAF_REQUIRE(!(s.epi.out_bf16 && s.epi.out_f32), "mega_run: job %d: one output per job", i);
    NERAF_REQUIRE(s.epi.act >= 0 && s.epi.act <= 2, "mega_run: job %d: unknown activation", i);
    DeviceJob& d = P.jobs[i];
    // K-major operand: matrix (MN rows, K cols), box (rows, 64 k).  MN-major: matrix (K rows, MN cols), box (64 k, 64 mn).
    if (!s.a_mn) NERAF_TRY(get_tensor_map_2d(s.A, 2, s.M, s.K, s.lda, BLOCK_M, BLOCK_K, &d.tmA));
    else NERAF_TRY(get_tensor_map_2d(s.A, 2, s.K, s.M, s.lda, BLOCK_K, 64, &d.tmA));
    if (!s.b_mn) NERAF_TRY(get_tensor_map_2d(s.B, 2, s.N, s.K, s.ldb, s.bn / 2, BLOCK_K, &d.tmB));
    else NERAF_TRY(get_tensor_map_2d(s.B, 2, s.K, s.N, s.ldb, BLOCK_K, 64, &d.tmB));
    d.a_mn = s.a_mn; d.b_mn = s.b_mn;
    d.M = (int)s.M; d.N = (int)s.N; d.K = (int)s.K; d.bn = s.bn;
    d.num_m = (int)ceil_div(s.M, 256); d.num_n = (int)ceil_div(s.N, s.bn);
    d.tile_start = tile; tile += d.num_m * d.num_n;
    d.cnt_off = cnt; cnt_off[i] = cnt; nrb[i] = d.num_m; num_n[i] = d.num_n; cnt += d.num_m;
    d.wait_job = s.wait_job; d.wait_all = s.wait_all;
    if (s.wait_job >= 0) {
      d.wait_target = num_n[s.wait_job] * 8;            // 4 epilogue warps x 2 CTAs report every tile
      d.wait_nrb = nrb[s.wait_job];
      d.wait_cnt_off = cnt_off[s.wait_job];
      if (!s.wait_all) NERAF_REQUIRE(nrb[s.wait_job] == d.num_m, "mega_run: job %d row blocks differ from its producer", i);
    } else { d.wait_target = 0; d.wait_nrb = 0; d.wait_cnt_off = 0; }
    d.act = s.epi.act; d.bias = s.epi.bias;
    d.gate = (const __nv_bfloat16*)s.epi.gate; d.ldg = s.epi.ldg;
    d.has_out_bf16 = s.epi.out_bf16 != nullptr;
    if (s.epi.out_bf16) NERAF_TRY(get_tensor_map_2d(s.epi.out_bf16, 2, s.M, s.N, s.epi.ld_bf16, 32, 64, &d.tmOutB));
    d.out_f32 = s.epi.out_f32; d.ld_f32 = s.epi.ld_f32;
    d.f32_tma = s.epi.out_f32 && (s.epi.ld_f32 % 4 == 0) && ((uintptr_t)s.epi.out_f32 % 16 == 0);
    if (d.f32_tma) NERAF_TRY(get_tensor_map_2d(s.epi.out_f32, 4, s.M, s.N, s.epi.ld_f32, 32, 32, &d.tmOutF));
    d.colsum = s.colsum;
  }
  P.num_tiles = tile;
  NERAF_REQUIRE(counters && counters_bytes >= (size_t)cnt * sizeof(unsigned int), "mega_run: counter buffer too small");
  P.counters = reinterpret_cast<unsigned int*>(counters);
  NERAF_CHECK_CUDA(cudaMemsetAsync(counters, 0, (size_t)cnt * sizeof(unsigned int), stream));

  static bool configured[64] = {false};
  int dev = 0;
  NERAF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    NERAF_CHECK_CUDA(cudaFuncSetAttribute(umma_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM_BYTES));
    configured[dev] = true;
  }
  const int units = sm_count() / 2;
  const int grid = (tile < units ? tile : units) * 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = MEGA_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, umma_mega_kernel, P));
  NERAF_CHECK_LAUNCH("umma_mega_kernel");
  return NERAF_OK;
}

}  // namespace neraf

extern "C" int neraf_gemm_bf16_jobs(const neraf_gemm_job* jobs, int n_jobs, void* counters, size_t counters_bytes,
                                    neraf_stream_t stream) {
  return neraf::mega_run(jobs, n_jobs, counters, counters_bytes, (cudaStream_t)stream);
}
