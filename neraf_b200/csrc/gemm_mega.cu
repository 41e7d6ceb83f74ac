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
// CTA stages its 128 rows of A and bn/2 rows of B per 64-deep k-block through a 5-stage TMA/mbarrier ring, the
// leader issues the MMAs, two 256-column TMEM accumulator stages overlap the epilogue with the next tile.
// 10 warps per CTA: TMA producer, MMA issuer, 8 epilogue warps (two per TMEM lane quarter, each takes half of the
// tile's 32-column chunks: at the reference batch a unit gets about one tile per layer, so the epilogue sits on
// the critical path of the layer chain and its latency is what matters).
//
// Operand layouts.  K-major operands (row-major with the contraction innermost) are the forward / dgrad case.
// The weight gradient dW = dZ^T X contracts over the BATCH, which is the OUTER dimension of the row-major
// activations: both operands are consumed MN-major (tcgen05 descriptor major bit, TMA boxes of 64 contiguous
// MN elements x 64 batch rows), so no transposed copy of any activation or gradient is ever written.
//
// Epilogue.  Per 32 x 32 chunk: tcgen05.ld (lane = row) with the NEXT chunk's bias / gate loads issued behind it
// -> bias (one coalesced load per chunk, shuffle broadcast) / LeakyReLU / 10 tanh / LeakyReLU' gate, activation
// switch hoisted out of the element loop -> 2 KB staging halves (64-byte swizzle) -> TMA stores.  The per-tile
// timeline (tools/mega_trace.py) shows what bounds it: TMEM -> registers runs at ~64 B/clk per SM (560 cycles per
// round of 8 warp-chunks), so everything else per round must stay below that: bias comes in as 8 uniform 16-byte
// loads (32 shuffles cost 600 cycles per round), the LeakyReLU' gate as ONE coalesced word of a transposed sign
// bit mask written by the forward epilogue (reading the bf16 activations row-wise costs 1000+ LSU cycles per
// round, as do row-wise register stores of the outputs: 32 distinct lines per instruction).  Bias gradients are column sums taken from the fp32 values with a
// recursive-halving shuffle reduction (31 shuffles per 32 x 32 block) and one atomic per column.
// Cross-SM visibility: TMA stores complete (bulk wait_group) -> __threadfence -> red.release(counter); consumer:
// ld.acquire spin -> fence.proxy.async -> TMA loads.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "mega_plan.h"
#include "umma_common.cuh"

namespace neraf {
namespace umma {

constexpr int MEGA_EPI_WARPS = 8;                              // two per TMEM lane quarter (column halves)
constexpr int MEGA_THREADS = 64 + 32 * MEGA_EPI_WARPS;         // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int MEGA_STAGES = 6;                                 // ring depth (MegaParams.stages may ask for fewer)
constexpr int MEGA_A_BYTES = BLOCK_M * BLOCK_K * 2;            // 16 KB
constexpr int MEGA_B_BYTES = 128 * BLOCK_K * 2;                // up to bn/2 = 128 rows: 16 KB
constexpr int MEGA_OUT_BYTES = 32 * 128;                       // per-warp 32x32 fp32 transpose tile (unaligned fp32 outputs only)
constexpr int MEGA_BIAS_BYTES = 256;                            // per-warp bias slice of the current chunk pair
constexpr int mega_smem_bytes(int stages) {
  return stages * (MEGA_A_BYTES + MEGA_B_BYTES) + MEGA_EPI_WARPS * (MEGA_OUT_BYTES + MEGA_BIAS_BYTES) + (2 * MEGA_STAGES + 4) * 8 + 16;
}
constexpr int MEGA_SMEM_BYTES = mega_smem_bytes(MEGA_STAGES);   // 226 KB: with the 1 KB the system reserves per CTA the SM is full
constexpr int MEGA_TMEM_COLS = 512;
constexpr int MEGA_ACC_COLS = 256;
constexpr int MN_BOX_BYTES = 64 * 128;                         // one MN-major box: 64 k-rows x 64 MN elements
constexpr unsigned FULL_MASK = 0xffffffffu;

struct alignas(64) DeviceJob {
  CUtensorMap tmA, tmB, tmOut, tmOut1;   // tmOut1: single-chunk (32-column) bf16 stores
  int M, N, K, bn;
  int tile_start, num_m, num_n, cnt_off;
  int merged, merge_na, merge_total;   // merged == 2: second job of an interleaved pair (first = previous job, na tiles)
  int wait_job, wait_all, wait_target, wait_nrb, wait_cnt_off;
  int act, a_mn, b_mn, b_static;
  int out_mode;                    // 0 none, 1 bf16 (TMA), 2 fp32 (TMA), 3 fp32 (rows not 16-byte tileable: smem transpose),
                                   // 4 fp32 reduced into every GPU's copy through the NVSwitch (multimem.red on out_mc)
  float* out_mc;
  int bias_vec;                    // bias may be read with 16-byte loads
  unsigned int* mask_out; const unsigned int* gate_mask; long long ld_mask;
  const float* bias;
  const __nv_bfloat16* gate; long long ldg;
  float* out_f32; long long ld_f32;
  float* colsum;
  unsigned int* notify;            // optional, one counter per row block: += 1 (gpu-scope release) per epilogue warp and tile,
                                   // never cleared -- a concurrent kernel (the data-parallel gradient exchange) learns from
                                   // them which rows of the job's output are complete
  const float* loss_gt; long long ld_gt; double* loss_sums;   // out_mode 3 only: spectral-loss partial sums (fused)
};

// Optional per-tile timeline (NERAF_MEGA_TRACE=<file>): 8 globaltimer stamps per tile, see tools/mega_trace.py.
enum { TR_DEP = 0, TR_LOADED, TR_MMA_START, TR_MMA_FIRST, TR_MMA_DONE, TR_EPI_START, TR_EPI_STORED, TR_EPI_DONE,
       TR_CK_START, TR_CK_LD, TR_CK_MATH, TR_CK_STORE, TR_CK_LOOP, TR_CK_FENCE, TR_PAD0, TR_PAD1, TR_SLOTS };   // TR_CK_*: clock64 of the tracer warp

struct MegaParams {
  int num_jobs, num_tiles;
  int num_counters;                // counters[0 .. num_counters) are dependency counters, counters[num_counters] a ticket:
                                   // the last CTA to finish clears them all, so the NEXT launch finds them zero
  unsigned int* counters;
  unsigned long long* trace;       // null unless tracing
  int stages;                      // depth of the operand ring (<= MEGA_STAGES): 5 leaves 32 KB of the SM's shared memory free,
                                   // which is what lets a CTA of another kernel (the gradient exchange) sit beside this one
  int pdl_late;                    // programmatic dependent launch: release the next kernel when this CTA's tiles are done
                                   // (1) or as soon as the grid is resident (0)
  const uint32_t* sched;           // explicit tile plan (mega_plan.h) or null = static stride over the tiles in job order:
                                   // int32 offsets[sched_units + 1], then per unit its tiles (job << 20 | tile of the job)
  int sched_units;
  DeviceJob jobs[NERAF_MEGA_MAX_JOBS];
};

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stamp(unsigned long long* trace, int tile, int slot) {
  if (trace == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  trace[(size_t)tile * TR_SLOTS + slot] = t;
}
__device__ __forceinline__ void stamp_clock(unsigned long long* trace, int tile, int slot) {
  if (trace == nullptr) return;
  trace[(size_t)tile * TR_SLOTS + slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target) {
  if (ld_acquire(p) >= target) return;
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (ld_acquire(p) < target) {
    __nanosleep(32);
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
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
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
      v[i] = keep + __shfl_xor_sync(FULL_MASK, send, s);
    }
  }
  return v[0];
}

// tcgen05.ld of one 32 x 32 fp32 block WITHOUT the wait (the caller overlaps independent loads with it).
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Tile index -> (job, tile of that job).  Jobs own contiguous index ranges, except an interleaved pair (A, B): both
// start at the same index and their tiles alternate in proportion na : nb over the merged range, so that e.g. the
// weight-gradient tiles whose epilogues push gradients over NVLink are mixed with tiles that do not.
__device__ __forceinline__ const DeviceJob& locate_tile(const MegaParams& P, int tile, int& j, int& local) {
  while (j + 1 < P.num_jobs && tile >= P.jobs[j + 1].tile_start) ++j;
  const DeviceJob& J = P.jobs[j];
  if (J.merged != 2) {
    local = tile - J.tile_start;
    return J;
  }
  const long long m = tile - J.tile_start;
  const int a_before = (int)((m * J.merge_na) / J.merge_total);
  const int a_after = (int)(((m + 1) * J.merge_na) / J.merge_total);
  if (a_after > a_before) {
    local = a_before;
    return P.jobs[j - 1];
  }
  local = (int)m - a_before;
  return J;
}

// The tiles of one unit, in execution order: the planner's list (the next code is fetched one tile ahead) or the static
// stride.  Every role of the CTA pair walks it on its own and sees the same sequence; `tile` is the tile's unique
// index (slot of the plan / position in job order), used by the timeline only.
struct TileWalk {
  int pos, end, stride, j;
  const uint32_t* codes;
  uint32_t nxt;
  __device__ __forceinline__ TileWalk(const MegaParams& P, int unit, int num_units) : j(0), nxt(0) {
    if (P.sched != nullptr) {
      const int32_t* off = reinterpret_cast<const int32_t*>(P.sched);
      pos = __ldg(off + unit); end = __ldg(off + unit + 1); stride = 1;
      codes = P.sched + (P.sched_units + 1);
      if (pos < end) nxt = __ldg(codes + pos);
    } else {
      pos = unit; end = P.num_tiles; stride = num_units; codes = nullptr;
    }
  }
  __device__ __forceinline__ bool next(const MegaParams& P, int& tile, const DeviceJob*& J, int& local) {
    if (pos >= end) return false;
    tile = pos;
    pos += stride;
    if (codes != nullptr) {
      const uint32_t c = nxt;
      if (pos < end) nxt = __ldg(codes + pos);
      J = &P.jobs[c >> plan::kLocalBits];
      local = (int)(c & plan::kLocalMask);
    } else {
      J = &locate_tile(P, tile, j, local);
    }
    return true;
  }
};

__device__ __forceinline__ float gate_factor(float x) { return x > 0.f ? 1.f : kLeakySlope; }

// Epilogue features a job list may need.  The kernel is compiled in a few feature sets (kF) and a launch takes the
// smallest one that covers its jobs: with every output form in one body the full kernel spills per-tile state (ptxas: 128
// bytes of stack at 168 registers, 80 bytes at the 152 registers of the beside-variant); the data-parallel backward's set
// (mask gate, column sums, bf16 outputs only) needs 142 registers and no stack at all.
enum : int { F_BIAS = 1, F_TANH = 2, F_MASKGATE = 4, F_GATE = 8, F_MASKOUT = 16, F_OUT_F32 = 32, F_OUT_ROWS = 64, F_OUT_MC = 128,
             F_LOSS = 256, F_COLSUM = 512, F_TRACE = 1024, F_ALL = 2047 };
constexpr int CFG_FWD = F_BIAS | F_TANH | F_MASKOUT | F_OUT_ROWS;
constexpr int CFG_BWD = F_MASKGATE | F_COLSUM | F_OUT_F32 | F_OUT_ROWS;
constexpr int CFG_LEAN = F_MASKGATE | F_COLSUM;

template <int kF>
__device__ __forceinline__ void mega_body(const MegaParams& P) {
  constexpr int CG = 2;
  // the kernel has no static shared memory, so the dynamic window starts 1024-byte aligned (checked below: the
  // 128-byte swizzle of TMA / UMMA needs it, and there is no room for an alignment slack)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int STAGES = P.stages;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * MEGA_A_BYTES;
  uint8_t* out_buf = smem + STAGES * (MEGA_A_BYTES + MEGA_B_BYTES);            // 1024-byte aligned
  float* bias_buf = reinterpret_cast<float*>(out_buf + MEGA_EPI_WARPS * MEGA_OUT_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(out_buf + MEGA_EPI_WARPS * (MEGA_OUT_BYTES + MEGA_BIAS_BYTES));
  uint64_t* empty_bar = full_bar + MEGA_STAGES;
  uint64_t* tmem_full = empty_bar + MEGA_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t cta_rank = cluster_ctarank();
  const bool is_leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, MEGA_EPI_WARPS * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc<CG>(tmem_slot, MEGA_TMEM_COLS); tmem_relinquish<CG>(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above overlapped the tail of the previous kernel of the stream; nothing
  // it produced has been touched yet.  This grid's CTAs are all resident by now, so the next kernel may start moving in.
  pdl_wait();
  if (!P.pdl_late) pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // Lane 0 issues every TMA; the other lanes only help to poll dependency counters (a "wait for every row block"
    // dependency checks up to 32 counters per round trip instead of one after the other: 8 sequential L2 round trips
    // cost ~3 us in front of every weight-gradient tile) -- and a job seen complete once is never polled again.
    int stage = 0; uint32_t phase = 0;
    uint32_t done_jobs = 0;                                      // bit i: every row block of job i is known complete
    TileWalk walk(P, unit, num_units);
    int tile, local;
    const DeviceJob* Jp;
    while (walk.next(P, tile, Jp, local)) {
      const DeviceJob& J = *Jp;
      const int mt = local / J.num_n, nt = local % J.num_n;          // row-block major: a row block completes early
      const int num_kb = (J.K + BLOCK_K - 1) / BLOCK_K;
      const int b_rows = J.bn / CG;
      const uint32_t stage_bytes = (uint32_t)(MEGA_A_BYTES + b_rows * BLOCK_K * 2) * CG;
      const int a_row0 = mt * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M;
      const int b_row0 = nt * J.bn + (int)cta_rank * b_rows;
      auto load_a = [&](int st, int kb) {
        uint8_t* sa = smem_a + st * MEGA_A_BYTES;
        if (!J.a_mn) {
          tma_load_2d<CG>(sa, &J.tmA, full_bar + st, kb * BLOCK_K, a_row0);
        } else {                                        // matrix is (K rows, MN cols): coordinates {mn, k}
          tma_load_2d<CG>(sa, &J.tmA, full_bar + st, a_row0, kb * BLOCK_K);
          tma_load_2d<CG>(sa + MN_BOX_BYTES, &J.tmA, full_bar + st, a_row0 + 64, kb * BLOCK_K);
        }
      };
      auto load_b = [&](int st, int kb) {
        uint8_t* sb = smem_b + st * MEGA_B_BYTES;
        if (!J.b_mn) {
          tma_load_2d<CG>(sb, &J.tmB, full_bar + st, kb * BLOCK_K, b_row0);
        } else {
          tma_load_2d<CG>(sb, &J.tmB, full_bar + st, b_row0, kb * BLOCK_K);
          if (b_rows > 64) tma_load_2d<CG>(sb + MN_BOX_BYTES, &J.tmB, full_bar + st, b_row0 + 64, kb * BLOCK_K);
        }
      };
      int kb0 = 0;
      const bool must_wait = J.wait_job >= 0 && !((done_jobs >> J.wait_job) & 1u);
      if (must_wait) {
        // Operands written by earlier tiles of this launch (other SMs).  The whole warp polls (wait_all: one counter
        // per lane).  While the dependency is unresolved, ring slots that become free are filled with the B operand
        // when it is flagged static (weights, data of an earlier launch): only A has to wait.
        auto ready = [&]() -> bool {
          bool ok = true;
          if (J.wait_all) {
            for (int rb0 = 0; rb0 < J.wait_nrb; rb0 += 32) {
              const int rb = rb0 + lane;
              if (rb < J.wait_nrb) ok = ok && ld_acquire(P.counters + J.wait_cnt_off + rb) >= (unsigned)J.wait_target;
            }
          } else if (lane == 0) {
            ok = ld_acquire(P.counters + J.wait_cnt_off + mt) >= (unsigned)J.wait_target;
          }
          return __all_sync(FULL_MASK, ok);
        };
        const int pre_max = J.b_static ? (num_kb < STAGES ? num_kb : STAGES) : 0;
        int pre = 0;
        int st = stage; uint32_t ph = phase;
        const long long t0 = clock64();
        unsigned spins = 0;
        while (!ready()) {
          bool issued = false;
          if (pre < pre_max) {
            if (lane == 0 && mbar_try_wait(empty_bar + st, ph ^ 1)) {
              if (is_leader) mbar_expect_tx(full_bar + st, stage_bytes);
              load_b(st, pre);
              issued = true;
            }
            issued = __shfl_sync(FULL_MASK, issued, 0);
            if (issued) {
              ++pre;
              if (++st == STAGES) { st = 0; ph ^= 1; }
            }
          }
          if (!issued) __nanosleep(32);
          if ((++spins & 255u) == 0 && clock64() - t0 > WATCHDOG_CYCLES) __trap();
        }
        if (J.wait_all) done_jobs |= 1u << J.wait_job;
        if (lane == 0) {
          fence_proxy_async_all();
          if (is_leader) if (kF & F_TRACE) stamp(P.trace, tile, TR_DEP);
          for (int kb = 0; kb < pre; ++kb) {
            load_a(stage, kb);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else {
          stage = st; phase = ph;
        }
        kb0 = pre;
      }
      if (lane == 0) {
        for (int kb = kb0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          if (is_leader) mbar_expect_tx(full_bar + stage, stage_bytes);
          load_a(stage, kb);
          load_b(stage, kb);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (is_leader) if (kF & F_TRACE) stamp(P.trace, tile, TR_LOADED);
      } else {
        for (int kb = kb0; kb < num_kb; ++kb)
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && is_leader) {
      int stage = 0; uint32_t phase = 0;
      TileWalk walk(P, unit, num_units);
      int tile, local;
      const DeviceJob* Jp;
      for (int it = 0; walk.next(P, tile, Jp, local); ++it) {
        const DeviceJob& J = *Jp;
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
        if (kF & F_TRACE) stamp(P.trace, tile, TR_MMA_START);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + stage, phase);
          tc_fence_after();
          if (kb == 0) if (kF & F_TRACE) stamp(P.trace, tile, TR_MMA_FIRST);
          const uint32_t sa = smem_u32(smem_a + stage * MEGA_A_BYTES);
          const uint32_t sb = smem_u32(smem_b + stage * MEGA_B_BYTES);
          const uint64_t adesc = a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
          const uint64_t bdesc = b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16<CG>(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, (kb | k) != 0);
          umma_commit<CG>(empty_bar + stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(tmem_full + as);
        if (kF & F_TRACE) stamp(P.trace, tile, TR_MMA_DONE);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps, warp (quarter, half)
    const int quarter = warp & 3;                                 // TMEM lanes this warp may read
    const int half = (warp - 2) >> 2;                             // which half of the tile's 32-column chunks
    uint8_t* wbuf = out_buf + (warp - 2) * MEGA_OUT_BYTES;
    float* wbuf_f = reinterpret_cast<float*>(wbuf);
    float* bias_s = bias_buf + (warp - 2) * (MEGA_BIAS_BYTES / 4);
    bool used_multicast = false;
    TileWalk walk(P, unit, num_units);
    int tile, local;
    const DeviceJob* Jp;
    for (int it = 0; walk.next(P, tile, Jp, local); ++it) {
      const DeviceJob& J = *Jp;
      const int mt = local / J.num_n, nt = local % J.num_n;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int M = J.M, N = J.N;
      const float* bias = J.bias;
      const int act = J.act;
      const __nv_bfloat16* gate = J.gate;
      const long long ldg = J.ldg;
      const int out_mode = J.out_mode;
      float* colsum = J.colsum;
      const int m_base = mt * (BLOCK_M * CG) + (int)cta_rank * BLOCK_M + quarter * 32;
      const int m = m_base + lane;
      const bool row_ok = m < M;
      const int per = J.bn / 64;                                  // chunks per warp: 1, 2 or 4
      const int c_first = half * per;
      const int n_first = nt * J.bn + c_first * 32;
      int nvalid = (N - n_first + 31) / 32;                       // my chunks that start inside N
      nvalid = nvalid < 0 ? 0 : (nvalid > per ? per : nvalid);
      const unsigned int* gate_mask = J.gate_mask;
      unsigned int* mask_out = J.mask_out;
      const long long ld_mask = J.ld_mask;
      const bool gate_vec_ok = gate != nullptr && (ldg % 8 == 0) && ((reinterpret_cast<uintptr_t>(gate) & 15) == 0);

      // Side operands, one coalesced load per chunk (at most 4 chunks per warp -> 8 registers):
      //  * bias (lane q holds column q): never produced inside the launch, so it is fetched BEFORE waiting for the
      //    accumulator; it reaches the other lanes through a 256-byte per-warp shared-memory slice read back as
      //    broadcast 16-byte loads (32 shuffles per chunk cost ~500 cycles per round of 8 warps: tools/micro/epi_bench);
      //  * the LeakyReLU' gate word of the lane's row in the transposed bit mask: may have been written by an earlier
      //    job of this launch, so it is read (L2-coherent) only once the accumulator is complete -- the MMAs consumed
      //    operands the TMA producer loaded after it had acquired the dependency counter.
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      unsigned int w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      auto pre_bias = [&](int jj, float& bj) {
        const int n0 = n_first + 32 * jj;
        if ((kF & F_BIAS) && jj < nvalid && bias != nullptr && n0 + lane < N) bj = __ldg(bias + n0 + lane);
      };
      auto pre_gate = [&](int jj, unsigned int& wj) {
        const int n0 = n_first + 32 * jj;
        if ((kF & F_MASKGATE) && jj < nvalid && gate_mask != nullptr && row_ok) wj = __ldcg(gate_mask + (long long)(n0 >> 5) * ld_mask + m);
      };
      pre_bias(0, b0); pre_bias(1, b1); pre_bias(2, b2); pre_bias(3, b3);

      mbar_wait(tmem_full + as, aphase);
      tc_fence_after();
      pre_gate(0, w0); pre_gate(1, w1); pre_gate(2, w2); pre_gate(3, w3);
      const bool tracer = is_leader && warp == 2 && lane == 0;
      if (tracer) { if (kF & F_TRACE) stamp(P.trace, tile, TR_EPI_START); if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_START); }
      if (nvalid == 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(tmem_empty + as);
          else mbar_arrive_remote(tmem_empty + as, 0);
        }
      }
      const uint32_t t_addr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * MEGA_ACC_COLS + c_first * 32);
      // Two chunks (64 columns) per iteration: both TMEM loads in flight before one wait, one proxy fence and one
      // TMA store (128-byte rows) per pair -- the per-chunk fixed costs measured by tools/micro/epi_bench (LDTM 216,
      // fence 170, TMA store 156 cycles per round of 8 warps) are what the 32-column version was made of.
#pragma unroll 1
      for (int k = 0; k < nvalid; k += 2) {
        const int nch = k + 1 < nvalid ? 2 : 1;
        const int n0 = nt * J.bn + (c_first + k) * 32;
        float v[64];
        {
          float (&va)[32] = *reinterpret_cast<float (*)[32]>(&v[0]);
          float (&vb)[32] = *reinterpret_cast<float (*)[32]>(&v[32]);
          tmem_ld_issue(t_addr0 + (uint32_t)(k * 32), va);
          if (nch == 2) tmem_ld_issue(t_addr0 + (uint32_t)(k * 32 + 32), vb);
        }
        // this pair's bias slice -> shared memory (while the TMEM loads are in flight)
        if ((kF & F_BIAS) && bias != nullptr) {
          bias_s[lane] = k == 0 ? b0 : b2;
          bias_s[32 + lane] = k == 0 ? b1 : b3;
          __syncwarp();
        }
        const unsigned int gm0 = k == 0 ? w0 : w2, gm1 = k == 0 ? w1 : w3;
        tmem_ld_wait();
        if (tracer && k == 0) if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_LD);
        if (k + 2 >= nvalid) {                                    // accumulator fully read: hand the TMEM stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(tmem_empty + as);
            else mbar_arrive_remote(tmem_empty + as, 0);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < nch) {
            float* vh = v + 32 * h;
            const int n0h = n0 + 32 * h;
            const bool full_chunk = n0h + 32 <= N;
            if ((kF & F_BIAS) && bias != nullptr) {
              const float4* bs = reinterpret_cast<const float4*>(bias_s + 32 * h);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = bs[q];
                vh[4 * q] += t.x; vh[4 * q + 1] += t.y; vh[4 * q + 2] += t.z; vh[4 * q + 3] += t.w;
              }
            }
            if (act == NERAF_ACT_LEAKY) {
              // LeakyReLU on the FMA pipe only: 0.55 v + 0.45 |v|  (= v for v > 0, 0.1 v for v < 0, a few ulp of fp32
              // off, invisible after the bf16 rounding).  max(v, 0.1 v) needs FMNMX, compare + predicated multiply
              // needs FSETP: both run on the half-rate ALU pipe, which is what bounds this epilogue.
#pragma unroll
              for (int q = 0; q < 32; ++q) vh[q] = fmaf(0.55f, vh[q], 0.45f * fabsf(vh[q]));
            } else if ((kF & F_TANH) && act == NERAF_ACT_TANH10) {
#pragma unroll
              for (int q = 0; q < 32; ++q) vh[q] = 10.f * tanhf(vh[q]);
            }
            if ((kF & F_MASKGATE) && gate_mask != nullptr) {
              const unsigned int gm = h == 0 ? gm0 : gm1;
#pragma unroll
              for (int q = 0; q < 32; ++q) vh[q] = (gm >> ((q & 1) * 16 + (q >> 1))) & 1u ? kLeakySlope * vh[q] : vh[q];
            } else if ((kF & F_GATE) && gate != nullptr && row_ok) {   // generic bf16 gate (not used by the field)
              const __nv_bfloat16* g = gate + (long long)m * ldg + n0h;
              if (gate_vec_ok && full_chunk) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(g) + q);
                  const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __bfloat1622float2(hh[e]);
                    vh[q * 8 + e * 2] *= gate_factor(f.x);
                    vh[q * 8 + e * 2 + 1] *= gate_factor(f.y);
                  }
                }
              } else {
#pragma unroll
                for (int q = 0; q < 32; ++q)
                  if (n0h + q < N) vh[q] *= gate_factor(__bfloat162float(__ldcg(g + q)));
              }
            }
          }
        }
        if (tracer && k == 0) if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_MATH);
        if (out_mode == 1) {
          // bf16 row-major through the 4 KB staging tile: a pair is 32 rows of 128 bytes (128-byte swizzle, one TMA
          // store), a single chunk 32 rows of 64 bytes (64-byte swizzle)
          if (lane == 0) bulk_wait_read0();                       // the previous store has left the staging tile
          __syncwarp();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nch) {
              unsigned int neg = 0;                               // sign bits: bit i <- element 2i, bit 16+i <- element 2i+1
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 pk;
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) hh[e] = __floats2bfloat162_rn(v[32 * h + i * 8 + 2 * e], v[32 * h + i * 8 + 2 * e + 1]);
                const int off = nch == 2 ? lane * 128 + (((h * 4 + i) ^ (lane & 7)) << 4)
                                         : lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4);
                *reinterpret_cast<uint4*>(wbuf + off) = pk;
                if ((kF & F_MASKOUT) && mask_out != nullptr) {
                  neg = (neg >> 1) | (pk.x & 0x80008000u);
                  neg = (neg >> 1) | (pk.y & 0x80008000u);
                  neg = (neg >> 1) | (pk.z & 0x80008000u);
                  neg = (neg >> 1) | (pk.w & 0x80008000u);
                }
              }
              // LeakyReLU keeps the sign and so does the bf16 rounding: the stored activations' sign bits ARE the
              // backward gate (bit set <=> x < 0), one coalesced word per lane in the transposed mask
              if ((kF & F_MASKOUT) && mask_out != nullptr && row_ok) mask_out[(long long)((n0 + 32 * h) >> 5) * ld_mask + m] = neg;
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(nch == 2 ? &J.tmOut : &J.tmOut1, wbuf, n0, m_base);
            bulk_commit();
          }
        } else if ((kF & F_OUT_F32) && out_mode == 2) {
          // fp32 row-major: one 32 x 32 tile (128-byte rows, 128-byte swizzle) per chunk
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nch) {
              if (lane == 0) bulk_wait_read0();
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4*>(wbuf + lane * 128 + ((i ^ (lane & 7)) << 4)) =
                    make_float4(v[32 * h + i * 4], v[32 * h + i * 4 + 1], v[32 * h + i * 4 + 2], v[32 * h + i * 4 + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&J.tmOut, wbuf, n0 + 32 * h, m_base);
                bulk_commit();
              }
            }
          }
        } else if ((kF & F_OUT_ROWS) && out_mode == 3) {
          // rows that TMA cannot tile (e.g. (B, 513) fp32 outputs): transpose through the staging tile
          // (XOR-swizzled 32 x 32 floats, conflict-free both ways) so that lanes write consecutive columns
          float* out_f32 = J.out_f32;
          const long long ld_f32 = J.ld_f32;
          const int rows = M - m_base < 32 ? M - m_base : 32;
          const float* loss_gt = J.loss_gt;
          double s_num = 0.0, s_den = 0.0, s_sq = 0.0, s_abs = 0.0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nch) {
#pragma unroll
              for (int q = 0; q < 32; ++q) wbuf_f[lane * 32 + (q ^ lane)] = v[32 * h + q];
              __syncwarp();
              const int n = n0 + 32 * h + lane;
              if (n < N) {
                if (!(kF & F_LOSS) || loss_gt == nullptr) {
#pragma unroll 4
                  for (int r = 0; r < rows; ++r) out_f32[(long long)(m_base + r) * ld_f32 + n] = wbuf_f[r * 32 + (lane ^ r)];
                } else {
                  // the prediction leaves through here exactly once: form the spectral loss's terms against the
                  // target on the way (same arithmetic as loss_sums_kernel, lanes = consecutive columns)
                  const long long ld_gt = J.ld_gt;
#pragma unroll 4
                  for (int r = 0; r < rows; ++r) {
                    const float x = wbuf_f[r * 32 + (lane ^ r)];
                    out_f32[(long long)(m_base + r) * ld_f32 + n] = x;
                    const float y = __ldg(loss_gt + (long long)(m_base + r) * ld_gt + n);
                    const float ex = exp_fma(x), ey = exp_fma(y);
                    const float dm = ey - ex, ym = ey - 1e-3f, d = y - x;
                    s_num += (double)(dm * dm);
                    s_den += (double)(ym * ym);
                    s_sq += (double)(d * d);
                    s_abs += (double)fabsf(d);
                  }
                }
              }
              __syncwarp();
            }
          }
          if ((kF & F_LOSS) && loss_gt != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              s_num += __shfl_xor_sync(FULL_MASK, s_num, o);
              s_den += __shfl_xor_sync(FULL_MASK, s_den, o);
              s_sq += __shfl_xor_sync(FULL_MASK, s_sq, o);
              s_abs += __shfl_xor_sync(FULL_MASK, s_abs, o);
            }
            if (lane == 0) {
              atomicAdd(J.loss_sums + 0, s_num);
              atomicAdd(J.loss_sums + 1, s_den);
              atomicAdd(J.loss_sums + 2, s_sq);
              atomicAdd(J.loss_sums + 3, s_abs);
            }
          }
        }
        else if ((kF & F_OUT_MC) && out_mode == 4) {
          // Fused all-reduce: the tile is ADDED into every rank's copy of the gradient by the NVSwitch
          // (multimem.red on the multicast alias of the output; NVLS).  Same shared-memory transpose as above so
          // that one instruction covers 128 contiguous bytes of a row.
          float* out_mc = J.out_mc;
          const long long ld_f32 = J.ld_f32;
          const int rows = M - m_base < 32 ? M - m_base : 32;
          const bool mc_vec = (ld_f32 % 4 == 0) && ((reinterpret_cast<uintptr_t>(out_mc) & 15) == 0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nch) {
#pragma unroll
              for (int q = 0; q < 32; ++q) wbuf_f[lane * 32 + (q ^ lane)] = v[32 * h + q];
              __syncwarp();
              const int nbase = n0 + 32 * h;
              if (mc_vec && nbase + 32 <= N) {
                // 16 bytes per lane: one instruction covers 4 rows x 128 contiguous bytes
                const int c4 = 4 * (lane & 7);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int r = i * 4 + (lane >> 3);
                  if (r < rows) {
                    const float* src = wbuf_f + r * 32;
                    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                     out_mc + (long long)(m_base + r) * ld_f32 + nbase + c4),
                                 "f"(src[(c4 + 0) ^ r]), "f"(src[(c4 + 1) ^ r]), "f"(src[(c4 + 2) ^ r]), "f"(src[(c4 + 3) ^ r])
                                 : "memory");
                  }
                }
              } else {
                const int n = nbase + lane;
                if (n < N) {
#pragma unroll 4
                  for (int r = 0; r < rows; ++r)
                    asm volatile("multimem.red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(out_mc + (long long)(m_base + r) * ld_f32 + n),
                                 "f"(wbuf_f[r * 32 + (lane ^ r)])
                                 : "memory");
                }
              }
              __syncwarp();
            }
          }
        }
        if (tracer && k == 0) if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_STORE);
        // ---- bias gradient: column sums of the fp32 values (destroys v)
        if ((kF & F_COLSUM) && colsum != nullptr) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < nch) {
              float (&vh)[32] = *reinterpret_cast<float (*)[32]>(&v[32 * h]);
              if (!row_ok) {
#pragma unroll
                for (int q = 0; q < 32; ++q) vh[q] = 0.f;
              }
              const float ssum = column_sums_32(vh, lane);
              if (n0 + 32 * h + lane < N) atomicAdd(colsum + n0 + 32 * h + lane, ssum);
            }
          }
        }
      }
      if (tracer) { if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_LOOP); if (kF & F_TRACE) stamp(P.trace, tile, TR_EPI_STORED); }
      // Publish the tile: lane 0 waits until its TMA stores have been performed, the warp's plain stores (bit mask,
      // unaligned fp32 rows, column-sum atomics) are ordered before lane 0 by the warp barrier, and the counter
      // update itself is a gpu-scope release (one fence instead of a full membar in every lane).
      if (lane == 0 && (out_mode == 1 || out_mode == 2)) { bulk_wait_all0(); fence_proxy_async_all(); }
      if (out_mode == 4) used_multicast = true;      // fenced once, at the end of the kernel
      __syncwarp();
      if (tracer) { if (kF & F_TRACE) stamp_clock(P.trace, tile, TR_CK_FENCE); if (kF & F_TRACE) stamp(P.trace, tile, TR_EPI_DONE); }
      if (lane == 0) {
        red_release_add(P.counters + J.cnt_off + mt, 1u);   // ... before the row block's progress is (16 arrivals per tile)
        if (J.notify != nullptr) red_release_add(J.notify + mt, 1u);
      }
    }
    // multimem.red reductions are fire-and-forget while the kernel runs (the NVLink queue drains behind the remaining
    // tiles); one system-scope fence per warp makes them performed before the kernel can end
    if (used_multicast) __threadfence_system();
  }

  if (P.pdl_late) pdl_launch_dependents();            // this CTA has no tile left: the next kernel's CTAs may be set up
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, MEGA_TMEM_COLS);
  }
  // Leave the counters as they were found: a CTA takes its ticket after its last poll, so whoever draws the last ticket
  // knows nobody reads the counters any more and clears them (and the ticket) for the next launch on this buffer.
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(P.counters + P.num_counters, 1u);
    if (t == gridDim.x - 1) {
      for (int i = 0; i <= P.num_counters; ++i) P.counters[i] = 0u;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(MEGA_THREADS, 1) umma_mega_kernel(const __grid_constant__ MegaParams P) { mega_body<F_ALL>(P); }
__global__ void __launch_bounds__(MEGA_THREADS, 1) umma_mega_kernel_fwd(const __grid_constant__ MegaParams P) { mega_body<CFG_FWD>(P); }
__global__ void __launch_bounds__(MEGA_THREADS, 1) umma_mega_kernel_bwd(const __grid_constant__ MegaParams P) { mega_body<CFG_BWD>(P); }

// The same kernel held to 152 registers per thread, for launches that another kernel must be able to sit BESIDE (the
// data-parallel gradient exchange).  Registers are handed out per SM sub-partition: this CTA's 10 warps land 3/3/2/2 on
// the four partitions, and with 160+ registers per thread the two partitions that hold three warps have no room left
// for a single warp of any other CTA -- measured (tools/micro/pdl_beside.cu): beside 320 threads x 159 registers NOTHING
// becomes resident, beside 320 x 151 a 128-thread x 48-register CTA does.
__global__ void __maxnreg__(152) umma_mega_kernel_slim(const __grid_constant__ MegaParams P) { mega_body<F_ALL>(P); }
__global__ void __maxnreg__(152) umma_mega_kernel_slim_lean(const __grid_constant__ MegaParams P) { mega_body<CFG_LEAN>(P); }

}  // namespace umma

int get_tensor_map_2d(const void* ptr, int elem_bytes, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols,
                      CUtensorMap* out);

// ---- tile plans: computed once per (job shapes, units, policy) and kept in device memory for the life of the process
namespace {
struct PlanEntry {
  uint32_t* dev = nullptr;            // null: the static stride is at least as good
  std::vector<uint32_t> codes;        // host copy (timeline)
  double static_us = 0.0, plan_us = 0.0;
  int policy = 0;
};
std::map<std::vector<int>, PlanEntry>& plan_cache() {
  static std::map<std::vector<int>, PlanEntry> cache;
  return cache;
}
}  // namespace

// NERAF_MEGA_PLAN = auto (default) | static | cp | rb : cp / rb force a policy even when the model predicts no gain
// (tests; A/B timing).  The plan table is uploaded with a synchronous copy the first time a job list is seen; a first
// sighting inside a stream capture (no warm-up call) falls back to the static stride for that graph.
static int attach_plan(umma::MegaParams& P, const MegaJob* jobs, int n_jobs, int units, int dev, cudaStream_t stream,
                       bool beside, std::vector<uint32_t>* trace_codes) {
  using namespace umma;
  P.sched = nullptr; P.sched_units = 0;
  const char* env = getenv("NERAF_MEGA_PLAN");
  int mode = -1;                                       // -1 auto
  if (env && env[0] == 's') mode = plan::STATIC_STRIDE;
  else if (env && env[0] == 'c') mode = plan::CRITICAL_PATH;
  else if (env && env[0] == 'r') mode = plan::ROW_BLOCK;
  bool merged = false;
  for (int i = 0; i < n_jobs; ++i) merged = merged || jobs[i].merge_next;
  if (trace_codes) {                                   // static order: slot = position in job order
    trace_codes->clear();
    if (!merged)
      for (int i = 0; i < n_jobs; ++i)
        for (int t = 0; t < P.jobs[i].num_m * P.jobs[i].num_n; ++t)
          trace_codes->push_back(((uint32_t)i << plan::kLocalBits) | (uint32_t)t);
  }
  if (mode == plan::STATIC_STRIDE || merged || units < 2) return NERAF_OK;
  if (beside && mode < 0 && !getenv("NERAF_MEGA_PLAN_BESIDE")) return NERAF_OK;
  std::vector<plan::PlanJob> pj(n_jobs);
  std::vector<int> key;
  key.reserve(n_jobs * 9 + 3);
  key.push_back(dev); key.push_back(units); key.push_back(mode);
  for (int i = 0; i < n_jobs; ++i) {
    const DeviceJob& d = P.jobs[i];
    plan::PlanJob& j = pj[i];
    j.num_m = d.num_m; j.num_n = d.num_n; j.kb = (d.K + BLOCK_K - 1) / BLOCK_K; j.bn = d.bn;
    j.a_mn = d.a_mn; j.b_mn = d.b_mn; j.wait_job = d.wait_job; j.wait_all = d.wait_all;
    if (d.out_mode == 3) j.kind = d.act == NERAF_ACT_TANH10 ? plan::EPI_ROWS_TANH : plan::EPI_ROWS;
    else if (d.gate_mask || d.gate) j.kind = plan::EPI_DGRAD;
    else if (d.a_mn) j.kind = plan::EPI_WGRAD;
    else j.kind = plan::EPI_ACT_BF16;
    NERAF_REQUIRE((unsigned)(d.num_m * d.num_n) <= plan::kLocalMask, "mega_run: job %d has too many tiles for a plan", i);
    const int rec[9] = {j.num_m, j.num_n, j.kb, j.bn, j.a_mn, j.b_mn, j.wait_job, j.wait_all, j.kind};
    key.insert(key.end(), rec, rec + 9);
  }
  auto& cache = plan_cache();
  auto it = cache.find(key);
  if (it == cache.end()) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    NERAF_CHECK_CUDA(cudaStreamIsCapturing(stream, &cap));
    if (cap != cudaStreamCaptureStatusNone) return NERAF_OK;      // no allocation / synchronous copy inside a capture
    PlanEntry e;
    e.static_us = plan::make_plan(pj.data(), n_jobs, units, plan::STATIC_STRIDE).makespan;
    plan::Plan best;
    if (mode >= 0) {
      best = plan::make_plan(pj.data(), n_jobs, units, (plan::Policy)mode);
      e.policy = mode;
    } else {
      best = plan::make_plan(pj.data(), n_jobs, units, plan::CRITICAL_PATH);
      e.policy = plan::CRITICAL_PATH;
      plan::Plan rb = plan::make_plan(pj.data(), n_jobs, units, plan::ROW_BLOCK);
      if (rb.makespan < best.makespan) { best = std::move(rb); e.policy = plan::ROW_BLOCK; }
    }
    e.plan_us = best.makespan;
    // auto: the model is good to ~10 %; a plan must promise clearly more than that noise
    const char* thr = getenv("NERAF_MEGA_PLAN_GAIN");
    const double need = thr ? atof(thr) : 0.05;
    if (mode >= 0 || best.makespan < e.static_us * (1.0 - need)) {
      std::vector<uint32_t> table((size_t)units + 1 + best.codes.size());
      for (int x = 0; x <= units; ++x) table[x] = (uint32_t)best.unit_off[x];
      std::copy(best.codes.begin(), best.codes.end(), table.begin() + units + 1);
      NERAF_CHECK_CUDA(cudaMalloc(&e.dev, table.size() * sizeof(uint32_t)));
      NERAF_CHECK_CUDA(cudaMemcpy(e.dev, table.data(), table.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
      e.codes = std::move(best.codes);
    }
    if (getenv("NERAF_MEGA_PLAN_VERBOSE"))
      fprintf(stderr, "neraf mega plan: %d jobs on %d units: static %.1f us, %s %.1f us -> %s\n", n_jobs, units, e.static_us,
              e.policy == plan::ROW_BLOCK ? "row-block" : "critical-path", e.plan_us, e.dev ? "planned" : "static stride");
    it = cache.emplace(std::move(key), std::move(e)).first;
  }
  const PlanEntry& e = it->second;
  if (e.dev) {
    P.sched = e.dev; P.sched_units = units;
    if (trace_codes) *trace_codes = e.codes;
  }
  return NERAF_OK;
}

int mega_run(const MegaJob* jobs, int n_jobs, void* counters, size_t counters_bytes, cudaStream_t stream, int max_ctas,
             bool counters_clean, bool pdl, unsigned int* notify_increment, bool release_dependents_early) {
  using namespace umma;
  NERAF_REQUIRE(jobs && n_jobs > 0 && n_jobs <= NERAF_MEGA_MAX_JOBS, "mega_run: 1..%d jobs", NERAF_MEGA_MAX_JOBS);
  static MegaParams P;          // large: build in static storage (single-threaded driver, see header conventions)
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  P.num_jobs = n_jobs;
  {
    const char* e = getenv("NERAF_PDL_TRIGGER");       // "early" / "late" (tuning; default late)
    P.pdl_late = (e && e[0] == 'e') ? 0 : 1;
    if (release_dependents_early) P.pdl_late = 0;      // a kernel that runs BESIDE this one is waiting to move in
    const char* st = getenv("NERAF_MEGA_STAGES");
    P.stages = st ? atoi(st) : (release_dependents_early ? MEGA_STAGES - 1 : MEGA_STAGES);
    if (P.stages < 2 || P.stages > MEGA_STAGES) P.stages = MEGA_STAGES;
  }
  // A shorter ring alone does not make room beside this kernel: the SM's shared-memory carve-out comes in steps (..., 196,
  // 228 KB) and is chosen to fit the resident CTA -- the 5-stage footprint lands 880 bytes under the 196 KB step, and a
  // second CTA needs 1 KB of its own.  4 KB of padding push the request over that step: the SM is configured with 228 KB
  // and ~28 KB stay free for the CTA of the kernel that runs beside.
  const int smem_bytes = mega_smem_bytes(P.stages) + (P.stages < MEGA_STAGES ? 4096 : 0);
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
    d.a_mn = s.a_mn; d.b_mn = s.b_mn; d.b_static = s.b_static;
    d.M = (int)s.M; d.N = (int)s.N; d.K = (int)s.K; d.bn = s.bn;
    d.num_m = (int)ceil_div(s.M, 256); d.num_n = (int)ceil_div(s.N, s.bn);
    d.merged = 0; d.merge_na = 0; d.merge_total = 0;
    if (i > 0 && jobs[i - 1].merge_next) {             // second job of an interleaved pair: shares the range of the first
      NERAF_REQUIRE(s.wait_job != i - 1 && !s.merge_next, "mega_run: job %d cannot be interleaved with the job it waits for", i);
      DeviceJob& a = P.jobs[i - 1];
      const int na = a.num_m * a.num_n, nb = d.num_m * d.num_n;
      a.merged = 1;
      d.merged = 2; d.merge_na = na; d.merge_total = na + nb;
      d.tile_start = a.tile_start; tile = a.tile_start + na + nb;
    } else {
      d.tile_start = tile; tile += d.num_m * d.num_n;
    }
    d.cnt_off = cnt; cnt_off[i] = cnt; nrb[i] = d.num_m; num_n[i] = d.num_n; cnt += d.num_m;
    d.wait_job = s.wait_job; d.wait_all = s.wait_all;
    if (s.wait_job >= 0) {
      d.wait_target = num_n[s.wait_job] * MEGA_EPI_WARPS * 2;   // every epilogue warp of both CTAs reports every tile
      d.wait_nrb = nrb[s.wait_job];
      d.wait_cnt_off = cnt_off[s.wait_job];
      if (!s.wait_all) NERAF_REQUIRE(nrb[s.wait_job] == d.num_m, "mega_run: job %d row blocks differ from its producer", i);
    } else { d.wait_target = 0; d.wait_nrb = 0; d.wait_cnt_off = 0; }
    d.act = s.epi.act; d.bias = s.epi.bias;
    d.gate = (const __nv_bfloat16*)s.epi.gate; d.ldg = s.epi.ldg;
    d.out_f32 = s.epi.out_f32; d.ld_f32 = s.epi.ld_f32;
    d.out_mode = 0;
    if (s.epi.out_bf16) {
      NERAF_REQUIRE(s.epi.ld_bf16 % 8 == 0 && ((uintptr_t)s.epi.out_bf16 % 16) == 0 && s.epi.ld_bf16 >= s.N,
                    "mega_run: job %d: out_bf16 needs ld %% 8 == 0, ld >= N and 16-byte alignment", i);
      d.out_mode = 1;
      NERAF_TRY(get_tensor_map_2d(s.epi.out_bf16, 2, s.M, s.N, s.epi.ld_bf16, 32, 64, &d.tmOut));
      NERAF_TRY(get_tensor_map_2d(s.epi.out_bf16, 2, s.M, s.N, s.epi.ld_bf16, 32, 32, &d.tmOut1));
    } else if (s.epi.out_f32) {
      NERAF_REQUIRE(s.epi.ld_f32 >= s.N, "mega_run: job %d: out_f32 row stride < N", i);
      // TMA stores clip with 16-byte granularity: only rows that end on a 16-byte boundary take the TMA path
      const bool tma_ok = (s.epi.ld_f32 % 4 == 0) && ((uintptr_t)s.epi.out_f32 % 16 == 0) && (s.N % 4 == 0);
      d.out_mode = (tma_ok && !s.epi.loss_gt) ? 2 : 3;
      d.out_mc = (float*)s.epi.out_f32_multicast;
      if (d.out_mc) d.out_mode = 4;
      NERAF_REQUIRE(!s.epi.loss_gt || (s.epi.loss_sums && !d.out_mc && s.epi.ld_gt >= s.N),
                    "mega_run: job %d: loss_gt needs loss_sums, ld_gt >= N and no multicast output", i);
      if (tma_ok) NERAF_TRY(get_tensor_map_2d(s.epi.out_f32, 4, s.M, s.N, s.epi.ld_f32, 32, 32, &d.tmOut));
    }
    d.bias_vec = s.epi.bias && ((uintptr_t)s.epi.bias % 16) == 0;
    d.mask_out = (unsigned int*)s.epi.mask_out; d.gate_mask = (const unsigned int*)s.epi.gate_mask; d.ld_mask = s.epi.ld_mask;
    NERAF_REQUIRE(!(s.epi.mask_out || s.epi.gate_mask) || s.epi.ld_mask >= s.M, "mega_run: job %d: ld_mask < M", i);
    NERAF_REQUIRE(!s.epi.mask_out || (s.epi.act == NERAF_ACT_LEAKY && s.epi.out_bf16),
                  "mega_run: job %d: mask_out needs the LeakyReLU epilogue and a bf16 output", i);
    NERAF_REQUIRE(!(s.epi.gate && s.epi.gate_mask), "mega_run: job %d: gate and gate_mask are exclusive", i);
    d.colsum = s.colsum;
    d.notify = s.notify;
    if (notify_increment) notify_increment[i] = (unsigned int)d.num_n * MEGA_EPI_WARPS * 2;   // per row-block counter
    d.loss_gt = s.epi.out_f32 ? s.epi.loss_gt : nullptr; d.ld_gt = s.epi.ld_gt; d.loss_sums = s.epi.loss_sums;
    NERAF_REQUIRE(!s.epi.loss_gt || s.epi.out_f32, "mega_run: job %d: loss_gt needs an fp32 output", i);
  }
  P.num_tiles = tile;
  NERAF_REQUIRE(counters && counters_bytes >= (size_t)(cnt + 1) * sizeof(unsigned int), "mega_run: counter buffer too small");
  P.counters = reinterpret_cast<unsigned int*>(counters);
  P.num_counters = cnt;
  // every launch leaves its counters zero (kernel epilogue); a caller that knows the buffer was cleared before, or last
  // used by this kernel, skips the memset node
  if (!counters_clean) NERAF_CHECK_CUDA(cudaMemsetAsync(counters, 0, (size_t)(cnt + 1) * sizeof(unsigned int), stream));
  // Debug timeline: NERAF_MEGA_TRACE=<file> makes every launch synchronous and appends its per-tile stamps.
  static const char* trace_path = getenv("NERAF_MEGA_TRACE");
  P.trace = nullptr;
  const size_t trace_bytes = (size_t)tile * TR_SLOTS * sizeof(unsigned long long);
  if (trace_path) {
    NERAF_CHECK_CUDA(cudaMalloc(&P.trace, trace_bytes));
    NERAF_CHECK_CUDA(cudaMemsetAsync(P.trace, 0, trace_bytes, stream));
  }

  static bool configured[64] = {false};
  int dev = 0;
  NERAF_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !configured[dev]) {
    for (auto kern : {umma_mega_kernel, umma_mega_kernel_fwd, umma_mega_kernel_bwd, umma_mega_kernel_slim, umma_mega_kernel_slim_lean}) {
      NERAF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, MEGA_SMEM_BYTES));
      NERAF_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    configured[dev] = true;
  }
  // feature set of this job list -> kernel variant
  int need = 0;
  for (int i = 0; i < n_jobs; ++i) {
    const DeviceJob& d = P.jobs[i];
    if (d.bias) need |= F_BIAS;
    if (d.act == NERAF_ACT_TANH10) need |= F_TANH;
    if (d.gate_mask) need |= F_MASKGATE;
    if (d.gate) need |= F_GATE;
    if (d.mask_out) need |= F_MASKOUT;
    if (d.out_mode == 2) need |= F_OUT_F32;
    if (d.out_mode == 3) need |= F_OUT_ROWS;
    if (d.out_mode == 4) need |= F_OUT_MC;
    if (d.loss_gt) need |= F_LOSS;
    if (d.colsum) need |= F_COLSUM;
  }
  if (trace_path) need |= F_TRACE;                   // the per-tile timeline lives in the full kernel only
  static const bool variants = getenv("NERAF_MEGA_VARIANTS") == nullptr || getenv("NERAF_MEGA_VARIANTS")[0] != '0';
  auto covers = [&](int cfg) { return variants && (need & ~cfg) == 0; };
  auto kernel = release_dependents_early ? (covers(CFG_LEAN) ? umma_mega_kernel_slim_lean : umma_mega_kernel_slim)
                                         : (covers(CFG_FWD) ? umma_mega_kernel_fwd
                                                            : (covers(CFG_BWD) ? umma_mega_kernel_bwd : umma_mega_kernel));
  int units = sm_count() / 2;
  if (max_ctas >= 2 && max_ctas / 2 < units) units = max_ctas / 2;   // leave SMs to a concurrent kernel (collectives)
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = (pdl && pdl_enabled() && !trace_path) ? 2 : 1;
  // The CTA pairs spin on counters other pairs write, with a static tile stride: every pair of the grid must be
  // co-resident.  Never launch more pairs than the device can hold at once for this kernel's footprint (asked once
  // per device; an SM held by a concurrent kernel at run time is the caller's business: max_ctas).
  static int resident_pairs[64] = {0};
  if (dev >= 0 && dev < 64 && resident_pairs[dev] == 0) {
    cfg.gridDim = dim3((unsigned)(units * 2));
    int n = 0;
    NERAF_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, umma_mega_kernel, &cfg));
    NERAF_REQUIRE(n > 0, "mega_run: no CTA pair of the job-list kernel fits on this device");
    resident_pairs[dev] = n;
  }
  if (dev >= 0 && dev < 64 && resident_pairs[dev] < units) units = resident_pairs[dev];
  if (tile < units) units = tile;
  // Which unit runs which tile: the planner's explicit lists when its model predicts a gain over the static stride
  // (mega_plan.h), else the stride.  Never beside another kernel by default: the data-parallel exchange wants the weight
  // gradients in the order the caller listed them.
  std::vector<uint32_t> trace_codes;
  NERAF_TRY(attach_plan(P, jobs, n_jobs, units, dev, stream, release_dependents_early, trace_path ? &trace_codes : nullptr));
  const int grid = units * 2;
  cfg.gridDim = dim3((unsigned)grid);
  NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, P));
  NERAF_CHECK_LAUNCH("umma_mega_kernel");
  if (trace_path) {
    std::vector<unsigned long long> host((size_t)tile * TR_SLOTS);
    NERAF_CHECK_CUDA(cudaStreamSynchronize(stream));
    NERAF_CHECK_CUDA(cudaMemcpy(host.data(), P.trace, trace_bytes, cudaMemcpyDeviceToHost));
    NERAF_CHECK_CUDA(cudaFree(P.trace));
    if (FILE* f = fopen(trace_path, "ab")) {
      // record: magic, n_jobs, n_tiles, units, then per job {tile_start, M, N, K, bn, a_mn, b_mn, wait_job}, then per
      // stamp slot the tile it belongs to (job << 20 | tile of the job; 0xffffffff: use tile_start), then the stamps
      const int hdr[4] = {0x4d454732, n_jobs, tile, units};
      fwrite(hdr, sizeof(int), 4, f);
      for (int i = 0; i < n_jobs; ++i) {
        const DeviceJob& d = P.jobs[i];
        const int rec[8] = {d.tile_start, d.M, d.N, d.K, d.bn, d.a_mn, d.b_mn, d.wait_job};
        fwrite(rec, sizeof(int), 8, f);
      }
      trace_codes.resize((size_t)tile, 0xffffffffu);
      fwrite(trace_codes.data(), sizeof(uint32_t), trace_codes.size(), f);
      fwrite(host.data(), sizeof(unsigned long long), host.size(), f);
      fclose(f);
    }
  }
  return NERAF_OK;
}

}  // namespace neraf

extern "C" int neraf_gemm_bf16_jobs(const neraf_gemm_job* jobs, int n_jobs, void* counters, size_t counters_bytes,
                                    neraf_stream_t stream) {
  return neraf::mega_run(jobs, n_jobs, counters, counters_bytes, (cudaStream_t)stream, 0);
}
