// Data-parallel gradient exchange as ONE kernel of this library, running BESIDE the backward's job-list launch.
//
// Why: at the reference batch (2048 columns per GPU) the gradient all-reduce is as long as a third of the step, and
// issued by the host after the backward it is fully exposed (SCALE_r01.json: 0.63 efficiency on 8 GPUs).  The weight
// gradients finish in backward order, so most of them can travel while the remaining GEMMs run -- if the exchange
// (a) needs no host call between the kernels (the step stays one CUDA graph) and (b) does not take SMs away from the
// persistent GEMM kernel.  This kernel does both: small CTAs (128 threads, <= 48 registers, no shared memory: one fits
// beside a job-list CTA on every SM), launched behind the job-list kernel as a programmatic dependent that never waits
// for it (it moves in once that grid is resident: field.cu), driven by the completion counters the job-list kernel
// advances as it stores gradient tiles (neraf_gemm_job.notify).
//
// Algorithm (two-shot all-reduce over peer-mapped "symmetric" memory; NVLink 5 / NVSwitch):
//   per chunk c (one weight-gradient matrix, bf16; last: the bias gradients, fp32), in the order the backward finishes
//   them:
//     herald (block 0, one thread)  waits until THIS rank has stored chunk c (notify counter, gpu-scope acquire), then
//                                   raises ready[c][rank] in every peer's signal buffer (system-scope release);
//     workers (all other blocks)    wait until ready[c][q] is raised for every rank q, then reduce this rank's 1/world
//                                   slice of the chunk: multimem.ld_reduce (the switch adds the ranks' copies, fp32
//                                   accumulation) -> multimem.st of the sum into every rank's copy (NVLS).  Without a
//                                   multicast mapping the same is done with plain peer loads and stores.
//   Pull mode (neraf_grad_exchange.pull; what the graphed data-parallel step uses): the second shot is turned around.
//   A worker stores the sums of its slice into its OWN copy only; the last worker of the rank to finish the slice
//   raises reduced[c][rank] everywhere; then every rank PULLS the other ranks' slices with plain peer loads and writes
//   them where the optimizer wants them: as fp32 into the parameter's .grad (chunk.dst; a compact row-strided block is
//   scattered into its columns of the layer-1 gradient).  Same bytes on the wire, one more flag hop per chunk (hidden
//   beside the backward for all but the last chunk) -- and the separate widening pass over all gradients that followed
//   the push form (28.7 MB read + 57.4 MB written after the exchange: 20 of the 34 us of the kernel behind the
//   backward, profiles/r02l_dp_parts_1rank.txt) is gone.
//   end: workers fence (system scope) and draw a ticket; the herald waits for all tickets, raises done[rank] everywhere
//   and waits for every rank's done flag: when the kernel ends, every rank holds the sums of every chunk.
// Flags carry the step number (kept in `state`, advanced by the herald), so nothing is ever reset and a flag that is
// ahead of a slow reader is still "raised".  Deadlock freedom: the job-list kernel never waits for this kernel; the
// herald only waits for counters of its own device and for peers that run the same graph.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace neraf {

// 128 threads x <= 48 registers: 4 warps x 1536 registers (warp allocations are rounded up to 512) = 6144, which is what
// is left beside a job-list CTA (10 warps x 5632 = 56320 of 65536); 256 threads x 40 registers did NOT fit -- the exchange
// then only moved in as CTA pairs of the GEMM kernel ran out of tiles (profiles/r02j_exchange_timeline.txt)
constexpr int kCommThreads = 128;
constexpr int kCommUnroll = 6;                 // 16-byte load-reduce requests in flight per thread
// Pull form, second shot: peer loads land in SHARED memory (cp.async, 16 bytes each, thread-private slots), which costs
// no registers -- the 48-register budget caps register-held loads at 6 per thread, i.e. 1.8 MB in flight per GPU, and at
// NVLink's ~3 us round trip that is what bounded the exchange (20.9 MB in 53 us, profiles/r02j_time_dp8.txt).  Two
// groups of kPullDepth slots per thread: one group lands while the other is converted and stored.  24 KB per CTA is
// what is left beside a 5-stage CTA of the job-list kernel (gemm_mega.cu: 228 KB carve-out - 199 KB - 4 KB padding).
constexpr int kPullDepth = 6;
constexpr int kPullSmemBytes = 2 * kPullDepth * kCommThreads * 16;
constexpr int kReadyOffset = 2048;             // u32 ready[NERAF_MAX_EXCHANGE_CHUNKS][NERAF_MAX_RANKS]
constexpr int kDoneOffset = 4096;              // u32 done[NERAF_MAX_RANKS]
constexpr int kReducedOffset = 4608;           // u32 reduced[NERAF_MAX_EXCHANGE_CHUNKS][NERAF_MAX_RANKS]  (pull mode)
constexpr int kChunkTickets = 2;               // state[2 + c]: workers of this rank that have reduced their part of chunk c
static_assert(kReadyOffset + NERAF_MAX_EXCHANGE_CHUNKS * NERAF_MAX_RANKS * 4 <= kDoneOffset, "signal buffer layout");
static_assert(kDoneOffset + NERAF_MAX_RANKS * 4 <= kReducedOffset, "signal buffer layout");
static_assert(kReducedOffset + NERAF_MAX_EXCHANGE_CHUNKS * NERAF_MAX_RANKS * 4 <= NERAF_EXCHANGE_BYTES, "signal buffer layout");
static_assert(kChunkTickets + NERAF_MAX_EXCHANGE_CHUNKS <= NERAF_EXCHANGE_STATE_WORDS, "state layout");
constexpr long long kCommSpinLimit = 4000000000LL;      // ~2 s: a lost peer traps instead of hanging the GPU

struct CommChunk {
  unsigned long long offset, bytes;            // of the exchange region; multiples of 16
  const unsigned int* notify; unsigned int count, increment;
  int f32;
  float* dst; long long dst_ld; int row_elems, src_ld;      // pull mode: where the sums go as fp32 (null: stay in the region)
};
struct CommArgs {
  int n_chunks, world, rank, pull;
  CommChunk ch[NERAF_MAX_EXCHANGE_CHUNKS];
  uint8_t* mc;                                 // multicast alias of the region (nullptr: peer loads / stores)
  uint8_t* peers[NERAF_MAX_RANKS];
  uint8_t* sig[NERAF_MAX_RANKS];
  unsigned int* state;                         // [0] steps completed  [1] worker tickets
  unsigned long long* trace;                   // optional: globaltimer stamps (see neraf_grad_exchange.trace)
  unsigned int sleep_ns;                       // back-off between polls of a flag / counter
  int one_fence;
};

__device__ __forceinline__ unsigned int comm_ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int comm_ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void comm_st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void comm_st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void comm_stamp(unsigned long long* trace, int slot) {
  if (trace == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  trace[slot] = t;
}
__device__ __forceinline__ bool reached(unsigned int value, unsigned int target) { return (int)(value - target) >= 0; }

__device__ __forceinline__ uint4 mm_ld_reduce_bf16(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 mm_ld_reduce_f32(const void* p) {
  uint4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mm_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ld_peer(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(void* p, const uint4& v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// sum of the ranks' 16-byte granules at byte offset `off` without NVLS: fp32 accumulation in rank order
__device__ __forceinline__ uint4 peer_sum(const CommArgs& A, unsigned long long off, bool f32) {
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int q = 0; q < A.world; ++q) {
    const uint4 v = ld_peer(A.peers[q] + off);
    const unsigned int w[4] = {v.x, v.y, v.z, v.w};
    if (f32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(w[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
  uint4 r;
  if (f32) {
    r = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
  } else {
    unsigned int o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
      o[i] = *reinterpret_cast<const unsigned int*>(&h);
    }
    r = make_uint4(o[0], o[1], o[2], o[3]);
  }
  return r;
}

__device__ __forceinline__ void st_local(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The 16-byte granule `g` of a chunk (8 bf16 or 4 fp32 sums) -> the chunk's fp32 destination.  Contiguous: element i of
// the chunk is dst[i].  Row-strided (row_elems > 0): the chunk is a (rows, src_ld) matrix of which row_elems columns are
// valid; element (r, c) goes to dst[r * dst_ld + c] (rows of the destination need not be 16-byte aligned).
__device__ __forceinline__ void deliver(const CommChunk& ch, unsigned int g, const uint4& v) {
  const unsigned int w[4] = {v.x, v.y, v.z, v.w};
  if (ch.f32) {
    if (ch.row_elems <= 0) {
      __stcs(reinterpret_cast<float4*>(ch.dst) + g, make_float4(__uint_as_float(w[0]), __uint_as_float(w[1]), __uint_as_float(w[2]), __uint_as_float(w[3])));
    } else {
      const unsigned int e0 = g * 4u, r = e0 / (unsigned int)ch.src_ld; const int c0 = (int)(e0 - r * (unsigned int)ch.src_ld);
#pragma unroll
      for (int i = 0; i < 4; ++i) if (c0 + i < ch.row_elems) ch.dst[(long long)r * ch.dst_ld + c0 + i] = __uint_as_float(w[i]);
    }
    return;
  }
  float f[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  if (ch.row_elems <= 0) {
    float4* d = reinterpret_cast<float4*>(ch.dst) + 2 * g;
    __stcs(d, make_float4(f[0], f[1], f[2], f[3]));
    __stcs(d + 1, make_float4(f[4], f[5], f[6], f[7]));
  } else {
    const unsigned int e0 = g * 8u, r = e0 / (unsigned int)ch.src_ld; const int c0 = (int)(e0 - r * (unsigned int)ch.src_ld);
#pragma unroll
    for (int i = 0; i < 8; ++i) if (c0 + i < ch.row_elems) ch.dst[(long long)r * ch.dst_ld + c0 + i] = f[i];
  }
}

template <bool kPull>
__global__ void __maxnreg__(48) grad_exchange_kernel(const CommArgs A) {
  const unsigned int seq = A.state[0] + 1u;               // this step's number: what a raised flag holds
  unsigned int* my_sig_ready = reinterpret_cast<unsigned int*>(A.sig[A.rank] + kReadyOffset);
  unsigned int* my_sig_done = reinterpret_cast<unsigned int*>(A.sig[A.rank] + kDoneOffset);

  if (blockIdx.x == 0) {
    // ------------------------------------------------------------------ herald
    if (threadIdx.x == 0) {
      comm_stamp(A.trace, 0);
      for (int c = 0; c < A.n_chunks; ++c) {
        const CommChunk& ch = A.ch[c];
        if (ch.notify != nullptr) {
          const unsigned int target = seq * ch.increment;   // counters only ever advance (mod 2^32 arithmetic)
          const long long t0 = clock64();
          for (unsigned int i = 0; i < ch.count; ++i)
            while (!reached(comm_ld_acquire_gpu(ch.notify + i), target)) {
              __nanosleep(A.sleep_ns);
              if (clock64() - t0 > kCommSpinLimit) __trap();
            }
        }
        // ONE system-scope fence (everything observed through the counters is ordered before the flags), then plain
        // flag stores: a release store per peer is a fence per peer -- measured ~3 us each, x 8 peers x 8 chunks on 8 GPUs
        __threadfence_system();
        for (int q = 0; q < A.world; ++q)
          comm_st_relaxed_sys(reinterpret_cast<unsigned int*>(A.sig[q] + kReadyOffset) + c * NERAF_MAX_RANKS + A.rank, seq);
        comm_stamp(A.trace, 4 + 4 * c);                     // this rank's chunk c announced
      }
      // every worker of this rank has reduced and broadcast its slices
      {
        const long long t0 = clock64();
        while (comm_ld_acquire_gpu(A.state + 1) != gridDim.x - 1) {
          __nanosleep(A.sleep_ns);
          if (clock64() - t0 > kCommSpinLimit) __trap();
        }
      }
      comm_stamp(A.trace, 1);                               // this rank's workers are done
      __threadfence_system();
      for (int q = 0; q < A.world; ++q)
        comm_st_relaxed_sys(reinterpret_cast<unsigned int*>(A.sig[q] + kDoneOffset) + A.rank, seq);
      for (int q = 0; q < A.world; ++q) {
        const long long t0 = clock64();
        while (!reached(comm_ld_acquire_sys(my_sig_done + q), seq)) {
          __nanosleep(A.sleep_ns);
          if (clock64() - t0 > kCommSpinLimit) __trap();
        }
      }
      comm_stamp(A.trace, 2);                               // every rank is done
      A.state[1] = 0u;
      A.state[0] = seq;
      __threadfence();
    }
    return;
  }

  // -------------------------------------------------------------------- workers
  const int workers = (int)gridDim.x - 1;
  const long long wtid = (long long)(blockIdx.x - 1) * kCommThreads + threadIdx.x;
  const long long wthreads = (long long)workers * kCommThreads;
  for (int c = 0; c < A.n_chunks; ++c) {
    const CommChunk& ch = A.ch[c];
    if (threadIdx.x < A.world) {
      const unsigned int* flag = my_sig_ready + c * NERAF_MAX_RANKS + threadIdx.x;
      const long long t0 = clock64();
      while (!reached(comm_ld_acquire_sys(flag), seq)) {
        __nanosleep(A.sleep_ns);
        if (clock64() - t0 > kCommSpinLimit) __trap();
      }
    }
    __syncthreads();
    if (blockIdx.x == 1 && threadIdx.x == 0) comm_stamp(A.trace, 4 + 4 * c + 1);     // every rank announced chunk c
    // this rank's slice of the chunk, in 16-byte granules (equal slices: the owner of a granule is g / per); 32-bit
    // indices throughout (a chunk is < 64 GB) -- the kernel lives on 48 registers
    const unsigned int granules = (unsigned int)(ch.bytes / 16);
    const unsigned int per = (granules + A.world - 1) / A.world;
    const unsigned int g0 = per * A.rank < granules ? per * A.rank : granules;
    const unsigned int g1 = g0 + per < granules ? g0 + per : granules;
    const unsigned int n = g1 - g0;
    const bool f32 = ch.f32 != 0;
    constexpr bool pull = kPull;
    const unsigned int wt = (unsigned int)wtid, wn = (unsigned int)wthreads;
    uint8_t* mine = A.peers[A.rank] + ch.offset;
    auto reduce16 = [&](unsigned int g) -> uint4 {
      const unsigned long long off = ch.offset + (unsigned long long)g * 16;
      if (A.mc != nullptr) return f32 ? mm_ld_reduce_f32(A.mc + off) : mm_ld_reduce_bf16(A.mc + off);
      return peer_sum(A, off, f32);
    };
    auto publish = [&](unsigned int g, const uint4& v) {              // the sum of granule g of the chunk
      if (!pull) {
        const unsigned long long off = ch.offset + (unsigned long long)g * 16;
        if (A.mc != nullptr) mm_st(A.mc + off, v);
        else for (int q = 0; q < A.world; ++q) st_peer(A.peers[q] + off, v);
      } else {
        st_local(mine + (unsigned long long)g * 16, v);               // the other ranks fetch it from here
        if (ch.dst != nullptr) deliver(ch, g, v);
      }
    };
    {
      unsigned int i = wt;
      for (; i + (kCommUnroll - 1) * wn < n; i += kCommUnroll * wn) {
        uint4 v[kCommUnroll];
#pragma unroll
        for (int u = 0; u < kCommUnroll; ++u) v[u] = reduce16(g0 + i + u * wn);
#pragma unroll
        for (int u = 0; u < kCommUnroll; ++u) publish(g0 + i + u * wn, v[u]);
      }
      for (; i < n; i += wn) publish(g0 + i, reduce16(g0 + i));
    }
    if (pull && A.world > 1) {
      // ---- second shot, turned around: announce this rank's reduced slice, then fetch the other ranks' slices
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(A.state + kChunkTickets + c, 1u);
        if (t == (unsigned int)workers - 1) {                         // every worker's part of the slice is stored
          A.state[kChunkTickets + c] = 0u;
          __threadfence_system();
          for (int q = 0; q < A.world; ++q)
            comm_st_relaxed_sys(reinterpret_cast<unsigned int*>(A.sig[q] + kReducedOffset) + c * NERAF_MAX_RANKS + A.rank, seq);
        }
      }
      if (threadIdx.x < A.world && threadIdx.x != A.rank) {
        const unsigned int* flag = reinterpret_cast<const unsigned int*>(A.sig[A.rank] + kReducedOffset) + c * NERAF_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (!reached(comm_ld_acquire_sys(flag), seq)) {
          __nanosleep(A.sleep_ns);
          if (clock64() - t0 > kCommSpinLimit) __trap();
        }
      }
      __syncthreads();
      const unsigned int others = granules - n;
      auto granule_of = [&](unsigned int j) { return j < g0 ? j : j + n; };     // j-th granule that is not this rank's
      auto keep = [&](unsigned int g, const uint4& v) {
        if (ch.dst != nullptr) deliver(ch, g, v);
        else st_local(mine + (unsigned long long)g * 16, v);
      };
      // thread-private slots: group h, slot u of thread t at ((h * kPullDepth + u) * kCommThreads + t) * 16
      extern __shared__ __align__(16) uint8_t pull_smem[];
      uint4* slots = reinterpret_cast<uint4*>(pull_smem) + threadIdx.x;
      auto issue = [&](int h, unsigned int j0) {                      // group h <- granules j0, j0 + wn, ... (those that exist)
#pragma unroll
        for (int u = 0; u < kPullDepth; ++u) {
          const unsigned int j = j0 + u * wn;
          if (j < others) {
            const unsigned int g = granule_of(j);
            cp_async16(slots + (h * kPullDepth + u) * kCommThreads, A.peers[g / per] + ch.offset + (unsigned long long)g * 16);
          }
        }
        cp_async_commit();
      };
      auto drain = [&](int h, unsigned int j0) {
#pragma unroll
        for (int u = 0; u < kPullDepth; ++u) {
          const unsigned int j = j0 + u * wn;
          if (j < others) keep(granule_of(j), slots[(h * kPullDepth + u) * kCommThreads]);
        }
      };
      const unsigned int step = kPullDepth * wn;
      unsigned int j = wt;
      int h = 0;
      if (j < others) issue(0, j);
      for (; j < others; j += step, h ^= 1) {
        const bool more = j + step < others;
        if (more) issue(h ^ 1, j + step);                             // the next group is in flight while this one drains
        if (more) cp_async_wait<1>(); else cp_async_wait<0>();
        drain(h, j);
      }
    }
    if (blockIdx.x == 1 && threadIdx.x == 0) comm_stamp(A.trace, 4 + 4 * c + 2);     // block 1 has issued its share of chunk c
  }
  // every store of this block is ordered before its ticket: the barrier orders the block's stores before thread 0, whose
  // single system fence is cumulative over them (the pattern of a grid-wide barrier, at system scope) -- 128 fences per
  // block instead cost ~5 us at the end of the kernel (profiles/r02al_*: 431 -> 428 us per step); NERAF_COMM_ONE_FENCE=0
  // restores a fence per thread
  if (!A.one_fence) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (A.one_fence) __threadfence_system();
    atomicAdd(A.state + 1, 1u);
  }
}

}  // namespace neraf

using namespace neraf;

namespace neraf {
int dp_exchange_grads(const neraf_grad_exchange* x, cudaStream_t stream, bool beside_previous) {
  NERAF_REQUIRE(x, "dp_exchange_grads: null argument");
  NERAF_REQUIRE(x->world >= 1 && x->world <= NERAF_MAX_RANKS && x->rank >= 0 && x->rank < x->world,
                "dp_exchange_grads: bad world / rank");
  NERAF_REQUIRE(x->n_chunks >= 1 && x->n_chunks <= NERAF_MAX_EXCHANGE_CHUNKS, "dp_exchange_grads: 1..%d chunks",
                NERAF_MAX_EXCHANGE_CHUNKS);
  NERAF_REQUIRE(x->state, "dp_exchange_grads: state is null");
  CommArgs A = {};
  A.n_chunks = x->n_chunks; A.world = x->world; A.rank = x->rank; A.pull = x->pull ? 1 : 0;
  A.mc = reinterpret_cast<uint8_t*>(x->multicast);
  A.state = x->state;
  A.trace = reinterpret_cast<unsigned long long*>(x->trace);
  {
    const char* e = getenv("NERAF_COMM_SLEEP_NS");        // tuning: back-off between polls (default 200 ns)
    A.sleep_ns = e ? (unsigned int)atoi(e) : 200u;
    const char* f = getenv("NERAF_COMM_ONE_FENCE");
    A.one_fence = (f && f[0] == '0') ? 0 : 1;
  }
  NERAF_REQUIRE(!A.mc || ((uintptr_t)A.mc & 15) == 0, "dp_exchange_grads: misaligned multicast mapping");
  for (int r = 0; r < x->world; ++r) {
    NERAF_REQUIRE(x->peers[r] && x->signals[r], "dp_exchange_grads: region / signal buffer of rank %d is null", r);
    NERAF_REQUIRE(((uintptr_t)x->peers[r] & 15) == 0, "dp_exchange_grads: misaligned region of rank %d", r);
    A.peers[r] = reinterpret_cast<uint8_t*>(x->peers[r]);
    A.sig[r] = reinterpret_cast<uint8_t*>(x->signals[r]);
  }
  for (int c = 0; c < x->n_chunks; ++c) {
    const neraf_exchange_chunk& s = x->chunks[c];
    NERAF_REQUIRE(s.offset % 16 == 0 && s.bytes % 16 == 0, "dp_exchange_grads: chunk %d is not 16-byte tileable", c);
    NERAF_REQUIRE(!s.notify || (s.notify_increment > 0 && s.notify_count > 0),
                  "dp_exchange_grads: chunk %d: notify without a counter count / increment", c);
    NERAF_REQUIRE(!s.dst || x->pull, "dp_exchange_grads: chunk %d: a destination needs the pull form", c);
    NERAF_REQUIRE(!s.dst || s.row_elems > 0 || ((uintptr_t)s.dst & 15) == 0, "dp_exchange_grads: chunk %d: misaligned destination", c);
    NERAF_REQUIRE(!s.dst || s.row_elems <= 0 || (s.src_ld >= s.row_elems && s.src_ld % (s.f32 ? 4 : 8) == 0 && s.dst_ld >= s.row_elems),
                  "dp_exchange_grads: chunk %d: bad row-strided destination", c);
    A.ch[c] = CommChunk{(unsigned long long)s.offset, (unsigned long long)s.bytes, s.notify, s.notify_count, s.notify_increment,
                        s.f32 ? 1 : 0, s.dst, (long long)s.dst_ld, s.dst ? s.row_elems : 0, s.src_ld};
  }
  // one small CTA per SM (it must fit BESIDE a CTA of the job-list kernel), never more than are resident at once
  int grid = x->max_ctas > 0 ? x->max_ctas : sm_count();
  if (grid > sm_count()) grid = sm_count();
  if (grid < 2) grid = 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kCommThreads); cfg.stream = stream;
  cfg.dynamicSmemBytes = (A.pull && A.world > 1) ? kPullSmemBytes : 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = beside_previous ? 1 : 0;
  if (A.pull) NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, grad_exchange_kernel<true>, A));
  else NERAF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, grad_exchange_kernel<false>, A));
  NERAF_CHECK_LAUNCH("grad_exchange_kernel");
  return NERAF_OK;
}
}  // namespace neraf

extern "C" int neraf_dp_exchange_grads(const neraf_grad_exchange* x, neraf_stream_t stream) {
  return neraf::dp_exchange_grads(x, (cudaStream_t)stream, false);
}
