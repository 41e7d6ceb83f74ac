// Microbenchmark (B200, sm_100a): what bounds the operand stream of the job-list GEMM kernel, and what TMA multicast buys.
// Every CTA runs the kernel's 6-stage TMA/mbarrier ring WITHOUT the MMA (a consumer thread frees a stage as soon as it is
// full): per stage one "A" box and one "B" box of 128 rows x 64 bf16 (16 KB each, 128-byte swizzle), 148 CTAs.
//   mode 0  every CTA streams its own A and its own B                         (no sharing: the L2 -> SM ceiling)
//   mode 1  groups of G CTAs read the SAME A rows with plain (unicast) loads  (what pairs on one row block do today)
//   mode 2  clusters of G CTAs: each loads 1/G of the A box and multicasts it to the whole cluster; B stays private
//   mode 3  like 2, and B is multicast too (every CTA of the cluster gets the same A and B: pure multicast ceiling)
// Prints delivered GB/s (bytes landing in shared memory) per mode; L2 read traffic is 1, 1 (or less if the L2
// de-duplicates), (1 + 1/G)/2 and 1/G of that.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int STAGES = 6;
constexpr int BOX_BYTES = 128 * 64 * 2;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\nmbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n"
               ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// tmFull: box 128 rows x 64 cols; tmSlice[g]: box (128 / G) rows x 64 cols
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tmFull,
                                                        const __grid_constant__ CUtensorMap tmSlice, int mode, int G, int num_kb,
                                                        int passes, int rows_total, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;
  uint8_t* sb = smem + STAGES * BOX_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * STAGES * BOX_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t crank = (mode >= 2) ? cluster_rank() : 0;
  const int cta = blockIdx.x;
  const int group = cta / G;                    // CTAs of one group share A
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, mode >= 2 ? G : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode >= 2) cluster_sync();
  const long long t0 = clock64();
  // row ranges: A of a group at [group * 128 ...), B of a CTA (or of a group in mode 3) in the second half of the matrix
  const int a_row = (mode == 0 ? cta : group) * 128 % (rows_total / 2);
  const int b_row = rows_total / 2 + ((mode == 3 ? group : cta) * 128) % (rows_total / 2);
  const int slice_rows = 128 / G;
  const uint16_t mask = (uint16_t)((1u << G) - 1);
  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < num_kb * passes; ++it) {
      const int kb = it % num_kb;
      mbar_wait(empty + stage, phase ^ 1);
      mbar_expect_tx(full + stage, 2 * BOX_BYTES);
      if (mode < 2) {
        tma_load(sa + stage * BOX_BYTES, &tmFull, full + stage, kb * 64, a_row);
        tma_load(sb + stage * BOX_BYTES, &tmFull, full + stage, kb * 64, b_row);
      } else {
        tma_load_mc(sa + stage * BOX_BYTES + crank * slice_rows * 128, &tmSlice, full + stage, kb * 64, a_row + crank * slice_rows, mask);
        if (mode == 2) tma_load(sb + stage * BOX_BYTES, &tmFull, full + stage, kb * 64, b_row);
        else tma_load_mc(sb + stage * BOX_BYTES + crank * slice_rows * 128, &tmSlice, full + stage, kb * 64, b_row + crank * slice_rows, mask);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < num_kb * passes; ++it) {
      mbar_wait(full + stage, phase);
      if (mode < 2) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(empty + stage)) : "memory");
      } else {
        for (int r = 0; r < G; ++r) mbar_arrive_cluster(empty + stage, (uint32_t)r);   // every CTA that writes into my stage
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (mode >= 2) cluster_sync();
  if (threadIdx.x == 0) cycles[cta] = (unsigned long long)(clock64() - t0);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeFn fn, void* ptr, int64_t rows, int64_t cols, int box_rows) {
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "encode failed %d\n", (int)r); exit(1); }
  return tm;
}

int main(int argc, char** argv) {
  // K = 64 * num_kb columns, walked `passes` times: (8, 64) keeps the 39 MB working set in L2 (the kernel's operands
  // are ~78 % L2 hits), (256, 2) streams 1.24 GB from HBM
  const int num_kb = argc > 1 ? atoi(argv[1]) : 8;
  const int passes = argc > 2 ? atoi(argv[2]) : 64;
  const int64_t cols = (int64_t)num_kb * 64, rows = 2 * 148 * 128;
  void* buf;
  cudaMalloc(&buf, (size_t)rows * cols * 2);
  cudaMemset(buf, 0, (size_t)rows * cols * 2);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  EncodeFn fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q);
  if (!fn) { fprintf(stderr, "no cuTensorMapEncodeTiled\n"); return 1; }
  const int smem = 2 * STAGES * BOX_BYTES + 2 * STAGES * 8 + 64;
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  CUtensorMap full = make_map(fn, buf, rows, cols, 128);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int gs[] = {1, 2, 4, 8};
  for (int mode = 0; mode < 4; ++mode) {
    for (int gi = 0; gi < 4; ++gi) {
      const int G = gs[gi];
      if ((mode == 0) != (G == 1)) continue;
      CUtensorMap slice = make_map(fn, buf, rows, cols, 128 / G);
      cudaLaunchConfig_t cfg = {};
      cfg.blockDim = dim3(64);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = mode >= 2 ? G : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int grid = 148;
      if (mode >= 2) {
        cfg.gridDim = dim3(148 / G * G);
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, stream_kernel, &cfg) != cudaSuccess) { printf("mode %d G %d: occupancy query failed: %s\n", mode, G, cudaGetErrorString(cudaGetLastError())); continue; }
        grid = (n < 148 / G ? n : 148 / G) * G;
        printf("mode %d G %d: %d clusters resident -> %d CTAs\n", mode, G, n, grid);
      }
      cfg.gridDim = dim3(grid);
      float best = 1e9f;
      for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        cudaError_t err = cudaLaunchKernelEx(&cfg, stream_kernel, full, slice, mode, G, num_kb, passes, (int)rows, cyc);
        cudaEventRecord(e1);
        if (err != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { printf("mode %d G %d failed: %s\n", mode, G, cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
      }
      const double delivered = (double)grid * num_kb * passes * 2 * BOX_BYTES;
      double l2 = delivered;
      if (mode == 2) l2 = delivered * (1.0 + 1.0 / G) / 2;
      if (mode == 3) l2 = delivered / G;
      printf("mode %d G %d: %d CTAs, %.1f us, delivered %.2f TB/s (%.1f B/clk/SM @1.9 GHz), distinct bytes read %.2f TB/s, "
             "%.3f us per 32 KB stage\n", mode, G, grid, best * 1e3, delivered / best / 1e9, delivered / best / 1e9 * 1e3 / 1.9 / grid * 1.0,
             l2 / best / 1e9, best * 1e3 / (num_kb * passes));
    }
  }
  return 0;
}
