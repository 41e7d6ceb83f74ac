// Microbenchmark of the epilogue building blocks (B200): per-iteration cycles of
//   LDTM.x32 + wait | + 16 cvt.bf16x2 + 4 STS.128 | + fence.proxy.async | + TMA store (2 KB) + commit + wait_read
// for 8 warps of one CTA.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bench epi_bench.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// level 0: LDTM only; 1: + cvt + STS; 2: + fence.proxy.async + syncwarp; 3: + TMA store + commit + wait_read0;
// 4: like 3 but two staging halves and wait_read 1; 5: level 1 + 32 shuffles + 32 fadd + leaky (the math of a chunk)
__global__ void __launch_bounds__(256, 1) bench(int level, int iters, const __grid_constant__ CUtensorMap tm,
                                               unsigned long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint8_t* wbuf = smem + warp * 4096;
  float acc = 0.f;
  const float bk = (float)lane;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[32];
    ld32(base + (uint32_t)((i * 32) & 255), v);
    if (level == 5) {
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] += __shfl_sync(0xffffffffu, bk, q);
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = fmaxf(v[q], 0.1f * v[q]);
    }
    if (level >= 1) {
      uint8_t* sb = wbuf + ((level == 4) ? (i & 1) * 2048 : 0);
      if (level >= 3) {
        if (lane == 0) {
          if (level == 4) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncwarp();
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
        *reinterpret_cast<uint4*>(sb + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = pk;
      }
      if (level >= 2) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
      }
      if (level >= 3 && lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm),
                     "r"(smem_u32(sb)), "r"((i * 32) & 255), "r"(warp * 32)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
#pragma unroll
      for (int q = 0; q < 32; ++q) acc += v[q];
    }
  }
  if (level >= 3 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[warp] = (unsigned long long)(t1 - t0);
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  unsigned long long* out; float* sink; __nv_bfloat16* dst;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&sink, 1024 * 4); cudaMalloc(&dst, 256 * 256 * 2);
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {256, 256}; const cuuint64_t gstride[1] = {512};
  const cuuint32_t box[2] = {32, 32}; const cuuint32_t estride[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dst, gdim, gstride, box, estride,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const char* names[] = {"LDTM.x32 + wait", "+ cvt + 4 STS.128", "+ fence.proxy.async + syncwarp", "+ TMA store, commit, wait_read 0",
                         "+ TMA store, 2 halves, wait_read 1", "LDTM + 32 SHFL/FADD + leaky + cvt + STS"};
  const int iters = 2000;
  for (int level = 0; level < 6; ++level) {
    bench<<<1, 256, 65536>>>(level, iters, tm, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("level %d: %s\n", level, cudaGetErrorString(e)); return 1; }
    unsigned long long h[8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0; for (int w = 0; w < 8; ++w) mx = h[w] > mx ? h[w] : mx;
    printf("%-42s %8.1f cycles / chunk (8 warps concurrently)\n", names[level], (double)mx / iters);
  }
  return 0;
}
