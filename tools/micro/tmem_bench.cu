// Microbenchmark: tcgen05.ld throughput per SM for different shapes / warp counts (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld32x32b(uint32_t taddr, uint32_t (&r)[X]);
template <>
__device__ __forceinline__ void ld32x32b<32>(uint32_t taddr, uint32_t (&r)[32]) {
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
template <>
__device__ __forceinline__ void ld32x32b<16>(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  // 16 lanes x 256 bit, x8 -> 32 registers per thread (covers 16 lanes x 64 columns)
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: 32x32b.x32 + wait each; 1: two x32 in flight before a wait; 2: x16 + wait; 3: 16x256b.x8 + wait
__global__ void bench(int mode, int iters, unsigned long long* out, uint32_t* sink) {
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
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = (uint32_t)((i * 32) & 255) + (uint32_t)((warp >> 2) * 256 & 255);
    if (mode == 0) {
      uint32_t r[32]; ld32x32b<32>(base + col, r); ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q];
    } else if (mode == 1) {
      uint32_t r[32], r2[32]; ld32x32b<32>(base + col, r); ld32x32b<32>(base + ((col + 32) & 255), r2); ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q] ^ r2[q];
    } else if (mode == 2) {
      uint32_t r[16]; ld32x32b<16>(base + col, r); ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc ^= r[q];
    } else {
      uint32_t r[32]; ld16x256b_x8(base + (col & 191), r); ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) acc ^= r[q];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (lane == 0) out[blockIdx.x * 32 + warp] = (unsigned long long)(t1 - t0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

int main() {
  unsigned long long* out; uint32_t* sink;
  cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
  const int iters = 2000;
  const char* names[] = {"32x32b.x32 + wait", "2 x (32x32b.x32) + wait", "32x32b.x16 + wait", "16x256b.x8 + wait"};
  const int bytes[] = {4096, 8192, 2048, 2048 * 2};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      bench<<<1, warps * 32>>>(mode, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d warps %d: %s\n", mode, warps, cudaGetErrorString(e)); return 1; }
      unsigned long long h[32]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      unsigned long long mx = 0; for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
      const double cyc = (double)mx / iters;
      printf("%-26s warps %2d: %7.1f cycles / iteration / warp  -> %6.1f B/clk per SM\n", names[mode], warps, cyc,
             (double)bytes[mode] * warps / cyc);
    }
  return 0;
}
