// Microbenchmark (B200): when does a programmatic dependent kernel that never waits actually start BESIDE its primary?
// Primary: persistent, one CTA (or CTA pair) per SM, `smem` bytes of dynamic shared memory, 320 threads, ~168 registers
// are not reproduced (registers are not the question), triggers launch_dependents at start and then spins for `us`.
// Dependent: 148 CTAs x 256 threads, stamps %globaltimer.  Prints dependent start - primary start for several footprints,
// launched directly and from a captured graph.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_beside pdl_beside.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(320, 1) primary(unsigned long long* stamps, int spin_us, int trigger_early) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0) smem[0] = 1;
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t0 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = t0;
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) __nanosleep(100);
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[1] = gtime();
}
__global__ void __launch_bounds__(256) dependent(unsigned long long* stamps) {
  if (threadIdx.x == 0) {
    const unsigned long long t = gtime();
    atomicMin(stamps + 2, t);
    atomicMax(stamps + 3, t);
  }
}

static void launch(cudaStream_t s, unsigned long long* d, int smem, int cluster, int early, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cluster; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, primary, d, 100, early));
  cudaLaunchConfig_t c2 = {};
  c2.gridDim = dim3(148); c2.blockDim = dim3(256); c2.stream = s;
  cudaLaunchAttribute b[1];
  b[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; b[0].val.programmaticStreamSerializationAllowed = 1;
  c2.attrs = b; c2.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&c2, dependent, d));
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 64));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  CK(cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  const int smems[] = {231568, 202896, 198800, 180000, 100000};
  for (int graph = 0; graph < 2; ++graph)
    for (int cluster = 1; cluster <= 2; ++cluster)
      for (int si = 0; si < 5; ++si)
        for (int mode = 0; mode < 3; ++mode) {           // 0: no PDL, 1: PDL + early trigger, 2: PDL, no trigger (implicit at exit)
          unsigned long long h[4] = {0, 0, ~0ull, 0};
          CK(cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice));
          if (!graph) {
            launch(s, d, smems[si], cluster, mode == 1, mode != 0);
          } else {
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            launch(s, d, smems[si], cluster, mode == 1, mode != 0);
            CK(cudaStreamEndCapture(s, &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, s));
            CK(cudaStreamSynchronize(s));
            CK(cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice));
            CK(cudaGraphLaunch(ge, s));
          }
          CK(cudaStreamSynchronize(s));
          CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
          printf("%s cluster %d smem %6d %-22s: primary ran %6.1f us; dependent CTAs started %6.1f .. %6.1f us after the primary\n",
                 graph ? "graph " : "stream", cluster, smems[si], mode == 0 ? "plain" : (mode == 1 ? "PDL, early trigger" : "PDL, no trigger"),
                 (h[1] - h[0]) / 1e3, ((double)h[2] - (double)h[0]) / 1e3, ((double)h[3] - (double)h[0]) / 1e3);
        }
  return 0;
}
