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

// HEAVY: ~160 live registers per thread, like the job-list GEMM kernel (168)
template <int LIVE>
__global__ void __launch_bounds__(320, 1) primary(unsigned long long* stamps, int spin_us, int trigger_early, float* sink) {
  extern __shared__ unsigned char smem[];
  if (threadIdx.x == 0) smem[0] = 1;
  if (trigger_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned long long t0 = gtime();
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = t0;
  constexpr bool HEAVY = LIVE > 1;
  float acc[HEAVY ? LIVE : 1];
#pragma unroll
  for (int i = 0; i < (HEAVY ? LIVE : 1); ++i) acc[i] = (float)(threadIdx.x + i);
  while (gtime() - t0 < (unsigned long long)spin_us * 1000ull) {
#pragma unroll
    for (int i = 0; i < (HEAVY ? LIVE : 1); ++i) acc[i] = fmaf(acc[i], 1.0001f, (float)i);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < (HEAVY ? LIVE : 1); ++i) s += acc[i];
  if (s == 12345.678f) sink[0] = s;
  if (blockIdx.x == 0 && threadIdx.x == 0) stamps[1] = gtime();
}
template <int THREADS, int REGS>
__global__ void __maxnreg__(REGS) dependent(unsigned long long* stamps, float* sink) {
  float v[REGS > 40 ? 40 : (REGS > 32 ? 16 : 8)];
#pragma unroll
  for (int i = 0; i < (int)(sizeof(v) / 4); ++i) v[i] = sink[threadIdx.x + 32 * i];
  if (threadIdx.x == 0) {
    const unsigned long long t = gtime();
    atomicMin(stamps + 2, t);
    atomicMax(stamps + 3, t);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < (int)(sizeof(v) / 4); ++i) s += v[i];
  if (s == 12345.678f) sink[0] = s;
}

static float* g_sink;
template <typename K> static void launch_dep(K kern, int threads, cudaStream_t s, unsigned long long* d, bool pdl) {
  cudaLaunchConfig_t c2 = {};
  c2.gridDim = dim3(148); c2.blockDim = dim3(threads); c2.stream = s;
  cudaLaunchAttribute b[1];
  b[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; b[0].val.programmaticStreamSerializationAllowed = 1;
  c2.attrs = b; c2.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&c2, kern, d, g_sink));
}
static int g_heavy_threads = 320;
static void launch_heavy(cudaStream_t s, unsigned long long* d, int smem, int variant) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(g_heavy_threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, primary<150>, d, 100, 1, g_sink));
  if (variant == 0) launch_dep(dependent<256, 40>, 256, s, d, true);
  if (variant == 1) launch_dep(dependent<128, 48>, 128, s, d, true);
  if (variant == 2) launch_dep(dependent<128, 32>, 128, s, d, true);
  if (variant == 3) launch_dep(dependent<64, 32>, 64, s, d, true);
  if (variant == 4) launch_dep(dependent<256, 32>, 256, s, d, true);
  if (variant == 5) launch_dep(dependent<32, 32>, 32, s, d, true);
}
static void launch(cudaStream_t s, unsigned long long* d, int smem, int cluster, int early, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute a[1];
  a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cluster; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, primary<1>, d, 100, early, g_sink));
  launch_dep(dependent<256, 40>, 256, s, d, pdl);
}

int main() {
  unsigned long long* d; CK(cudaMalloc(&d, 64));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  CK(cudaFuncSetAttribute(primary<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaFuncSetAttribute(primary<150>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  CK(cudaMalloc(&g_sink, 1 << 20)); CK(cudaMemset(g_sink, 0, 1 << 20));
  const int smems[] = {231568, 202896, 198800, 180000, 100000};
  {
    auto run = [&](auto kern, const char* what) {
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
      cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
      for (int v = 0; v < 3; ++v) {
        unsigned long long h[4] = {0, 0, ~0ull, 0};
        CK(cudaMemcpy(d, h, 32, cudaMemcpyHostToDevice));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = 202896; cfg.stream = s;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        cfg.attrs = a; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, kern, d, 100, 1, g_sink));
        const char* dn = v == 0 ? "128 thr x 48 regs" : (v == 1 ? "64 thr x 48 regs" : "32 thr x 32 regs");
        if (v == 0) launch_dep(dependent<128, 48>, 128, s, d, true);
        if (v == 1) launch_dep(dependent<64, 48>, 64, s, d, true);
        if (v == 2) launch_dep(dependent<32, 32>, 32, s, d, true);
        CK(cudaStreamSynchronize(s));
        CK(cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost));
        printf("primary 320 threads x %3d registers (%s), dependent %-18s: started %6.1f .. %6.1f us after the primary\n", fa.numRegs, what,
               dn, ((double)h[2] - (double)h[0]) / 1e3, ((double)h[3] - (double)h[0]) / 1e3);
      }
    };
    run(primary<150>, "150 live"); run(primary<142>, "142 live"); run(primary<134>, "134 live"); run(primary<126>, "126 live");
    run(primary<118>, "118 live"); run(primary<110>, "110 live"); run(primary<100>, "100 live"); run(primary<80>, "80 live");
  }
  for (int graph = 2; graph < 2; ++graph)
    for (int cluster = 1; cluster <= 2; ++cluster)
      for (int si = 1; si < 3; ++si)
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
