// Misc C-ABI entry points: version, error text, device probe.
#include "common.cuh"

#include <cstring>

namespace neraf {

std::atomic<long long> g_launch_count{0};

char* last_error_buffer() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace neraf

extern "C" {

int neraf_version(void) { return NERAF_ABI_VERSION; }

const char* neraf_last_error(void) { return neraf::last_error_buffer(); }

size_t neraf_abi_sizeof(const char* name) {
  if (!name) return 0;
#define NERAF_SIZEOF(T) if (strcmp(name, #T) == 0) return sizeof(T)
  NERAF_SIZEOF(neraf_field_dims); NERAF_SIZEOF(neraf_queries); NERAF_SIZEOF(neraf_multicast);
  NERAF_SIZEOF(neraf_rank_exchange); NERAF_SIZEOF(neraf_loss_grad); NERAF_SIZEOF(neraf_dp_options);
  NERAF_SIZEOF(neraf_exchange_chunk); NERAF_SIZEOF(neraf_grad_exchange); NERAF_SIZEOF(neraf_gl_params);
  NERAF_SIZEOF(neraf_metric_params); NERAF_SIZEOF(neraf_gemm_epilogue); NERAF_SIZEOF(neraf_gemm_job);
  NERAF_SIZEOF(neraf_window3d);
#undef NERAF_SIZEOF
  return 0;
}

long long neraf_launch_count(void) { return neraf::g_launch_count.load(std::memory_order_relaxed); }

int neraf_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
