// Shared helpers for the neraf_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/neraf_b200.h"

namespace neraf {

// Thread-local text of the last error, returned by neraf_last_error().
char* last_error_buffer();
int set_error(int code, const char* fmt, ...);
// Number of kernels this library has launched (process-wide); read through neraf_launch_count().
extern std::atomic<long long> g_launch_count;

#define NERAF_CHECK_CUDA(expr)                                                                   \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::neraf::set_error(NERAF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                     \
  } while (0)

#define NERAF_CHECK_LAUNCH(name)                                                                 \
  do {                                                                                           \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return ::neraf::set_error(NERAF_ERR_CUDA, "launch of %s failed: %s (%s:%d)", name,         \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                     \
    ::neraf::g_launch_count.fetch_add(1, std::memory_order_relaxed);                             \
  } while (0)

#define NERAF_REQUIRE(cond, ...)                                                                 \
  do {                                                                                           \
    if (!(cond)) return ::neraf::set_error(NERAF_ERR_INVALID, __VA_ARGS__);                      \
  } while (0)

#define NERAF_TRY(expr)                                                                          \
  do {                                                                                           \
    int _rc = (expr);                                                                            \
    if (_rc != NERAF_OK) return _rc;                                                             \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

int sm_count();

constexpr float kLeakySlope = 0.1f;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NERAF_ACT_LEAKY) return v > 0.f ? v : kLeakySlope * v;
  if (act == NERAF_ACT_TANH10) return 10.f * tanhf(v);
  return v;
}

}  // namespace neraf
