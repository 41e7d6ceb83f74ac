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

// e^x for the spectral loss (two per element in every K2 kernel).  libdevice's expf spends ~8 instructions per call,
// three of them on the half-rate integer pipe (scaling by 2^j): with two calls per element the loss kernels were
// INSTRUCTION-bound at 0.73 of the HBM roofline (profiles/r02k bench: forward 56 us for 268 MB).  MUFU.EX2 scales by
// itself, so all that is needed is the argument x log2(e) to better than fp32: t = fl(x L), r = the exact rounding
// error of that product plus x (log2 e - L), and 2^(t + r) = 2^t (1 + r ln 2) to first order (|r| < 2^-20).  Five
// FMA-pipe instructions + one MUFU; relative error <= 2^-22 (the MUFU's), the same 2 ulp class as expf.
__device__ __forceinline__ float exp_fma(float x) {
  constexpr float kL2E = 1.44269502162933349609375f;          // float(log2 e)
  constexpr float kL2E_lo = 1.92596299112661746e-08f;         // log2 e - float(log2 e)
  const float t = x * kL2E;
  const float r = fmaf(x, kL2E_lo, fmaf(x, kL2E, -t));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return fmaf(e, r * 0.693147180559945309f, e);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NERAF_ACT_LEAKY) return v > 0.f ? v : kLeakySlope * v;
  if (act == NERAF_ACT_TANH10) return 10.f * tanhf(v);
  return v;
}

}  // namespace neraf
