// Shared helpers for libddn_b200.so: error reporting, launch accounting, small device utilities.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/ddn_b200.h"

namespace ddn {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_cuda(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DDN_ERR_CUDA;
  }
  return DDN_OK;
}

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
inline int after_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaPeekAtLastError(), name);
}

#define DDN_REQUIRE(cond, msg)                      \
  do {                                              \
    if (!(cond)) {                                  \
      ::ddn::set_error("invalid argument: %s", msg); \
      return DDN_ERR_INVALID_ARGUMENT;              \
    }                                               \
  } while (0)

#define DDN_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != DDN_OK) return _rc; \
  } while (0)

constexpr int kNumSMs = 148;  // B200

inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Order-preserving float <-> int encoding for atomicMin/atomicMax on floats.
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
  return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

}  // namespace ddn
