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

// Optional per-launch timing (ddn_profile_enable): an event after every launch that passes its stream.
void profile_mark(const char* name, cudaStream_t st);
extern std::atomic<int> g_profile;

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
inline int after_launch(const char* name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaPeekAtLastError(), name);
}
inline int after_launch(const char* name, cudaStream_t st) {
  if (g_profile.load(std::memory_order_relaxed)) profile_mark(name, st);
  return after_launch(name);
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

// World position of a pixel from P = d*x, Q = d*y, d and the view's 16-float src_table row (rows of
// R_s^T Kinv_s with the camera centre appended; scripts/test.py:79-90 then :233).  K4 (filter.cu) and the
// bounding-box epilogue of K3 (align.cu) both call this, so the box encloses K4's points bit for bit.
__device__ __forceinline__ void backproject_pqd(const float* __restrict__ s, float P, float Q, float d, float& X, float& Y,
                                                float& Z) {
  X = fmaf(s[0], P, fmaf(s[1], Q, fmaf(s[2], d, s[3])));
  Y = fmaf(s[4], P, fmaf(s[5], Q, fmaf(s[6], d, s[7])));
  Z = fmaf(s[8], P, fmaf(s[9], Q, fmaf(s[10], d, s[11])));
}

}  // namespace ddn
