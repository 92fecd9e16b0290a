// Shared helpers for libctcvr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/ctcvr.h"

namespace ctcvr {

void set_error(const char* fmt, ...);
void count_launch();   // bumps the counter behind ctcvr_launch_count()

#define CTCVR_CHECK_CUDA(expr)                                                            \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::ctcvr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                         cudaGetErrorString(_e));                                         \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

#define CTCVR_REQUIRE(cond, ...)                                                          \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::ctcvr::set_error(__VA_ARGS__);                                                    \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

#define CTCVR_LAUNCH_CHECK()                                                              \
  do {                                                                                    \
    ::ctcvr::count_launch();                                                              \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::ctcvr::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                         cudaGetErrorString(_e));                                         \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// log(exp(a)+exp(b)), -inf safe.
__device__ __forceinline__ float log_add_exp(float a, float b) {
  float m = fmaxf(a, b);
  if (m == kNegInf) return kNegInf;
  return m + log1pf(expf(-fabsf(a - b)));
}
__device__ __forceinline__ double log_add_exp(double a, double b) {
  double m = fmax(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1p(exp(-fabs(a - b)));
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace ctcvr
