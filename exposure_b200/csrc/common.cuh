// Shared helpers for the exposure_b200 C-ABI library (error reporting, launch checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/exposure_b200.h"

namespace expo {

// thread-local message returned by exp_last_error()
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define EXP_CHECK_ARG(cond, ...)                                    \
  do {                                                              \
    if (!(cond)) return ::expo::set_error(EXP_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define EXP_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess)                                                           \
      return ::expo::set_error(EXP_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace expo
