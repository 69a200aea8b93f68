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

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// A kernel launched through launch_pdl() may start while its predecessor on the stream (or in the
// captured CUDA graph) is still running: its CTAs become resident as soon as every CTA of the
// predecessor has executed pdl_trigger(), run their prologue (barrier init, TMEM allocation,
// tensor-map prefetch) and block in pdl_wait() until the predecessor has COMPLETED and its writes
// are visible.  Rule: no global-memory access before pdl_wait().  Both instructions are no-ops in a
// kernel that was launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define EXP_PDL_ENTRY()   \
  do {                    \
    ::expo::pdl_trigger(); \
    ::expo::pdl_wait();    \
  } while (0)

bool pdl_enabled();   // process-wide switch: exp_set_pdl() / EXPOSURE_PDL (filters.cu)

// kern<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute when
// PDL is enabled.  ONLY for kernels that execute pdl_wait() before touching global memory.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (pdl_enabled()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace expo
