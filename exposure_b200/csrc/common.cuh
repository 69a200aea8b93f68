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

// Same, as one thread-block cluster of `cluster_x` CTAs along x per (blockIdx.y, blockIdx.z).
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster_x,
                                      cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// Sum N doubles over every thread of every CTA of the cluster (of the CTA alone when the kernel was launched without
// one): warp shuffles, one shared-memory pass per CTA, then EVERY CTA adds the per-CTA totals in rank order out of
// distributed shared memory -- all CTAs of the cluster end up with the same bits.  sh: THREADS / 32 * N doubles,
// xch: N doubles, both __shared__.  The per-image reductions of the critic statistics run on this with 8 CTAs per image:
// one CTA per image is a serial walk of P / 256 dependent loads per thread (17 us for 64 x 64 pixels).
template <int THREADS, int N>
__device__ __forceinline__ void cluster_sum(double (&v)[N], double* sh, double* xch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) sh[warp * N + i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double t = 0.0;
    for (int w = 0; w < THREADS / 32; ++w) t += sh[w * N + threadIdx.x];
    xch[threadIdx.x] = t;
  }
  unsigned nb;
  asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(nb));
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = 0.0;
  const uint32_t local = (uint32_t)__cvta_generic_to_shared(xch);
  for (unsigned r = 0; r < nb; ++r) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double t;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(t) : "r"(remote + 8u * i) : "memory");
      v[i] += t;
    }
  }
  // nobody may leave (or reuse xch) while a peer still reads its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_rank_x() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_size_x() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(r)); return r; }
#endif

#ifdef __CUDACC__
// for (i = threadIdx.x; i < n; i += nthreads) store(i, load(i)) with U loads IN FLIGHT per thread: a staging loop
// whose every iteration waits for its own global load pays one L2 round trip (~600 clocks) per element and thread.
template <int U, class Load, class Store>
__device__ __forceinline__ void batched_fill(int n, int nthreads, Load load, Store store) {
  for (int i0 = threadIdx.x; i0 < n; i0 += U * nthreads) {
    decltype(load(0)) v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n) v[u] = load(i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * nthreads;
      if (i < n) store(i, v[u]);
    }
  }
}
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace expo
