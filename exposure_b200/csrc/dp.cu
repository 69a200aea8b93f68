// Data-parallel optimizer step over NVLink / NVSwitch peer memory: ONE kernel per optimizer step that
// all-reduces the flat gradient buffer across the GPUs of the box AND applies Adam to the local replica.
//
// The reference has no multi-GPU code; the data-parallel step is new: every loss of net.py:92-199 is
// a batch mean, so the mean over ranks of the per-shard gradients is the global-batch gradient, and
// both optimizers of the generator step run in the same sess.run (net.py:330-331) -> theta_g and theta_v share
// one flat buffer and one exchange.
//
// One process per GPU.  Every rank maps every peer's gradient buffer, reduction scratch and flag block into its
// address space (CUDA IPC handles exchanged once by the host; exposure_b200/dp.py) and the kernel does
//
//   barrier 0   every rank has entered the kernel => its backward kernels (earlier on its stream) are complete
//   phase 1     reduce-scatter: rank r sums slice r of all `world` gradient buffers IN RANK ORDER (peer loads
//               over NVLink, 16-byte vectors) into slice r of its own scratch
//   barrier 1   all slices are reduced and visible (system-scope fence)
//   phase 2     all-gather fused with Adam: every rank reads slice s from rank s's scratch (peer loads) and
//               updates its own params / m / v -- the gathered gradient is never stored
//
// Each element is reduced by exactly one rank and read by all: the replicas stay BIT-IDENTICAL and the result does
// not depend on timing.  Per rank and step the NVLink traffic is 2 (world-1)/world of the buffer, like a ring
// all-reduce, but in two hops; no NCCL call, no host involvement, plain kernel node in the step's CUDA graph.
// Barriers: monotonic epochs in flag words written with st.release.sys / polled with ld.acquire.sys by block 0;
// the other blocks of the (co-resident, one per SM) grid wait on a device-scope flag.  No barrier is needed at
// the end: a rank overwrites its gradient buffer only after barrier 1 of this step (peers are done reading it)
// and its scratch only after barrier 0 of the next step on the same buffer (peers have left this kernel).
#include <string.h>

#include "common.cuh"

namespace expo {

constexpr int kDpMaxWorld = 8;
constexpr int kDpThreads = 512;
// flag block layout (uint32 words); the host zero-fills it once
constexpr int kDpSlot0 = 0;                      // [world] barrier-0 arrivals, written by the peers
constexpr int kDpSlot1 = kDpMaxWorld;            // [world] barrier-1 arrivals
constexpr int kDpEpoch = 2 * kDpMaxWorld;        // completed launches on this flag block
constexpr int kDpGridA = kDpEpoch + 1;           // grid release after barrier 0
constexpr int kDpGridB = kDpEpoch + 2;           // grid release after barrier 1
constexpr int kDpCntA = kDpEpoch + 3;            // blocks that finished phase 1 (monotonic)
constexpr int kDpCntDone = kDpEpoch + 4;         // blocks that finished the kernel (monotonic)
constexpr int kDpFlagWords = 32;

struct DpArgs {
  const float* grads[kDpMaxWorld];               // rank q's gradient buffer (entry `rank` is local memory)
  float* red[kDpMaxWorld];                       // rank q's reduction scratch
  unsigned* flags[kDpMaxWorld];                  // rank q's flag block
  float* p; float* m; float* v;                  // local replica
  const float* hyper_a; const float* hyper_b;    // lr_t of the two segments (device scalars)
  size_t n4, n4_a;                               // float4 elements in total / in segment A
  int world, rank;
  float beta1, beta2, eps, inv_world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data: system-scope relaxed loads (never served from a stale non-coherent cache line)
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// cross-GPU barrier executed by block 0 (thread q talks to rank q), then the rest of the grid is released
__device__ __forceinline__ void dp_barrier(const DpArgs& A, int slot, int grid_flag, unsigned e) {
  unsigned* mine = A.flags[A.rank];
  if (blockIdx.x == 0) {
    const int q = threadIdx.x;
    if (q < A.world) {
      st_release_sys(A.flags[q] + slot + A.rank, e);
      while (ld_acquire_sys(mine + slot + q) < e) {}
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(mine + grid_flag, e);
  } else {
    if (threadIdx.x == 0) {
      while (ld_acquire_gpu(mine + grid_flag) < e) {}
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kDpThreads, 1) dp_allreduce_adam_kernel(const DpArgs A) {
  unsigned* mine = A.flags[A.rank];
  __shared__ unsigned e_s;
  if (threadIdx.x == 0) e_s = ld_acquire_gpu(mine + kDpEpoch) + 1u;   // the epoch word only moves when the LAST block of a launch retires
  __syncthreads();
  const unsigned e = e_s;
  const size_t gtid = (size_t)blockIdx.x * kDpThreads + threadIdx.x, gstride = (size_t)gridDim.x * kDpThreads;

  dp_barrier(A, kDpSlot0, kDpGridA, e);

  // ---- phase 1: reduce-scatter of my slice, rank order ----
  // A peer load is an NVLink round trip (~2 us) and the SM issues in order: a load whose value is consumed right
  // away costs the whole trip.  So every thread first ISSUES its `world` loads of up to 2 elements, then adds them
  // (in rank order: the result does not depend on the batching).
  const size_t chunk = A.n4 / A.world;                       // the host pads n to a multiple of 4 * world
  float* red_mine = A.red[A.rank];
  {
    const size_t lo = (size_t)A.rank * chunk, hi = lo + chunk;
    for (size_t i = lo + gtid; i < hi; i += 2 * gstride) {
      const size_t i1 = i + gstride;
      const bool two = i1 < hi;
      float4 t0[kDpMaxWorld], t1[kDpMaxWorld];
#pragma unroll
      for (int q = 0; q < kDpMaxWorld; ++q)
        if (q < A.world) {
          t0[q] = ld_sys_v4(A.grads[q] + 4 * i);
          if (two) t1[q] = ld_sys_v4(A.grads[q] + 4 * i1);
        }
      float4 s0 = t0[0], s1 = two ? t1[0] : t0[0];
#pragma unroll
      for (int q = 1; q < kDpMaxWorld; ++q)
        if (q < A.world) {
          s0.x += t0[q].x; s0.y += t0[q].y; s0.z += t0[q].z; s0.w += t0[q].w;
          if (two) { s1.x += t1[q].x; s1.y += t1[q].y; s1.z += t1[q].z; s1.w += t1[q].w; }
        }
      *reinterpret_cast<float4*>(red_mine + 4 * i) = s0;
      if (two) *reinterpret_cast<float4*>(red_mine + 4 * i1) = s1;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      atomicAdd(mine + kDpCntA, 1u);
      while (ld_acquire_gpu(mine + kDpCntA) < e * gridDim.x) {}     // every block of this GPU has published its part
    }
    __syncthreads();
  } else if (threadIdx.x == 0) {
    atomicAdd(mine + kDpCntA, 1u);
  }
  dp_barrier(A, kDpSlot1, kDpGridB, e);

  // ---- phase 2: all-gather fused with Adam ----
  const float lr_a = A.hyper_a[0], lr_b = A.hyper_b ? A.hyper_b[0] : lr_a;
  const float b1 = A.beta1, b2 = A.beta2;
  // 4 elements per round: the 4 peer loads (and the 12 local ones) are in flight together
  for (size_t i0 = gtid; i0 < A.n4; i0 += 4 * gstride) {
    float4 g[4], pm[4], pv[4], pp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + (size_t)u * gstride;
      if (i < A.n4) g[u] = ld_sys_v4(A.red[(int)(i / chunk)] + 4 * i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + (size_t)u * gstride;
      if (i < A.n4) {
        pm[u] = *reinterpret_cast<const float4*>(A.m + 4 * i);
        pv[u] = *reinterpret_cast<const float4*>(A.v + 4 * i);
        pp[u] = *reinterpret_cast<const float4*>(A.p + 4 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t i = i0 + (size_t)u * gstride;
      if (i >= A.n4) continue;
      const float lr_t = i < A.n4_a ? lr_a : lr_b;
#define EXP_ADAM1(c)                                                    \
      {                                                                 \
        const float gi = g[u].c * A.inv_world;                          \
        const float mi = b1 * pm[u].c + (1.f - b1) * gi;                \
        const float vi = b2 * pv[u].c + (1.f - b2) * gi * gi;           \
        pm[u].c = mi; pv[u].c = vi;                                     \
        pp[u].c -= lr_t * mi / (sqrtf(vi) + A.eps);                     \
      }
      EXP_ADAM1(x) EXP_ADAM1(y) EXP_ADAM1(z) EXP_ADAM1(w)
#undef EXP_ADAM1
      *reinterpret_cast<float4*>(A.m + 4 * i) = pm[u];
      *reinterpret_cast<float4*>(A.v + 4 * i) = pv[u];
      *reinterpret_cast<float4*>(A.p + 4 * i) = pp[u];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(mine + kDpCntDone, 1u);
    if (t == e * gridDim.x - 1u) st_release_gpu(mine + kDpEpoch, e);    // last block of this launch: the next launch sees epoch e
  }
}

}  // namespace expo

using namespace expo;

extern "C" {

// ---- CUDA IPC plumbing: map a peer process's buffer into THIS device's address space ----------------------
// (cudaIpcMemLazyEnablePeerAccess: the mapping is opened in the current device's context and peer access to the
// owning GPU is enabled on first use -- no context is ever created on a foreign device.)
size_t exp_dp_ipc_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int exp_dp_ipc_export(const void* dev_ptr, void* handle_out, size_t* offset_out) {
  EXP_CHECK_ARG(dev_ptr && handle_out && offset_out, "null pointer");
  // the handle names the whole cudaMalloc allocation that contains dev_ptr (the caller's allocator may carve
  // many tensors out of one): report the offset of dev_ptr inside it
  typedef int (*GetRangeFn)(unsigned long long*, size_t*, unsigned long long);
  static GetRangeFn get_range = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<GetRangeFn>(f);
  }();
  if (!get_range) return set_error(EXP_ERR_UNSUPPORTED, "cuMemGetAddressRange not available");
  unsigned long long base = 0;
  size_t size = 0;
  if (get_range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
    return set_error(EXP_ERR_CUDA, "cuMemGetAddressRange failed (not a device allocation?)");
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (size_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return EXP_OK;
}

int exp_dp_ipc_open(const void* handle, void** base_out) {
  EXP_CHECK_ARG(handle && base_out, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return EXP_OK;
}

int exp_dp_ipc_close(void* base) {
  EXP_CHECK_ARG(base, "null pointer");
  const cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  return EXP_OK;
}

size_t exp_dp_flag_bytes(void) { return kDpFlagWords * sizeof(unsigned); }
int exp_dp_max_world(void) { return kDpMaxWorld; }

int exp_dp_allreduce_adam(float* params, float* m, float* v, const float* const* grads_host, float* const* red_host,
                          unsigned* const* flags_host, int world, int rank, const float* hyper_a, size_t n_a,
                          const float* hyper_b, size_t n, float beta1, float beta2, float eps, void* stream) {
  EXP_CHECK_ARG(params && m && v && grads_host && red_host && flags_host && hyper_a, "null pointer");
  EXP_CHECK_ARG(world >= 2 && world <= kDpMaxWorld && rank >= 0 && rank < world, "world must be in [2, %d] (got %d), rank %d", kDpMaxWorld, world, rank);
  EXP_CHECK_ARG(n > 0 && n % (4 * (size_t)world) == 0, "n = %zu must be a multiple of 4 * world", n);
  EXP_CHECK_ARG(n_a <= n && n_a % 4 == 0 && (n_a == n || hyper_b), "segment A must be a multiple of 4 floats; segment B needs hyper_b");
  DpArgs A{};
  for (int q = 0; q < world; ++q) {
    EXP_CHECK_ARG(grads_host[q] && red_host[q] && flags_host[q], "null buffer of rank %d", q);
    if (!aligned16(grads_host[q]) || !aligned16(red_host[q])) return set_error(EXP_ERR_ALIGNMENT, "buffers of rank %d must be 16-byte aligned", q);
    A.grads[q] = grads_host[q]; A.red[q] = red_host[q]; A.flags[q] = flags_host[q];
  }
  if (!aligned16(params) || !aligned16(m) || !aligned16(v)) return set_error(EXP_ERR_ALIGNMENT, "params / m / v must be 16-byte aligned");
  A.p = params; A.m = m; A.v = v; A.hyper_a = hyper_a; A.hyper_b = hyper_b;
  A.n4 = n / 4; A.n4_a = n_a / 4; A.world = world; A.rank = rank;
  A.beta1 = beta1; A.beta2 = beta2; A.eps = eps; A.inv_world = 1.0f / (float)world;
  // one block per SM: the spin barriers need the whole grid resident; the grid size is part of the protocol
  // (arrival counters advance by gridDim.x per launch), so it is the same for every launch on a flag block
  static int sms[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return set_error(EXP_ERR_CUDA, "cudaGetDevice failed");
  if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
  dp_allreduce_adam_kernel<<<dim3(sms[dev] > 0 ? sms[dev] : 148), dim3(kDpThreads), 0, (cudaStream_t)stream>>>(A);
  EXP_CHECK_LAUNCH("exp_dp_allreduce_adam");
  return EXP_OK;
}

}  // extern "C"
