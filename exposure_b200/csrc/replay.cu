// Device-side replay memory: the selection logic of replay_memory.py:187-273 as kernels, so that a
// whole train iteration -- draw the generator batch, train, re-insert the outputs, draw five critic batches, train --
// is device work only: no host list handling, no index upload, capturable as ONE CUDA graph.
// The logic itself is replay_logic.cuh (shared with the host harness); the pool has 128 records, so one thread does
// it in a few microseconds and the images / states follow by gathers over the index lists written here.
// Also here: the per-step random draws of the train step (z[:,0], the two tf.nn.dropout masks, alpha) as one
// counter-based kernel, so that a replayed graph draws fresh numbers.
#include "common.cuh"
#include "replay_logic.cuh"

namespace expo {

// control block (int32 words) kept in device memory by the caller
constexpr int kRlCallLo = 0, kRlCallHi = 1;      // 64-bit call counter of the random streams
constexpr int kRlNRest = 2, kRlFreshUsed = 3;    // between draw_generator and replace
constexpr int kRlError = 4;                      // != 0: a critic batch was requested while no record had terminated
constexpr int kRlLastTerm = 5;                   // terminated records seen by the last critic draw (diagnostics)
constexpr int kRlCtlWords = 8;
constexpr int kRlMaxPool = 1024;

__device__ __forceinline__ uint64_t rl_next_call(int* ctl) {
  const uint64_t c = ((uint64_t)(uint32_t)ctl[kRlCallHi] << 32) | (uint32_t)ctl[kRlCallLo];
  const uint64_t n = c + 1;
  ctl[kRlCallLo] = (int)(uint32_t)n;
  ctl[kRlCallHi] = (int)(uint32_t)(n >> 32);
  return c;
}

// random order of the pool: one thread per record (keys, then ranks); returns this launch's call number
__device__ __forceinline__ uint64_t block_shuffle(int* perm, uint32_t* keys, int P, uint64_t seed, int* ctl, uint32_t stream) {
  __shared__ uint64_t call_s;
  if (threadIdx.x == 0) call_s = rl_next_call(ctl);
  __syncthreads();
  const uint64_t call = call_s;
  for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = rl::shuffle_key(seed, call, stream, i);
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) perm[rl::shuffle_rank(keys, P, i)] = i;
  __syncthreads();
  return call;
}

__global__ void replay_draw_generator_kernel(const float* __restrict__ pool_states, int S, int P, int B, uint64_t seed,
                                             int* __restrict__ ctl, long long* __restrict__ batch_src, int* __restrict__ rest_src) {
  EXP_PDL_ENTRY();
  __shared__ int perm[kRlMaxPool];
  __shared__ uint32_t keys[kRlMaxPool];
  __shared__ float stopped[kRlMaxPool];
  __shared__ int rest_s[kRlMaxPool];
  for (int i = threadIdx.x; i < P; i += blockDim.x) stopped[i] = pool_states[(size_t)i * S + rl::kStateStopped];
  block_shuffle(perm, keys, P, seed, ctl, 1u);
  __shared__ int n_rest_s;
  __shared__ long long batch_s[kRlMaxPool];
  if (threadIdx.x == 0) {
    int n_rest, fresh_used;
    rl::draw_generator(stopped, P, B, perm, batch_s, rest_s, &n_rest, &fresh_used);
    ctl[kRlNRest] = n_rest;
    ctl[kRlFreshUsed] = fresh_used;
    n_rest_s = n_rest;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) batch_src[i] = batch_s[i];
  for (int i = threadIdx.x; i < n_rest_s; i += blockDim.x) rest_src[i] = rest_s[i];
}

__global__ void replay_replace_kernel(const float* __restrict__ new_states, int S, int P, int B, int max_traj_len, float keep_prob,
                                      uint64_t seed, int* __restrict__ ctl, const int* __restrict__ rest_src,
                                      long long* __restrict__ new_pool_src) {
  EXP_PDL_ENTRY();
  __shared__ float step[kRlMaxPool];
  __shared__ int rest_s[kRlMaxPool];
  __shared__ long long pool_s[kRlMaxPool];
  const int n_rest = ctl[kRlNRest];
  for (int i = threadIdx.x; i < B; i += blockDim.x) step[i] = new_states[(size_t)i * S + rl::kStateStep];
  for (int i = threadIdx.x; i < n_rest; i += blockDim.x) rest_s[i] = rest_src[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    rl::Philox g;
    rl::philox_init(g, seed, rl_next_call(ctl), 2u);
    rl::replace(step, P, B, max_traj_len, keep_prob, g, rest_s, n_rest, ctl[kRlFreshUsed], pool_s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) new_pool_src[i] = pool_s[i];
}

__global__ void replay_draw_critic_kernel(const float* __restrict__ pool_states, int S, int P, int B, uint64_t seed,
                                          int* __restrict__ ctl, long long* __restrict__ batch_src) {
  EXP_PDL_ENTRY();
  __shared__ int perm[kRlMaxPool];
  __shared__ int term[kRlMaxPool];
  __shared__ uint32_t keys[kRlMaxPool];
  __shared__ float stopped[kRlMaxPool];
  __shared__ long long batch_s[kRlMaxPool];
  for (int i = threadIdx.x; i < P; i += blockDim.x) stopped[i] = pool_states[(size_t)i * S + rl::kStateStopped];
  block_shuffle(perm, keys, P, seed, ctl, 3u);
  if (threadIdx.x == 0) {
    const int nt = rl::draw_critic(stopped, P, B, perm, term, batch_s);
    ctl[kRlLastTerm] = nt;
    if (nt == 0) ctl[kRlError] = 1;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B; i += blockDim.x) batch_src[i] = batch_s[i];
}

// out[i] for i < n_uniform: U[0,1);  then n_mask dropout multipliers floor(keep + U) / keep (tf.nn.dropout, agent.py:36)
__global__ void train_draws_kernel(uint64_t seed, int* __restrict__ ctl, float* __restrict__ uniform, int n_uniform,
                                   float* __restrict__ mask, size_t n_mask, float keep) {
  EXP_PDL_ENTRY();
  __shared__ uint64_t call_s;
  // every block must see the SAME call number: block 0 cannot bump the counter before the others have read it, so the
  // counter is bumped by a separate one-thread launch (train_draws_bump_kernel) that follows this kernel on the stream
  if (threadIdx.x == 0) call_s = ((uint64_t)(uint32_t)ctl[kRlCallHi] << 32) | (uint32_t)ctl[kRlCallLo];
  __syncthreads();
  const size_t total4 = ((size_t)n_uniform + n_mask + 3) / 4;
  const float inv_keep = 1.0f / keep;
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total4; q += (size_t)gridDim.x * blockDim.x) {
    rl::Philox g;
    rl::philox_init(g, seed, call_s, 4u);
    g.ctr[0] = (uint32_t)q;
    g.ctr[1] ^= (uint32_t)(q >> 32) << 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float u = rl::philox_uniform(g);
      const size_t i = 4 * q + k;
      if (i < (size_t)n_uniform) uniform[i] = u;
      else if (i - n_uniform < n_mask) mask[i - n_uniform] = floorf(keep + u) * inv_keep;
    }
  }
}
__global__ void train_draws_bump_kernel(int* __restrict__ ctl) {
  EXP_PDL_ENTRY();
  if (threadIdx.x == 0) rl_next_call(ctl);
}

// dst[i] = src[idx[i]] for rows of `row` floats (row % 4 == 0): the image / state gathers behind the index lists
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ src, const long long* __restrict__ idx,
                                                          float4* __restrict__ dst, int row4) {
  EXP_PDL_ENTRY();
  const long long s = idx[blockIdx.y];
  const float4* a = src + (size_t)s * row4;
  float4* o = dst + (size_t)blockIdx.y * row4;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < row4; i += gridDim.x * 256) o[i] = __ldg(a + i);
}
__global__ void gather_small_kernel(const float* __restrict__ src, const long long* __restrict__ idx, float* __restrict__ dst,
                                    int row, int n) {
  EXP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * row) return;
  const int r = i / row, c = i - r * row;
  dst[i] = src[(size_t)idx[r] * row + c];
}

}  // namespace expo

using namespace expo;

extern "C" {

int exp_replay_ctl_words(void) { return kRlCtlWords; }

int exp_replay_draw_generator(const float* pool_states, int n_states, int pool, int batch, unsigned long long seed, int* ctl,
                              long long* batch_src, int* rest_src, void* stream) {
  EXP_CHECK_ARG(pool_states && ctl && batch_src && rest_src, "null pointer");
  EXP_CHECK_ARG(pool > 0 && pool <= kRlMaxPool && batch > 0 && batch <= pool && n_states > rl::kStateStep, "bad sizes (pool <= %d, batch <= pool)", kRlMaxPool);
  launch_pdl(replay_draw_generator_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, pool_states, n_states, pool, batch,
             (uint64_t)seed, ctl, batch_src, rest_src);
  EXP_CHECK_LAUNCH("exp_replay_draw_generator");
  return EXP_OK;
}

int exp_replay_replace(const float* new_states, int n_states, int pool, int batch, int max_traj_len, float keep_prob,
                       unsigned long long seed, int* ctl, const int* rest_src, long long* new_pool_src, void* stream) {
  EXP_CHECK_ARG(new_states && ctl && rest_src && new_pool_src, "null pointer");
  EXP_CHECK_ARG(pool > 0 && pool <= kRlMaxPool && batch > 0 && batch <= pool && n_states > rl::kStateStep, "bad sizes");
  launch_pdl(replay_replace_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, new_states, n_states, pool, batch, max_traj_len,
             keep_prob, (uint64_t)seed, ctl, rest_src, new_pool_src);
  EXP_CHECK_LAUNCH("exp_replay_replace");
  return EXP_OK;
}

int exp_replay_draw_critic(const float* pool_states, int n_states, int pool, int batch, unsigned long long seed, int* ctl,
                           long long* batch_src, void* stream) {
  EXP_CHECK_ARG(pool_states && ctl && batch_src, "null pointer");
  EXP_CHECK_ARG(pool > 0 && pool <= kRlMaxPool && batch > 0 && batch <= kRlMaxPool && n_states > rl::kStateStep, "bad sizes");
  launch_pdl(replay_draw_critic_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, pool_states, n_states, pool, batch,
             (uint64_t)seed, ctl, batch_src);
  EXP_CHECK_LAUNCH("exp_replay_draw_critic");
  return EXP_OK;
}

int exp_train_draws(unsigned long long seed, int* ctl, float* uniform, int n_uniform, float* mask, size_t n_mask, float keep,
                    void* stream) {
  EXP_CHECK_ARG(ctl && (uniform || n_uniform == 0) && (mask || n_mask == 0) && n_uniform >= 0, "bad args");
  EXP_CHECK_ARG(n_mask == 0 || (keep > 0.f && keep <= 1.f), "keep probability must be in (0, 1]");
  const size_t total4 = ((size_t)n_uniform + n_mask + 3) / 4;
  if (total4 == 0) return EXP_OK;
  unsigned blocks = (unsigned)((total4 + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  launch_pdl(train_draws_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (uint64_t)seed, ctl, uniform, n_uniform, mask,
             n_mask, keep);
  launch_pdl(train_draws_bump_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, ctl);
  EXP_CHECK_LAUNCH("exp_train_draws");
  return EXP_OK;
}

int exp_gather_rows(const float* src, const long long* idx, float* dst, int n, int row, void* stream) {
  EXP_CHECK_ARG(src && idx && dst && n > 0 && n <= 65535 && row > 0, "bad args");
  if (row % 4 == 0 && aligned16(src) && aligned16(dst)) {
    int gx = (row / 4 + 255) / 256;
    if (gx > 8) gx = 8;
    launch_pdl(gather_rows_kernel, dim3(gx, n), dim3(256), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(src), idx,
               reinterpret_cast<float4*>(dst), row / 4);
  } else {
    launch_pdl(gather_small_kernel, dim3((n * row + 255) / 256), dim3(256), 0, (cudaStream_t)stream, src, idx, dst, row, n);
  }
  EXP_CHECK_LAUNCH("exp_gather_rows");
  return EXP_OK;
}

}  // extern "C"
