// Small fused kernels of the train step that replace chains of framework element-wise / indexing launches
// (round 1 profile: ~200 torch glue launches and ~150 colsum / reduce launches per iteration, each 2-8 us of a
// 6.2 ms iteration).  Everything here is per-image or per-head bookkeeping around the conv / FC / filter kernels:
//
//   exp_critic_inputs     real | fake | real + alpha (fake - real) into ONE batch      net.py:174-179 (+ the concat)
//   exp_critic_scalars    emd, gradient penalty, c_loss, critic_gradient_norm, c_average and the zero-debiased
//                         moving average of net.py:119-120, 164-168, 185-187, 268-269
//   exp_heads_fc2_fwd     the 8 filter heads' fc2 layers as one launch                 filters.py:39-44 via agent.py:58-72
//   exp_heads_select      one-hot selection of the chosen head's outputs               agent.py:113-125
//   exp_heads_fc2_bwd     their backward (dgrad through lrelu, wgrad, bias grad), one launch
//   exp_colsum_multi      up to 8 bias-gradient column sums in one launch, deterministic last-block finish
//   exp_stats_bwd_gin     image gradient of the critic input statistics straight from the layer-1 input gradient
//                         (critics.py:48-87 through tf.gradients): no channel slicing / summing launches
#include "common.cuh"

namespace expo {

constexpr int kGlueThreads = 256;
constexpr float kLR2 = 0.27f, kLG2 = 0.67f, kLB2 = 0.06f;

__device__ __forceinline__ float lrelu_g(float v) { return 0.6f * v + 0.4f * fabsf(v); }
__device__ __forceinline__ float dlrelu_g(float a) { return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f); }

template <int THREADS>
__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < THREADS / 32; ++w) t += sh[w];
  return t;
}

// ---- X[0:B] = real, X[B:2B] = fake, X[2B:3B] = real + alpha[b] (fake - real) ------------------------------
__global__ void __launch_bounds__(kGlueThreads) critic_inputs_kernel(const float4* __restrict__ real, const float4* __restrict__ fake,
                                                                     const float* __restrict__ alpha, float4* __restrict__ X,
                                                                     int B, int n4) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.y;
  const float a = alpha[b];
  const size_t base = (size_t)b * n4, total = (size_t)B * n4;
  for (int i = blockIdx.x * kGlueThreads + threadIdx.x; i < n4; i += gridDim.x * kGlueThreads) {
    const float4 r = __ldg(real + base + i), f = __ldg(fake + base + i);
    X[base + i] = r;
    X[total + base + i] = f;
    float4 o;                                    // net.py:177-179: differences = fake - real; real + alpha * differences
    o.x = r.x + a * (f.x - r.x); o.y = r.y + a * (f.y - r.y); o.z = r.z + a * (f.z - r.z); o.w = r.w + a * (f.w - r.w);
    X[2 * total + base + i] = o;
  }
}

// ---- logging scalars of the critic step + moving average (one block) ---------------------------------------
// out[0] emd = mean(real_logit) - mean(fake_logit)   out[1] gradient penalty   out[2] mean ||grad||
// out[3] c_loss = -emd + gp                           out[4] c_average = (mean real + mean fake) / 2
// ema (nullable) = [debiased value, biased accumulator, local_step]: moving_averages._zero_debias
__global__ void __launch_bounds__(kGlueThreads) critic_scalars_kernel(const float* __restrict__ logits, const float* __restrict__ norm,
                                                                      int B, float lambda, float* __restrict__ ema, float decay,
                                                                      float* __restrict__ out) {
  EXP_PDL_ENTRY();
  __shared__ double sh[kGlueThreads / 32];
  double sr = 0.0, sf = 0.0, sn = 0.0, sp = 0.0;
  for (int i = threadIdx.x; i < B; i += kGlueThreads) {
    sr += (double)logits[i];
    sf += (double)logits[B + i];
    const float nv = norm[i];
    sn += (double)nv;
    const float ex = fmaxf(nv - 1.0f, 0.f);
    sp += (double)(ex * ex);
  }
  sr = block_sum_d<kGlueThreads>(sr, sh);
  sf = block_sum_d<kGlueThreads>(sf, sh);
  sn = block_sum_d<kGlueThreads>(sn, sh);
  sp = block_sum_d<kGlueThreads>(sp, sh);
  if (threadIdx.x == 0) {
    const float mr = (float)(sr / B), mf = (float)(sf / B);
    const float emd = mr - mf, gp = lambda * (float)(sp / B), cav = (mr + mf) * 0.5f;
    out[0] = emd; out[1] = gp; out[2] = (float)(sn / B); out[3] = -emd + gp; out[4] = cav;
    if (ema) {
      const float biased = ema[1] * decay + cav * (1.0f - decay);
      const float step = ema[2] + 1.0f;
      ema[1] = biased;
      ema[2] = step;
      ema[0] = biased / (1.0f - powf(decay, step));
    }
  }
}

// ---- the 8 filter heads' fc2 layers ------------------------------------------------------------------------
constexpr int kMaxHeads = 8;
struct HeadsArgs {
  const float* params;             // flat parameter buffer of the generator
  float* grads;                    // flat gradient buffer (backward)
  int w_off[kMaxHeads], b_off[kMaxHeads];   // offsets of fc2 weights [fc1, dim_j] / biases [dim_j] inside it
  int dim[kMaxHeads];              // fc2 width of head j = n_j filter parameters + mask parameters
  int npar[kMaxHeads];             // n_j
  int n_heads, fc1, ostride, B, ldh, nmask;
};

// O[b, j, o] = sum_k H[b, j*fc1 + k] W_j[k, o] + bias_j[o]  (o < dim_j), 0 above: every entry of O is written
__global__ void __launch_bounds__(kGlueThreads) heads_fc2_fwd_kernel(const HeadsArgs A, const float* __restrict__ H, float* __restrict__ O) {
  EXP_PDL_ENTRY();
  extern __shared__ float hrow[];                          // [n_heads * fc1] activations of image b
  const int b = blockIdx.x;
  const int width = A.n_heads * A.fc1;
  for (int i = threadIdx.x; i < width; i += kGlueThreads) hrow[i] = H[(size_t)b * A.ldh + i];
  __syncthreads();
  for (int i = threadIdx.x; i < A.n_heads * A.ostride; i += kGlueThreads) {
    const int j = i / A.ostride, o = i - j * A.ostride;
    float acc = 0.f;
    if (o < A.dim[j]) {
      const float* W = A.params + A.w_off[j] + o;
      const float* h = hrow + j * A.fc1;
      acc = A.params[A.b_off[j] + o];
      for (int k = 0; k < A.fc1; ++k) acc = fmaf(h[k], __ldg(W + (size_t)k * A.dim[j]), acc);
    }
    O[((size_t)b * A.n_heads + j) * A.ostride + o] = acc;
  }
}

// sel[b, 0:24] = O[b, id, 0:n_id] (0 above n_id), msel[b, 0:nmask] = O[b, id, n_id : n_id + nmask]; id -1 -> zeros
__global__ void heads_select_kernel(const HeadsArgs A, const float* __restrict__ O, const int* __restrict__ ids,
                                    float* __restrict__ sel, int selstride, float* __restrict__ msel) {
  EXP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = selstride + A.nmask;
  if (i >= A.B * per) return;
  const int b = i / per, o = i - b * per;
  const int id = ids[b];
  const bool ok = id >= 0 && id < A.n_heads;
  const float* row = O + ((size_t)b * A.n_heads + (ok ? id : 0)) * A.ostride;
  if (o < selstride) sel[(size_t)b * selstride + o] = (ok && o < A.npar[id]) ? row[o] : 0.f;
  else if (msel) msel[(size_t)b * A.nmask + (o - selstride)] = ok ? row[A.npar[id] + (o - selstride)] : 0.f;
}

// backward of both: dy_j[b, o] = [ids[b] == j] * (o < n_j ? gsel[b, o] : gmsel[b, o - n_j])
//   blocks [0, B):            dH[b, j*fc1 + k] = lrelu'(H) * sum_o dy_j[b, o] W_j[k, o]       (all heads; zero rows for j != id)
//   blocks [B, B + n_heads):  gW_j[k, o] = sum_b H[b, j*fc1 + k] dy_j[b, o],  gb_j[o] = sum_b dy_j[b, o]   (batch order: deterministic)
__global__ void __launch_bounds__(kGlueThreads) heads_fc2_bwd_kernel(const HeadsArgs A, const float* __restrict__ H,
                                                                     const int* __restrict__ ids, const float* __restrict__ gsel,
                                                                     int selstride, const float* __restrict__ gmsel,
                                                                     float* __restrict__ dH) {
  EXP_PDL_ENTRY();
  __shared__ float dy[64];
  if ((int)blockIdx.x < A.B) {
    const int b = blockIdx.x;
    const int id = ids[b];
    const bool ok = id >= 0 && id < A.n_heads;
    const int n = ok ? A.npar[id] : 0, dim = ok ? A.dim[id] : 0;
    if ((int)threadIdx.x < dim)
      dy[threadIdx.x] = (int)threadIdx.x < n ? gsel[(size_t)b * selstride + threadIdx.x]
                                             : (gmsel ? gmsel[(size_t)b * A.nmask + (threadIdx.x - n)] : 0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < A.n_heads * A.fc1; i += kGlueThreads) {
      const int j = i / A.fc1, k = i - j * A.fc1;
      float acc = 0.f;
      if (ok && j == id) {
        const float* W = A.params + A.w_off[j] + (size_t)k * dim;
        for (int o = 0; o < dim; ++o) acc = fmaf(dy[o], __ldg(W + o), acc);
        acc *= dlrelu_g(H[(size_t)b * A.ldh + i]);
      }
      dH[(size_t)b * A.ldh + i] = acc;
    }
    return;
  }
  const int j = blockIdx.x - A.B;
  const int dim = A.dim[j], n = A.npar[j];
  float* gW = A.grads + A.w_off[j];
  float* gb = A.grads + A.b_off[j];
  for (int i = threadIdx.x; i < (A.fc1 + 1) * dim; i += kGlueThreads) {
    const int k = i / dim, o = i - k * dim;               // row fc1 is the bias
    float acc = 0.f;
    for (int b = 0; b < A.B; ++b) {
      if (ids[b] != j) continue;
      const float d = o < n ? gsel[(size_t)b * selstride + o] : (gmsel ? gmsel[(size_t)b * A.nmask + (o - n)] : 0.f);
      acc = fmaf(k < A.fc1 ? H[(size_t)b * A.ldh + j * A.fc1 + k] : 1.0f, d, acc);
    }
    if (k < A.fc1) gW[(size_t)k * dim + o] = acc;
    else gb[o] = acc;
  }
}

// ---- several column sums in one launch ----------------------------------------------------------------------
constexpr int kMaxColsumTasks = 8;
constexpr int kColsumChunk = 256;                 // rows per block
// The ticket counters occupy a FIXED block at the start of the workspace, whatever the task list: a later launch
// with more column blocks must never read an earlier launch's partial sums as tickets.
constexpr int kColsumMaxCounters = 1024;
constexpr size_t kColsumCounterBytes = kColsumMaxCounters * sizeof(unsigned);
struct ColsumTasks {
  const float* src[kMaxColsumTasks];
  float* dst[kMaxColsumTasks];
  int rows[kMaxColsumTasks], cols[kMaxColsumTasks], accumulate[kMaxColsumTasks];
  int blk0[kMaxColsumTasks + 1];                  // first block of task t (blocks = colblocks * chunks)
  int part0[kMaxColsumTasks + 1];                 // first partial row of task t
  int cnt0[kMaxColsumTasks + 1];                  // first ticket counter of task t (one per column block)
  int n;
  float* partials;                                // [sum chunks_t][cols_t] laid out task after task
  unsigned* counters;
};

__global__ void __launch_bounds__(kGlueThreads) colsum_multi_kernel(const ColsumTasks T) {
  EXP_PDL_ENTRY();
  __shared__ float red[8][33];
  __shared__ unsigned ticket;
  int t = 0;
  while (t + 1 < T.n && (int)blockIdx.x >= T.blk0[t + 1]) ++t;
  const int rows = T.rows[t], cols = T.cols[t];
  const int chunks = (rows + kColsumChunk - 1) / kColsumChunk;
  const int local = blockIdx.x - T.blk0[t];
  const int cb = local / chunks, chunk = local - cb * chunks;
  const int col = cb * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
  const int r0 = chunk * kColsumChunk, r1 = min(rows, r0 + kColsumChunk);
  const float* a = T.src[t];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < cols) {
    int r = r0 + ry;
    for (; r + 56 < r1; r += 64) {                          // 8 loads in flight per thread
      const float* q = a + (size_t)r * cols + col;
      const float t0 = q[0], t1 = q[(size_t)8 * cols], t2 = q[(size_t)16 * cols], t3 = q[(size_t)24 * cols];
      const float t4 = q[(size_t)32 * cols], t5 = q[(size_t)40 * cols], t6 = q[(size_t)48 * cols], t7 = q[(size_t)56 * cols];
      s0 += t0; s1 += t1; s2 += t2; s3 += t3;
      s0 += t4; s1 += t5; s2 += t6; s3 += t7;
    }
    for (; r < r1; r += 8) s0 += a[(size_t)r * cols + col];
  }
  red[ry][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  float* part = T.partials + (size_t)T.part0[t];         // task t's partial rows start at element offset part0[t]
  if (ry == 0 && col < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    part[(size_t)chunk * cols + col] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(T.counters + T.cnt0[t] + cb, 1u);
  __syncthreads();
  if (ticket != (unsigned)(chunks - 1)) return;
  __threadfence();
  // last block of this column block: row-lane ry adds chunks ry, ry + 8, ... (4 loads in flight), then the 8 lanes
  // are added in lane order -- a fixed order for a given shape, so the result is reproducible
  float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
  if (col < cols) {
    int c = ry;
    for (; c + 24 < chunks; c += 32) {
      v0 += __ldcg(part + (size_t)c * cols + col);
      v1 += __ldcg(part + (size_t)(c + 8) * cols + col);
      v2 += __ldcg(part + (size_t)(c + 16) * cols + col);
      v3 += __ldcg(part + (size_t)(c + 24) * cols + col);
    }
    for (; c < chunks; c += 8) v0 += __ldcg(part + (size_t)c * cols + col);
  }
  __syncthreads();
  red[ry][threadIdx.x & 31] = (v0 + v1) + (v2 + v3);
  __syncthreads();
  if (ry == 0 && col < cols) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    float* o = T.dst[t] + col;
    *o = T.accumulate[t] ? *o + v : v;
  }
  if (threadIdx.x == 0) T.counters[T.cnt0[t] + cb] = 0u;   // ready for the next launch
}

// ---- image gradient of the critic statistics from the layer-1 input gradient ---------------------------------
// g_in [B, P, cin]: channels [0,3) = image, the last three = tiled (mean lum, var lum, mean sat) channels.
// g_out[b, p, c] = g_in[b, p, c] + (J_stats^T g_stat[b])[p, c],  g_stat[b] = sum_p g_in[b, p, cin-3 .. cin)
__device__ __forceinline__ float sat_w(const float* p, float* w) {
  const float c0 = fminf(fmaxf(p[0], 0.f), 1.f), c1 = fminf(fmaxf(p[1], 0.f), 1.f), c2 = fminf(fmaxf(p[2], 0.f), 1.f);
  const float mx = fmaxf(c0, fmaxf(c1, c2)), mn = fminf(c0, fminf(c1, c2));
  const float s = mx + mn, t = 2.0f - mx - mn;
  const bool d_is_s = s <= t;
  const float den = (d_is_s ? s : t) + 1e-2f;
  const float r = mx - mn;
  const float dd = d_is_s ? 1.f : -1.f;
  const float common = r / (den * den) * dd;
  const float dmx = 1.f / den - common, dmn = -1.f / den - common;
  const float c[3] = {c0, c1, c2};
  int nmx = 0, nmn = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { nmx += c[i] == mx; nmn += c[i] == mn; }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const bool pass = p[i] >= 0.f && p[i] <= 1.f;
    w[i] = pass ? ((c[i] == mx ? dmx / nmx : 0.f) + (c[i] == mn ? dmn / nmn : 0.f)) : 0.f;
  }
  return r / den;
}

// grid (kGinCluster, B), one cluster per image: CTA r sums and then writes the r-th slice of the pixels
constexpr unsigned kGinCluster = 8;
__global__ void __launch_bounds__(kGlueThreads) stats_bwd_gin_kernel(const float* __restrict__ img, const float* __restrict__ stats,
                                                                     const float* __restrict__ g_in, int cin,
                                                                     float* __restrict__ g_out, int P) {
  EXP_PDL_ENTRY();
  __shared__ double sh[kGlueThreads / 32 * 3], xch[3];
  const int b = blockIdx.y;
  const unsigned r = cluster_rank_x(), n = cluster_size_x();
  const int lo = (int)((long long)P * r / n), hi = (int)((long long)P * (r + 1) / n);
  const float* gi = g_in + (size_t)b * P * cin;
  double a[3] = {0.0, 0.0, 0.0};
  for (int p = lo + threadIdx.x; p < hi; p += kGlueThreads) {
    const float* q = gi + (size_t)p * cin + (cin - 3);
    a[0] += (double)q[0]; a[1] += (double)q[1]; a[2] += (double)q[2];
  }
  cluster_sum<kGlueThreads, 3>(a, sh, xch);
  const float mean = stats[b * 3], gm = (float)a[0], gv = (float)a[1], gs = (float)a[2];
  const float invP = 1.0f / (float)P;
  for (int p = lo + threadIdx.x; p < hi; p += kGlueThreads) {
    const size_t o = ((size_t)b * P + p) * 3;
    float w[3];
    sat_w(img + o, w);
    const float lum = img[o] * kLR2 + img[o + 1] * kLG2 + img[o + 2] * kLB2 + 1e-5f;
    const float gl = (gm + gv * 2.f * (lum - mean)) * invP;
    const float coef[3] = {kLR2, kLG2, kLB2};
#pragma unroll
    for (int c = 0; c < 3; ++c) g_out[o + c] = gi[(size_t)p * cin + c] + coef[c] * gl + gs * invP * w[c];
  }
}

// dst[0..n) = the n <= 16 values passed BY VALUE: one launch instead of n scalar fills, no host buffer to keep alive
struct FloatVals { float v[16]; };
__global__ void set_floats_kernel(float* __restrict__ dst, const FloatVals V, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = V.v[threadIdx.x];
}

}  // namespace expo

using namespace expo;

extern "C" {

int exp_set_floats(float* dst, const float* host_vals, int n, void* stream) {
  EXP_CHECK_ARG(dst && host_vals && n > 0 && n <= 16, "1..16 values (got %d)", n);
  FloatVals V{};
  for (int i = 0; i < n; ++i) V.v[i] = host_vals[i];
  set_floats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(dst, V, n);
  EXP_CHECK_LAUNCH("exp_set_floats");
  return EXP_OK;
}

int exp_critic_inputs(const float* real, const float* fake, const float* alpha, float* X, int B, int n, void* stream) {
  EXP_CHECK_ARG(real && fake && alpha && X && B > 0 && B <= 65535 && n > 0 && n % 4 == 0, "bad args (n must be a multiple of 4)");
  if (!aligned16(real) || !aligned16(fake) || !aligned16(X)) return set_error(EXP_ERR_ALIGNMENT, "images must be 16-byte aligned");
  const int n4 = n / 4;
  int gx = (n4 + kGlueThreads - 1) / kGlueThreads;
  if (gx > 16) gx = 16;
  launch_pdl(critic_inputs_kernel, dim3(gx, B), dim3(kGlueThreads), 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(real),
             reinterpret_cast<const float4*>(fake), alpha, reinterpret_cast<float4*>(X), B, n4);
  EXP_CHECK_LAUNCH("exp_critic_inputs");
  return EXP_OK;
}

int exp_critic_scalars(const float* logits, const float* norm, int B, float lambda, float* ema_state, float decay, float* out,
                       void* stream) {
  EXP_CHECK_ARG(logits && norm && out && B > 0, "bad args");
  launch_pdl(critic_scalars_kernel, dim3(1), dim3(kGlueThreads), 0, (cudaStream_t)stream, logits, norm, B, lambda, ema_state, decay, out);
  EXP_CHECK_LAUNCH("exp_critic_scalars");
  return EXP_OK;
}

static int fill_heads(HeadsArgs& A, const float* params, float* grads, const int* w_off, const int* b_off, const int* dims,
                      const int* npar, int n_heads, int fc1, int ostride, int B, int ldh, int nmask) {
  EXP_CHECK_ARG(params && w_off && b_off && dims && npar, "null pointer");
  EXP_CHECK_ARG(n_heads > 0 && n_heads <= kMaxHeads && fc1 > 0 && B > 0 && ldh >= n_heads * fc1, "bad head geometry");
  A.params = params; A.grads = grads; A.n_heads = n_heads; A.fc1 = fc1; A.ostride = ostride; A.B = B; A.ldh = ldh; A.nmask = nmask;
  for (int j = 0; j < n_heads; ++j) {
    EXP_CHECK_ARG(dims[j] > 0 && dims[j] <= ostride && dims[j] <= 64 && npar[j] >= 0 && npar[j] + nmask <= dims[j], "bad dims of head %d", j);
    A.w_off[j] = w_off[j]; A.b_off[j] = b_off[j]; A.dim[j] = dims[j]; A.npar[j] = npar[j];
  }
  return EXP_OK;
}

int exp_heads_fc2_fwd(const float* params, const int* w_off_host, const int* b_off_host, const int* dims_host, const int* npar_host,
                      int n_heads, int fc1, int nmask, const float* H, int ldh, float* O, int ostride, int B, void* stream) {
  EXP_CHECK_ARG(H && O, "null pointer");
  HeadsArgs A{};
  const int rc = fill_heads(A, params, nullptr, w_off_host, b_off_host, dims_host, npar_host, n_heads, fc1, ostride, B, ldh, nmask);
  if (rc) return rc;
  launch_pdl(heads_fc2_fwd_kernel, dim3(B), dim3(kGlueThreads), (size_t)n_heads * fc1 * sizeof(float), (cudaStream_t)stream, A, H, O);
  EXP_CHECK_LAUNCH("exp_heads_fc2_fwd");
  return EXP_OK;
}

int exp_heads_select(const float* O, int ostride, const int* ids, const int* npar_host, int n_heads, int nmask, float* sel,
                     int selstride, float* msel, int B, void* stream) {
  EXP_CHECK_ARG(O && ids && npar_host && sel && B > 0 && n_heads > 0 && n_heads <= kMaxHeads && selstride > 0, "bad args");
  HeadsArgs A{};
  A.n_heads = n_heads; A.ostride = ostride; A.B = B; A.nmask = nmask;
  for (int j = 0; j < n_heads; ++j) {
    EXP_CHECK_ARG(npar_host[j] >= 0 && npar_host[j] <= selstride && npar_host[j] + nmask <= ostride, "bad npar of head %d", j);
    A.npar[j] = npar_host[j];
  }
  const int total = B * (selstride + nmask);
  launch_pdl(heads_select_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream, A, O, ids, sel, selstride, msel);
  EXP_CHECK_LAUNCH("exp_heads_select");
  return EXP_OK;
}

int exp_heads_fc2_bwd(const float* params, float* grads, const int* w_off_host, const int* b_off_host, const int* dims_host,
                      const int* npar_host, int n_heads, int fc1, int nmask, const float* H, int ldh, const int* ids,
                      const float* gsel, int selstride, const float* gmsel, float* dH, int B, void* stream) {
  EXP_CHECK_ARG(grads && H && ids && gsel && dH, "null pointer");
  HeadsArgs A{};
  const int rc = fill_heads(A, params, grads, w_off_host, b_off_host, dims_host, npar_host, n_heads, fc1, 64, B, ldh, nmask);
  if (rc) return rc;
  launch_pdl(heads_fc2_bwd_kernel, dim3(B + n_heads), dim3(kGlueThreads), 0, (cudaStream_t)stream, A, H, ids, gsel, selstride, gmsel, dH);
  EXP_CHECK_LAUNCH("exp_heads_fc2_bwd");
  return EXP_OK;
}

size_t exp_colsum_multi_workspace_bytes(const int* rows_host, const int* cols_host, int n) {
  if (!rows_host || !cols_host || n <= 0 || n > kMaxColsumTasks) return 0;
  size_t floats = 0, counters = 0;
  for (int t = 0; t < n; ++t) {
    if (rows_host[t] <= 0 || cols_host[t] <= 0) return 0;
    floats += (size_t)((rows_host[t] + kColsumChunk - 1) / kColsumChunk) * cols_host[t];
    counters += (size_t)(cols_host[t] + 31) / 32;
  }
  if (counters > (size_t)kColsumMaxCounters) return 0;
  return kColsumCounterBytes + floats * sizeof(float);
}

int exp_colsum_multi(const float* const* src_host, float* const* dst_host, const int* rows_host, const int* cols_host,
                     const int* accumulate_host, int n, void* workspace, size_t workspace_bytes, void* stream) {
  EXP_CHECK_ARG(src_host && dst_host && rows_host && cols_host && workspace, "null pointer");
  EXP_CHECK_ARG(n > 0 && n <= kMaxColsumTasks, "1..%d tasks per launch (got %d)", kMaxColsumTasks, n);
  const size_t need = exp_colsum_multi_workspace_bytes(rows_host, cols_host, n);
  if (need == 0) return set_error(EXP_ERR_INVALID_ARG, "bad rows / cols (at most %d column blocks of 32 per launch)", kColsumMaxCounters);
  if (workspace_bytes < need) return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (!aligned16(workspace)) return set_error(EXP_ERR_ALIGNMENT, "workspace must be 16-byte aligned");
  ColsumTasks T{};
  T.n = n;
  int blk = 0, part = 0, cnt = 0;
  for (int t = 0; t < n; ++t) {
    EXP_CHECK_ARG(src_host[t] && dst_host[t], "null task pointer");
    T.src[t] = src_host[t]; T.dst[t] = dst_host[t]; T.rows[t] = rows_host[t]; T.cols[t] = cols_host[t];
    T.accumulate[t] = accumulate_host ? accumulate_host[t] : 0;
    const int chunks = (rows_host[t] + kColsumChunk - 1) / kColsumChunk, cbs = (cols_host[t] + 31) / 32;
    T.blk0[t] = blk; T.part0[t] = part; T.cnt0[t] = cnt;
    blk += chunks * cbs; part += chunks * cols_host[t]; cnt += cbs;
  }
  T.blk0[n] = blk; T.part0[n] = part; T.cnt0[n] = cnt;
  T.counters = reinterpret_cast<unsigned*>(workspace);
  T.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kColsumCounterBytes);
  launch_pdl(colsum_multi_kernel, dim3(blk), dim3(kGlueThreads), 0, (cudaStream_t)stream, T);
  EXP_CHECK_LAUNCH("exp_colsum_multi");
  return EXP_OK;
}

int exp_stats_bwd_gin(const float* img, const float* stats, const float* g_in, int cin, float* g_out, int B, int H, int W,
                      void* stream) {
  EXP_CHECK_ARG(img && stats && g_in && g_out && B > 0 && B <= 65535 && H > 0 && W > 0 && cin >= 6, "bad args (cin >= 6: 3 image + 3 statistic channels)");
  launch_pdl_cluster(stats_bwd_gin_kernel, dim3(kGinCluster, B), dim3(kGlueThreads), 0, kGinCluster, (cudaStream_t)stream, img, stats, g_in, cin, g_out, H * W);
  EXP_CHECK_LAUNCH("exp_stats_bwd_gin");
  return EXP_OK;
}

}  // extern "C"
