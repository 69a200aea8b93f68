// Per-image "head" math of the Exposure train step: critic input statistics (critics.py:48-76)
// with their backward and forward-mode tangent, the action-selection head of the policy
// (agent.py:100-122, 208-252, pdf_sample_layer.py:5-10), the over-exposure penalty, the RL / GAN
// loss seeds (net.py:92-163), the WGAN-GP helpers (net.py:174-187) and fused Adam
// (config_example.py:158, tf.train.AdamOptimizer).  All tiny compared with the filter and
// conv kernels; they exist so that a whole train step is a fixed sequence of launches with no
// host round trip (CUDA-graph capturable).
#include "common.cuh"

namespace expo {

constexpr float kLR = 0.27f, kLG = 0.67f, kLB = 0.06f;

template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  // warp shuffle on the two halves of the double
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < THREADS / 32; ++w) t += sh[w];
  return t;
}

__device__ __forceinline__ float lum_of(const float* p) {           // critics.py:48-49
  return p[0] * kLR + p[1] * kLG + p[2] * kLB + 1e-5f;
}
// saturation of one pixel and its derivative w.r.t. the three channels (TF tie rules:
// clip passes on ties; reduce_max / reduce_min split the gradient evenly among ties)
__device__ __forceinline__ float sat_of(const float* p, float* w /*nullable [3]*/) {
  const float c0 = fminf(fmaxf(p[0], 0.f), 1.f), c1 = fminf(fmaxf(p[1], 0.f), 1.f), c2 = fminf(fmaxf(p[2], 0.f), 1.f);
  const float mx = fmaxf(c0, fmaxf(c1, c2)), mn = fminf(c0, fminf(c1, c2));
  const float s = mx + mn, t = 2.0f - mx - mn;
  const bool d_is_s = s <= t;
  const float den = (d_is_s ? s : t) + 1e-2f;
  const float r = mx - mn;
  const float sat = r / den;
  if (w) {
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
  }
  return sat;
}

// stats[b] = (mean lum, population variance of lum, mean saturation)
// grid (kStatsCluster, B), one cluster per image: CTA r takes the r-th slice of the pixels (cluster_sum, common.cuh)
constexpr unsigned kStatsCluster = 8;
__device__ __forceinline__ void stats_slice(int P, int& lo, int& hi) {
  const unsigned r = cluster_rank_x(), n = cluster_size_x();
  lo = (int)((long long)P * r / n);
  hi = (int)((long long)P * (r + 1) / n);
}
__global__ void __launch_bounds__(256) stats_fwd_kernel(const float* __restrict__ img, float* __restrict__ stats, int P) {
  EXP_PDL_ENTRY();
  __shared__ double sh[8 * 3], xch[3];
  const float* x = img + (size_t)blockIdx.y * P * 3;
  int lo, hi;
  stats_slice(P, lo, hi);
  // one pass, one cluster reduction: sum l, sum l^2 (exact products in double), sum sat; the population variance is
  // E[l^2] - E[l]^2 in double (l <= 4, P <= 2^29: the cancellation costs ~1e-16 mean^2 / var, far below float rounding)
  double a[3] = {0.0, 0.0, 0.0};
  for (int p = lo + threadIdx.x; p < hi; p += 256) {
    const double l = (double)lum_of(x + 3 * (size_t)p);
    a[0] += l;
    a[1] += l * l;
    a[2] += (double)sat_of(x + 3 * (size_t)p, nullptr);
  }
  cluster_sum<256, 3>(a, sh, xch);
  if (threadIdx.x == 0 && cluster_rank_x() == 0) {
    const double mean = a[0] / P, var = a[1] / P - mean * mean;
    stats[blockIdx.y * 3 + 0] = (float)mean;
    stats[blockIdx.y * 3 + 1] = (float)(var > 0.0 ? var : 0.0);
    stats[blockIdx.y * 3 + 2] = (float)(a[2] / P);
  }
}

// g_out = (g_direct ? g_direct : 0) + J_stats^T g_stat     grid (chunks, B)
__global__ void __launch_bounds__(256) stats_bwd_kernel(const float* __restrict__ img, const float* __restrict__ stats,
                                                        const float* __restrict__ g_stat,
                                                        const float* __restrict__ g_direct, float* __restrict__ g_out,
                                                        int P) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.y;
  const float mean = stats[b * 3], gm = g_stat[b * 3], gv = g_stat[b * 3 + 1], gs = g_stat[b * 3 + 2];
  const float invP = 1.0f / (float)P;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < P; p += gridDim.x * 256) {
    const size_t o = ((size_t)b * P + p) * 3;
    float w[3];
    sat_of(img + o, w);
    const float gl = (gm + gv * 2.f * (lum_of(img + o) - mean)) * invP;
    const float coef[3] = {kLR, kLG, kLB};
#pragma unroll
    for (int c = 0; c < 3; ++c)
      g_out[o + c] = (g_direct ? g_direct[o + c] : 0.f) + coef[c] * gl + gs * invP * w[c];
  }
}

// dstat[b] = J_stats u   (forward-mode tangent)   grid (kStatsCluster, B), one cluster per image
__global__ void __launch_bounds__(256) stats_jvp_kernel(const float* __restrict__ img, const float* __restrict__ stats,
                                                        const float* __restrict__ u, float* __restrict__ dstat, int P) {
  EXP_PDL_ENTRY();
  __shared__ double sh[8 * 3], xch[3];
  const size_t base = (size_t)blockIdx.y * P * 3;
  const float mean = stats[blockIdx.y * 3];
  int lo, hi;
  stats_slice(P, lo, hi);
  double a[3] = {0.0, 0.0, 0.0};
  for (int p = lo + threadIdx.x; p < hi; p += 256) {
    const size_t o = base + 3 * (size_t)p;
    float w[3];
    sat_of(img + o, w);
    const float cu = kLR * u[o] + kLG * u[o + 1] + kLB * u[o + 2];
    a[0] += (double)cu;
    a[1] += (double)((lum_of(img + o) - mean) * cu);
    a[2] += (double)(w[0] * u[o] + w[1] * u[o + 1] + w[2] * u[o + 2]);
  }
  cluster_sum<256, 3>(a, sh, xch);
  if (threadIdx.x == 0 && cluster_rank_x() == 0) {
    dstat[blockIdx.y * 3 + 0] = (float)(a[0] / P);
    dstat[blockIdx.y * 3 + 1] = (float)(2.0 * a[1] / P);
    dstat[blockIdx.y * 3 + 2] = (float)(a[2] / P);
  }
}

// ---- action selection head (one thread per image) ---------------------------------------
struct HeadCfg {
  int n_filters, n_states, is_train, test_steps;
  float exploration, exploration_penalty, filter_usage_penalty;
  const float* progress;   // device scalar (a captured CUDA graph can be replayed with a new value)
};
constexpr int kMaxFilters = 16;

__device__ __forceinline__ void head_pdf(const float* l, const HeadCfg& c, float* sm, float* pdf, float* Zout) {
  float mxl = l[0];
  for (int k = 1; k < c.n_filters; ++k) mxl = fmaxf(mxl, l[k]);
  float se = 0.f;
  for (int k = 0; k < c.n_filters; ++k) { sm[k] = expf(l[k] - mxl); se += sm[k]; }
  float Z = 0.f;
  for (int k = 0; k < c.n_filters; ++k) {
    sm[k] = sm[k] / se;
    pdf[k] = (sm[k] + 1e-37f) * (1.f - c.exploration) + c.exploration * 1.0f / c.n_filters;   // agent.py:100-104
    Z += pdf[k];
  }
  Z += 1e-30f;
  for (int k = 0; k < c.n_filters; ++k) pdf[k] = pdf[k] / Z;                                   // agent.py:107
  *Zout = Z;
}

__global__ void policy_head_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                       const float* __restrict__ states, HeadCfg c, int B, float* __restrict__ pdf_out,
                                       int* __restrict__ id_out, float* __restrict__ surrogate, float* __restrict__ entropy,
                                       float* __restrict__ penalty_head, float* __restrict__ new_states) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sm[kMaxFilters], pdf[kMaxFilters], Z;
  head_pdf(logits + (size_t)b * c.n_filters, c, sm, pdf, &Z);
  float ent = 0.f, tot = 0.f;
  int amax = 0;
  for (int k = 0; k < c.n_filters; ++k) {
    ent += -pdf[k] * logf(pdf[k]);                                  // agent.py:108-109
    tot += pdf[k];
    if (pdf[k] > pdf[amax]) amax = k;
    pdf_out[(size_t)b * c.n_filters + k] = pdf[k];
  }
  // pdf_sample (pdf_sample_layer.py:5-10): exclusive cumsum, count cdf_k < u, minus one
  const float u = noise[b];
  float cdf = 0.f;
  int cnt = 0;
  for (int k = 0; k < c.n_filters; ++k) {
    cnt += cdf < u;
    cdf += pdf[k] / (tot + 1e-36f);
  }
  const int id = c.is_train ? cnt - 1 : amax;                       // agent.py:113-116
  id_out[b] = id;
  surrogate[b] = id >= 0 ? logf(pdf[id] + 1e-10f) : 0.f;            // agent.py:121-122
  entropy[b] = ent;
  const float* st = states + (size_t)b * c.n_states;
  float* ns = new_states + (size_t)b * c.n_states;
  const float is_last = fabsf(st[2] + 1.f - (float)c.test_steps) < 1e-4f ? 1.f : 0.f;   // agent.py:210-214
  ns[0] = is_last; ns[1] = is_last; ns[2] = st[2] + 1.f;
  float usage_pen = 0.f;
  for (int k = 0; k < c.n_filters; ++k) {
    const float oh = k == id ? 1.f : 0.f;
    usage_pen += st[3 + k] * oh;                                    // agent.py:230-233
    ns[3 + k] = fmaxf(st[3 + k], oh);
  }
  const float ent_pen = (1.0f - __ldg(c.progress)) * c.exploration_penalty * (-ent + logf((float)c.n_filters));
  penalty_head[b] = ent_pen + usage_pen * c.filter_usage_penalty;   // agent.py:246-252 (early stop term == 0)
}

// g_logits from d/d surrogate and d/d penalty (through the entropy term)
__global__ void policy_head_bwd_kernel(const float* __restrict__ logits, const int* __restrict__ ids,
                                       const float* __restrict__ g_surrogate, const float* __restrict__ g_penalty,
                                       HeadCfg c, int B, float* __restrict__ g_logits) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sm[kMaxFilters], pdf[kMaxFilters], Z;
  head_pdf(logits + (size_t)b * c.n_filters, c, sm, pdf, &Z);
  const int id = ids[b];
  const float gs = g_surrogate[b];
  const float ge = g_penalty[b] * (1.0f - __ldg(c.progress)) * c.exploration_penalty;   // d pen / d(-entropy)
  float gp[kMaxFilters];
  float dot_pq = 0.f;
  for (int k = 0; k < c.n_filters; ++k) {
    gp[k] = ge * (logf(pdf[k]) + 1.f) + (k == id ? gs / (pdf[k] + 1e-10f) : 0.f);
    dot_pq += gp[k] * pdf[k];                       // sum_j g_pdf_j q_j / Z  == sum g_pdf_j pdf_j
  }
  float gsum = 0.f;
  float gq[kMaxFilters];
  for (int k = 0; k < c.n_filters; ++k) {
    gq[k] = (gp[k] - dot_pq) / Z * (1.f - c.exploration);   // through q/Z and s*(1-expl)
    gsum += gq[k] * sm[k];
  }
  for (int k = 0; k < c.n_filters; ++k) g_logits[(size_t)b * c.n_filters + k] = sm[k] * (gq[k] - gsum);
}

// pen[b] = mean_{h,w,c} max(x-1,0)^2   (agent.py:247)          one CTA per image
__global__ void __launch_bounds__(256) overexposure_fwd_kernel(const float* __restrict__ img, float* __restrict__ pen, int n) {
  EXP_PDL_ENTRY();
  __shared__ double sh[8];
  const float* x = img + (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = fmaxf(x[i] - 1.f, 0.f);
    s += (double)(d * d);
  }
  s = block_sum<256>(s, sh);
  if (threadIdx.x == 0) pen[blockIdx.x] = (float)(s / n);
}
// g_out = (g_in ? g_in : 0) + g_pen[b] * 2 max(x-1,0) / n
__global__ void __launch_bounds__(256) overexposure_bwd_kernel(const float* __restrict__ img, const float* __restrict__ g_pen,
                                                               const float* __restrict__ g_in, float* __restrict__ g_out,
                                                               int n) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.y;
  const float k = g_pen[b] * 2.f / (float)n;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const size_t o = (size_t)b * n + i;
    g_out[o] = (g_in ? g_in[o] : 0.f) + k * fmaxf(img[o] - 1.f, 0.f);
  }
}

// ---- RL / GAN loss seeds of the generator + value step (net.py:92-163) ------------------
struct RlCfg {
  float all_reward, critic_logit_multiplier, discount_factor, parameter_lr_mul;
  int max_traj_len, n_states, use_penalty;
};
// out_seeds[5][B]: g_fake_logit, g_new_value, g_old_value, g_penalty, g_surrogate ;
// out_losses[2] = (g_loss, v_loss)                                   single CTA
__global__ void __launch_bounds__(256) rl_losses_kernel(const float* __restrict__ fake_logit, const float* __restrict__ fake_input_logit,
                                                        const float* __restrict__ old_value, const float* __restrict__ new_value,
                                                        const float* __restrict__ penalty, const float* __restrict__ surrogate,
                                                        const float* __restrict__ new_states, RlCfg c, int B,
                                                        float* __restrict__ seeds, float* __restrict__ losses) {
  EXP_PDL_ENTRY();
  __shared__ double sh[8];
  double gl = 0.0, vl = 0.0;
  const float invB = 1.0f / (float)B;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float stopped = new_states[(size_t)b * c.n_states + 1];
    const float clear_final = new_states[(size_t)b * c.n_states + 2] > (float)c.max_traj_len ? 1.f : 0.f;
    const float nv = new_value[b] * (1.0f - clear_final);
    const float rw = c.all_reward + (1.f - c.all_reward) * stopped;
    const float raw = rw * (fake_logit[b] - fake_input_logit[b]) * c.critic_logit_multiplier;
    const float reward = c.use_penalty ? raw - penalty[b] : raw;
    const float q = reward + (1.0f - stopped) * c.discount_factor * nv;
    const float adv = q - old_value[b];
    vl += (double)(adv * adv);
    gl += (double)(-q * c.parameter_lr_mul + surrogate[b] * (-adv));
    const float gq = -c.parameter_lr_mul * invB;
    seeds[0 * B + b] = gq * rw * c.critic_logit_multiplier;
    seeds[1 * B + b] = gq * (1.0f - stopped) * c.discount_factor * (1.0f - clear_final);
    seeds[2 * B + b] = -2.f * adv * invB;
    seeds[3 * B + b] = c.use_penalty ? -gq : 0.f;
    seeds[4 * B + b] = -adv * invB;
  }
  gl = block_sum<256>(gl, sh);
  vl = block_sum<256>(vl, sh);
  if (threadIdx.x == 0) { losses[0] = (float)(gl / B); losses[1] = (float)(vl / B); }
}

// ---- WGAN-GP helpers (net.py:174-187) -----------------------------------------------------
__global__ void __launch_bounds__(256) interpolate_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                                          const float* __restrict__ alpha, float* __restrict__ out, int n) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.y;
  const float a = alpha[b];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const size_t o = (size_t)b * n + i;
    out[o] = real[o] + a * (fake[o] - real[o]);
  }
}
// norm[b] = sqrt(1e-6 + sum g^2); u = g * lambda * 2 max(norm-1,0) / (B norm)   (u may alias g)
__global__ void __launch_bounds__(256) gp_scale_kernel(const float* __restrict__ g, float* __restrict__ u,
                                                       float* __restrict__ norm, float lambda, int B, int n) {
  EXP_PDL_ENTRY();
  __shared__ double sh[8];
  const size_t base = (size_t)blockIdx.x * n;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)g[base + i] * (double)g[base + i];
  s = block_sum<256>(s, sh);
  const float nrm = sqrtf(1e-6f + (float)s);
  const float coef = lambda * 2.f * fmaxf(nrm - 1.f, 0.f) / ((float)B * nrm);
  for (int i = threadIdx.x; i < n; i += 256) u[base + i] = g[base + i] * coef;
  if (threadIdx.x == 0) norm[blockIdx.x] = nrm;
}

// ---- fused Adam over a flat parameter buffer (tf.train.AdamOptimizer) ---------------------
// hyper[0] = lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)  (device scalar so that a captured
// CUDA graph can be replayed with a new learning rate), grad_scale multiplies the gradient
// (1/world_size after the all-reduce sum).
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ hyper, float beta1, float beta2, float eps, float grad_scale, size_t n) {
  EXP_PDL_ENTRY();
  const float lr_t = hyper[0];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace expo

using namespace expo;

extern "C" {

int exp_stats_fwd(const float* img, float* stats, int B, int H, int W, void* stream) {
  EXP_CHECK_ARG(img && stats && B > 0 && B <= 65535 && H > 0 && W > 0, "bad args");
  launch_pdl_cluster(stats_fwd_kernel, dim3(kStatsCluster, B), dim3(256), 0, kStatsCluster, (cudaStream_t)stream, img, stats, H * W);
  EXP_CHECK_LAUNCH("exp_stats_fwd");
  return EXP_OK;
}
int exp_stats_bwd(const float* img, const float* stats, const float* g_stat, const float* g_direct, float* g_out,
                  int B, int H, int W, void* stream) {
  EXP_CHECK_ARG(img && stats && g_stat && g_out && B > 0 && H > 0 && W > 0 && B <= 65535, "bad args");
  const int P = H * W;
  dim3 grid(min((P + 255) / 256, 64), B);
  launch_pdl(stats_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, img, stats, g_stat, g_direct, g_out, P);
  EXP_CHECK_LAUNCH("exp_stats_bwd");
  return EXP_OK;
}
int exp_stats_jvp(const float* img, const float* stats, const float* u, float* dstat, int B, int H, int W, void* stream) {
  EXP_CHECK_ARG(img && stats && u && dstat && B > 0 && B <= 65535 && H > 0 && W > 0, "bad args");
  launch_pdl_cluster(stats_jvp_kernel, dim3(kStatsCluster, B), dim3(256), 0, kStatsCluster, (cudaStream_t)stream, img, stats, u, dstat, H * W);
  EXP_CHECK_LAUNCH("exp_stats_jvp");
  return EXP_OK;
}

int exp_policy_head_fwd(const float* logits, const float* noise, const float* states, int B, int n_filters, int n_states,
                        int is_train, int test_steps, float exploration, float exploration_penalty,
                        float filter_usage_penalty, const float* progress, float* pdf, int* ids, float* surrogate,
                        float* entropy, float* penalty_head, float* new_states, void* stream) {
  EXP_CHECK_ARG(logits && noise && states && progress && pdf && ids && surrogate && entropy && penalty_head && new_states,
                "null pointer");
  EXP_CHECK_ARG(B > 0 && n_filters > 0 && n_filters <= kMaxFilters && n_states == 3 + n_filters, "bad sizes");
  HeadCfg c{n_filters, n_states, is_train, test_steps, exploration, exploration_penalty, filter_usage_penalty, progress};
  launch_pdl(policy_head_fwd_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)stream, logits, noise, states, c, B, pdf, ids, surrogate,
                                                                         entropy, penalty_head, new_states);
  EXP_CHECK_LAUNCH("exp_policy_head_fwd");
  return EXP_OK;
}
int exp_policy_head_bwd(const float* logits, const int* ids, const float* g_surrogate, const float* g_penalty, int B,
                        int n_filters, float exploration, float exploration_penalty, const float* progress, float* g_logits,
                        void* stream) {
  EXP_CHECK_ARG(logits && ids && g_surrogate && g_penalty && progress && g_logits, "null pointer");
  EXP_CHECK_ARG(B > 0 && n_filters > 0 && n_filters <= kMaxFilters, "bad sizes");
  HeadCfg c{n_filters, 3 + n_filters, 1, 0, exploration, exploration_penalty, 0.f, progress};
  launch_pdl(policy_head_bwd_kernel, dim3((B + 63) / 64), dim3(64), 0, (cudaStream_t)stream, logits, ids, g_surrogate, g_penalty, c, B, g_logits);
  EXP_CHECK_LAUNCH("exp_policy_head_bwd");
  return EXP_OK;
}

int exp_overexposure_fwd(const float* img, float* pen, int B, int H, int W, void* stream) {
  EXP_CHECK_ARG(img && pen && B > 0 && H > 0 && W > 0, "bad args");
  launch_pdl(overexposure_fwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, img, pen, H * W * 3);
  EXP_CHECK_LAUNCH("exp_overexposure_fwd");
  return EXP_OK;
}
int exp_overexposure_bwd(const float* img, const float* g_pen, const float* g_in, float* g_out, int B, int H, int W,
                         void* stream) {
  EXP_CHECK_ARG(img && g_pen && g_out && B > 0 && H > 0 && W > 0 && B <= 65535, "bad args");
  const int n = H * W * 3;
  dim3 grid(min((n + 255) / 256, 64), B);
  launch_pdl(overexposure_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, img, g_pen, g_in, g_out, n);
  EXP_CHECK_LAUNCH("exp_overexposure_bwd");
  return EXP_OK;
}

int exp_rl_losses(const float* fake_logit, const float* fake_input_logit, const float* old_value, const float* new_value,
                  const float* penalty, const float* surrogate, const float* new_states, int B, int n_states,
                  float all_reward, float critic_logit_multiplier, float discount_factor, float parameter_lr_mul,
                  int max_traj_len, int use_penalty, float* seeds, float* losses, void* stream) {
  EXP_CHECK_ARG(fake_logit && fake_input_logit && old_value && new_value && penalty && surrogate && new_states && seeds && losses,
                "null pointer");
  EXP_CHECK_ARG(B > 0 && n_states >= 3, "bad sizes");
  RlCfg c{all_reward, critic_logit_multiplier, discount_factor, parameter_lr_mul, max_traj_len, n_states, use_penalty};
  launch_pdl(rl_losses_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, fake_logit, fake_input_logit, old_value, new_value, penalty, surrogate,
                                                        new_states, c, B, seeds, losses);
  EXP_CHECK_LAUNCH("exp_rl_losses");
  return EXP_OK;
}

int exp_interpolate(const float* real, const float* fake, const float* alpha, float* out, int B, int n, void* stream) {
  EXP_CHECK_ARG(real && fake && alpha && out && B > 0 && n > 0 && B <= 65535, "bad args");
  dim3 grid(min((n + 255) / 256, 64), B);
  launch_pdl(interpolate_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, real, fake, alpha, out, n);
  EXP_CHECK_LAUNCH("exp_interpolate");
  return EXP_OK;
}
int exp_gp_scale(const float* g, float* u, float* norm, float lambda, int B, int n, void* stream) {
  EXP_CHECK_ARG(g && u && norm && B > 0 && n > 0, "bad args");
  launch_pdl(gp_scale_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, g, u, norm, lambda, B, n);
  EXP_CHECK_LAUNCH("exp_gp_scale");
  return EXP_OK;
}

int exp_adam(float* params, const float* grads, float* m, float* v, const float* hyper, float beta1, float beta2,
             float eps, float grad_scale, size_t n, void* stream) {
  EXP_CHECK_ARG(params && grads && m && v && hyper && n > 0, "bad args");
  const unsigned blocks = (unsigned)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256);
  launch_pdl(adam_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, params, grads, m, v, hyper, beta1, beta2, eps, grad_scale, n);
  EXP_CHECK_LAUNCH("exp_adam");
  return EXP_OK;
}

}  // extern "C"
