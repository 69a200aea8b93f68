// Per-pixel arithmetic of the eight Exposure filters (forward + hand-derived backward).
//
// Reference: yuanming-hu/exposure filters.py (process() of each Filter subclass) and the
// TF-1.6 kernels it lowers to; see include/exposure_b200.h for the file:line map and
// DESIGN.md section 4 for the derivations of the backward closed forms.
//
// Numerics policy: chains that the reference's fp32 formula makes ill-conditioned
// (SaturationPlus' 1-s', the curve prefix sums) are written with explicit __f*_rn
// intrinsics in the reference's op order so that ptxas cannot contract them into FMAs and
// the result tracks the fp32 restatement to ~1 ulp.  Everything else may use FMA.
#pragma once
#ifndef EXPO_HOST_MATH   // tests/host_math/px_harness.cpp compiles this header for the host
#include "common.cuh"
#endif

namespace expo {

constexpr int kCurveSteps = 8;              // cfg.curve_steps (config_example.py:27)
constexpr float kInvL = 0.125f;             // 1 / curve_steps
constexpr float kLn2 = 0.69314718f;         // float32(np.log(2))      filters.py:182
constexpr float kPi = 3.14159274f;          // float32(math.pi)        filters.py:417
constexpr float kLumR = 0.27f, kLumG = 0.67f, kLumB = 0.06f;   // util.py:271-274
constexpr int kAccStride = 32;              // floats per partial-sum record

__host__ __device__ constexpr int num_params(int fid) {
  return fid == EXP_FILTER_WB ? 3 : fid == EXP_FILTER_TONE ? 8 : fid == EXP_FILTER_COLOR ? 24 : fid == EXP_FILTER_LEVEL ? 2 : 1;
}
// number of per-thread accumulators the backward of filter `fid` carries
__host__ __device__ constexpr int num_acc(int fid) {
  return fid == EXP_FILTER_WB ? 3 : fid == EXP_FILTER_TONE ? 8 : fid == EXP_FILTER_COLOR ? 24 : fid == EXP_FILTER_LEVEL ? 2 : 1;
}

// cfg-driven ranges of the filter_param_regressors (filters.py:179 cfg.exposure_range, :202 cfg.gamma_range,
// :261 cfg.color_curve_range, :309 cfg.tone_curve_range).  Set process-wide through exp_set_filter_ranges and
// passed BY VALUE in every launch's arguments; the defaults are config_example.py:27-33.
struct FilterRanges {
  float exposure;            // tanh_range(-r, r, initial=0)
  float gamma_log;           // ln(cfg.gamma_range)
  float tone_lo, tone_hi;    // tanh_range(*cfg.tone_curve_range)
  float color_lo, color_hi;  // tanh_range(*cfg.color_curve_range, initial=1)
  float color_bias;          // atanh(2 (1 - lo)/(hi - lo) - 1)  (util.py:285-286); 0 for ranges centred on 1
};
__host__ __device__ inline FilterRanges default_ranges() {
  return FilterRanges{3.5f, 1.0986123f /* float32(np.log(3)) */, 0.5f, 2.f, 0.90f, 1.10f, 0.f};
}

// Per-image constants, built once per CTA in shared memory from params[b, :].
struct __align__(16) FilterConsts {
  float p[EXP_MAX_FILTER_PARAMS];   // regressed parameters
  float cum[3][kCurveSteps + 1];    // curve prefix sums  sum_{i<k} t_i / L   (T uses row 0)
  float scale[3];                   // L / (sum_i t_i + 1e-30)
  float S[3];                       // sum_i t_i + 1e-30
  // backward slope lookup, index = (k+1) + (L x == k ? 10 : 0) with k = clamp(floor(L x), -1, L):
  //   [0..9]   = {0, t_0 .. t_7, 0} L/S            x strictly inside segment k (0 outside [0,1])
  //   [10..19] = {0, t_0, t_0+t_1, .., t_6+t_7, t_7} L/S   x exactly on knot k (TF tie rule)
  float slope[3][2 * (kCurveSteps + 2)];
  float e;                          // Exposure: exp(p * ln2) ; Level: upper - lower + 1e-6
  float raw[EXP_MAX_FILTER_PARAMS]; // raw regressor logits (EXP_OPT_LOGITS mode only)
  FilterRanges rg;                  // regressor ranges of this launch
};

// ---- filter_param_regressor of one image (filters.py:177-179, 201-203, 223-235, 256-262,
// 306-310, 411-413, 435-436, 481-482; util.tanh_range util.py:281-294 with bias == 0, which
// holds for every range the configs use).  Single-thread helpers: n <= 24 values.
__device__ __forceinline__ float tanh_range_f(float f, float l, float r, float* dpdf) {
  const float a = tanhf(f);
  *dpdf = 0.5f * (r - l) * (1.f - a * a);
  return (a * 0.5f + 0.5f) * (r - l) + l;
}
__device__ __forceinline__ float sigmoid_f(float f, float* d) {
  const float s = 1.f / (1.f + expf(-f));
  *d = s * (1.f - s);
  return s;
}
// p[0..n) = regress(f[0..n)); BWD: gf[0..n) = J^T gp
template <bool BWD>
__device__ __forceinline__ void regress_image(int fid, const float* f, float* po, const float* gp, float* gf,
                                              const FilterRanges& rg) {
  float d;
  switch (fid) {
    case EXP_FILTER_EXPOSURE: {
      const float p = tanh_range_f(f[0], -rg.exposure, rg.exposure, &d);
      if (BWD) gf[0] = gp[0] * d; else po[0] = p;
    } break;
    case EXP_FILTER_GAMMA: {
      const float lg = rg.gamma_log;                 // float32(np.log(cfg.gamma_range))
      const float g = expf(tanh_range_f(f[0], -lg, lg, &d));
      if (BWD) gf[0] = gp[0] * g * d; else po[0] = g;
    } break;
    case EXP_FILTER_WB: {
      float s[3], ds[3];
      for (int c = 0; c < 3; ++c) {
        const float fm = c == 0 ? 0.f : f[c];        // mask (0,1,1)
        s[c] = expf(tanh_range_f(fm, -0.5f, 0.5f, &ds[c]));
        ds[c] = c == 0 ? 0.f : ds[c] * s[c];
      }
      const float D = 1e-5f + kLumR * s[0] + kLumG * s[1] + kLumB * s[2];
      const float inv = 1.0f / D;
      if (!BWD) {
        for (int c = 0; c < 3; ++c) po[c] = s[c] * inv;
      } else {
        const float dot = (gp[0] * s[0] + gp[1] * s[1] + gp[2] * s[2]) * inv * inv;
        const float coef[3] = {kLumR, kLumG, kLumB};
        for (int c = 0; c < 3; ++c) gf[c] = (gp[c] * inv - dot * coef[c]) * ds[c];
      }
    } break;
    case EXP_FILTER_SATPLUS:
    case EXP_FILTER_WNB:
    case EXP_FILTER_VIGNET: {                        // filters.py:481-482, 435-436, 348-349
      const float p = sigmoid_f(f[0], &d);
      if (BWD) gf[0] = gp[0] * d; else po[0] = p;
    } break;
    case EXP_FILTER_LEVEL:                           // filters.py:456-457
      for (int i = 0; i < 2; ++i) {
        const float p = sigmoid_f(f[i], &d);
        if (BWD) gf[i] = gp[i] * d; else po[i] = p;
      }
      break;
    case EXP_FILTER_CONTRAST: {
      const float a = tanhf(f[0]);
      if (BWD) gf[0] = gp[0] * (1.f - a * a); else po[0] = a;
    } break;
    case EXP_FILTER_TONE:
      for (int i = 0; i < 8; ++i) {
        const float p = tanh_range_f(f[i], rg.tone_lo, rg.tone_hi, &d);
        if (BWD) gf[i] = gp[i] * d; else po[i] = p;
      }
      break;
    case EXP_FILTER_COLOR:
      for (int i = 0; i < 24; ++i) {
        const float p = tanh_range_f(f[i] + rg.color_bias, rg.color_lo, rg.color_hi, &d);
        if (BWD) gf[i] = gp[i] * d; else po[i] = p;
      }
      break;
    default: break;
  }
}

// Executed by the first warp of a CTA; followed by __syncthreads() at the call site.
// logits != 0: prow holds raw regressor logits; the regressed parameters are computed here
// (fused filter_param_regressor) and the logits kept for the backward's chain rule.
// `t` = lane of the warp doing the set-up (the fused-chain kernel gives every step its own warp).
__device__ __forceinline__ void setup_consts_lane(FilterConsts& sc, const float* __restrict__ prow, int fid, int logits, int t,
                                                  const FilterRanges& rg = default_ranges()) {
  const int n = num_params(fid);
  if (t == 0) sc.rg = rg;
  if (logits) {
    if (t < EXP_MAX_FILTER_PARAMS) { sc.raw[t] = (t < n) ? prow[t] : 0.f; sc.p[t] = 0.f; }
    __syncwarp();
    if (fid == EXP_FILTER_TONE || fid == EXP_FILTER_COLOR) {      // 8 / 24 independent tanh_range: one lane each
      if (t < n) {
        float d;
        sc.p[t] = fid == EXP_FILTER_TONE ? tanh_range_f(sc.raw[t], rg.tone_lo, rg.tone_hi, &d)
                                         : tanh_range_f(sc.raw[t] + rg.color_bias, rg.color_lo, rg.color_hi, &d);
      }
    } else if (t == 0) {
      regress_image<false>(fid, sc.raw, sc.p, nullptr, nullptr, rg);
    }
  } else {
    if (t < EXP_MAX_FILTER_PARAMS) sc.p[t] = (t < n) ? prow[t] : 0.f;
  }
  __syncwarp();
  if (t == 0) {
    if (fid == EXP_FILTER_LEVEL)                                  // filters.py:460-464: upper = p1 + 1
      sc.e = __fadd_rn(__fsub_rn(__fadd_rn(sc.p[1], 1.f), sc.p[0]), 1e-6f);
    else
      sc.e = expf(sc.p[0] * kLn2);                                // filters.py:182
  }
  if (t < 3) {                                                    // filters.py:264-273 / 312-322
    float cum = 0.f, sum = 0.f;
    sc.cum[t][0] = 0.f;
#pragma unroll
    for (int i = 0; i < kCurveSteps; ++i) {
      const float ti = sc.p[t * kCurveSteps + i];
      cum = __fadd_rn(cum, __fmul_rn(kInvL, ti));
      sum = __fadd_rn(sum, ti);
      sc.cum[t][i + 1] = cum;
    }
    const float S = __fadd_rn(sum, 1e-30f);
    sc.S[t] = S;
    const float scale = __fdiv_rn((float)kCurveSteps, S);
    sc.scale[t] = scale;
    sc.slope[t][0] = sc.slope[t][kCurveSteps + 1] = sc.slope[t][kCurveSteps + 2] = 0.f;
#pragma unroll
    for (int i = 0; i < kCurveSteps; ++i) {
      const float ti = sc.p[t * kCurveSteps + i];
      sc.slope[t][i + 1] = ti * scale;
      sc.slope[t][kCurveSteps + 3 + i] = (i ? ti + sc.p[t * kCurveSteps + i - 1] : ti) * scale;
    }
    sc.slope[t][2 * kCurveSteps + 3] = sc.p[t * kCurveSteps + kCurveSteps - 1] * scale;
  }
}

__device__ __forceinline__ void setup_consts(FilterConsts& sc, const float* __restrict__ prow, int fid, int logits = 0,
                                             const FilterRanges& rg = default_ranges()) {
  setup_consts_lane(sc, prow, fid, logits, (int)threadIdx.x, rg);
}

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

// ---- pow(xc, g) of GammaFilter as exp2(g * log2 xc), xc >= 1e-3 (filters.py:205-206) ------------------
// log2 by exponent split: xc = m 2^e with m in [0.75, 1.5), log2 xc = e + lg2(m) with ONE MUFU.LG2 whose
// absolute error on that interval is <= 2^-22 (CUDA C Programming Guide, __log2f on [0.5, 2]); together with
// the rounding of the sum the error of the logarithm is <= 7.2e-7 for |log2 xc| < 16, i.e. a relative error of
// the power of <= ln2 * g * 7.2e-7 <= 1.5e-6 at g = 3 -- the same class as log2f's 1 ulp (9.5e-7 at |L| >= 8)
// at a third of its instructions (log2f is a ~20-instruction software polynomial).  exp2 is MUFU.EX2; the
// result is never denormal (>= 1e-9).
// rcp_fast: MUFU.RCP (1 ulp, flush-to-zero) for the BACKWARD formulas, whose bar is 1e-4 (no bit tracking of
// a reference op order there): `1.f / x` and __fdividef compile to 6-10 instructions with denormal checks and
// a slow-path branch.  Callers guarantee a normal, non-zero argument.
// cos_pi / sincos_pi: cos(pi x), sin(pi x) by exact range reduction in x (no Cody-Waite constants, no slow
// path): half the instructions of cosf(pi_f * x) and within its rounding of it on [0, 1].
#ifdef EXPO_HOST_MATH
__device__ __forceinline__ float lg2_pos(float x) { return log2f(x); }
__device__ __forceinline__ float ex2_fast(float t) { return exp2f(t); }
__device__ __forceinline__ float rcp_fast(float x) { return 1.0f / x; }
__device__ __forceinline__ float cos_pi(float x) { return (float)cos(3.14159265358979323846 * (double)x); }
__device__ __forceinline__ void sincos_pi(float x, float* s, float* c) {
  *s = (float)sin(3.14159265358979323846 * (double)x);
  *c = (float)cos(3.14159265358979323846 * (double)x);
}
#else
__device__ __forceinline__ float rcp_fast(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float cos_pi(float x) { return cospif(x); }
__device__ __forceinline__ void sincos_pi(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ float lg2_pos(float x) {
  const int ix = __float_as_int(x);
  const int e = (ix - 0x3f400000) >> 23;                   // floor(log2(x / 0.75))
  const float m = __int_as_float(ix - (e << 23));          // in [0.75, 1.5)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m));
  return (float)e + r;
}
__device__ __forceinline__ float ex2_fast(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}
#endif

// ---- monotone piecewise-linear curve (ToneFilter / ColorFilter) ---------------------
// y = (L/S) * sum_i clip(x - i/L, 0, 1/L) * t_i ; evaluated through the prefix sums:
// segment k = floor(L * clamp(x,0,1)), y = (cum[k] + (xc - k/L) * t_k) * (L/S).
// `k` is returned for the backward's slope lookup.
__device__ __forceinline__ float curve_eval(float x, const FilterConsts& sc, int row, int* kout = nullptr) {
  const float xc = __saturatef(x);
  const int k = min((int)(xc * (float)kCurveSteps), kCurveSteps - 1);
  const float frac = xc - (float)k * kInvL;                 // exact (Sterbenz)
  const float tot = __fadd_rn(sc.cum[row][k], __fmul_rn(frac, sc.p[row * kCurveSteps + k]));
  if (kout) *kout = k;
  return __fmul_rn(tot, sc.scale[row]);
}

// ---- TF colour-space round trip of SaturationPlusFilter (filters.py:484-498) ----------
// Returns xm = min(x,1) and full = hsv_to_rgb(h, s', v).  tensorflow/core/kernels/colorspace_op.h computes a hue and
// turns it back into three ramps d_c = clamp(...) in [0, 1]; whichever hue branch is taken, those ramps are
//     d_c = (c - min) / (max - min)        (1 for the largest channel(s), 0 for the smallest, linear in between)
// so  full_c = V ((1 - s') + s' d_c)  needs no hue: ~35 instructions instead of ~78, and the same closed form the
// backward uses (px_bwd).  A grey pixel (range 0) has hue 0 in TF, i.e. d = (1, 0, 0): its saturation is raised towards
// RED -- the reference's behaviour, kept.  Two MUFU reciprocals (1 ulp) replace TF's IEEE divisions: the result differs
// from the op-by-op restatement by a few ulp, inside the 1e-5 relative bar of the parity tests.
// A range below FLT_MIN counts as grey (rcp.approx flushes it to zero), as in the backward.
__device__ __forceinline__ void satplus_full(const float (&x)[3], float (&xm)[3], float (&full)[3]) {
  const float r = fminf(x[0], 1.f), g = fminf(x[1], 1.f), b = fminf(x[2], 1.f);
  xm[0] = r; xm[1] = g; xm[2] = b;
  const float V = fmaxf(r, fmaxf(g, b));
  const float m = fminf(r, fminf(g, b));
  const float rng = V - m;
  const bool col = rng >= 1.17549435e-38f;
  const float S = (col && V > 0.f) ? rng * rcp_fast(V) : 0.f;
  const float kk = 0.5f - fabsf(0.5f - V);
  const float s2 = S + (1.f - S) * kk * 0.8f;
  const float inv = col ? rcp_fast(rng) : 0.f;
  const float dr = col ? (r - m) * inv : 1.f;
  const float dg = (g - m) * inv;                          // 0 for a grey pixel (inv = 0)
  const float db = (b - m) * inv;
  const float one_s = 1.f - s2;
  full[0] = (one_s + s2 * dr) * V;
  full[1] = (one_s + s2 * dg) * V;
  full[2] = (one_s + s2 * db) * V;
}

// =======================================================================================
// forward of one pixel
// =======================================================================================
template <int FID>
__device__ __forceinline__ void px_fwd(const float (&x)[3], float (&y)[3], const FilterConsts& sc) {
  if constexpr (FID == EXP_FILTER_EXPOSURE) {          // filters.py:181-182
    const float e = sc.e;
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = x[c] * e;
  } else if constexpr (FID == EXP_FILTER_GAMMA) {      // filters.py:205-206  max(x,1e-3)^gamma
    const float gm = sc.p[0];
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = ex2_fast(gm * lg2_pos(fmaxf(x[c], 0.001f)));
  } else if constexpr (FID == EXP_FILTER_WB) {         // filters.py:237-238
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = x[c] * sc.p[c];
  } else if constexpr (FID == EXP_FILTER_SATPLUS) {    // filters.py:484-498
    float xm[3], full[3];
    satplus_full(x, xm, full);
    const float p = sc.p[0];
    const float q = __fsub_rn(1.f, p);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = __fadd_rn(__fmul_rn(xm[c], q), __fmul_rn(full[c], p));
  } else if constexpr (FID == EXP_FILTER_TONE) {       // filters.py:312-322
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = curve_eval(x[c], sc, 0);
  } else if constexpr (FID == EXP_FILTER_COLOR) {      // filters.py:264-273
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = curve_eval(x[c], sc, c);
  } else if constexpr (FID == EXP_FILTER_CONTRAST) {   // filters.py:415-419
    const float p = sc.p[0];
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(kLumR, x[0]), __fmul_rn(kLumG, x[1])), __fmul_rn(kLumB, x[2]));
    const float l = clamp01(lum);
    // cos(pi l): the reference's fp32 constant pi_f = 3.14159274 differs from pi by 8.7e-8, which moves the
    // cosine by < 1 ulp(1) on [0, 1] -- inside the allowance the tests give this formula's own cancellation
    const float cl = __fadd_rn(__fmul_rn(-cos_pi(l), 0.5f), 0.5f);
    // x / (l + 1e-6) * cl as x * (cl * rcp(l + 1e-6)): one MUFU.RCP instead of three IEEE divisions; <= 3 ulp of
    // x / (l + 1e-6) away from the op-by-op restatement, inside the slack the forward tests grant this formula for the
    // cancellation in cl (tests/test_filters_gpu.py _fwd_tol)
    const float w = __fmul_rn(cl, rcp_fast(__fadd_rn(l, 1e-6f)));
    const float q = __fsub_rn(1.f, p);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = __fadd_rn(__fmul_rn(q, x[c]), __fmul_rn(p, __fmul_rn(x[c], w)));
  } else if constexpr (FID == EXP_FILTER_LEVEL) {      // filters.py:459-464
    const float lo = sc.p[0], d = sc.e;
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = clamp01(__fdiv_rn(__fsub_rn(x[c], lo), d));
  } else if constexpr (FID == EXP_FILTER_VIGNET) {     // filters.py:351-352  `img * 0`
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = x[c] * 0.f;
  } else {                                             // WNB  filters.py:438-440
    const float p = sc.p[0];
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(kLumR, x[0]), __fmul_rn(kLumG, x[1])), __fmul_rn(kLumB, x[2]));
    const float q = __fsub_rn(1.f, p);
#pragma unroll
    for (int c = 0; c < 3; ++c) y[c] = __fadd_rn(__fmul_rn(q, x[c]), __fmul_rn(p, lum));
  }
}

// =======================================================================================
// backward of one pixel: gx = (dy/dx)^T gy ; acc += per-parameter partial sums
// (final transform of the sums is finalize_gparams()).
// =======================================================================================
// HAS_Y: `yf` holds the forward output of this pixel (the whole-chain kernel has it in registers: it is the
// next step's input); filters whose backward would recompute it (Gamma) read it instead -- same bits.
template <int FID, bool HAS_GX, bool HAS_Y = false>
__device__ __forceinline__ void px_bwd(const float (&x)[3], const float (&gy)[3], float (&gx)[3],
                                       float* __restrict__ acc, const FilterConsts& sc, const float* yf = nullptr) {
  if constexpr (FID == EXP_FILTER_EXPOSURE) {
    // y = x e ; dy/dp = y ln2 (ln2 applied at finalize) ; dy/dx = e
    const float e = sc.e;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s = fmaf(gy[c] * x[c], e, s);
      if (HAS_GX) gx[c] = gy[c] * e;
    }
    acc[0] += s;
  } else if constexpr (FID == EXP_FILTER_GAMMA) {
    // y = xc^g ; dy/dg = y ln xc ; dy/dx = g y / xc [x >= 1e-3]   (TF MaximumGrad tie -> x)
    const float gm = sc.p[0];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float xc = fmaxf(x[c], 0.001f);
      const float l2 = lg2_pos(xc);
      const float y = HAS_Y ? yf[c] : ex2_fast(gm * l2);
      s = fmaf(gy[c] * y, l2, s);
      if (HAS_GX) gx[c] = x[c] >= 0.001f ? gy[c] * gm * (y * rcp_fast(xc)) : 0.f;
    }
    acc[0] = fmaf(s, kLn2, acc[0]);
  } else if constexpr (FID == EXP_FILTER_WB) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc[c] = fmaf(gy[c], x[c], acc[c]);
      if (HAS_GX) gx[c] = gy[c] * sc.p[c];
    }
  } else if constexpr (FID == EXP_FILTER_SATPLUS) {
    // closed form of the HSV round trip:
    //   full_c = xm_c - k(V) m u_c,  u_c = (V - xm_c)/(V - m),  k = 0.8 (0.5 - |0.5 - V|)
    const float p = sc.p[0];
    const float r = fminf(x[0], 1.f), g = fminf(x[1], 1.f), b = fminf(x[2], 1.f);
    const bool is_r = (r >= g) && (r >= b);
    const bool is_g = !is_r && (g >= b);
    const bool mn_r = (r <= g) && (r <= b);
    const bool mn_g = !mn_r && (g <= b);
    const float V = is_r ? r : (is_g ? g : b);
    const float m = mn_r ? r : (mn_g ? g : b);
    const float rng = V - m;
    const float k = (0.5f - fabsf(0.5f - V)) * 0.8f;
    const float kp = V < 0.5f ? 0.8f : (V > 0.5f ? -0.8f : 0.f);
    const float xm[3] = {r, g, b};
    float gF[3] = {gy[0] * p, gy[1] * p, gy[2] * p};
    float gxm[3], gV, gm;
    float dp;   // sum_c gy_c (full_c - xm_c)
    if (rng >= 1.17549435e-38f) {                    // a denormal range is a grey pixel (rcp_fast flushes it to zero)
      const float inv = rcp_fast(rng);
      const float Q = k * m;
      float Tt = 0.f, sg = 0.f, sgy_u = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float u = (V - xm[c]) * inv;
        Tt = fmaf(gF[c], u, Tt);
        sgy_u = fmaf(gy[c], u, sgy_u);
        sg += gF[c];
        gxm[c] = fmaf(Q * inv, gF[c], gy[c]);
      }
      dp = -Q * sgy_u;
      gV = -Tt * kp * m - Q * (sg - Tt) * inv;
      gm = -Tt * k - Q * Tt * inv;
    } else {
      // grey pixel: TF's hue == 0 gives full = (V, (1-k)V, (1-k)V)
#pragma unroll
      for (int c = 0; c < 3; ++c) gxm[c] = gy[c] * (1.f - p);
      gV = gF[0] + (gF[1] + gF[2]) * (1.f - k - kp * V);
      gm = 0.f;
      dp = -(gy[1] + gy[2]) * k * V;
    }
    acc[0] += dp;
    if (HAS_GX) {
      gxm[0] += (is_r ? gV : 0.f) + (mn_r ? gm : 0.f);
      gxm[1] += (is_g ? gV : 0.f) + (mn_g ? gm : 0.f);
      gxm[2] += ((!is_r && !is_g) ? gV : 0.f) + ((!mn_r && !mn_g) ? gm : 0.f);
#pragma unroll
      for (int c = 0; c < 3; ++c) gx[c] = x[c] <= 1.f ? gxm[c] : 0.f;
    }
  } else if constexpr (FID == EXP_FILTER_TONE || FID == EXP_FILTER_COLOR) {
    // y = (1/S) sum_i c_i t_i with c_i = L clip_i(x) = sat(L x - i)  (scaling by L = 8 is exact, so
    // c_i equals clamp(x - i/L, 0, 1/L) * L bit for bit; one FADD.SAT).
    //   dy/dt_j = (c_j - y)/S    ->  acc[row*8 + j] = A_j = sum gy c_j ; sum gy y = (sum_i t_i A_i)/S
    //                                is a linear combination of the A_j and is formed once per image
    //                                in finalize_gparams (fp64), not per pixel.
    //   dy/dx   = (L/S) sum_{j passes} t_j : TF clip_by_value passes the gradient on ties
    //             (x - j/L == 0 or == 1/L), so on an exact interior knot BOTH neighbouring segments
    //             pass; every case (inside / on a knot / outside [0,1]) is ONE branch-free lookup in
    //             sc.slope.  row = 0 for Tone.
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int row = (FID == EXP_FILTER_COLOR) ? c : 0;
      float* a = acc + row * kCurveSteps;
      const float xs = x[c] * (float)kCurveSteps;
#pragma unroll
      for (int j = 0; j < kCurveSteps; ++j) a[j] = fmaf(gy[c], __saturatef(xs - (float)j), a[j]);
      if (HAS_GX) {
        const int k = min(max(__float2int_rd(xs), -1), kCurveSteps);
        gx[c] = gy[c] * sc.slope[row][xs == (float)k ? k + kCurveSteps + 3 : k + 1];
      }
    }
  } else if constexpr (FID == EXP_FILTER_CONTRAST) {
    // y_c = (1-p) x_c + p x_c w(l),  w = cl/(l+eps),  cl = sin^2(pi l / 2)  (== -cos(pi l)/2 + 1/2)
    const float p = sc.p[0];
    const float lum = fmaf(kLumB, x[2], fmaf(kLumG, x[1], kLumR * x[0]));
    const float l = clamp01(lum);
    float sn, cs;
    sincos_pi(0.5f * l, &sn, &cs);
    const float cl = sn * sn;
    const float dcl = kPi * sn * cs;                 // 0.5 pi sin(pi l)
    const float iden = rcp_fast(l + 1e-6f);
    const float w = cl * iden;
    const float dw = (dcl - w) * iden;
    const float sgx = fmaf(gy[2], x[2], fmaf(gy[1], x[1], gy[0] * x[0]));
    acc[0] = fmaf(sgx, w - 1.f, acc[0]);
    if (HAS_GX) {
      const float a = fmaf(p, w, 1.f - p);
      const float bq = (lum >= 0.f && lum <= 1.f) ? p * sgx * dw : 0.f;
      gx[0] = fmaf(bq, kLumR, gy[0] * a);
      gx[1] = fmaf(bq, kLumG, gy[1] * a);
      gx[2] = fmaf(bq, kLumB, gy[2] * a);
    }
  } else if constexpr (FID == EXP_FILTER_LEVEL) {
    // v = (x - lo)/d, y = clip(v,0,1), d = up - lo + 1e-6 ; on 0 <= v <= 1 (ties pass, TF clip):
    // dy/dx = 1/d ; dy/dlo = (v - 1)/d ; dy/dup = -v/d   (the 1/d is applied at finalize)
    const float lo = sc.p[0], d = sc.e;
    const float inv = 1.f / d;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn(__fsub_rn(x[c], lo), d);
      const float g = (v >= 0.f && v <= 1.f) ? gy[c] : 0.f;
      acc[0] = fmaf(g, v - 1.f, acc[0]);
      acc[1] = fmaf(g, -v, acc[1]);
      if (HAS_GX) gx[c] = g * inv;
    }
  } else if constexpr (FID == EXP_FILTER_VIGNET) {
    // y = x * 0: no parameter dependence; dy/dx = 0
    if (HAS_GX) gx[0] = gx[1] = gx[2] = 0.f;
  } else {                                           // WNB
    const float p = sc.p[0];
    const float lum = fmaf(kLumB, x[2], fmaf(kLumG, x[1], kLumR * x[0]));
    const float sg = gy[0] + gy[1] + gy[2];
    acc[0] += sg * lum - fmaf(gy[2], x[2], fmaf(gy[1], x[1], gy[0] * x[0]));
    if (HAS_GX) {
      const float q = 1.f - p;
      const float ps = p * sg;
      gx[0] = fmaf(ps, kLumR, q * gy[0]);
      gx[1] = fmaf(ps, kLumG, q * gy[1]);
      gx[2] = fmaf(ps, kLumB, q * gy[2]);
    }
  }
}

// Final transform from reduced sums (double) to dL/dparam for image b.  `sum[a]` holds the
// image-wide total of accumulator a.  Writes n = num_params(fid) floats.
__device__ __forceinline__ void finalize_gparams(int fid, const double* sum, const float* p, float* out) {
  if (fid == EXP_FILTER_EXPOSURE) {
    out[0] = (float)(sum[0] * (double)kLn2);
  } else if (fid == EXP_FILTER_TONE || fid == EXP_FILTER_COLOR) {
    const int rows = fid == EXP_FILTER_TONE ? 1 : 3;
    for (int r = 0; r < rows; ++r) {
      float s = 0.f;
      for (int i = 0; i < kCurveSteps; ++i) s = __fadd_rn(s, p[r * kCurveSteps + i]);
      const double S = (double)__fadd_rn(s, 1e-30f);
      double Bs = 0.0;                           // sum gy y = (sum_i t_i A_i) / S
      for (int i = 0; i < kCurveSteps; ++i) Bs += (double)p[r * kCurveSteps + i] * sum[r * kCurveSteps + i];
      Bs /= S;
      for (int j = 0; j < kCurveSteps; ++j)      // A_j already carries the factor L
        out[r * kCurveSteps + j] = (float)((sum[r * kCurveSteps + j] - Bs) / S);
    }
  } else if (fid == EXP_FILTER_LEVEL) {
    const double d = (double)__fadd_rn(__fsub_rn(__fadd_rn(p[1], 1.f), p[0]), 1e-6f);
    out[0] = (float)(sum[0] / d);
    out[1] = (float)(sum[1] / d);
  } else {
    const int n = num_params(fid);
    for (int i = 0; i < n; ++i) out[i] = (float)sum[i];
  }
}

// Same, followed (EXP_OPT_LOGITS mode) by the regressor's chain rule: out = dL/dlogits.
__device__ __forceinline__ void finalize_grads(int fid, const double* sum, const FilterConsts& sc, int logits, float* out) {
  if (!logits) {
    finalize_gparams(fid, sum, sc.p, out);
    return;
  }
  float gp[EXP_MAX_FILTER_PARAMS];
  finalize_gparams(fid, sum, sc.p, gp);
  regress_image<true>(fid, sc.raw, nullptr, gp, out, sc.rg);
}

}  // namespace expo
