// Spatial mask of Filter.apply (cfg.masking == True) and its backward.
//
// Reference: filters.py:62-99 (apply: out = lerp(img, process(img, param), mask)),
// filters.py:110-148 (Filter.get_mask, 6 parameters) and filters.py:354-396
// (VignetFilter.get_mask, 5 parameters), util.py:281-294 (tanh_range), util.py:307-308 (lerp).
//
//   m    = tanh_range(-5, 5, initial=0)(mask_logits)                        per image
//   g0,g1 = ((row | col) + (se - (H | W))/2)/se - 0.5, se = min(H, W)       centred unit grid
//   kind 0: pre = g0 m0 + g1 m1 + m2 (lum(img) - 0.5) + 2 m3 ; sharp = max_sharpness m4 / 5
//           mask = sigmoid(pre sharp) (m5/5 0.5 + 0.5) (1 - min_strength) + min_strength
//   kind 1: pre = (g0 m0)^2 + (g1 m1)^2 + m2 - 5            ; sharp = max_sharpness m3 / 5
//           mask = sigmoid(pre sharp) (m4/5 0.5 + 0.5)       (then mask*0+1 when masking is off)
//   out  = (1 - mask) img + mask proc
//
// Backward (gy = dL/dout), per pixel:  gmask = sum_c gy_c (proc_c - img_c),
//   ginp = gmask strength' sig (1 - sig), and with the six running sums
//     kind 0: A0 = S ginp g0, A1 = S ginp g1, A2 = S ginp (lum - .5), A3 = S ginp, A4 = S ginp pre,
//             A5 = S gmask sig
//             dL/dm = (sharp A0, sharp A1, sharp A2, 2 sharp A3, max_sharp/5 A4, (1 - min_s)/10 A5)
//     kind 1: A0 = S ginp g0^2, A1 = S ginp g1^2, A2 = S ginp, A3 = S ginp pre, A4 = S gmask sig
//             dL/dm = (2 m0 sharp A0, 2 m1 sharp A1, sharp A2, max_sharp/5 A3, A4/10)
//   dL/dlogit_i = dL/dm_i 5 (1 - tanh^2 f_i);  the filter's own parameter sums are px_bwd's with
//   gy' = mask gy, and dL/dimg = (1 - mask) gy + J_proc^T (mask gy) + ginp sharp m2 lumcoef (kind 0).
#pragma once
#include "filter_math.cuh"

namespace expo {

constexpr int kMaskParams = 6;          // Filter.get_num_mask_parameters (filters.py:107-108)
constexpr float kMaskRange = 5.f;       // filter_input_range (filters.py:121)

struct __align__(16) MaskConsts {
  float m[kMaskParams];     // tanh_range(-5,5)(logits)
  float dm[kMaskParams];    // dm/dlogit
  float sharp;              // max_sharpness * m[4 | 3] / 5
  float strength;           // kind 0: (m5/5*.5+.5)*(1-min_s) ; kind 1: m4/5*.5+.5
  float floor_;             // kind 0: min_strength ; kind 1: 0
  float off_i, off_j, se;   // grid = (i + off)/se - 0.5
  float max_sharp, one_minus_min;
  int kind;                 // 0 Filter.get_mask, 1 VignetFilter.get_mask
  int on;                   // 0: masking disabled -> mask == 1 (Filter.get_mask / Vignet `mask*0+1`)
};

// executed by one thread (t == 0) of the CTA
__device__ __forceinline__ void setup_mask(MaskConsts& mc, const float* __restrict__ lrow, int fid, int H, int W,
                                           float max_sharp, float min_strength, int on) {
  const int kind = fid == EXP_FILTER_VIGNET ? 1 : 0;
  const int n = kind ? 5 : 6;
  for (int i = 0; i < kMaskParams; ++i) {
    const float f = (i < n && lrow) ? lrow[i] : 0.f;
    const float a = tanhf(f);
    mc.m[i] = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(a, 0.5f), 0.5f), 2.f * kMaskRange), -kMaskRange);
    mc.dm[i] = i < n ? kMaskRange * (1.f - a * a) : 0.f;
  }
  const float ms = kind ? 0.f : min_strength;
  mc.sharp = __fdiv_rn(__fmul_rn(max_sharp, mc.m[kind ? 3 : 4]), kMaskRange);
  const float st = __fadd_rn(__fmul_rn(__fdiv_rn(mc.m[kind ? 4 : 5], kMaskRange), 0.5f), 0.5f);
  mc.strength = st;                       // the (1 - min_s) factor is applied per pixel in the reference's order
  mc.one_minus_min = __fsub_rn(1.f, ms);
  mc.floor_ = ms;
  const int se = min(H, W);
  mc.se = (float)se;
  mc.off_i = (float)(se - H) * 0.5f;
  mc.off_j = (float)(se - W) * 0.5f;
  mc.max_sharp = max_sharp;
  mc.kind = kind;
  mc.on = on;
}

struct MaskPx {
  float mask, sig, pre, g0, g1, lumc;   // lumc = lum - 0.5 (kind 0)
};

// mask value of pixel (i, j) with colour x  (forward order of filters.py:134-147 / 376-389)
__device__ __forceinline__ MaskPx mask_eval(const MaskConsts& mc, int i, int j, const float (&x)[3]) {
  MaskPx r;
  r.g0 = __fsub_rn(__fdiv_rn(__fadd_rn((float)i, mc.off_i), mc.se), 0.5f);
  r.g1 = __fsub_rn(__fdiv_rn(__fadd_rn((float)j, mc.off_j), mc.se), 0.5f);
  float pre;
  if (mc.kind == 0) {
    const float lum = __fadd_rn(__fadd_rn(__fmul_rn(kLumR, x[0]), __fmul_rn(kLumG, x[1])), __fmul_rn(kLumB, x[2]));
    r.lumc = __fsub_rn(lum, 0.5f);
    pre = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r.g0, mc.m[0]), __fmul_rn(r.g1, mc.m[1])), __fmul_rn(mc.m[2], r.lumc)),
                    __fmul_rn(mc.m[3], 2.f));
  } else {
    const float a = __fmul_rn(r.g0, mc.m[0]), b = __fmul_rn(r.g1, mc.m[1]);
    r.lumc = 0.f;
    pre = __fsub_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), mc.m[2]), kMaskRange);
  }
  r.pre = pre;
  r.sig = 1.f / (1.f + expf(-__fmul_rn(pre, mc.sharp)));
  const float mk = __fadd_rn(__fmul_rn(__fmul_rn(r.sig, mc.strength), mc.one_minus_min), mc.floor_);
  r.mask = mc.on ? mk : 1.f;
  return r;
}

// out = lerp(x, proc, mask)  (util.py:307-308: (1 - l) a + l b)
__device__ __forceinline__ void mask_blend(const float (&x)[3], const float (&proc)[3], float mask, float (&out)[3]) {
  const float q = __fsub_rn(1.f, mask);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c] = __fadd_rn(__fmul_rn(q, x[c]), __fmul_rn(mask, proc[c]));
}

// backward of one masked pixel.  macc[0..5] += the running sums A0..A5 above; acc = px_bwd's sums.
template <int FID, bool HAS_GX>
__device__ __forceinline__ void px_bwd_masked(const float (&x)[3], const float (&gy)[3], float (&gx)[3], float* acc,
                                              float* macc, const FilterConsts& sc, const MaskConsts& mc, int i, int j) {
  const MaskPx r = mask_eval(mc, i, j, x);
  float proc[3], gyp[3], gxp[3];
  px_fwd<FID>(x, proc, sc);
#pragma unroll
  for (int c = 0; c < 3; ++c) gyp[c] = gy[c] * r.mask;
  px_bwd<FID, HAS_GX>(x, gyp, gxp, acc, sc);
  float ginp = 0.f;
  if (mc.on) {
    const float gmask = gy[0] * (proc[0] - x[0]) + gy[1] * (proc[1] - x[1]) + gy[2] * (proc[2] - x[2]);
    ginp = gmask * mc.strength * mc.one_minus_min * r.sig * (1.f - r.sig);
    if (mc.kind == 0) {
      macc[0] = fmaf(ginp, r.g0, macc[0]);
      macc[1] = fmaf(ginp, r.g1, macc[1]);
      macc[2] = fmaf(ginp, r.lumc, macc[2]);
      macc[3] += ginp;
      macc[4] = fmaf(ginp, r.pre, macc[4]);
      macc[5] = fmaf(gmask, r.sig, macc[5]);
    } else {
      macc[0] = fmaf(ginp, r.g0 * r.g0, macc[0]);
      macc[1] = fmaf(ginp, r.g1 * r.g1, macc[1]);
      macc[2] += ginp;
      macc[3] = fmaf(ginp, r.pre, macc[3]);
      macc[4] = fmaf(gmask, r.sig, macc[4]);
    }
  }
  if (HAS_GX) {
    const float q = 1.f - r.mask;
    const float gl = mc.kind == 0 ? ginp * mc.sharp * mc.m[2] : 0.f;
    gx[0] = fmaf(q, gy[0], gxp[0]) + gl * kLumR;
    gx[1] = fmaf(q, gy[1], gxp[1]) + gl * kLumG;
    gx[2] = fmaf(q, gy[2], gxp[2]) + gl * kLumB;
  }
}

// reduced sums (fp64) -> dL/dmask_logits[0..5]
__device__ __forceinline__ void finalize_mask_grads(const double* A, const MaskConsts& mc, float* out) {
  double gm[kMaskParams] = {0, 0, 0, 0, 0, 0};
  if (mc.on) {
    const double sh = (double)mc.sharp;
    if (mc.kind == 0) {
      gm[0] = sh * A[0];
      gm[1] = sh * A[1];
      gm[2] = sh * A[2];
      gm[3] = 2.0 * sh * A[3];
      gm[4] = (double)mc.max_sharp / (double)kMaskRange * A[4];
      gm[5] = (double)mc.one_minus_min * 0.5 / (double)kMaskRange * A[5];
    } else {
      gm[0] = 2.0 * (double)mc.m[0] * sh * A[0];
      gm[1] = 2.0 * (double)mc.m[1] * sh * A[1];
      gm[2] = sh * A[2];
      gm[3] = (double)mc.max_sharp / (double)kMaskRange * A[3];
      gm[4] = 0.5 / (double)kMaskRange * A[4];
    }
  }
  for (int i = 0; i < kMaskParams; ++i) out[i] = (float)(gm[i] * (double)mc.dm[i]);
}

}  // namespace expo
