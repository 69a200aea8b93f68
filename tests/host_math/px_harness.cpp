// TEST INFRASTRUCTURE ONLY -- never shipped, never imported by exposure_b200/.
//
// Compiles the product's per-pixel device math (exposure_b200/csrc/filter_math.cuh and
// filter_mask.cuh: px_fwd / px_bwd / px_bwd_masked / setup_consts / finalize_*) for the HOST with
// g++ by mapping the CUDA intrinsics they use onto plain C (-ffp-contract=off keeps the explicit
// __f*_rn chains uncontracted, as ptxas does).  tests/test_host_math.py drives it against the CPU
// oracle, so the closed-form backward formulas, the curve slope table, the Level / Vignet filters
// and the mask are checked here, without a GPU, before the kernels run on one.  The kernels
// themselves (tiling, TMA ring, reductions) are only exercised by the `-m gpu` tests.
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define EXPO_HOST_MATH 1

// one host thread per lane of the warp that runs setup_consts(); __syncwarp() is a real barrier
static thread_local struct { int x; } threadIdx;
static pthread_barrier_t g_warp_barrier;
static inline void __syncwarp() { pthread_barrier_wait(&g_warp_barrier); }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __saturatef(float a) { return a != a ? 0.f : fminf(fmaxf(a, 0.f), 1.f); }
static inline int __float2int_rd(float a) {
  if (a != a) return 0;
  const float f = floorf(a);
  return f >= 2147483648.f ? 2147483647 : f <= -2147483648.f ? (-2147483647 - 1) : (int)f;
}
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline float warp_sum(float v) { return v; }

// the two product headers under test (common.cuh is replaced by the lines above)
#define EXPOSURE_B200_COMMON_CUH_SKIP 1
#include "../../include/exposure_b200.h"
#include "../../exposure_b200/csrc/filter_math.cuh"
#include "../../exposure_b200/csrc/filter_mask.cuh"

using namespace expo;

// regressor ranges the harness hands to setup_consts (hm_set_ranges; the kernels receive them as launch arguments)
static FilterRanges g_hm_ranges = default_ranges();
struct LaneArg { FilterConsts* sc; const float* prow; int fid, logits, lane; };
static void* lane_main(void* p) {
  LaneArg* a = (LaneArg*)p;
  threadIdx.x = a->lane;
  setup_consts(*a->sc, a->prow, a->fid, a->logits, g_hm_ranges);
  return nullptr;
}
static void host_setup(FilterConsts& sc, const float* prow, int fid, int logits) {
  memset(&sc, 0, sizeof(sc));
  pthread_barrier_init(&g_warp_barrier, nullptr, 32);
  pthread_t th[32];
  LaneArg args[32];
  for (int t = 0; t < 32; ++t) {
    args[t] = LaneArg{&sc, prow, fid, logits, t};
    pthread_create(&th[t], nullptr, lane_main, &args[t]);
  }
  for (int t = 0; t < 32; ++t) pthread_join(th[t], nullptr);
  pthread_barrier_destroy(&g_warp_barrier);
}

template <int FID>
static void fwd_t(const float* x, float* y, const float* params, int pstride, int B, int P, int logits) {
  FilterConsts sc;
  for (int b = 0; b < B; ++b) {
    host_setup(sc, params + (size_t)b * pstride, FID, logits);
    for (int q = 0; q < P; ++q) {
      const float* p = x + ((size_t)b * P + q) * 3;
      const float px[3] = {p[0], p[1], p[2]};
      float py[3];
      px_fwd<FID>(px, py, sc);
      memcpy(y + ((size_t)b * P + q) * 3, py, 12);
    }
  }
}

template <int FID>
static void bwd_t(const float* x, const float* gy, float* gx, float* gparams, const float* params, int pstride, int B,
                  int P, int logits) {
  FilterConsts sc;
  constexpr int NA = num_acc(FID);
  for (int b = 0; b < B; ++b) {
    host_setup(sc, params + (size_t)b * pstride, FID, logits);
    double tot[32] = {0};
    for (int q = 0; q < P; ++q) {
      const size_t o = ((size_t)b * P + q) * 3;
      const float px[3] = {x[o], x[o + 1], x[o + 2]}, pg[3] = {gy[o], gy[o + 1], gy[o + 2]};
      float pgx[3] = {0, 0, 0}, acc[NA + 1];
      for (int a = 0; a < NA; ++a) acc[a] = 0.f;
      px_bwd<FID, true>(px, pg, pgx, acc, sc);
      for (int a = 0; a < NA; ++a) tot[a] += (double)acc[a];
      if (gx) memcpy(gx + o, pgx, 12);
    }
    finalize_grads(FID, tot, sc, logits, gparams + (size_t)b * pstride);
  }
}

template <int FID>
static void mfwd_t(const float* x, float* y, float* mask_out, const float* params, int pstride, const float* ml,
                   int mstride, int B, int H, int W, float ms, float mn, int masking, int logits) {
  FilterConsts sc;
  MaskConsts mc;
  const int P = H * W;
  for (int b = 0; b < B; ++b) {
    host_setup(sc, params + (size_t)b * pstride, FID, logits);
    setup_mask(mc, ml ? ml + (size_t)b * mstride : nullptr, FID, H, W, ms, mn, masking);
    for (int q = 0; q < P; ++q) {
      const size_t o = ((size_t)b * P + q) * 3;
      const float px[3] = {x[o], x[o + 1], x[o + 2]};
      float proc[3], py[3];
      px_fwd<FID>(px, proc, sc);
      const MaskPx r = mask_eval(mc, q / W, q % W, px);
      mask_blend(px, proc, r.mask, py);
      memcpy(y + o, py, 12);
      if (mask_out) mask_out[(size_t)b * P + q] = r.mask;
    }
  }
}

template <int FID>
static void mbwd_t(const float* x, const float* gy, float* gx, float* gparams, float* gmask, const float* params,
                   int pstride, const float* ml, int mstride, int B, int H, int W, float ms, float mn, int masking,
                   int logits) {
  FilterConsts sc;
  MaskConsts mc;
  constexpr int NA = num_acc(FID);
  const int P = H * W;
  for (int b = 0; b < B; ++b) {
    host_setup(sc, params + (size_t)b * pstride, FID, logits);
    setup_mask(mc, ml ? ml + (size_t)b * mstride : nullptr, FID, H, W, ms, mn, masking);
    double tot[40] = {0};
    for (int q = 0; q < P; ++q) {
      const size_t o = ((size_t)b * P + q) * 3;
      const float px[3] = {x[o], x[o + 1], x[o + 2]}, pg[3] = {gy[o], gy[o + 1], gy[o + 2]};
      float pgx[3] = {0, 0, 0}, acc[NA + kMaskParams];
      for (int a = 0; a < NA + kMaskParams; ++a) acc[a] = 0.f;
      px_bwd_masked<FID, true>(px, pg, pgx, acc, acc + NA, sc, mc, q / W, q % W);
      for (int a = 0; a < NA + kMaskParams; ++a) tot[a] += (double)acc[a];
      if (gx) memcpy(gx + o, pgx, 12);
    }
    finalize_grads(FID, tot, sc, logits, gparams + (size_t)b * pstride);
    finalize_mask_grads(tot + NA, mc, gmask + (size_t)b * mstride);
  }
}

#define DISPATCH(fid, CALL)                                                                        \
  switch (fid) {                                                                                   \
    case 0: CALL(0); break; case 1: CALL(1); break; case 2: CALL(2); break; case 3: CALL(3); break; \
    case 4: CALL(4); break; case 5: CALL(5); break; case 6: CALL(6); break; case 7: CALL(7); break; \
    case 8: CALL(8); break; case 9: CALL(9); break; default: return -1;                            \
  }

extern "C" {
// same arithmetic as exp_set_filter_ranges (csrc/filters.cu); exposure_range <= 0 restores the defaults
void hm_set_ranges(float exposure_range, float gamma_range, float tone_lo, float tone_hi, float color_lo, float color_hi) {
  if (exposure_range <= 0.f) { g_hm_ranges = default_ranges(); return; }
  FilterRanges g;
  g.exposure = exposure_range;
  g.gamma_log = (float)log((double)gamma_range);
  g.tone_lo = tone_lo; g.tone_hi = tone_hi; g.color_lo = color_lo; g.color_hi = color_hi;
  g.color_bias = (float)atanh(2.0 * (1.0 - (double)color_lo) / ((double)color_hi - (double)color_lo) - 1.0);
  if (fabsf(g.color_bias) < 1e-7f) g.color_bias = 0.f;
  g_hm_ranges = g;
}
int hm_fwd(int fid, const float* x, float* y, const float* params, int pstride, int B, int P, int logits) {
#define C_(F) fwd_t<F>(x, y, params, pstride, B, P, logits)
  DISPATCH(fid, C_)
#undef C_
  return 0;
}
int hm_bwd(int fid, const float* x, const float* gy, float* gx, float* gparams, const float* params, int pstride,
           int B, int P, int logits) {
#define C_(F) bwd_t<F>(x, gy, gx, gparams, params, pstride, B, P, logits)
  DISPATCH(fid, C_)
#undef C_
  return 0;
}
int hm_masked_fwd(int fid, const float* x, float* y, float* mask_out, const float* params, int pstride,
                  const float* ml, int mstride, int B, int H, int W, float ms, float mn, int masking, int logits) {
#define C_(F) mfwd_t<F>(x, y, mask_out, params, pstride, ml, mstride, B, H, W, ms, mn, masking, logits)
  DISPATCH(fid, C_)
#undef C_
  return 0;
}
int hm_masked_bwd(int fid, const float* x, const float* gy, float* gx, float* gparams, float* gmask,
                  const float* params, int pstride, const float* ml, int mstride, int B, int H, int W, float ms,
                  float mn, int masking, int logits) {
#define C_(F) mbwd_t<F>(x, gy, gx, gparams, gmask, params, pstride, ml, mstride, B, H, W, ms, mn, masking, logits)
  DISPATCH(fid, C_)
#undef C_
  return 0;
}
}
