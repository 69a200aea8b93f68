// Fused per-pixel filter step of the Exposure hot path: kernels + C-ABI entry points.
//
// One launch = one filter step over a whole batch (forward: read x, write y = 24 B/pixel;
// backward: read x, read gy, write gx = 36 B/pixel, per-image parameter gradients reduced
// warp-shuffle -> shared memory -> per-CTA partial -> fixed-order fp64 finish by the last
// CTA of each image).  HBM-bandwidth bound by construction; see DESIGN.md section 5.
//
// Replaces the unfused TF-1.6 Eigen launches behind filters.py process() (2..43 launches
// per filter) and the 8-way stack/one_hot/reduce_sum select of agent.py:77,118-129.
#include <cmath>
#include <algorithm>
#include <cstdlib>

#include "filter_math.cuh"
#include "filter_mask.cuh"

namespace expo {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPixPerBlockFwd = 2048;    // 256 threads x 4 pixels x 2 iterations
constexpr int kPixPerBlockBwd = 4096;    // fatter CTAs: one partial record per CTA
// Workspace layout: [kMaxImages ticket counters][partial-sum records].  The counter block has
// a FIXED size and position so that records of an earlier, differently shaped launch can
// never be mistaken for tickets.
constexpr int kMaxImages = 65536;
constexpr size_t kCounterBytes = (size_t)kMaxImages * sizeof(unsigned);
constexpr int kMaxPersistentCtas = 4096;  // upper bound on the TMA variant's grid

// process-wide PDL switch (common.cuh launch_pdl): default OFF -- measured on B200 the train iteration is
// GPU-throughput bound (93 % busy, profiles/r1d_train_timeline.md), early-resident dependents only take
// SM slots from the parallel graph branches (6.31 -> 6.51 ms); EXPOSURE_PDL=1 / exp_set_pdl(1) turns it on
static int g_pdl = [] { const char* e = getenv("EXPOSURE_PDL"); return e ? atoi(e) : 0; }();
bool pdl_enabled() { return g_pdl != 0; }
void set_pdl(int on) { g_pdl = on; }

// process-wide regressor ranges (exp_set_filter_ranges); copied by value into every launch's arguments
static FilterRanges g_ranges = default_ranges();
FilterRanges host_ranges() { return g_ranges; }

static thread_local char g_err[512] = "";
char* last_error_buf() { return g_err; }
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct FilterArgs {
  const float* x;
  const float* gy;       // backward only
  float* out;            // y (forward) or gx (backward, nullable)
  const float* params;
  int pstride;
  const int* ids;        // per-image filter ids or nullptr
  int P;                 // pixels per image
  int pix_per_block;
  float* partials;       // [B][nblk][kAccStride]
  unsigned* counters;    // [B]
  float* gparams;        // [B][pstride]
  int logits;            // params are raw regressor logits (EXP_OPT_LOGITS)
  // masked steps only (exp_filter_masked_*):
  const float* mask_logits;   // [B][mstride] raw fc2 outputs [:, n:]  (nullable: zeros)
  float* gmask;               // [B][mstride] dL/dmask_logits (backward)
  float* mask_out;            // [B][P] (forward, nullable)
  int mstride, H, W, uniform_id, masking;
  float max_sharp, min_strength;
  FilterRanges rg;            // cfg-driven regressor ranges (exp_set_filter_ranges)
};

// ---- 4 pixels <-> 3 float4 ------------------------------------------------------------
struct Px4 { float4 a, b, c; };
__device__ __forceinline__ Px4 load_px4(const float* __restrict__ base, size_t q) {
  const float4* p = reinterpret_cast<const float4*>(base) + q * 3;
  Px4 v;
  v.a = __ldg(p);
  v.b = __ldg(p + 1);
  v.c = __ldg(p + 2);
  return v;
}
__device__ __forceinline__ void store_px4(float* __restrict__ base, size_t q, const Px4& v) {
  float4* p = reinterpret_cast<float4*>(base) + q * 3;
  p[0] = v.a;
  p[1] = v.b;
  p[2] = v.c;
}
__device__ __forceinline__ void unpack(const Px4& v, float (&px)[4][3]) {
  px[0][0] = v.a.x; px[0][1] = v.a.y; px[0][2] = v.a.z;
  px[1][0] = v.a.w; px[1][1] = v.b.x; px[1][2] = v.b.y;
  px[2][0] = v.b.z; px[2][1] = v.b.w; px[2][2] = v.c.x;
  px[3][0] = v.c.y; px[3][1] = v.c.z; px[3][2] = v.c.w;
}
__device__ __forceinline__ Px4 pack(const float (&px)[4][3]) {
  Px4 v;
  v.a = make_float4(px[0][0], px[0][1], px[0][2], px[1][0]);
  v.b = make_float4(px[1][1], px[1][2], px[2][0], px[2][1]);
  v.c = make_float4(px[2][2], px[3][0], px[3][1], px[3][2]);
  return v;
}

}  // namespace expo
#include "filters_tma.cuh"
namespace expo {

// ---- per-CTA reduction of the parameter-gradient accumulators + last-CTA finish --------
template <int FID, int NEXTRA = 0>
__device__ __forceinline__ void reduce_and_finish(float* acc, const FilterArgs& A, const FilterConsts& sc,
                                                  float (*red)[kAccStride], int b, int nblk,
                                                  const MaskConsts* mc = nullptr) {
  constexpr int NACC = num_acc(FID) + NEXTRA;
  static_assert(NACC <= kAccStride, "partial-sum record too small");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const float v = warp_sum(acc[a]);
    if (lane == 0) red[warp][a] = v;
  }
  __syncthreads();
  float* rec = A.partials + ((size_t)b * nblk + blockIdx.x) * kAccStride;
  if (threadIdx.x < NACC) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += red[w][threadIdx.x];
    rec[threadIdx.x] = s;
  }
  __shared__ unsigned ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(A.counters + b, 1u);
  __syncthreads();
  if (ticket != (unsigned)(nblk - 1)) return;
  // last CTA of image b: fixed-order fp64 sum over the CTA partials (deterministic)
  __threadfence();
  __shared__ double tot[kAccStride];
  if (threadIdx.x < NACC) {
    const float* base = A.partials + (size_t)b * nblk * kAccStride + threadIdx.x;
    double s = 0.0;
    for (int i = 0; i < nblk; ++i) s += (double)__ldcg(base + (size_t)i * kAccStride);
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    finalize_grads(FID, tot, sc, A.logits, A.gparams + (size_t)b * A.pstride);
    if constexpr (NEXTRA > 0) finalize_mask_grads(tot + num_acc(FID), *mc, A.gmask + (size_t)b * A.mstride);
    A.counters[b] = 0u;      // leave the workspace ready for the next launch
  }
}

// ---- the CTA body: VEC = 4 pixels / thread via 3 float4, else 1 pixel / thread ---------
template <int FID, bool BWD, bool HAS_GX, bool VEC>
__device__ __forceinline__ void filter_body(const FilterArgs& A) {
  __shared__ FilterConsts sc;
  __shared__ float red[BWD ? kWarps : 1][kAccStride];
  const int b = blockIdx.y;
  if (threadIdx.x < 32) setup_consts(sc, A.params + (size_t)b * A.pstride, FID, A.logits, A.rg);
  __syncthreads();

  const size_t img = (size_t)b * A.P * 3;
  const float* __restrict__ x = A.x + img;
  const float* __restrict__ gy = BWD ? A.gy + img : nullptr;
  float* __restrict__ out = (!BWD || HAS_GX) ? A.out + img : nullptr;
  const int p0 = blockIdx.x * A.pix_per_block;
  const int p1 = min(A.P, p0 + A.pix_per_block);

  constexpr int NACC = num_acc(FID);
  float acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = 0.f;

  if constexpr (VEC) {
    const int q1 = p1 >> 2;
#pragma unroll 2
    for (int q = (p0 >> 2) + threadIdx.x; q < q1; q += kThreads) {
      float px[4][3], py[4][3];
      unpack(load_px4(x, q), px);
      if constexpr (BWD) {
        float pg[4][3];
        unpack(load_px4(gy, q), pg);
#pragma unroll
        for (int i = 0; i < 4; ++i) px_bwd<FID, HAS_GX>(px[i], pg[i], py[i], acc, sc);
        if constexpr (HAS_GX) store_px4(out, q, pack(py));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) px_fwd<FID>(px[i], py[i], sc);
        store_px4(out, q, pack(py));
      }
    }
  } else {
    for (int q = p0 + threadIdx.x; q < p1; q += kThreads) {
      float px[3] = {x[3 * (size_t)q], x[3 * (size_t)q + 1], x[3 * (size_t)q + 2]};
      float py[3];
      if constexpr (BWD) {
        float pg[3] = {gy[3 * (size_t)q], gy[3 * (size_t)q + 1], gy[3 * (size_t)q + 2]};
        px_bwd<FID, HAS_GX>(px, pg, py, acc, sc);
      } else {
        px_fwd<FID>(px, py, sc);
      }
      if constexpr (!BWD || HAS_GX) {
        out[3 * (size_t)q] = py[0];
        out[3 * (size_t)q + 1] = py[1];
        out[3 * (size_t)q + 2] = py[2];
      }
    }
  }
  if constexpr (BWD) reduce_and_finish<FID>(acc, A, sc, red, b, gridDim.x);
}

// id -1 (pdf_sample's u == 0 quirk, pdf_sample_layer.py:5-10: an all-zero one-hot row): the reference's
// one-hot sum (agent.py:124-125) yields a BLACK image and no gradient.  The kernels write those zeros
// themselves, so the caller's output buffers may be uninitialised.
template <bool BWD, bool HAS_GX>
__device__ __forceinline__ void zero_body(const FilterArgs& A) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * A.pix_per_block;
  const int p1 = min(A.P, p0 + A.pix_per_block);
  if constexpr (!BWD || HAS_GX) {
    float* __restrict__ out = A.out + (size_t)b * A.P * 3;
    for (int i = 3 * p0 + threadIdx.x; i < 3 * p1; i += kThreads) out[i] = 0.f;
  }
  if constexpr (BWD) {
    if (blockIdx.x == 0) {
      for (int i = threadIdx.x; i < A.pstride; i += kThreads) A.gparams[(size_t)b * A.pstride + i] = 0.f;
      if (A.gmask) for (int i = threadIdx.x; i < A.mstride; i += kThreads) A.gmask[(size_t)b * A.mstride + i] = 0.f;
    }
  } else {
    if (A.mask_out) for (int q = p0 + threadIdx.x; q < p1; q += kThreads) A.mask_out[(size_t)b * A.P + q] = 0.f;
  }
}

// One kernel per filter for uniform steps (tight register allocation per filter) ...
template <int FID, bool BWD, bool HAS_GX, bool VEC>
__global__ void __launch_bounds__(kThreads) filter_step_kernel(const FilterArgs A) {
  EXP_PDL_ENTRY();
  filter_body<FID, BWD, HAS_GX, VEC>(A);
}

// ... and one dispatching kernel for per-image filter ids (agent.py:113-125 selection).
template <bool BWD, bool HAS_GX, bool VEC>
__global__ void __launch_bounds__(kThreads) filter_step_select_kernel(const FilterArgs A) {
  EXP_PDL_ENTRY();
  switch (A.ids[blockIdx.y]) {
    case 0: filter_body<0, BWD, HAS_GX, VEC>(A); break;
    case 1: filter_body<1, BWD, HAS_GX, VEC>(A); break;
    case 2: filter_body<2, BWD, HAS_GX, VEC>(A); break;
    case 3: filter_body<3, BWD, HAS_GX, VEC>(A); break;
    case 4: filter_body<4, BWD, HAS_GX, VEC>(A); break;
    case 5: filter_body<5, BWD, HAS_GX, VEC>(A); break;
    case 6: filter_body<6, BWD, HAS_GX, VEC>(A); break;
    case 7: filter_body<7, BWD, HAS_GX, VEC>(A); break;
    case 8: filter_body<8, BWD, HAS_GX, VEC>(A); break;
    case 9: filter_body<9, BWD, HAS_GX, VEC>(A); break;
    default: zero_body<BWD, HAS_GX>(A); break;   // id -1: black output / zero gradients, written here
  }
}


// ---- masked step (cfg.masking == True, filters.py:62-99 + get_mask): out = lerp(x, proc, mask) ----
template <int FID, bool BWD, bool HAS_GX, bool VEC>
__device__ __forceinline__ void masked_body(const FilterArgs& A) {
  __shared__ FilterConsts sc;
  __shared__ MaskConsts mc;
  __shared__ float red[BWD ? kWarps : 1][kAccStride];
  const int b = blockIdx.y;
  if (threadIdx.x < 32) setup_consts(sc, A.params + (size_t)b * A.pstride, FID, A.logits, A.rg);
  if (threadIdx.x == 32)
    setup_mask(mc, A.mask_logits ? A.mask_logits + (size_t)b * A.mstride : nullptr, FID, A.H, A.W, A.max_sharp,
               A.min_strength, A.masking);
  __syncthreads();

  const size_t img = (size_t)b * A.P * 3;
  const float* __restrict__ x = A.x + img;
  const float* __restrict__ gy = BWD ? A.gy + img : nullptr;
  float* __restrict__ out = (!BWD || HAS_GX) ? A.out + img : nullptr;
  float* __restrict__ mout = (!BWD && A.mask_out) ? A.mask_out + (size_t)b * A.P : nullptr;
  const int p0 = blockIdx.x * A.pix_per_block;
  const int p1 = min(A.P, p0 + A.pix_per_block);

  constexpr int NF = num_acc(FID);
  float acc[NF + kMaskParams];
#pragma unroll
  for (int a = 0; a < NF + kMaskParams; ++a) acc[a] = 0.f;

  auto one = [&](const float (&px)[3], const float* pg, float (&py)[3], int q) {
    const int i = q / A.W, j = q - i * A.W;
    if constexpr (BWD) {
      const float g3[3] = {pg[0], pg[1], pg[2]};
      px_bwd_masked<FID, HAS_GX>(px, g3, py, acc, acc + NF, sc, mc, i, j);
    } else {
      float proc[3];
      px_fwd<FID>(px, proc, sc);
      const MaskPx r = mask_eval(mc, i, j, px);
      mask_blend(px, proc, r.mask, py);
      if (mout) mout[q] = r.mask;
    }
  };

  if constexpr (VEC) {
    const int q1 = p1 >> 2;
    for (int q = (p0 >> 2) + threadIdx.x; q < q1; q += kThreads) {
      float px[4][3], py[4][3], pg[4][3];
      unpack(load_px4(x, q), px);
      if constexpr (BWD) unpack(load_px4(gy, q), pg);
#pragma unroll
      for (int k = 0; k < 4; ++k) one(px[k], pg[k], py[k], 4 * q + k);
      if constexpr (!BWD || HAS_GX) store_px4(out, q, pack(py));
    }
  } else {
    for (int q = p0 + threadIdx.x; q < p1; q += kThreads) {
      float px[3] = {x[3 * (size_t)q], x[3 * (size_t)q + 1], x[3 * (size_t)q + 2]};
      float py[3], pg[3] = {0.f, 0.f, 0.f};
      if constexpr (BWD) {
        pg[0] = gy[3 * (size_t)q]; pg[1] = gy[3 * (size_t)q + 1]; pg[2] = gy[3 * (size_t)q + 2];
      }
      one(px, pg, py, q);
      if constexpr (!BWD || HAS_GX) {
        out[3 * (size_t)q] = py[0];
        out[3 * (size_t)q + 1] = py[1];
        out[3 * (size_t)q + 2] = py[2];
      }
    }
  }
  if constexpr (BWD) reduce_and_finish<FID, kMaskParams>(acc, A, sc, red, b, gridDim.x, &mc);
}

template <bool BWD, bool HAS_GX, bool VEC>
__global__ void __launch_bounds__(kThreads) filter_step_masked_kernel(const FilterArgs A) {
  EXP_PDL_ENTRY();
  switch (A.ids ? A.ids[blockIdx.y] : A.uniform_id) {
#define EXP_CASE(F) case F: masked_body<F, BWD, HAS_GX, VEC>(A); break;
    EXP_CASE(0) EXP_CASE(1) EXP_CASE(2) EXP_CASE(3) EXP_CASE(4) EXP_CASE(5) EXP_CASE(6) EXP_CASE(7) EXP_CASE(8) EXP_CASE(9)
#undef EXP_CASE
    default: zero_body<BWD, HAS_GX>(A); break;   // id -1: black output / zero gradients, written here
  }
}

// Filter.get_mask alone (debug_info['mask'], filters.py:85-87): mask [B][P]
__global__ void __launch_bounds__(kThreads) mask_only_kernel(const FilterArgs A) {
  __shared__ MaskConsts mc;
  const int b = blockIdx.y;
  const int fid = A.ids ? A.ids[b] : A.uniform_id;
  if (threadIdx.x == 0)
    setup_mask(mc, A.mask_logits ? A.mask_logits + (size_t)b * A.mstride : nullptr, fid, A.H, A.W, A.max_sharp,
               A.min_strength, A.masking);
  __syncthreads();
  const float* __restrict__ x = A.x + (size_t)b * A.P * 3;
  const int p0 = blockIdx.x * A.pix_per_block;
  const int p1 = min(A.P, p0 + A.pix_per_block);
  for (int q = p0 + threadIdx.x; q < p1; q += kThreads) {
    const float px[3] = {x[3 * (size_t)q], x[3 * (size_t)q + 1], x[3 * (size_t)q + 2]};
    const int i = q / A.W;
    A.mask_out[(size_t)b * A.P + q] = mask_eval(mc, i, q - i * A.W, px).mask;
  }
}

template <bool BWD, bool HAS_GX, bool VEC>
static void launch_uniform(int fid, dim3 grid, cudaStream_t st, const FilterArgs& A) {
  switch (fid) {
#define EXP_CASE(F) \
  case F: launch_pdl(filter_step_kernel<F, BWD, HAS_GX, VEC>, grid, dim3(kThreads), 0, st, A); break;
    EXP_CASE(0) EXP_CASE(1) EXP_CASE(2) EXP_CASE(3) EXP_CASE(4) EXP_CASE(5) EXP_CASE(6) EXP_CASE(7) EXP_CASE(8) EXP_CASE(9)
#undef EXP_CASE
  }
}

template <bool BWD, bool HAS_GX>
static void launch_step(bool vec, const int* ids, int fid, dim3 grid, cudaStream_t st, const FilterArgs& A) {
  if (ids) {
    if (vec) launch_pdl(filter_step_select_kernel<BWD, HAS_GX, true>, grid, dim3(kThreads), 0, st, A);
    else launch_pdl(filter_step_select_kernel<BWD, HAS_GX, false>, grid, dim3(kThreads), 0, st, A);
  } else {
    if (vec) launch_uniform<BWD, HAS_GX, true>(fid, grid, st, A);
    else launch_uniform<BWD, HAS_GX, false>(fid, grid, st, A);
  }
}

// ---- S filter steps fused into ONE pass over the pixels (inference / high-resolution path) --
// y[b] = f_{ids[S-1][b]}( ... f_{ids[0][b]}(x[b]) ... ): intermediates stay in registers, so the
// whole episode costs 24 B/pixel of HBM traffic instead of 24*S (net.py:796-820 applies the
// test_steps selected filters to the full-resolution image one sess.run at a time).
constexpr int kMaxChain = 8;
struct ChainArgs {
  const float* x; float* y; const float* params; const int* ids;   // params [S][B][pstride], ids [S][B]
  int S, B, P, pstride, pix_per_block, logits;
  FilterRanges rg;
};

template <int FID>
__device__ __forceinline__ void chain_apply(float (&px)[4][3], const FilterConsts& sc, int npx) {
  float py[3];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i < npx) {
      px_fwd<FID>(px[i], py, sc);
      px[i][0] = py[0]; px[i][1] = py[1]; px[i][2] = py[2];
    }
  }
}
__device__ __forceinline__ void chain_apply_any(int fid, float (&px)[4][3], const FilterConsts& sc, int npx) {
  switch (fid) {
    case 0: chain_apply<0>(px, sc, npx); break;
    case 1: chain_apply<1>(px, sc, npx); break;
    case 2: chain_apply<2>(px, sc, npx); break;
    case 3: chain_apply<3>(px, sc, npx); break;
    case 4: chain_apply<4>(px, sc, npx); break;
    case 5: chain_apply<5>(px, sc, npx); break;
    case 6: chain_apply<6>(px, sc, npx); break;
    case 7: chain_apply<7>(px, sc, npx); break;
    case 8: chain_apply<8>(px, sc, npx); break;
    case 9: chain_apply<9>(px, sc, npx); break;
    default:                                   // id -1: all-zero one-hot -> black (pdf_sample quirk)
#pragma unroll
      for (int i = 0; i < 4; ++i) px[i][0] = px[i][1] = px[i][2] = 0.f;
      break;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) filter_chain_fwd_kernel(const ChainArgs A) {
  __shared__ FilterConsts sc[kMaxChain];
  __shared__ int fids[kMaxChain];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = warp; s < A.S; s += kWarps) {               // one warp per step: the S set-ups run in parallel
    const int f = A.ids ? A.ids[s * A.B + b] : -1;
    if (lane == 0) fids[s] = f;
    if (f >= 0 && f < EXP_NUM_FILTER_KINDS)
      setup_consts_lane(sc[s], A.params + ((size_t)s * A.B + b) * A.pstride, f, A.logits, lane, A.rg);
  }
  __syncthreads();
  const size_t img = (size_t)b * A.P * 3;
  const float* __restrict__ x = A.x + img;
  float* __restrict__ y = A.y + img;
  const int p0 = blockIdx.x * A.pix_per_block;
  const int p1 = min(A.P, p0 + A.pix_per_block);
  if constexpr (VEC) {
    for (int q = (p0 >> 2) + threadIdx.x; q < (p1 >> 2); q += kThreads) {
      float px[4][3];
      unpack(load_px4(x, q), px);
      for (int s = 0; s < A.S; ++s) chain_apply_any(fids[s], px, sc[s], 4);
      store_px4(y, q, pack(px));
    }
  } else {
    for (int q = p0 + threadIdx.x; q < p1; q += kThreads) {
      float px[4][3];
      px[0][0] = x[3 * (size_t)q]; px[0][1] = x[3 * (size_t)q + 1]; px[0][2] = x[3 * (size_t)q + 2];
      for (int s = 0; s < A.S; ++s) chain_apply_any(fids[s], px, sc[s], 1);
      y[3 * (size_t)q] = px[0][0]; y[3 * (size_t)q + 1] = px[0][1]; y[3 * (size_t)q + 2] = px[0][2];
    }
  }
}

// ---- whole chain forward + backward in ONE pass over the pixels ------------------------------
// y = f_{S-1}(.. f_0(x)), gx = (dy/dx)^T gy and the parameter gradients of every step, from one read of
// x and gy: 24 B/pixel in, 12 (+12 for y) out for the WHOLE chain instead of 60 B/pixel/step
// (SURVEY 8d "fully chain-fused variant").  Per 1024-pixel tile: the forward sweeps the S steps with
// the pixels in registers and parks every step's input in shared memory (S x 12 KiB); the backward
// walks the steps in reverse, reading its input back from shared memory, so nothing but x, gy, gx (and
// y) ever touches HBM and no step is recomputed.  The kernel is instruction-bound, not HBM-bound.
// Filter ids are per image and per step (run time): the step loop is a `switch` over the per-filter
// bodies, so the parameter-gradient accumulators cannot live in registers across steps.  They are
// reduced per tile-step with a multi-value butterfly (8 values in 4+2+1+1+1 = 9 shuffles instead of 40)
// into per-warp shared-memory slots owned by one lane each -> deterministic; CTA partial records and
// the last-CTA fixed-order fp64 finish are those of the per-step kernels.
constexpr int kChainRec = kMaxChain * EXP_MAX_FILTER_PARAMS;   // floats per CTA partial record (worst case)
struct ChainBwdArgs {
  const float* x; const float* gy; float* y; float* gx;          // y, gx nullable
  const float* params; const int* ids; float* gparams;           // [S][B][pstride], [S][B], [S][B][pstride]
  float* partials; unsigned* counters;
  int S, B, P, pstride, logits, nblk, ntiles;
  FilterRanges rg;
  int uniform_ids[8];                                            // used when ids == nullptr (kMaxChain entries)
};

// Sum N (power of two) per-lane values over the warp: at every stage a lane keeps one half of its values
// (which half is chosen by its lane bit), sends the other half to its partner and adds what it receives, so
// the number of values it carries halves; once one is left the remaining stages are plain butterflies.
// Afterwards every lane holds the warp total of value `idx` (lanes that differ only in their low
// log2(32/N) bits hold the same one: the caller lets the lane with those bits zero write).  Fixed order.
template <int N>
__device__ __forceinline__ float warp_multi_sum(float (&v)[N], int lane, int& idx) {
  idx = 0;
  int n = N;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    if (n > 1) {
      const bool up = (lane & d) != 0;
      const int h = n >> 1;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < h) {
          const float send = up ? v[i] : v[i + h];
          const float keep = up ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
      }
      if (up) idx += h;
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], d);
    }
  }
  return v[0];
}

template <int FID, int NPX>
__device__ __forceinline__ void chain_bwd_step(const float (&px)[4][3], float (&g)[4][3], const FilterConsts& sc,
                                               float* __restrict__ slots, int lane) {
  constexpr int NACC = num_acc(FID);
  float acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll
  for (int i = 0; i < NPX; ++i) {
    float gx[3];
    px_bwd<FID, true>(px[i], g[i], gx, acc, sc);
    g[i][0] = gx[0]; g[i][1] = gx[1]; g[i][2] = gx[2];
  }
  // rows of <= 8 accumulators: one butterfly per row (E,G,S+,Ct,BW,V: 1 value; W: 3 -> 4; Le: 2; T: 8; C: 3 x 8)
  constexpr int ROW = NACC >= 8 ? 8 : NACC == 3 ? 4 : NACC;
  constexpr int ROWS = (NACC + ROW - 1) / ROW;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float v[ROW];
#pragma unroll
    for (int i = 0; i < ROW; ++i) v[i] = (r * ROW + i < NACC) ? acc[r * ROW + i] : 0.f;
    int idx;
    const float tot = warp_multi_sum<ROW>(v, lane, idx);
    if ((lane & (32 / ROW - 1)) == 0 && r * ROW + idx < NACC) slots[r * ROW + idx] += tot;   // one owner lane per slot
  }
}

template <int NPX>
__device__ __forceinline__ void chain_bwd_any(int fid, const float (&px)[4][3], float (&g)[4][3], const FilterConsts& sc,
                                              float* __restrict__ slots, int lane) {
  switch (fid) {
#define EXP_CASE(F) case F: chain_bwd_step<F, NPX>(px, g, sc, slots, lane); break;
    EXP_CASE(0) EXP_CASE(1) EXP_CASE(2) EXP_CASE(3) EXP_CASE(4) EXP_CASE(5) EXP_CASE(6) EXP_CASE(7) EXP_CASE(8) EXP_CASE(9)
#undef EXP_CASE
    default:                                   // id -1: the output is black whatever the input -> zero gradient
#pragma unroll
      for (int i = 0; i < 4; ++i) g[i][0] = g[i][1] = g[i][2] = 0.f;
      break;
  }
}

// ---- CTA record (warps summed in fixed order), ticket, last CTA of the image finishes in fp64 ----
// slots[w][a]: warp w's partial sum of accumulator a (a < off[S]); `scratch`: >= kChainRec * 8 doubles of
// shared memory that is free by now (the parking area).  Shared by the run-time and the compile-time chain kernels.
template <int SLOTS>
__device__ __forceinline__ void chain_finish(const ChainBwdArgs& A, int b, const FilterConsts* sc, const int* fids,
                                             const int* off, float (*slots)[SLOTS], unsigned char* scratch) {
  const int tid = threadIdx.x;
  const int ntot = off[A.S];
  float* rec = A.partials + ((size_t)b * A.nblk + blockIdx.x) * kChainRec;
  for (int a = tid; a < ntot; a += kThreads) {
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sum += slots[w][a];
    rec[a] = sum;
  }
  __shared__ unsigned ticket;
  __threadfence();
  __syncthreads();
  if (tid == 0) ticket = atomicAdd(A.counters + b, 1u);
  __syncthreads();
  if (ticket != (unsigned)(A.nblk - 1)) return;
  __threadfence();
  constexpr int PARTS = 8;                                  // records summed in 8 interleaved fp64 streams, combined in order
  double* part = reinterpret_cast<double*>(scratch);       // [ntot][PARTS]
  __shared__ double tot[kChainRec];
  const float* base = A.partials + (size_t)b * A.nblk * kChainRec;
  for (int i = tid; i < ntot * PARTS; i += kThreads) {
    const int a = i / PARTS, p = i - a * PARTS;
    double sum = 0.0;
    for (int r = p; r < A.nblk; r += PARTS) sum += (double)__ldcg(base + (size_t)r * kChainRec + a);
    part[i] = sum;
  }
  __syncthreads();
  for (int a = tid; a < ntot; a += kThreads) {
    double sum = 0.0;
#pragma unroll
    for (int p = 0; p < PARTS; ++p) sum += part[a * PARTS + p];
    tot[a] = sum;
  }
  __syncthreads();
  if (tid < A.S) {
    float* out = A.gparams + ((size_t)tid * A.B + b) * A.pstride;
    if (fids[tid] >= 0) finalize_grads(fids[tid], tot + off[tid], sc[tid], A.logits, out);
    else for (int i = 0; i < EXP_MAX_FILTER_PARAMS; ++i) out[i] = 0.f;
  }
  if (tid == 0) A.counters[b] = 0u;
}

// Instruction footprint: the hot loop (10 forward + 10 backward bodies x 4 pixels, ~100 KB) is three times
// the 32 KB instruction cache (ncu: icc hit rate 83 %, stall_no_instruction 0.9 per issue).  A CTA barrier per
// step, to keep the 8 warps inside the same body, was measured and is slower (0.87 vs 0.82 ms: hit rate only
// 86 %, barrier stalls x3) -- the loop is re-streamed from L2 per tile either way.  Two pixels per thread
// (half the code per body, 3 CTAs/SM) was measured too: hit rate 91 %, issue-active 63 -> 75 %, but the
// per-thread fixed work (switch, butterflies, parking) is then paid per 2 pixels: +19 % instructions, 0.84 ms.
// A thread's work item: NPX consecutive pixels = NPX*3 floats (4 -> 3 x float4, 1 -> 3 floats).
template <int NPX>
__device__ __forceinline__ void group_load(const float* __restrict__ base, size_t q, float (&px)[4][3]) {
  if constexpr (NPX == 4) {
    unpack(load_px4(base, q), px);
  } else {
    px[0][0] = base[3 * q]; px[0][1] = base[3 * q + 1]; px[0][2] = base[3 * q + 2];
  }
}
template <int NPX>
__device__ __forceinline__ void group_store(float* __restrict__ base, size_t q, const float (&px)[4][3]) {
  if constexpr (NPX == 4) {
    store_px4(base, q, pack(px));
  } else {
    base[3 * q] = px[0][0]; base[3 * q + 1] = px[0][1]; base[3 * q + 2] = px[0][2];
  }
}
// parking area of one step: 3 columns of kThreads vectors (float4 / float), conflict-free
template <int NPX>
__device__ __forceinline__ void park_store(float* __restrict__ area, int tid, const float (&px)[4][3]) {
  if constexpr (NPX == 4) {
    float4* pk = reinterpret_cast<float4*>(area) + tid;
    const Px4 v = pack(px);
    pk[0] = v.a; pk[kThreads] = v.b; pk[2 * kThreads] = v.c;
  } else {
    float* pk = area + tid;
    pk[0] = px[0][0]; pk[kThreads] = px[0][1]; pk[2 * kThreads] = px[0][2];
  }
}
template <int NPX>
__device__ __forceinline__ void park_load(const float* __restrict__ area, int tid, float (&px)[4][3]) {
  if constexpr (NPX == 4) {
    const float4* pk = reinterpret_cast<const float4*>(area) + tid;
    Px4 v;
    v.a = pk[0]; v.b = pk[kThreads]; v.c = pk[2 * kThreads];
    unpack(v, px);
  } else {
    const float* pk = area + tid;
    px[0][0] = pk[0]; px[0][1] = pk[kThreads]; px[0][2] = pk[2 * kThreads];
  }
}

template <int NPX, bool PF>
__global__ void __launch_bounds__(kThreads, 2) filter_chain_fwd_bwd_kernel(const ChainBwdArgs A) {
  extern __shared__ __align__(16) unsigned char chain_smem[];
  __shared__ FilterConsts sc[kMaxChain];
  __shared__ int fids[kMaxChain], off[kMaxChain + 1];
  __shared__ float slots[kWarps][kChainRec];
  // parked step inputs: [S][3][kThreads] vectors of NPX floats
  float* const park = reinterpret_cast<float*>(chain_smem);
  constexpr int kArea = 3 * NPX * kThreads;                // floats per step
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  for (int s = warp; s < A.S; s += kWarps) {               // one warp per step: the S set-ups run in parallel
    const int f = A.ids ? A.ids[s * A.B + b] : A.uniform_ids[s];
    if (lane == 0) fids[s] = (f >= 0 && f < EXP_NUM_FILTER_KINDS) ? f : -1;
    if (f >= 0 && f < EXP_NUM_FILTER_KINDS)
      setup_consts_lane(sc[s], A.params + ((size_t)s * A.B + b) * A.pstride, f, A.logits, lane, A.rg);
  }
  for (int i = tid; i < kWarps * kChainRec; i += kThreads) (&slots[0][0])[i] = 0.f;
  __syncthreads();
  if (tid == 0) {
    int o = 0;
    for (int s = 0; s < A.S; ++s) { off[s] = o; o += fids[s] >= 0 ? num_acc(fids[s]) : 0; }
    off[A.S] = o;
  }
  __syncthreads();

  const size_t img = (size_t)b * A.P * 3;
  const float* __restrict__ x = A.x + img;
  const float* __restrict__ gy = A.gy + img;
  float* __restrict__ y = A.y ? A.y + img : nullptr;
  float* __restrict__ gxo = A.gx ? A.gx + img : nullptr;
  const int t0 = (int)((long long)blockIdx.x * A.ntiles / A.nblk);
  const int t1 = (int)((long long)(blockIdx.x + 1) * A.ntiles / A.nblk);
  const int nq = A.P / NPX;                                // work items (NPX-pixel groups) per image

  // PF: the tile's dL/dy and the NEXT tile's x are requested before the forward sweep starts, so both DRAM
  // round trips are covered by the sweep's arithmetic instead of stalling the warp twice per tile (ncu:
  // long_scoreboard 0.56 cycles per issued instruction without it); +21 live registers.  Measured: 0.829 ->
  // 0.810 ms at 64x512x512, 3.189 -> 3.107 ms at 256x512x512.
  float nx[4][3];
  bool nlive = false;
  if constexpr (PF) {
#pragma unroll
    for (int i = 0; i < 4; ++i) nx[i][0] = nx[i][1] = nx[i][2] = 0.f;
    const int q0 = t0 * kThreads + tid;
    nlive = t0 < t1 && q0 < nq;
    if (nlive) group_load<NPX>(x, q0, nx);
  }
  for (int t = t0; t < t1; ++t) {
    const int q = t * kThreads + tid;
    const bool live = q < nq;                              // all lanes stay in the loop: the butterflies need the full warp
    float px[4][3], g[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) px[i][0] = px[i][1] = px[i][2] = g[i][0] = g[i][1] = g[i][2] = 0.f;
    if constexpr (PF) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { px[i][0] = nx[i][0]; px[i][1] = nx[i][1]; px[i][2] = nx[i][2]; }
      if (live) group_load<NPX>(gy, q, g);
      const int qn = q + kThreads;
      nlive = t + 1 < t1 && qn < nq;
      if (nlive) {
        group_load<NPX>(x, qn, nx);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i][0] = nx[i][1] = nx[i][2] = 0.f;
      }
    } else {
      if (live) group_load<NPX>(x, q, px);
    }
    // ---- forward sweep: park the input of every step ----
    for (int s = 0; s < A.S; ++s) {
      park_store<NPX>(park + (size_t)s * kArea, tid, px);
      chain_apply_any(fids[s], px, sc[s], NPX);
    }
    if (live) {
      if (y) group_store<NPX>(y, q, px);
      if constexpr (!PF) group_load<NPX>(gy, q, g);
    }
    // ---- backward sweep (dead lanes carry g == 0: every accumulator term is a multiple of g) ----
    for (int s = A.S - 1; s >= 0; --s) {
      park_load<NPX>(park + (size_t)s * kArea, tid, px);
      chain_bwd_any<NPX>(fids[s], px, g, sc[s], &slots[warp][off[s]], lane);
    }
    if (live && gxo) group_store<NPX>(gxo, q, g);
  }

  __syncthreads();
  chain_finish<kChainRec>(A, b, sc, fids, off, slots, chain_smem);
}

// ---- the same whole-chain pass for a chain that is known at COMPILE time ---------------------------------
// When every image runs the same filter sequence (the benchmark chain E,G,W,S+,T,Ct,BW,C of BASELINE configs[1] /
// [4]; ids uniform over the batch) the sequence becomes a template pack and the generality the run-time kernel
// pays for disappears (profiles/r1k_chain_fused_ncu.md: 1 117 instructions/pixel, 65 % issue-active, i-cache 83 %):
//   * no `switch` per step and tile, no merge MOVs, a third of the code (8 + 8 bodies instead of 10 + 10 behind
//     two jump tables);
//   * the parameter-gradient accumulators (40 for the canonical chain) stay in REGISTERS across the whole tile
//     loop; one multi-value butterfly per CTA instead of one per tile and step, no shared-memory slots;
//   * the input of a step that follows a pure per-channel scale (Exposure, WhiteBalance) is not parked: it is
//     recomputed from the scale's own parked input with the same 3 FMULs (bit-identical) -- 6 parking areas
//     instead of 8 for the canonical chain;
//   * a step's backward gets its forward OUTPUT for free (it is the next step's input, already in registers):
//     Gamma's backward no longer recomputes the power.
// Same px_fwd / px_bwd functions, same record / ticket / fp64 finish: y and gx are bit-identical to the
// run-time kernel and to the per-step kernels; parameter gradients agree to reduction order.
template <int... F> struct FidSeq {
  static constexpr int n = sizeof...(F);
  static constexpr int v[sizeof...(F) ? sizeof...(F) : 1] = {F...};
};
template <class Q> __host__ __device__ constexpr bool seq_is_scale(int i) {
  return Q::v[i] == EXP_FILTER_EXPOSURE || Q::v[i] == EXP_FILTER_WB;
}
// step i's input is parked unless step i-1 is a per-channel scale whose own input is parked
template <class Q> __host__ __device__ constexpr bool seq_parked(int i) {
  return i == 0 || !(seq_is_scale<Q>(i - 1) && seq_parked<Q>(i - 1));
}
template <class Q> __host__ __device__ constexpr int seq_slot(int i) {          // parking area index of step i
  int n = 0;
  for (int k = 0; k < i; ++k) n += seq_parked<Q>(k) ? 1 : 0;
  return n;
}
template <class Q> __host__ __device__ constexpr int seq_acc_off(int i) {       // first accumulator of step i (record order)
  int n = 0;
  for (int k = 0; k < i; ++k) n += num_acc(Q::v[k]);
  return n;
}
// Accumulator placement: the curve filters carry 8 (Tone) / 24 (Color) sums -- pinned in registers next to the
// 8 scalar ones they push the kernel past 128 registers (ptxas: 568 B of spills).  They live in per-thread
// SHARED memory instead ([vector][thread] float4 columns, conflict free): a step sums its 4 pixels in temporary
// registers and adds them with one LDS.128 / STS.128 pair per 4 sums and tile.  Sums of fewer than 8 stay in
// registers for the whole tile loop.
__host__ __device__ constexpr bool acc_in_smem(int fid) { return num_acc(fid) >= 8; }
template <class Q> __host__ __device__ constexpr int seq_reg_off(int i) {       // register accumulators before step i
  int n = 0;
  for (int k = 0; k < i; ++k) n += acc_in_smem(Q::v[k]) ? 0 : num_acc(Q::v[k]);
  return n;
}
template <class Q> __host__ __device__ constexpr int seq_sm_off4(int i) {       // float4 vectors of smem accumulators before step i
  int n = 0;
  for (int k = 0; k < i; ++k) n += acc_in_smem(Q::v[k]) ? num_acc(Q::v[k]) / 4 : 0;
  return n;
}

template <class Q, int I, int NPX>
__device__ __forceinline__ void static_fwd_sweep(float (&px)[4][3], float* __restrict__ park, int tid, const FilterConsts* sc) {
  if constexpr (I < Q::n) {
    if constexpr (seq_parked<Q>(I)) park_store<NPX>(park + (size_t)seq_slot<Q>(I) * (3 * NPX * kThreads), tid, px);
    chain_apply<Q::v[I]>(px, sc[I], NPX);
    static_fwd_sweep<Q, I + 1, NPX>(px, park, tid, sc);
  }
}
// yn: the output of step I (= input of step I+1), g: gradient in / out, acc: register accumulators,
// accsm: this thread's column of the shared-memory accumulators
template <class Q, int I, int NPX>
__device__ __forceinline__ void static_bwd_sweep(float (&yn)[4][3], float (&g)[4][3], float* __restrict__ acc,
                                                 float4* __restrict__ accsm, const float* __restrict__ park, int tid,
                                                 const FilterConsts* sc) {
  if constexpr (I >= 0) {
    constexpr int FID = Q::v[I];
    float px[4][3];
    if constexpr (seq_parked<Q>(I)) {
      park_load<NPX>(park + (size_t)seq_slot<Q>(I) * (3 * NPX * kThreads), tid, px);
    } else {                                               // x_I = scale(x_{I-1}): the same FMULs as the forward sweep
      park_load<NPX>(park + (size_t)seq_slot<Q>(I - 1) * (3 * NPX * kThreads), tid, px);
      chain_apply<Q::v[I - 1]>(px, sc[I - 1], NPX);
    }
    if constexpr (acc_in_smem(FID)) {
      constexpr int NA = num_acc(FID);
      float a[NA];
#pragma unroll
      for (int k = 0; k < NA; ++k) a[k] = 0.f;
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        float gx[3];
        px_bwd<FID, true, true>(px[i], g[i], gx, a, sc[I], yn[i]);
        g[i][0] = gx[0]; g[i][1] = gx[1]; g[i][2] = gx[2];
      }
      float4* sa = accsm + (size_t)seq_sm_off4<Q>(I) * kThreads;
#pragma unroll
      for (int v = 0; v < NA / 4; ++v) {
        float4 t = sa[(size_t)v * kThreads];
        t.x += a[4 * v]; t.y += a[4 * v + 1]; t.z += a[4 * v + 2]; t.w += a[4 * v + 3];
        sa[(size_t)v * kThreads] = t;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPX; ++i) {
        float gx[3];
        px_bwd<FID, true, true>(px[i], g[i], gx, acc + seq_reg_off<Q>(I), sc[I], yn[i]);
        g[i][0] = gx[0]; g[i][1] = gx[1]; g[i][2] = gx[2];
      }
    }
    static_bwd_sweep<Q, I - 1, NPX>(px, g, acc, accsm, park, tid, sc);
  }
}
// all accumulators of the chain in record order (registers + this thread's shared-memory column)
template <class Q, int I>
__device__ __forceinline__ void static_collect(float* __restrict__ all, const float* __restrict__ acc, const float4* __restrict__ accsm) {
  if constexpr (I < Q::n) {
    constexpr int FID = Q::v[I];
    constexpr int NA = num_acc(FID);
    if constexpr (acc_in_smem(FID)) {
#pragma unroll
      for (int v = 0; v < NA / 4; ++v) {
        const float4 t = accsm[(size_t)(seq_sm_off4<Q>(I) + v) * kThreads];
        all[seq_acc_off<Q>(I) + 4 * v] = t.x; all[seq_acc_off<Q>(I) + 4 * v + 1] = t.y;
        all[seq_acc_off<Q>(I) + 4 * v + 2] = t.z; all[seq_acc_off<Q>(I) + 4 * v + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < NA; ++k) all[seq_acc_off<Q>(I) + k] = acc[seq_reg_off<Q>(I) + k];
    }
    static_collect<Q, I + 1>(all, acc, accsm);
  }
}

constexpr int kStaticSlots = 64;      // accumulators per warp record of the compile-time kernel (40 for the canonical chain)

template <int NPX, bool PF, int... F>
__global__ void __launch_bounds__(kThreads, 2) filter_chain_static_kernel(const ChainBwdArgs A) {
  using Q = FidSeq<F...>;
  constexpr int S = Q::n;
  constexpr int NACC = seq_acc_off<Q>(S);
  constexpr int NREG = seq_reg_off<Q>(S) > 0 ? seq_reg_off<Q>(S) : 1;
  constexpr int NSM4 = seq_sm_off4<Q>(S);
  constexpr int kArea = 3 * NPX * kThreads;
  static_assert(NACC <= kStaticSlots, "warp record too small");
  extern __shared__ __align__(16) unsigned char chain_smem[];
  __shared__ FilterConsts sc[S];
  __shared__ int fids[S], off[S + 1];
  __shared__ float slots[kWarps][kStaticSlots];
  float* const park = reinterpret_cast<float*>(chain_smem);
  float4* const accsm = reinterpret_cast<float4*>(park + (size_t)seq_slot<Q>(S) * kArea) + threadIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int vf[S] = {F...};                                 // run-time indexable copy of the pack (prologue only)
  for (int s = warp; s < S; s += kWarps)                    // one warp per step: the S set-ups run in parallel
    setup_consts_lane(sc[s], A.params + ((size_t)s * A.B + b) * A.pstride, vf[s], A.logits, lane, A.rg);
  if (tid <= S) {
    int o = 0;
    for (int s = 0; s < tid; ++s) o += num_acc(vf[s]);
    off[tid] = o;
    if (tid < S) fids[tid] = vf[tid];
  }
#pragma unroll
  for (int v = 0; v < NSM4; ++v) accsm[(size_t)v * kThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  const size_t img = (size_t)b * A.P * 3;
  const float* __restrict__ x = A.x + img;
  const float* __restrict__ gy = A.gy + img;
  float* __restrict__ y = A.y ? A.y + img : nullptr;
  float* __restrict__ gxo = A.gx ? A.gx + img : nullptr;
  const int t0 = (int)((long long)blockIdx.x * A.ntiles / A.nblk);
  const int t1 = (int)((long long)(blockIdx.x + 1) * A.ntiles / A.nblk);
  const int nq = A.P / NPX;

  float acc[NREG];
#pragma unroll
  for (int a = 0; a < NREG; ++a) acc[a] = 0.f;

  float nx[4][3];
#pragma unroll
  for (int i = 0; i < 4; ++i) nx[i][0] = nx[i][1] = nx[i][2] = 0.f;
  if constexpr (PF) {
    const int q0 = t0 * kThreads + tid;
    if (t0 < t1 && q0 < nq) group_load<NPX>(x, q0, nx);
  }
  for (int t = t0; t < t1; ++t) {
    const int q = t * kThreads + tid;
    const bool live = q < nq;
    float px[4][3], g[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      px[i][0] = nx[i][0]; px[i][1] = nx[i][1]; px[i][2] = nx[i][2];
      g[i][0] = g[i][1] = g[i][2] = 0.f;
    }
    if constexpr (!PF) {
      if (live) group_load<NPX>(x, q, px);
    }
    if (live) group_load<NPX>(gy, q, g);                     // dL/dy and the NEXT tile's x are requested before the sweep
    if constexpr (PF) {
      const int qn = q + kThreads;
      if (t + 1 < t1 && qn < nq) {
        group_load<NPX>(x, qn, nx);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) nx[i][0] = nx[i][1] = nx[i][2] = 0.f;
      }
    }
    static_fwd_sweep<Q, 0, NPX>(px, park, tid, sc);
    if (live && y) group_store<NPX>(y, q, px);
    // dead lanes carry g == 0: every accumulator term is a multiple of g
    static_bwd_sweep<Q, S - 1, NPX>(px, g, acc, accsm, park, tid, sc);
    if (live && gxo) group_store<NPX>(gxo, q, g);
  }

  // ---- one multi-value butterfly per CTA: rows of <= 8 accumulators -> per-warp slots -> shared finish ----
  float all[NACC];
  static_collect<Q, 0>(all, acc, accsm);
#pragma unroll
  for (int r = 0; r < (NACC + 7) / 8; ++r) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (r * 8 + i < NACC) ? all[r * 8 + i] : 0.f;
    int idx;
    const float tot = warp_multi_sum<8>(v, lane, idx);
    if ((lane & 3) == 0 && r * 8 + idx < NACC) slots[warp][r * 8 + idx] = tot;
  }
  __syncthreads();
  chain_finish<kStaticSlots>(A, b, sc, fids, off, slots, chain_smem);
}

// chains with a compile-time instantiation (the host falls back to the run-time kernel for everything else)
template <int NPX, bool PF, int... F>
static cudaError_t launch_chain_static(const ChainBwdArgs& A, cudaStream_t st) {
  using Q = FidSeq<F...>;
  constexpr size_t smem = (size_t)seq_slot<Q>(Q::n) * 3 * NPX * kThreads * sizeof(float) +
                          (size_t)seq_sm_off4<Q>(Q::n) * kThreads * sizeof(float4);
  auto kern = filter_chain_static_kernel<NPX, PF, F...>;
  static bool attr_set[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  kern<<<dim3(A.nblk, A.B), kThreads, smem, st>>>(A);
  return cudaGetLastError();
}

// ---- filter_param_regressor kernels (one thread per image; math in filter_math.cuh) ------
template <bool BWD>
__global__ void regress_kernel(const float* __restrict__ logits, int lstride, float* __restrict__ params,
                               const float* __restrict__ gparams, int pstride, float* __restrict__ glogits,
                               const int* __restrict__ ids, int uniform_id, int B, const FilterRanges rg) {
  EXP_PDL_ENTRY();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int fid = ids ? ids[b] : uniform_id;
  const float* f = logits + (size_t)b * lstride;
  float* po = BWD ? nullptr : params + (size_t)b * pstride;
  const float* gp = BWD ? gparams + (size_t)b * pstride : nullptr;
  float* gf = BWD ? glogits + (size_t)b * lstride : nullptr;
  if (BWD) for (int i = 0; i < lstride; ++i) gf[i] = 0.f;
  if (fid < 0 || fid >= EXP_NUM_FILTER_KINDS) return;
  regress_image<BWD>(fid, f, po, gp, gf, rg);
}

// ---- host side of the persistent TMA variant ---------------------------------------------
template <int FID, bool BWD, bool HAS_GX>
static int launch_tma_one(const TmaArgs& A0, cudaStream_t st) {
  constexpr int STAGES = 3;
  constexpr size_t smem = (size_t)STAGES * kTileBytes * (BWD ? 2 : 1);
  auto kern = filter_step_tma_kernel<FID, BWD, HAS_GX, STAGES>;
  static int ctas_per_sm[64];          // per device
  static int num_sms[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
    return set_error(EXP_ERR_CUDA, "cudaGetDevice failed");
  if (ctas_per_sm[dev] == 0) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int nb = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kTmaThreads, smem);
    if (e != cudaSuccess || nb < 1) return set_error(EXP_ERR_CUDA, "occupancy query failed: %s", cudaGetErrorString(e));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    num_sms[dev] = sms;
    ctas_per_sm[dev] = nb;
  }
  int grid = num_sms[dev] * ctas_per_sm[dev];
  if (grid > kMaxPersistentCtas) grid = kMaxPersistentCtas;
  if (grid > A0.total_tiles) grid = A0.total_tiles;
  launch_pdl(kern, dim3(grid), dim3(kTmaThreads), smem, st, A0);
  return EXP_OK;
}

template <bool BWD, bool HAS_GX>
static int launch_tma(int fid, const TmaArgs& A, cudaStream_t st) {
  switch (fid) {
#define EXP_CASE(F) case F: return launch_tma_one<F, BWD, HAS_GX>(A, st);
    EXP_CASE(0) EXP_CASE(1) EXP_CASE(2) EXP_CASE(3) EXP_CASE(4) EXP_CASE(5) EXP_CASE(6) EXP_CASE(7) EXP_CASE(8) EXP_CASE(9)
#undef EXP_CASE
  }
  return set_error(EXP_ERR_INVALID_ARG, "bad filter id %d", fid);
}

static TmaArgs make_tma_args(const float* x, const float* gy, float* out, const float* params, int pstride,
                             int B, int P) {
  TmaArgs T{};
  T.x = x; T.gy = gy; T.out = out; T.params = params; T.pstride = pstride; T.P = P; T.B = B;
  T.rg = host_ranges();
  T.tiles_per_image = (P + kTilePx - 1) / kTilePx;
  T.total_tiles = B * T.tiles_per_image;
  return T;
}

static int check_common(const void* x, const float* params, int pstride, const int* ids, int uniform_id,
                        int B, int H, int W) {
  EXP_CHECK_ARG(x && params, "null image or params pointer");
  EXP_CHECK_ARG(B > 0 && H > 0 && W > 0, "bad shape B=%d H=%d W=%d", B, H, W);
  EXP_CHECK_ARG((long long)H * W < (1ll << 29), "image too large: %dx%d", H, W);
  EXP_CHECK_ARG(B <= 65535, "B=%d exceeds gridDim.y limit 65535", B);
  EXP_CHECK_ARG(pstride >= 1, "pstride=%d", pstride);
  if (!ids) {
    EXP_CHECK_ARG(uniform_id >= 0 && uniform_id < EXP_NUM_FILTER_KINDS, "bad filter id %d", uniform_id);
    EXP_CHECK_ARG(pstride >= num_params(uniform_id), "pstride=%d < %d params of filter %d", pstride,
                  num_params(uniform_id), uniform_id);
  } else {
    EXP_CHECK_ARG(pstride >= EXP_MAX_FILTER_PARAMS, "per-image ids need pstride >= %d (got %d)",
                  EXP_MAX_FILTER_PARAMS, pstride);
  }
  return EXP_OK;
}

// decide vector (4 px / thread) vs scalar path
static int pick_vec(int variant, int P, const void* a, const void* b, const void* c, bool* vec) {
  const bool ok = (P % 4 == 0) && aligned16(a) && (!b || aligned16(b)) && (!c || aligned16(c));
  if (variant == EXP_VARIANT_SCALAR) { *vec = false; return EXP_OK; }
  if (variant == EXP_VARIANT_DIRECT || variant == EXP_VARIANT_TMA) {
    if (!ok) return set_error(EXP_ERR_ALIGNMENT,
                              "vector variants need H*W %% 4 == 0 and 16-byte aligned images (H*W=%d)", P);
    *vec = true;
    return EXP_OK;
  }
  if (variant != EXP_VARIANT_AUTO) return set_error(EXP_ERR_INVALID_ARG, "unknown variant %d", variant);
  *vec = ok;
  return EXP_OK;
}

}  // namespace expo

using namespace expo;

extern "C" {

int exp_version(void) { return 1; }
int exp_set_pdl(int enable) {
  set_pdl(enable ? 1 : 0);
  return EXP_OK;
}
const char* exp_last_error(void) { return last_error_buf(); }
int exp_set_filter_ranges(const exp_filter_ranges* r, int curve_steps) {
  if (!r) { g_ranges = default_ranges(); return EXP_OK; }
  if (curve_steps != kCurveSteps)
    return set_error(EXP_ERR_UNSUPPORTED, "cfg.curve_steps = %d: the curve kernels are built for %d knots", curve_steps, kCurveSteps);
  EXP_CHECK_ARG(r->exposure_range > 0.f && r->gamma_range > 1.f, "exposure_range must be > 0 and gamma_range > 1");
  EXP_CHECK_ARG(r->tone_lo < r->tone_hi && r->color_lo < r->color_hi, "curve ranges must be (lo, hi) with lo < hi");
  EXP_CHECK_ARG(r->color_lo < 1.f && 1.f < r->color_hi, "color_curve_range must contain its initial value 1 (util.py:285-286)");
  FilterRanges g;
  g.exposure = r->exposure_range;
  g.gamma_log = (float)log((double)r->gamma_range);        // filters.py:202 np.log(cfg.gamma_range)
  g.tone_lo = r->tone_lo; g.tone_hi = r->tone_hi;
  g.color_lo = r->color_lo; g.color_hi = r->color_hi;
  // util.py:285-286 (python doubles): bias = atanh(2 (initial - l)/(r - l) - 1), initial = 1
  g.color_bias = (float)atanh(2.0 * (1.0 - (double)r->color_lo) / ((double)r->color_hi - (double)r->color_lo) - 1.0);
  if (fabsf(g.color_bias) < 1e-7f) g.color_bias = 0.f;     // ranges centred on 1 (every shipped config)
  g_ranges = g;
  return EXP_OK;
}
int exp_get_filter_ranges(exp_filter_ranges* r) {
  EXP_CHECK_ARG(r, "null pointer");
  r->exposure_range = g_ranges.exposure; r->gamma_range = (float)exp((double)g_ranges.gamma_log);
  r->tone_lo = g_ranges.tone_lo; r->tone_hi = g_ranges.tone_hi; r->color_lo = g_ranges.color_lo; r->color_hi = g_ranges.color_hi;
  return EXP_OK;
}
int exp_num_filter_params(int fid) {
  if (fid < 0 || fid >= EXP_NUM_FILTER_KINDS) return set_error(EXP_ERR_INVALID_ARG, "bad filter id %d", fid);
  return num_params(fid);
}

int exp_filter_regress_fwd(const float* logits, int lstride, float* params, int pstride, const int* ids,
                           int uniform_id, int B, void* stream) {
  EXP_CHECK_ARG(logits && params && B > 0, "null pointer or B=%d", B);
  const int need = ids ? EXP_MAX_FILTER_PARAMS : exp_num_filter_params(uniform_id);
  if (need < 0) return need;
  EXP_CHECK_ARG(lstride >= need && pstride >= need, "strides (%d,%d) < %d", lstride, pstride, need);
  launch_pdl(regress_kernel<false>, dim3((B + 127) / 128), dim3(128), 0, (cudaStream_t)stream,
      logits, lstride, params, nullptr, pstride, nullptr, ids, uniform_id, B, host_ranges());
  EXP_CHECK_LAUNCH("exp_filter_regress_fwd");
  return EXP_OK;
}

int exp_filter_regress_bwd(const float* logits, int lstride, const float* gparams, int pstride,
                           float* glogits, const int* ids, int uniform_id, int B, void* stream) {
  EXP_CHECK_ARG(logits && gparams && glogits && B > 0, "null pointer or B=%d", B);
  const int need = ids ? EXP_MAX_FILTER_PARAMS : exp_num_filter_params(uniform_id);
  if (need < 0) return need;
  EXP_CHECK_ARG(lstride >= need && pstride >= need, "strides (%d,%d) < %d", lstride, pstride, need);
  launch_pdl(regress_kernel<true>, dim3((B + 127) / 128), dim3(128), 0, (cudaStream_t)stream,
      logits, lstride, nullptr, gparams, pstride, glogits, ids, uniform_id, B, host_ranges());
  EXP_CHECK_LAUNCH("exp_filter_regress_bwd");
  return EXP_OK;
}

int exp_filter_fwd(const float* x, float* y, const float* params, int pstride, const int* ids,
                   int uniform_id, int B, int H, int W, int options, void* stream) {
  const int logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  int variant = options & 0xFF;
  int rc = check_common(x, params, pstride, ids, uniform_id, B, H, W);
  if (rc) return rc;
  EXP_CHECK_ARG(y, "null output pointer");
  const int P = H * W;
  bool vec;
  rc = pick_vec(variant, P, x, y, nullptr, &vec);
  if (rc) return rc;
  if (variant == EXP_VARIANT_AUTO && vec && !ids) variant = EXP_VARIANT_TMA;   // measured faster
  if (variant == EXP_VARIANT_TMA) {
    if (ids) return set_error(EXP_ERR_UNSUPPORTED, "the TMA variant needs a uniform filter id");
    if ((long long)B * ((P + kTilePx - 1) / kTilePx) > 0x7fffffffll) return set_error(EXP_ERR_UNSUPPORTED, "too many tiles");
    TmaArgs T = make_tma_args(x, nullptr, y, params, pstride, B, P);
    T.logits = logits;
    rc = launch_tma<false, false>(uniform_id, T, (cudaStream_t)stream);
    if (rc) return rc;
    EXP_CHECK_LAUNCH("exp_filter_fwd[tma]");
    return EXP_OK;
  }
  FilterArgs A{};
  A.rg = host_ranges();
  A.x = x; A.out = y; A.params = params; A.pstride = pstride; A.ids = ids; A.P = P;
  A.pix_per_block = kPixPerBlockFwd;
  A.logits = logits;
  dim3 grid((P + kPixPerBlockFwd - 1) / kPixPerBlockFwd, B);
  launch_step<false, false>(vec, ids, uniform_id, grid, (cudaStream_t)stream, A);
  EXP_CHECK_LAUNCH("exp_filter_fwd");
  return EXP_OK;
}

int exp_filter_chain_fwd(const float* x, float* y, const float* params, int pstride, const int* ids, int S, int B,
                         int H, int W, int options, void* stream) {
  EXP_CHECK_ARG(x && y && params && ids, "null pointer");
  EXP_CHECK_ARG(S >= 1 && S <= kMaxChain, "S must be in [1, %d] (got %d)", kMaxChain, S);
  EXP_CHECK_ARG(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 29), "bad shape");
  EXP_CHECK_ARG(pstride >= EXP_MAX_FILTER_PARAMS, "per-image ids need pstride >= %d (got %d)", EXP_MAX_FILTER_PARAMS, pstride);
  const int P = H * W;
  bool vec;
  int variant = options & 0xFF;
  if (variant == EXP_VARIANT_TMA) variant = EXP_VARIANT_DIRECT;
  int rc = pick_vec(variant, P, x, y, nullptr, &vec);
  if (rc) return rc;
  ChainArgs A{};
  A.rg = host_ranges();
  A.x = x; A.y = y; A.params = params; A.ids = ids; A.S = S; A.B = B; A.P = P; A.pstride = pstride;
  A.logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  // The per-CTA set-up (tanh_range of up to 24 parameters per step) costs as much as ~2000 pixels of filter math, so a
  // CTA takes a long run of pixels: ~8 waves of resident CTAs over the whole batch, never less than kPixPerBlockFwd.
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int per_image = std::max(1, (sms * 8 * 4 + B - 1) / B);        // 8 CTAs of 40 registers per SM, ~4 waves
  A.pix_per_block = (((P + per_image - 1) / per_image + 1023) / 1024) * 1024;
  if (A.pix_per_block < kPixPerBlockFwd) A.pix_per_block = kPixPerBlockFwd;
  dim3 grid((P + A.pix_per_block - 1) / A.pix_per_block, B);
  if (vec) filter_chain_fwd_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(A);
  else filter_chain_fwd_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(A);
  EXP_CHECK_LAUNCH("exp_filter_chain_fwd");
  return EXP_OK;
}

// CTAs per image of the fused chain: ~8 waves of 2 CTAs/SM over the whole batch, at most one per tile
static int chain_nblk(int B, int ntiles) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const int waves = [] { const char* e = getenv("EXPOSURE_CHAIN_WAVES"); return e && atoi(e) > 0 ? atoi(e) : 8; }();   // tuning aid
  // measured at 64x512x512, run-time kernel (round 1): 4 waves 0.844 ms (tail), 8 waves 0.810, 16 waves 0.907 (set-up per
  // CTA); compile-time kernel (round 2): 4 waves 0.534, 6 waves 0.524, 8 waves 0.523, 12 waves 0.554
  const int target = sms * 2 * waves;
  int nblk = (target + B - 1) / B;
  if (nblk > ntiles) nblk = ntiles;
  if (nblk > 65535) nblk = 65535;
  return nblk < 1 ? 1 : nblk;
}
static int chain_ntiles(int P, int npx) { return (P / npx + kThreads - 1) / kThreads; }
size_t exp_filter_chain_fwd_bwd_workspace_bytes(int S, int B, int H, int W) {
  if (S <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  const int P = H * W;
  int nblk = 1;
  for (int npx = 1; npx <= 4; npx *= 4) {                 // whichever variant the launch ends up using
    if (P % npx) break;
    const int n = chain_nblk(B, chain_ntiles(P, npx));
    if (n > nblk) nblk = n;
  }
  return kCounterBytes + (size_t)B * nblk * kChainRec * sizeof(float);
}

static int chain_fwd_bwd_impl(const float* x, const float* gy, float* y, float* gx, const float* params, int pstride,
                              const int* ids, const int* ids_host, int S, int B, int H, int W, float* gparams,
                              void* workspace, size_t workspace_bytes, int options, void* stream, const char* what) {
  EXP_CHECK_ARG(x && gy && params && (ids || ids_host) && gparams && workspace, "null pointer");
  EXP_CHECK_ARG(S >= 1 && S <= kMaxChain, "S must be in [1, %d] (got %d)", kMaxChain, S);
  EXP_CHECK_ARG(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 29), "bad shape");
  EXP_CHECK_ARG(pstride >= EXP_MAX_FILTER_PARAMS, "per-image ids need pstride >= %d (got %d)", EXP_MAX_FILTER_PARAMS, pstride);
  const size_t need = exp_filter_chain_fwd_bwd_workspace_bytes(S, B, H, W);
  if (workspace_bytes < need)
    return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (!aligned16(workspace)) return set_error(EXP_ERR_ALIGNMENT, "workspace must be 16-byte aligned");
  const int P = H * W;
  bool vec;
  int variant = options & 0xFF;
  if (variant == EXP_VARIANT_TMA) variant = EXP_VARIANT_DIRECT;
  int rc = pick_vec(variant, P, x, gy, gx, &vec);
  if (rc) return rc;
  if (vec && y && !aligned16(y)) {
    if (variant == EXP_VARIANT_DIRECT) return set_error(EXP_ERR_ALIGNMENT, "y must be 16-byte aligned for the DIRECT variant");
    vec = false;
  }
  ChainBwdArgs A{};
  A.rg = host_ranges();
  A.x = x; A.gy = gy; A.y = y; A.gx = gx; A.params = params; A.ids = ids; A.gparams = gparams;
  A.counters = reinterpret_cast<unsigned*>(workspace);
  A.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
  A.S = S; A.B = B; A.P = P; A.pstride = pstride; A.logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  const int npx = vec ? 4 : 1;
  A.ntiles = chain_ntiles(P, npx);
  A.nblk = chain_nblk(B, A.ntiles);
  if (ids_host) {
    bool canonical = S == EXP_NUM_FILTERS;
    for (int s = 0; s < S; ++s) {
      EXP_CHECK_ARG(ids_host[s] >= -1 && ids_host[s] < EXP_NUM_FILTER_KINDS, "bad filter id %d at step %d", ids_host[s], s);
      A.uniform_ids[s] = ids_host[s];
      canonical = canonical && ids_host[s] == s;
    }
    // the shipped cfg.filters order E,G,W,S+,T,Ct,BW,C (config_example.py:22-25) has a compile-time instantiation
    if (canonical && vec && !(options & EXP_OPT_NO_STATIC_CHAIN)) {
      // PF (next tile's x requested before the sweep, +12 live registers) measured on B200 at 64x512x512:
      // 0.560 ms with, 0.523 ms without (128 registers either way; the prefetch costs spills) -> off
      const cudaError_t e = launch_chain_static<4, false, 0, 1, 2, 3, 4, 5, 6, 7>(A, (cudaStream_t)stream);
      if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "%s[static]: %s", what, cudaGetErrorString(e));
      return EXP_OK;
    }
  }
  const size_t smem = (size_t)S * 3 * npx * kThreads * sizeof(float);
  auto kern = npx == 4 ? filter_chain_fwd_bwd_kernel<4, true> : filter_chain_fwd_bwd_kernel<1, false>;
  static bool attr_set[2][64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return set_error(EXP_ERR_CUDA, "cudaGetDevice failed");
  if (!attr_set[npx >> 2][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kMaxChain * 12 * kThreads * sizeof(float)));
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set[npx >> 2][dev] = true;
  }
  kern<<<dim3(A.nblk, B), kThreads, smem, (cudaStream_t)stream>>>(A);
  EXP_CHECK_LAUNCH(what);
  return EXP_OK;
}

int exp_filter_chain_fwd_bwd(const float* x, const float* gy, float* y, float* gx, const float* params, int pstride,
                             const int* ids, int S, int B, int H, int W, float* gparams, void* workspace,
                             size_t workspace_bytes, int options, void* stream) {
  EXP_CHECK_ARG(ids, "null ids pointer");
  return chain_fwd_bwd_impl(x, gy, y, gx, params, pstride, ids, nullptr, S, B, H, W, gparams, workspace, workspace_bytes,
                            options, stream, "exp_filter_chain_fwd_bwd");
}

int exp_filter_chain_fwd_bwd_uniform(const float* x, const float* gy, float* y, float* gx, const float* params, int pstride,
                                     const int* ids_host, int S, int B, int H, int W, float* gparams, void* workspace,
                                     size_t workspace_bytes, int options, void* stream) {
  EXP_CHECK_ARG(ids_host, "null ids_host pointer");
  return chain_fwd_bwd_impl(x, gy, y, gx, params, pstride, nullptr, ids_host, S, B, H, W, gparams, workspace,
                            workspace_bytes, options, stream, "exp_filter_chain_fwd_bwd_uniform");
}

size_t exp_filter_bwd_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const size_t nblk = ((size_t)H * W + kPixPerBlockBwd - 1) / kPixPerBlockBwd;
  size_t recs = (size_t)B * nblk;
  if (recs < (size_t)B + kMaxPersistentCtas) recs = (size_t)B + kMaxPersistentCtas;
  return kCounterBytes + recs * kAccStride * sizeof(float);
}

int exp_filter_bwd(const float* x, const float* gy, float* gx, float* gparams, const float* params,
                   int pstride, const int* ids, int uniform_id, int B, int H, int W, void* workspace,
                   size_t workspace_bytes, int options, void* stream) {
  const int logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  int variant = options & 0xFF;
  int rc = check_common(x, params, pstride, ids, uniform_id, B, H, W);
  if (rc) return rc;
  EXP_CHECK_ARG(gy && gparams && workspace, "null gy / gparams / workspace pointer");
  const size_t need = exp_filter_bwd_workspace_bytes(B, H, W);
  if (workspace_bytes < need)
    return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (!aligned16(workspace)) return set_error(EXP_ERR_ALIGNMENT, "workspace must be 16-byte aligned");
  const int P = H * W;
  bool vec;
  rc = pick_vec(variant, P, x, gy, gx, &vec);
  if (rc) return rc;
  if (variant == EXP_VARIANT_AUTO && vec && !ids) variant = EXP_VARIANT_TMA;
  if (variant == EXP_VARIANT_TMA) {
    if (ids) return set_error(EXP_ERR_UNSUPPORTED, "the TMA variant needs a uniform filter id");
    TmaArgs T = make_tma_args(x, gy, gx, params, pstride, B, P);
    T.counters = reinterpret_cast<unsigned*>(workspace);
    T.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
    T.gparams = gparams;
    T.logits = logits;
    rc = gx ? launch_tma<true, true>(uniform_id, T, (cudaStream_t)stream)
            : launch_tma<true, false>(uniform_id, T, (cudaStream_t)stream);
    if (rc) return rc;
    EXP_CHECK_LAUNCH("exp_filter_bwd[tma]");
    return EXP_OK;
  }
  const int nblk = (P + kPixPerBlockBwd - 1) / kPixPerBlockBwd;
  FilterArgs A{};
  A.rg = host_ranges();
  A.x = x; A.gy = gy; A.out = gx; A.params = params; A.pstride = pstride; A.ids = ids; A.P = P;
  A.pix_per_block = kPixPerBlockBwd;
  A.logits = logits;
  A.counters = reinterpret_cast<unsigned*>(workspace);
  A.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
  A.gparams = gparams;
  dim3 grid(nblk, B);
  if (gx) launch_step<true, true>(vec, ids, uniform_id, grid, (cudaStream_t)stream, A);
  else launch_step<true, false>(vec, ids, uniform_id, grid, (cudaStream_t)stream, A);
  EXP_CHECK_LAUNCH("exp_filter_bwd");
  return EXP_OK;
}

/* ---- masked step: Filter.apply with cfg.masking == True ------------------------------------ */
static void fill_mask_args(FilterArgs& A, const float* mask_logits, int mstride, int H, int W, int uniform_id,
                           float max_sharpness, float min_strength, int masking) {
  A.mask_logits = mask_logits; A.mstride = mstride; A.H = H; A.W = W; A.uniform_id = uniform_id;
  A.max_sharp = max_sharpness; A.min_strength = min_strength; A.masking = masking ? 1 : 0;
}

int exp_filter_masked_fwd(const float* x, float* y, float* mask_out, const float* params, int pstride,
                          const float* mask_logits, int mstride, const int* ids, int uniform_id, int B, int H,
                          int W, float max_sharpness, float min_strength, int masking, int options, void* stream) {
  EXP_CHECK_ARG(x && (y || mask_out), "null image pointer, or neither y nor mask_out given");
  EXP_CHECK_ARG(B > 0 && H > 0 && W > 0 && B <= 65535 && (long long)H * W < (1ll << 29), "bad shape B=%d H=%d W=%d", B, H, W);
  EXP_CHECK_ARG(!mask_logits || mstride >= kMaskParams, "mstride=%d < %d", mstride, kMaskParams);
  if (!ids) EXP_CHECK_ARG(uniform_id >= 0 && uniform_id < EXP_NUM_FILTER_KINDS, "bad filter id %d", uniform_id);
  const int P = H * W;
  FilterArgs A{};
  A.rg = host_ranges();
  A.x = x; A.out = y; A.params = params; A.pstride = pstride; A.ids = ids; A.P = P;
  A.pix_per_block = kPixPerBlockFwd;
  A.logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  A.mask_out = mask_out;
  fill_mask_args(A, mask_logits, mstride, H, W, uniform_id, max_sharpness, min_strength, masking);
  dim3 grid((P + kPixPerBlockFwd - 1) / kPixPerBlockFwd, B);
  if (!y) {                                   // get_mask only
    mask_only_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(A);
    EXP_CHECK_LAUNCH("exp_filter_masked_fwd[mask only]");
    return EXP_OK;
  }
  int rc = check_common(x, params, pstride, ids, uniform_id, B, H, W);
  if (rc) return rc;
  bool vec;
  int variant = options & 0xFF;
  if (variant == EXP_VARIANT_TMA) variant = EXP_VARIANT_DIRECT;
  rc = pick_vec(variant, P, x, y, nullptr, &vec);
  if (rc) return rc;
  if (vec) launch_pdl(filter_step_masked_kernel<false, false, true>, grid, dim3(kThreads), 0, (cudaStream_t)stream, A);
  else launch_pdl(filter_step_masked_kernel<false, false, false>, grid, dim3(kThreads), 0, (cudaStream_t)stream, A);
  EXP_CHECK_LAUNCH("exp_filter_masked_fwd");
  return EXP_OK;
}

int exp_filter_masked_bwd(const float* x, const float* gy, float* gx, float* gparams, float* gmask_logits,
                          const float* params, int pstride, const float* mask_logits, int mstride, const int* ids,
                          int uniform_id, int B, int H, int W, float max_sharpness, float min_strength, int masking,
                          void* workspace, size_t workspace_bytes, int options, void* stream) {
  int rc = check_common(x, params, pstride, ids, uniform_id, B, H, W);
  if (rc) return rc;
  EXP_CHECK_ARG(gy && gparams && gmask_logits && workspace, "null gy / gparams / gmask_logits / workspace pointer");
  EXP_CHECK_ARG(mstride >= kMaskParams, "mstride=%d < %d", mstride, kMaskParams);
  const size_t need = exp_filter_bwd_workspace_bytes(B, H, W);
  if (workspace_bytes < need)
    return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (!aligned16(workspace)) return set_error(EXP_ERR_ALIGNMENT, "workspace must be 16-byte aligned");
  const int P = H * W;
  bool vec;
  int variant = options & 0xFF;
  if (variant == EXP_VARIANT_TMA) variant = EXP_VARIANT_DIRECT;
  rc = pick_vec(variant, P, x, gy, gx, &vec);
  if (rc) return rc;
  const int nblk = (P + kPixPerBlockBwd - 1) / kPixPerBlockBwd;
  FilterArgs A{};
  A.rg = host_ranges();
  A.x = x; A.gy = gy; A.out = gx; A.params = params; A.pstride = pstride; A.ids = ids; A.P = P;
  A.pix_per_block = kPixPerBlockBwd;
  A.logits = (options & EXP_OPT_LOGITS) ? 1 : 0;
  A.counters = reinterpret_cast<unsigned*>(workspace);
  A.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
  A.gparams = gparams;
  A.gmask = gmask_logits;
  fill_mask_args(A, mask_logits, mstride, H, W, uniform_id, max_sharpness, min_strength, masking);
  dim3 grid(nblk, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (gx) {
    if (vec) launch_pdl(filter_step_masked_kernel<true, true, true>, grid, dim3(kThreads), 0, st, A);
    else launch_pdl(filter_step_masked_kernel<true, true, false>, grid, dim3(kThreads), 0, st, A);
  } else {
    if (vec) launch_pdl(filter_step_masked_kernel<true, false, true>, grid, dim3(kThreads), 0, st, A);
    else launch_pdl(filter_step_masked_kernel<true, false, false>, grid, dim3(kThreads), 0, st, A);
  }
  EXP_CHECK_LAUNCH("exp_filter_masked_bwd");
  return EXP_OK;
}

}  // extern "C"
