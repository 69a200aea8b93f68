// Convolution primitives on the TMA-fed tcgen05 engine (tc_engine_tma.cuh, backend 4).
//
// Reference semantics: the 4x4 stride-2 SAME convolutions of agent.py:12-41 / critic.py (ly.conv2d)
// and their gradients, as in nn.cu; this file only changes HOW the operands reach the tensor core:
//
//   fprop  y[(b,oy,ox), co] = sum_{tap,ci} x[b, 2oy-1+ky, 2ox-1+kx, ci] W[tap,ci,co]
//          A K-major : one 4-D box {32 ci, 2tw, 2th, tb} per K step, element strides {1,2,2,1}
//          B MN-major: W[K][Cout] boxes {32 co, 32 k}
//   dgrad  dx[(b,a,c) of parity class z, ci] = sum_{t,co} dy[b, a+oy_off, c+ox_off, co] W[tap(t,z),ci,co]
//          A K-major : dense 4-D box {32 co, tw, th, tb};  B K-major: 3-D box {32 co, BN ci, 1 tap}
//   wgrad  gW[(tap,ci), co] = sum_p x[p shifted by tap, ci] dy[p, co]        (K = pixels, split over z)
//          A MN-major: four boxes {32 ci, 2tw, 2th, tb} (32 pixels each);  B MN-major: {32 co, 32 pixels}
//
// Zero padding, image borders and batch / channel tails are the TMA out-of-bounds fill.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>

#include "common.cuh"
#include "nn_tc.h"
#include "tc_engine_tma.cuh"

namespace expo {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// fp32 tensor map; dims / box / estr innermost first, strides in bytes for dims 1..rank-1
bool make_map(CUtensorMap* m, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
              const cuuint32_t* box, const cuuint32_t* estr, bool mn_major) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }
int imin(int a, int b) { return a < b ? a : b; }

// pixel tile of `count` output pixels over an OH x OW grid: tw x th x tb
struct PixTile { int tw, th, tb; };
PixTile pix_tile(int count, int OH, int OW) {
  PixTile t;
  t.tw = imin(OW, count);
  t.th = imin(OH, count / t.tw);
  t.tb = count / (t.tw * t.th);
  return t;
}

// NHWC activation map with the stride-2 im2col box of `count` output pixels
bool make_im2col_map(CUtensorMap* m, const float* x, int B, int IH, int IW, int C, int count, bool mn_major) {
  const PixTile t = pix_tile(count, IH / 2, IW / 2);
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)IW * C * 4, (cuuint64_t)IH * IW * C * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)(2 * t.tw), (cuuint32_t)(2 * t.th), (cuuint32_t)t.tb};
  const cuuint32_t estr[4] = {1, 2, 2, 1};
  return make_map(m, x, 4, dims, strides, box, estr, mn_major);
}

__device__ __forceinline__ float lrelu(float v) { return 0.6f * v + 0.4f * fabsf(v); }
__device__ __forceinline__ float dlrelu(float a) { return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f); }

}  // namespace

// ------------------------------------------------------------------------------------------
struct TmaConvFprop {
  CUtensorMap ta, tb;
  const float* bias; const float* mask_ref; const float* post_mul; float* y; float* y2;
  int M, Cin, Cout, OW, lgOW, lgOHW, mode, splits;
  static constexpr bool kAMn = false, kBMn = true;
  __device__ int k_iters(int) const { return 16 * Cin / tma::kBK; }
  template <int BN>
  __device__ void load(int ki, int, int m0, int n0, unsigned char* a_dst, unsigned char* b_dst, uint64_t* bar) const {
    const int k = ki * tma::kBK;
    const int tap = k / Cin, ci0 = k - tap * Cin;
    const int b0 = m0 >> lgOHW, rem = m0 & ((1 << lgOHW) - 1);
    const int oy0 = rem >> lgOW, ox0 = rem & (OW - 1);
    tma::tma_load_4d(a_dst, &ta, ci0, 2 * ox0 - 1 + (tap & 3), 2 * oy0 - 1 + (tap >> 2), b0, bar);
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) tma::tma_load_2d(b_dst + j * 4096, &tb, n0 + 32 * j, k, bar);
  }
  __device__ void store4(int, int m, int n, float4 o) const {
    if (m >= M || n >= Cout) return;                           // Cout % 32 == 0: whole groups of 4
    const size_t idx = (size_t)m * Cout + n;
    if (mode == 0) {
      if (bias) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n));
        o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
      }
      o.x = lrelu(o.x); o.y = lrelu(o.y); o.z = lrelu(o.z); o.w = lrelu(o.w);
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(mask_ref + idx));
      o.x *= dlrelu(a.x); o.y *= dlrelu(a.y); o.z *= dlrelu(a.z); o.w *= dlrelu(a.w);
    }
    *reinterpret_cast<float4*>(y + idx) = o;
    if (y2) {
      const float4 pm = __ldg(reinterpret_cast<const float4*>(post_mul + idx));
      *reinterpret_cast<float4*>(y2 + idx) = make_float4(o.x * pm.x, o.y * pm.y, o.z * pm.z, o.w * pm.w);
    }
  }
  __device__ void store16(int z, int m, int n0, const float (&v)[16]) const {
#pragma unroll
    for (int g = 0; g < 4; ++g) store4(z, m, n0 + 4 * g, make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
  }
};

bool tma_conv_fwd_supported(const float* x, int Cx, int Cv, float shift, const float* W, const float* bias,
                            const float* mask_ref, const float* post_mul, const float* y, const float* y2, int Cout) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return encode_fn() && Cv == 0 && shift == 0.f && Cx % 32 == 0 && Cout % 32 == 0 && al(x) && al(W) && al(bias) &&
         al(mask_ref) && al(post_mul) && al(y) && al(y2);
}

cudaError_t tma_conv_fwd(const float* x, int Cx, const float* W, const float* bias, const float* mask_ref,
                         const float* post_mul, float* y, float* y2, int B, int IH, int IW, int Cout, int mode,
                         cudaStream_t st) {
  TmaConvFprop p{};
  const int OH = IH / 2, OW = IW / 2;
  if (!make_im2col_map(&p.ta, x, B, IH, IW, Cx, tma::kBM, false)) return cudaErrorInvalidValue;
  const cuuint64_t wd[2] = {(cuuint64_t)Cout, (cuuint64_t)16 * Cx};
  const cuuint64_t ws[1] = {(cuuint64_t)Cout * 4};
  const cuuint32_t wb[2] = {32, 32}, we[2] = {1, 1};
  if (!make_map(&p.tb, W, 2, wd, ws, wb, we, true)) return cudaErrorInvalidValue;
  p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
  p.M = B * OH * OW; p.Cin = Cx; p.Cout = Cout; p.OW = OW; p.lgOW = ilog2(OW); p.lgOHW = ilog2(OH * OW); p.mode = mode;
  const int mt = (p.M + tma::kBM - 1) / tma::kBM, ki = 16 * Cx / tma::kBK;
  if (Cout % 128 == 0) {
    p.splits = tma::pick_splits(mt * (Cout / 128), ki, 128);
    return tma::launch_tma_gemm<TmaConvFprop, 128>(p, p.M, Cout, p.splits, st);
  }
  if (Cout % 64 == 0) {
    p.splits = tma::pick_splits(mt * (Cout / 64), ki, 64);
    return tma::launch_tma_gemm<TmaConvFprop, 64>(p, p.M, Cout, p.splits, st);
  }
  p.splits = tma::pick_splits(mt * (Cout / 32), ki, 32);
  return tma::launch_tma_gemm<TmaConvFprop, 32>(p, p.M, Cout, p.splits, st);
}

// ------------------------------------------------------------------------------------------
struct TmaConvDgrad {
  CUtensorMap ta, tb;
  const float* a_in; float* dx;
  int M, IH, IW, Cin, Cout, lgW2, lgHW2, splits;
  static constexpr bool kAMn = false, kBMn = false;
  __device__ int k_iters(int) const { return 4 * Cout / tma::kBK; }
  template <int BN>
  __device__ void load(int ki, int z, int m0, int n0, unsigned char* a_dst, unsigned char* b_dst, uint64_t* bar) const {
    const int py = z >> 1, px = z & 1;
    const int k0 = ki * tma::kBK;
    const int t = k0 / Cout, co0 = k0 - t * Cout;
    const int j = t >> 1, l = t & 1;
    const int ky = py == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
    const int kx = px == 0 ? (l == 0 ? 1 : 3) : (l == 0 ? 0 : 2);
    const int oy_off = py == 0 ? (j == 0 ? 0 : -1) : (j == 0 ? 1 : 0);
    const int ox_off = px == 0 ? (l == 0 ? 0 : -1) : (l == 0 ? 1 : 0);
    const int b0 = m0 >> lgHW2, rem = m0 & ((1 << lgHW2) - 1);
    const int a0 = rem >> lgW2, c0 = rem & ((IW / 2) - 1);
    tma::tma_load_4d(a_dst, &ta, co0, c0 + ox_off, a0 + oy_off, b0, bar);
    tma::tma_load_3d(b_dst, &tb, co0, n0, ky * 4 + kx, bar);
  }
  __device__ void store4(int z, int m, int n, float4 o) const {
    if (m >= M || n >= Cin) return;                            // Cin % 4 == 0
    const int py = z >> 1, px = z & 1;
    const int b = m >> lgHW2, rem = m & ((1 << lgHW2) - 1);
    const int iy = 2 * (rem >> lgW2) + py, ix = 2 * (rem & ((IW / 2) - 1)) + px;
    const size_t idx = ((size_t)(b * IH + iy) * IW + ix) * Cin + n;
    if (a_in) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(a_in + idx));
      o.x *= dlrelu(a.x); o.y *= dlrelu(a.y); o.z *= dlrelu(a.z); o.w *= dlrelu(a.w);
    }
    *reinterpret_cast<float4*>(dx + idx) = o;
  }
  __device__ void store16(int z, int m, int n0, const float (&v)[16]) const {
#pragma unroll
    for (int g = 0; g < 4; ++g) store4(z, m, n0 + 4 * g, make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]));
  }
};

bool tma_conv_dgrad_supported(const float* dy, const float* W, const float* a_in, const float* dx, int Cin, int Cout) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return encode_fn() && Cout % 32 == 0 && Cin % 4 == 0 && al(dy) && al(W) && al(a_in) && al(dx);
}

cudaError_t tma_conv_dgrad(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                           int Cout, cudaStream_t st) {
  TmaConvDgrad p{};
  const int OH = IH / 2, OW = IW / 2;
  const PixTile t = pix_tile(tma::kBM, OH, OW);
  const cuuint64_t ad[4] = {(cuuint64_t)Cout, (cuuint64_t)OW, (cuuint64_t)OH, (cuuint64_t)B};
  const cuuint64_t as[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)OW * Cout * 4, (cuuint64_t)OH * OW * Cout * 4};
  const cuuint32_t ab[4] = {32, (cuuint32_t)t.tw, (cuuint32_t)t.th, (cuuint32_t)t.tb}, ae[4] = {1, 1, 1, 1};
  if (!make_map(&p.ta, dy, 4, ad, as, ab, ae, false)) return cudaErrorInvalidValue;
  const int N = Cin;
  const int BN = N > 64 ? 128 : (N > 32 ? 64 : 32);
  const cuuint64_t wd[3] = {(cuuint64_t)Cout, (cuuint64_t)Cin, 16};
  const cuuint64_t ws[2] = {(cuuint64_t)Cout * 4, (cuuint64_t)Cin * Cout * 4};
  const cuuint32_t wb[3] = {32, (cuuint32_t)BN, 1}, we[3] = {1, 1, 1};
  if (!make_map(&p.tb, W, 3, wd, ws, wb, we, false)) return cudaErrorInvalidValue;
  p.a_in = a_in; p.dx = dx; p.M = B * OH * OW; p.IH = IH; p.IW = IW; p.Cin = Cin; p.Cout = Cout;
  p.lgW2 = ilog2(OW); p.lgHW2 = ilog2(OH * OW);
  p.splits = tma::pick_splits(4 * ((p.M + tma::kBM - 1) / tma::kBM) * ((N + BN - 1) / BN), 4 * Cout / tma::kBK, BN);
  if (BN == 128) return tma::launch_tma_gemm<TmaConvDgrad, 128>(p, p.M, N, 4 * p.splits, st);
  if (BN == 64) return tma::launch_tma_gemm<TmaConvDgrad, 64>(p, p.M, N, 4 * p.splits, st);
  return tma::launch_tma_gemm<TmaConvDgrad, 32>(p, p.M, N, 4 * p.splits, st);
}

// ------------------------------------------------------------------------------------------
struct TmaConvWgrad {
  CUtensorMap ta, tb;
  float* part;
  int Cin, Cout, OW, lgOW, lgOHW, steps_per_split, total_steps, splits;   // splits: cluster split-K, unused (= 1): z already splits K
  static constexpr bool kAMn = true, kBMn = true;
  __device__ int k_iters(int z) const {
    const int left = total_steps - z * steps_per_split;
    return left < steps_per_split ? (left > 0 ? left : 0) : steps_per_split;
  }
  template <int BN>
  __device__ void load(int ki, int z, int m0, int n0, unsigned char* a_dst, unsigned char* b_dst, uint64_t* bar) const {
    const int p0 = (z * steps_per_split + ki) * tma::kBK;
    const int b0 = p0 >> lgOHW, rem = p0 & ((1 << lgOHW) - 1);
    const int oy0 = rem >> lgOW, ox0 = rem & (OW - 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + 32 * j;                    // row (tap, ci) of the weight gradient
      const int tap = r / Cin, ci0 = r - tap * Cin; // rows past 16*Cin: tap >= 16 -> iy far out of bounds -> zeros
      tma::tma_load_4d(a_dst + j * 4096, &ta, ci0, 2 * ox0 - 1 + (tap & 3), 2 * oy0 - 1 + (tap >> 2), b0, bar);
    }
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) tma::tma_load_2d(b_dst + j * 4096, &tb, n0 + 32 * j, p0, bar);
  }
  __device__ void store4(int z, int m, int n, float4 o) const {
    if (m >= 16 * Cin || n >= Cout) return;
    *reinterpret_cast<float4*>(part + ((size_t)z * 16 * Cin + m) * Cout + n) = o;
  }
  __device__ void store16(int z, int m, int n0, const float (&v)[16]) const {
    if (m >= 16 * Cin || n0 >= Cout) return;
    float4* dst = reinterpret_cast<float4*>(part + ((size_t)z * 16 * Cin + m) * Cout + n0);
#pragma unroll
    for (int g = 0; g < 4; ++g) dst[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
  }
};

bool tma_conv_wgrad_supported(const float* x, int Cx, int Cv, float shift, const float* dy, int Cout) {
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  return encode_fn() && Cv == 0 && shift == 0.f && Cx % 32 == 0 && Cout % 32 == 0 && al(x) && al(dy);
}

int tma_wgrad_splits(int B, int OH, int OW, int Cin, int Cout) {
  const int steps = (B * OH * OW + tma::kBK - 1) / tma::kBK;
  const int BN = Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32);
  const int tiles = ((16 * Cin + tma::kBM - 1) / tma::kBM) * ((Cout + BN - 1) / BN);
  int splits = (2 * 148 + tiles - 1) / tiles;                   // ~2 waves of CTAs
  if (splits > steps / 4) splits = steps / 4;                   // at least 4 K steps per CTA
  if (splits < 1) splits = 1;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  return splits;
}

cudaError_t tma_conv_wgrad_partials(const float* x, int Cx, const float* dy, float* part, int B, int IH, int IW, int Cout,
                                    int splits, cudaStream_t st) {
  TmaConvWgrad p{};
  const int OH = IH / 2, OW = IW / 2, P = B * OH * OW;
  if (!make_im2col_map(&p.ta, x, B, IH, IW, Cx, tma::kBK, true)) return cudaErrorInvalidValue;
  const cuuint64_t dd[2] = {(cuuint64_t)Cout, (cuuint64_t)P};
  const cuuint64_t ds[1] = {(cuuint64_t)Cout * 4};
  const cuuint32_t db[2] = {32, 32}, de[2] = {1, 1};
  if (!make_map(&p.tb, dy, 2, dd, ds, db, de, true)) return cudaErrorInvalidValue;
  p.part = part; p.Cin = Cx; p.Cout = Cout; p.OW = OW; p.lgOW = ilog2(OW); p.lgOHW = ilog2(OH * OW);
  p.total_steps = (P + tma::kBK - 1) / tma::kBK;
  p.splits = 1;
  p.steps_per_split = (p.total_steps + splits - 1) / splits;
  const int M = 16 * Cx;
  if (Cout % 128 == 0) return tma::launch_tma_gemm<TmaConvWgrad, 128>(p, M, Cout, splits, st);
  if (Cout % 64 == 0) return tma::launch_tma_gemm<TmaConvWgrad, 64>(p, M, Cout, splits, st);
  return tma::launch_tma_gemm<TmaConvWgrad, 32>(p, M, Cout, splits, st);
}

}  // namespace expo
