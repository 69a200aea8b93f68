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
#include "nn_internal.h"
#include "tc_engine_tma_persistent.cuh"

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
  int first;   // 1: first layer on the zero-bordered 16-channel staging copy (see exp_conv1_pad_input)
  static constexpr bool kAMn = false, kBMn = true;
  static constexpr bool kPlainRows = false;
  __device__ float* row_ptr(int, int, int) const { return nullptr; }
  __host__ __device__ int k_iters(int) const { return 16 * Cin / tma::kBK; }
  template <int BN>
  __device__ void load(int ki, int, int m0, int n0, unsigned char* a_dst, unsigned char* b_dst, uint64_t* bar) const {
    const int k = ki * tma::kBK;
    const int tap = k / Cin, ci0 = k - tap * Cin;
    const int b0 = m0 >> lgOHW, rem = m0 & ((1 << lgOHW) - 1);
    const int oy0 = rem >> lgOW, ox0 = rem & (OW - 1);
    // first layer: K step ki = (ky, half); the 32 floats are padded pixels 2(ox+half), 2(ox+half)+1
    // x 16 channels = chunk ox+half of the row viewed as 32-float chunks, row 2oy+ky
    if (first) tma::tma_load_4d(a_dst, &ta, 0, ox0 + (ki & 1), 2 * oy0 + (ki >> 1), b0, bar);
    else tma::tma_load_4d(a_dst, &ta, ci0, 2 * ox0 - 1 + (tap & 3), 2 * oy0 - 1 + (tap >> 2), b0, bar);
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) tma::tma_load_2d(b_dst + j * 4096, &tb, n0 + 32 * j, k, bar);
  }
  // the epilogue of one float4 in two halves, so that the store pass can put the global loads of several elements in
  // flight before it consumes any (tc_engine_tma.cuh): epi_load issues them, epi_store finishes the element
  struct Aux { float4 a, b; };
  __device__ Aux epi_load(int, int m, int n) const {
    Aux x;
    x.a = x.b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m >= M || n >= Cout) return x;
    const size_t idx = (size_t)m * Cout + n;
    if (mode == 0) { if (bias) x.a = __ldg(reinterpret_cast<const float4*>(bias + n)); }
    else x.a = __ldg(reinterpret_cast<const float4*>(mask_ref + idx));
    if (y2) x.b = __ldg(reinterpret_cast<const float4*>(post_mul + idx));
    return x;
  }
  __device__ void epi_store(int, int m, int n, float4 o, const Aux& x) const {
    if (m >= M || n >= Cout) return;
    const size_t idx = (size_t)m * Cout + n;
    if (mode == 0) {
      o.x = lrelu(o.x + x.a.x); o.y = lrelu(o.y + x.a.y); o.z = lrelu(o.z + x.a.z); o.w = lrelu(o.w + x.a.w);
    } else {
      o.x *= dlrelu(x.a.x); o.y *= dlrelu(x.a.y); o.z *= dlrelu(x.a.z); o.w *= dlrelu(x.a.w);
    }
    *reinterpret_cast<float4*>(y + idx) = o;
    if (y2) *reinterpret_cast<float4*>(y2 + idx) = make_float4(o.x * x.b.x, o.y * x.b.y, o.z * x.b.z, o.w * x.b.w);
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
    return tma::launch_tma_auto<TmaConvFprop, 128>(p, p.M, Cout, p.splits, st);
  }
  if (Cout % 64 == 0) {
    p.splits = tma::pick_splits(mt * (Cout / 64), ki, 64);
    return tma::launch_tma_auto<TmaConvFprop, 64>(p, p.M, Cout, p.splits, st);
  }
  p.splits = tma::pick_splits(mt * (Cout / 32), ki, 32);
  return tma::launch_tma_auto<TmaConvFprop, 32>(p, p.M, Cout, p.splits, st);
}

// ------------------------------------------------------------------------------------------
struct TmaConvDgrad {
  CUtensorMap ta, tb;
  const float* a_in; float* dx;
  int M, IH, IW, Cin, Cout, lgW2, lgHW2, splits;
  static constexpr bool kAMn = false, kBMn = false;
  static constexpr bool kPlainRows = false;
  __device__ float* row_ptr(int, int, int) const { return nullptr; }
  __host__ __device__ int k_iters(int) const { return 4 * Cout / tma::kBK; }
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
  struct Aux { float4 a; size_t idx; };
  __device__ Aux epi_load(int z, int m, int n) const {
    Aux x;
    x.a = make_float4(1.f, 1.f, 1.f, 1.f);
    x.idx = 0;
    if (m >= M || n >= Cin) return x;
    const int py = z >> 1, px = z & 1;
    const int b = m >> lgHW2, rem = m & ((1 << lgHW2) - 1);
    const int iy = 2 * (rem >> lgW2) + py, ix = 2 * (rem & ((IW / 2) - 1)) + px;
    x.idx = ((size_t)(b * IH + iy) * IW + ix) * Cin + n;
    if (a_in) x.a = __ldg(reinterpret_cast<const float4*>(a_in + x.idx));
    return x;
  }
  __device__ void epi_store(int, int m, int n, float4 o, const Aux& x) const {
    if (m >= M || n >= Cin) return;
    if (a_in) { o.x *= dlrelu(x.a.x); o.y *= dlrelu(x.a.y); o.z *= dlrelu(x.a.z); o.w *= dlrelu(x.a.w); }
    *reinterpret_cast<float4*>(dx + x.idx) = o;
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
  if (BN == 128) return tma::launch_tma_auto<TmaConvDgrad, 128>(p, p.M, N, 4 * p.splits, st);
  if (BN == 64) return tma::launch_tma_auto<TmaConvDgrad, 64>(p, p.M, N, 4 * p.splits, st);
  return tma::launch_tma_auto<TmaConvDgrad, 32>(p, p.M, N, 4 * p.splits, st);
}

// ------------------------------------------------------------------------------------------
struct TmaConvWgrad {
  CUtensorMap ta, tb;
  float* part;
  int Cin, Cout, OW, lgOW, lgOHW, steps_per_split, total_steps, splits;   // splits: cluster split-K, unused (= 1): z already splits K
  int first;   // 1: first layer on the zero-bordered 16-channel staging copy
  static constexpr bool kAMn = true, kBMn = true;
  static constexpr bool kPlainRows = true;                     // partial tiles are stored as they are
  __device__ float* row_ptr(int z, int m, int n0) const {
    return (m < 16 * Cin && n0 < Cout) ? part + ((size_t)z * 16 * Cin + m) * Cout + n0 : nullptr;
  }
  __host__ __device__ int k_iters(int z) const {
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
      const int r = m0 + 32 * j;                    // row (tap, ci) of the weight gradient (16*Cin % 128 == 0: no tail)
      const int tap = r / Cin, ci0 = r - tap * Cin;
      // first layer (staging copy, Cin = 16): the 32 rows are (ky, 2 taps x 16 ch) = chunk ox+half of row 2oy+ky
      if (first) tma::tma_load_4d(a_dst + j * 4096, &ta, 0, ox0 + ((r >> 5) & 1), 2 * oy0 + (r >> 6), b0, bar);
      else tma::tma_load_4d(a_dst + j * 4096, &ta, ci0, 2 * ox0 - 1 + (tap & 3), 2 * oy0 - 1 + (tap >> 2), b0, bar);
    }
#pragma unroll
    for (int j = 0; j < BN / 32; ++j) tma::tma_load_2d(b_dst + j * 4096, &tb, n0 + 32 * j, p0, bar);
  }
  struct Aux {};
  __device__ Aux epi_load(int, int, int) const { return Aux{}; }
  __device__ void epi_store(int z, int m, int n, float4 o, const Aux&) const { store4(z, m, n, o); }
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
  // ONE wave: a CTA costs ~10 us before its first K step (profiles/r1_tma_gemm_ncu.md), so a second
  // wave of shorter CTAs is slower than one wave of longer ones
  int splits = 148 / tiles;
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
  if (Cout % 128 == 0) return tma::launch_tma_auto<TmaConvWgrad, 128>(p, M, Cout, splits, st);
  if (Cout % 64 == 0) return tma::launch_tma_auto<TmaConvWgrad, 64>(p, M, Cout, splits, st);
  return tma::launch_tma_auto<TmaConvWgrad, 32>(p, M, Cout, splits, st);
}

// ------------------------------------------------------------------------------------------
// First layer (Cin = 3 + states = 6 / 14 / 17 channels of which all but 3 are per-image constants,
// util.py enrich_image_input; ly.conv2d SAME padding).  Cin is not a multiple of 32, so the layer
// runs on a STAGING COPY xp[B][IH+2][IW+2][16]: zero border (the SAME padding made explicit), the
// enriched input minus `shift` in channels [0,Cin), zeros above.  One padded row is a sequence of
// 32-float chunks (2 pixels x 16 channels); output pixel ox and tap pair `half` read chunk ox+half,
// so the im2col box needs no element stride in x and K = 4 ky x 2 halves x 32 = 256 with weights
// padded to Wp[4][4][16][Cout].
// ------------------------------------------------------------------------------------------
constexpr int kC1 = 16;

static bool make_conv1_map(CUtensorMap* m, const float* xp, int B, int IH, int IW, int count, bool mn_major) {
  const PixTile t = pix_tile(count, IH / 2, IW / 2);
  const cuuint64_t dims[4] = {32, (cuuint64_t)(IW + 2) / 2, (cuuint64_t)IH + 2, (cuuint64_t)B};
  const cuuint64_t strides[3] = {128, (cuuint64_t)(IW + 2) * kC1 * 4, (cuuint64_t)(IH + 2) * (IW + 2) * kC1 * 4};
  const cuuint32_t box[4] = {32, (cuuint32_t)t.tw, (cuuint32_t)(2 * t.th), (cuuint32_t)t.tb};
  const cuuint32_t estr[4] = {1, 1, 2, 1};
  return make_map(m, xp, 4, dims, strides, box, estr, mn_major);
}

// out[b][y][x][CP] = concat(x, tile(vec)) - shift in channels [0, Cx+Cv), 0 above; BORDER adds a
// one-pixel zero frame (out is [B][IH+2][IW+2][CP])
template <int CP, int BORDER>
__global__ void conv_stage_input_kernel(const float* __restrict__ x, const float* __restrict__ vec, float shift,
                                        float* __restrict__ xp, int B, int IH, int IW, int Cx, int Cv) {
  EXP_PDL_ENTRY();
  // one float4 of one (padded) pixel per thread: a warp stores 512 contiguous bytes (one thread per pixel stored 64
  // bytes at a 64-byte stride per lane: 12 us for the 17.8 MB of a batch-64 layer)
  constexpr int G = CP / 4;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int PW = IW + 2 * BORDER, PH = IH + 2 * BORDER;
  const size_t total = (size_t)B * PH * PW * G;
  if (i >= total) return;
  const int g = (int)(i % G);
  const size_t pix = i / G;
  const int xx = (int)(pix % PW) - BORDER, yy = (int)((pix / PW) % PH) - BORDER, b = (int)(pix / ((size_t)PW * PH));
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (xx >= 0 && xx < IW && yy >= 0 && yy < IH) {
    const float* px = x + (((size_t)b * IH + yy) * IW + xx) * Cx;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 4 * g + k;
      if (c < Cx) v[k] = __ldg(px + c) - shift;
      else if (c < Cx + Cv) v[k] = __ldg(vec + (size_t)b * Cv + (c - Cx)) - shift;
    }
  }
  reinterpret_cast<float4*>(xp)[i] = make_float4(v[0], v[1], v[2], v[3]);
}

// Wp[tap][cp][co] = W[tap][c][co] for c < Cin, 0 above
__global__ void conv_pad_weights_kernel(const float* __restrict__ W, float* __restrict__ Wp, int Cin, int Cout, int CP) {
  EXP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;                     // over [16 taps][CP][Cout]
  if (i >= 16 * CP * Cout) return;
  const int co = i % Cout, c = (i / Cout) % CP, tap = i / (Cout * CP);
  Wp[i] = c < Cin ? __ldg(W + ((size_t)tap * Cin + c) * Cout + co) : 0.f;
}

// gW[tap][ci][co] (=|+=) sum_z part[z][tap*16 + ci][co], ci < Cin, in z order
__global__ void conv1_wgrad_reduce_kernel(const float* __restrict__ part, int splits, int Cin, int Cout,
                                          float* __restrict__ gW, int accumulate) {
  EXP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 16 * Cin * Cout) return;
  const int co = i % Cout, ci = (i / Cout) % Cin, tap = i / (Cout * Cin);
  const size_t src = ((size_t)tap * kC1 + ci) * Cout + co, stride = (size_t)16 * kC1 * Cout;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[src + (size_t)z * stride];
  gW[i] = accumulate ? gW[i] + s : s;
}

static int conv1_wgrad_splits(int B, int OH, int OW, int Cout) {
  const int steps = (B * OH * OW + tma::kBK - 1) / tma::kBK;
  const int tiles = 2 * ((Cout + 127) / 128);                      // M = 256 rows
  int splits = (148 + tiles - 1) / tiles;
  if (splits > steps / 4) splits = steps / 4;
  if (splits < 1) splits = 1;
  if (splits > 128) splits = 128;
  return splits;
}

}  // namespace expo

using namespace expo;

extern "C" {

size_t exp_conv1_padded_input_elems(int B, int IH, int IW) {
  if (B <= 0 || IH <= 0 || IW <= 0) return 0;
  return (size_t)B * (IH + 2) * (IW + 2) * kC1;
}

int exp_conv1_pad_input(const float* x, int Cx, const float* vec, int Cv, float shift, float* xp, int B, int IH, int IW,
                        void* stream) {
  EXP_CHECK_ARG(x && xp && B > 0 && IH >= 2 && IW >= 2 && IW % 2 == 0, "bad args");
  EXP_CHECK_ARG(Cx > 0 && Cv >= 0 && (Cv == 0 || vec) && Cx + Cv <= kC1, "first-layer staging holds at most %d channels", kC1);
  EXP_CHECK_ARG((reinterpret_cast<uintptr_t>(xp) & 15u) == 0, "xp must be 16-byte aligned");
  const size_t total = (size_t)B * (IH + 2) * (IW + 2) * (kC1 / 4);           // one float4 per thread
  launch_pdl(conv_stage_input_kernel<kC1, 1>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, vec, shift, xp, B, IH, IW, Cx, Cv);
  EXP_CHECK_LAUNCH("exp_conv1_pad_input");
  return EXP_OK;
}

int exp_conv1_pad_weights(const float* W, int Cin, int Cout, float* Wp, void* stream) {
  EXP_CHECK_ARG(W && Wp && Cin > 0 && Cin <= kC1 && Cout > 0, "bad args");
  const int n = 16 * kC1 * Cout;
  launch_pdl(conv_pad_weights_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, W, Wp, Cin, Cout, kC1);
  EXP_CHECK_LAUNCH("exp_conv1_pad_weights");
  return EXP_OK;
}

int exp_conv_enrich32(const float* x, int Cx, const float* vec, int Cv, float shift, float* out, int B, int IH, int IW,
                      void* stream) {
  EXP_CHECK_ARG(x && out && B > 0 && IH > 0 && IW > 0, "bad args");
  EXP_CHECK_ARG(Cx > 0 && Cv >= 0 && (Cv == 0 || vec) && Cx + Cv <= 32, "at most 32 channels");
  EXP_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15u) == 0, "out must be 16-byte aligned");
  const size_t total = (size_t)B * IH * IW * (32 / 4);                        // one float4 per thread
  launch_pdl(conv_stage_input_kernel<32, 0>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, x, vec, shift, out, B, IH, IW, Cx, Cv);
  EXP_CHECK_LAUNCH("exp_conv_enrich32");
  return EXP_OK;
}

int exp_conv_pad_weights32(const float* W, int Cin, int Cout, float* Wp, void* stream) {
  EXP_CHECK_ARG(W && Wp && Cin > 0 && Cin <= 32 && Cout > 0, "bad args");
  const int n = 16 * 32 * Cout;
  launch_pdl(conv_pad_weights_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, W, Wp, Cin, Cout, 32);
  EXP_CHECK_LAUNCH("exp_conv_pad_weights32");
  return EXP_OK;
}

int exp_conv1_supported(int Cin, int Cout) { return encode_fn() != nullptr && Cin > 0 && Cin <= kC1 && Cout > 0 && Cout % 32 == 0; }

int exp_conv1_fwd(const float* xp, const float* Wp, const float* bias, const float* mask_ref, const float* post_mul, float* y,
                  float* y2, int B, int IH, int IW, int Cout, int mode, void* stream) {
  EXP_CHECK_ARG(xp && Wp && y && B > 0 && IH >= 2 && IW >= 2, "bad args");
  EXP_CHECK_ARG((IH & (IH - 1)) == 0 && (IW & (IW - 1)) == 0, "IH/IW must be powers of two");
  EXP_CHECK_ARG(Cout % 32 == 0, "Cout must be a multiple of 32");
  EXP_CHECK_ARG(mode == 0 || (mode == 1 && mask_ref), "mode 1 needs mask_ref");
  EXP_CHECK_ARG(!y2 || post_mul, "y2 needs post_mul");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!(al(xp) && al(Wp) && al(bias) && al(mask_ref) && al(post_mul) && al(y) && al(y2)))
    return set_error(EXP_ERR_ALIGNMENT, "exp_conv1_fwd: pointers must be 16-byte aligned");
  if (!encode_fn()) return set_error(EXP_ERR_UNSUPPORTED, "exp_conv1_fwd: cuTensorMapEncodeTiled not available");
  TmaConvFprop p{};
  const int OH = IH / 2, OW = IW / 2;
  const cuuint64_t wd[2] = {(cuuint64_t)Cout, (cuuint64_t)16 * kC1};
  const cuuint64_t ws[1] = {(cuuint64_t)Cout * 4};
  const cuuint32_t wb[2] = {32, 32}, we[2] = {1, 1};
  if (!make_conv1_map(&p.ta, xp, B, IH, IW, tma::kBM, false) || !make_map(&p.tb, Wp, 2, wd, ws, wb, we, true))
    return set_error(EXP_ERR_CUDA, "exp_conv1_fwd: tensor map encode failed");
  p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
  p.M = B * OH * OW; p.Cin = kC1; p.Cout = Cout; p.OW = OW; p.lgOW = ilog2(OW); p.lgOHW = ilog2(OH * OW); p.mode = mode;
  p.first = 1;
  const int mt = (p.M + tma::kBM - 1) / tma::kBM, ki = 16 * kC1 / tma::kBK;
  cudaError_t e;
  if (Cout % 128 == 0) {
    p.splits = tma::pick_splits(mt * (Cout / 128), ki, 128);
    e = tma::launch_tma_auto<TmaConvFprop, 128>(p, p.M, Cout, p.splits, (cudaStream_t)stream);
  } else if (Cout % 64 == 0) {
    p.splits = tma::pick_splits(mt * (Cout / 64), ki, 64);
    e = tma::launch_tma_auto<TmaConvFprop, 64>(p, p.M, Cout, p.splits, (cudaStream_t)stream);
  } else {
    p.splits = tma::pick_splits(mt * (Cout / 32), ki, 32);
    e = tma::launch_tma_auto<TmaConvFprop, 32>(p, p.M, Cout, p.splits, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv1_fwd: %s", cudaGetErrorString(e));
  return EXP_OK;
}

size_t exp_conv1_wgrad_workspace_bytes(int B, int IH, int IW, int Cout) {
  if (B <= 0 || IH < 2 || IW < 2 || Cout <= 0) return 0;
  return (size_t)conv1_wgrad_splits(B, IH / 2, IW / 2, Cout) * 16 * kC1 * Cout * sizeof(float);
}

int exp_conv1_wgrad(const float* xp, const float* dy, float* gW, int Cin, int B, int IH, int IW, int Cout, int accumulate,
                    void* workspace, size_t workspace_bytes, void* stream) {
  EXP_CHECK_ARG(xp && dy && gW && workspace && B > 0 && IH >= 2 && IW >= 2, "bad args");
  EXP_CHECK_ARG((IH & (IH - 1)) == 0 && (IW & (IW - 1)) == 0, "IH/IW must be powers of two");
  EXP_CHECK_ARG(Cin > 0 && Cin <= kC1 && Cout % 32 == 0, "Cin <= %d and Cout %% 32 == 0 required", kC1);
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (!(al(xp) && al(dy) && al(workspace))) return set_error(EXP_ERR_ALIGNMENT, "exp_conv1_wgrad: pointers must be 16-byte aligned");
  if (!encode_fn()) return set_error(EXP_ERR_UNSUPPORTED, "exp_conv1_wgrad: cuTensorMapEncodeTiled not available");
  const int OH = IH / 2, OW = IW / 2, P = B * OH * OW;
  const int splits = conv1_wgrad_splits(B, OH, OW, Cout);
  const size_t need = (size_t)splits * 16 * kC1 * Cout * sizeof(float);
  if (workspace_bytes < need) return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  TmaConvWgrad p{};
  const cuuint64_t dd[2] = {(cuuint64_t)Cout, (cuuint64_t)P};
  const cuuint64_t ds[1] = {(cuuint64_t)Cout * 4};
  const cuuint32_t db[2] = {32, 32}, de[2] = {1, 1};
  if (!make_conv1_map(&p.ta, xp, B, IH, IW, tma::kBK, true) || !make_map(&p.tb, dy, 2, dd, ds, db, de, true))
    return set_error(EXP_ERR_CUDA, "exp_conv1_wgrad: tensor map encode failed");
  p.part = reinterpret_cast<float*>(workspace); p.Cin = kC1; p.Cout = Cout; p.OW = OW; p.lgOW = ilog2(OW); p.lgOHW = ilog2(OH * OW);
  p.total_steps = (P + tma::kBK - 1) / tma::kBK;
  p.splits = 1;
  p.first = 1;
  p.steps_per_split = (p.total_steps + splits - 1) / splits;
  const int M = 16 * kC1;
  cudaError_t e;
  if (Cout % 128 == 0) e = tma::launch_tma_auto<TmaConvWgrad, 128>(p, M, Cout, splits, (cudaStream_t)stream);
  else if (Cout % 64 == 0) e = tma::launch_tma_auto<TmaConvWgrad, 64>(p, M, Cout, splits, (cudaStream_t)stream);
  else e = tma::launch_tma_auto<TmaConvWgrad, 32>(p, M, Cout, splits, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv1_wgrad: %s", cudaGetErrorString(e));
  const int n = 16 * Cin * Cout;
  launch_pdl(conv1_wgrad_reduce_kernel, dim3((n + 255) / 256), dim3(256), 0, (cudaStream_t)stream, p.part, splits, Cin, Cout, gW, accumulate);
  EXP_CHECK_LAUNCH("exp_conv1_wgrad[reduce]");
  return EXP_OK;
}

#ifdef EXPO_TMA_TRACE
// development build only (tools/tma_trace.py): copies the stamps of the last persistent launch to the host
int exp_debug_tma_trace_cta(long long* cta) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(cta, tma::g_tma_trace_cta, sizeof(tma::g_tma_trace_cta)) == cudaSuccess ? 0 : -1;
}
int exp_debug_tma_trace(long long* steps, long long* tiles) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(steps, tma::g_tma_trace, sizeof(tma::g_tma_trace)) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(tiles, tma::g_tma_trace_epi, sizeof(tma::g_tma_trace_epi)) != cudaSuccess) return -1;
  return 0;
}
#endif

}  // extern "C"
