// tcgen05 instantiations of the network primitives (see tc_engine.cuh): conv fprop (forward and
// tangent), conv dgrad, conv wgrad, FC forward.  Called from nn.cu when the GEMM backend allows
// it and the shape fits the 128 x BN x 32 tensor-core tile; otherwise nn.cu's exact-fp32 CUDA-core
// engine runs.  Both backends implement the same C-ABI entry points (include/exposure_b200.h).
#include "nn_tc.h"
#include "tc_engine.cuh"

namespace expo {

__device__ __forceinline__ float tc_lrelu(float v) { return 0.6f * v + 0.4f * fabsf(v); }
__device__ __forceinline__ float tc_dlrelu(float a) { return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f); }

// ------------------------------------------------------------------------------------------
// conv fprop.  K index = tap * Cin + ci, padded to a multiple of 32.
// FAST (Cin % 32 == 0, no per-image vector, no shift): one aligned float4 per chunk.
// ------------------------------------------------------------------------------------------
template <bool FAST>
struct TcConvFprop {
  const float* x; const float* vec; const float* W; const float* bias; const float* mask_ref;
  const float* post_mul; float* y; float* y2;
  int B, IH, IW, Cx, Cv, Cin, Cout, OH, OW, Ktot, mode, lgOW, lgOHW;
  float shift;
  struct RowA { int b, iy0, ix0; };
  struct KS { int k0; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return (Ktot + tc::kBK - 1) / tc::kBK; }
  __device__ RowA row_a(int m) const {
    RowA r;
    if (m >= B * OH * OW) { r.b = -1; r.iy0 = r.ix0 = 0; return r; }
    r.b = m >> lgOHW;
    const int rem = m & ((1 << lgOHW) - 1);
    r.iy0 = 2 * (rem >> lgOW) - 1;
    r.ix0 = 2 * (rem & (OW - 1)) - 1;
    return r;
  }
  __device__ KS kstate(int ki) const { KS s; s.k0 = ki * tc::kBK; return s; }
  __device__ float a_elem(const RowA& r, int k) const {
    if (k >= Ktot) return 0.f;
    const int tap = k / Cin, ci = k - tap * Cin;
    const int iy = r.iy0 + (tap >> 2), ix = r.ix0 + (tap & 3);
    if ((unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return 0.f;
    const float v = ci < Cx ? __ldg(x + ((size_t)(r.b * IH + iy) * IW + ix) * Cx + ci)
                            : __ldg(vec + (size_t)r.b * Cv + (ci - Cx));
    return v - shift;
  }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.b < 0) return v;
    const int k = s.k0 + 4 * c;
    if (FAST) {
      const int tap = k / Cin, ci = k - tap * Cin;
      const int iy = r.iy0 + (tap >> 2), ix = r.ix0 + (tap & 3);
      if ((unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return v;
      return __ldg(reinterpret_cast<const float4*>(x + ((size_t)(r.b * IH + iy) * IW + ix) * Cx + ci));
    }
    v.x = a_elem(r, k); v.y = a_elem(r, k + 1); v.z = a_elem(r, k + 2); v.w = a_elem(r, k + 3);
    return v;
  }
  __device__ float4 load_b4(const KS& s, int n, int c) const {       // B[n][k] = W[k][n]
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= Cout) return v;
    const int k = s.k0 + 4 * c;
    if (k + 0 < Ktot) v.x = __ldg(W + (size_t)(k + 0) * Cout + n);
    if (k + 1 < Ktot) v.y = __ldg(W + (size_t)(k + 1) * Cout + n);
    if (k + 2 < Ktot) v.z = __ldg(W + (size_t)(k + 2) * Cout + n);
    if (k + 3 < Ktot) v.w = __ldg(W + (size_t)(k + 3) * Cout + n);
    return v;
  }
  __device__ void store16(int m, int n0, const float (&v)[16]) const {
    if (m >= B * OH * OW) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= Cout) break;
      const size_t idx = (size_t)m * Cout + n;
      float o = v[i];
      if (mode == 0) o = tc_lrelu(o + (bias ? __ldg(bias + n) : 0.f));
      else o *= tc_dlrelu(__ldg(mask_ref + idx));
      y[idx] = o;
      if (y2) y2[idx] = o * __ldg(post_mul + idx);
    }
  }
};

// ------------------------------------------------------------------------------------------
// FC forward with split-K: part[z][M][N] = x[M, kz] W[kz, N]
// ------------------------------------------------------------------------------------------
struct TcFcFwd {
  const float* x; const float* W; float* part;
  int M, K, N, ldx, k_per_split, k_begin;
  struct RowA { int m; };
  struct KS { int k0; };
  __device__ void init(int z) { k_begin = z * k_per_split; part += (size_t)z * M * N; }
  __device__ int k_iters() const { return k_per_split / tc::kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.m = m < M ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.k0 = k_begin + ki * tc::kBK; return s; }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.m < 0) return v;
    const int k = s.k0 + 4 * c;
    const float* p = x + (size_t)r.m * ldx + k;
    if (k + 3 < K && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0)) return __ldg(reinterpret_cast<const float4*>(p));
    if (k + 0 < K) v.x = __ldg(p + 0);
    if (k + 1 < K) v.y = __ldg(p + 1);
    if (k + 2 < K) v.z = __ldg(p + 2);
    if (k + 3 < K) v.w = __ldg(p + 3);
    return v;
  }
  __device__ float4 load_b4(const KS& s, int n, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= N) return v;
    const int k = s.k0 + 4 * c;
    if (k + 0 < K) v.x = __ldg(W + (size_t)(k + 0) * N + n);
    if (k + 1 < K) v.y = __ldg(W + (size_t)(k + 1) * N + n);
    if (k + 2 < K) v.z = __ldg(W + (size_t)(k + 2) * N + n);
    if (k + 3 < K) v.w = __ldg(W + (size_t)(k + 3) * N + n);
    return v;
  }
  __device__ void store16(int m, int n0, const float (&v)[16]) const {
    if (m >= M) return;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (n0 + i < N) part[(size_t)m * N + n0 + i] = v[i];
  }
};

static int host_ilog2_tc(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

template <class P>
static cudaError_t launch_by_bn(const P& p, int M, int N, int Z, cudaStream_t st) {
  if (N <= 32) return tc::launch_tc_gemm<P, 32>(p, M, N, Z, st);
  if (N <= 64) return tc::launch_tc_gemm<P, 64>(p, M, N, Z, st);
  if (N <= 128) return tc::launch_tc_gemm<P, 128>(p, M, N, Z, st);
  return tc::launch_tc_gemm<P, 256>(p, M, N, Z, st);
}

bool tc_conv_fwd_supported(int Cout) { return Cout >= 16 && Cout <= 1024; }

cudaError_t tc_conv_fwd(const float* x, int Cx, const float* vec, int Cv, float shift, const float* W,
                        const float* bias, const float* mask_ref, const float* post_mul, float* y, float* y2, int B,
                        int IH, int IW, int Cout, int mode, cudaStream_t st) {
  const bool fast = (Cv == 0) && (Cx % 32 == 0) && shift == 0.f && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  const int OH = IH / 2, OW = IW / 2, M = B * OH * OW;
  if (fast) {
    TcConvFprop<true> p{};
    p.x = x; p.vec = vec; p.W = W; p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
    p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cx + Cv; p.Cout = Cout; p.OH = OH; p.OW = OW;
    p.Ktot = 16 * p.Cin; p.mode = mode; p.shift = shift; p.lgOW = host_ilog2_tc(OW); p.lgOHW = host_ilog2_tc(OH * OW);
    return launch_by_bn(p, M, Cout, 1, st);
  }
  TcConvFprop<false> p{};
  p.x = x; p.vec = vec; p.W = W; p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
  p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cx + Cv; p.Cout = Cout; p.OH = OH; p.OW = OW;
  p.Ktot = 16 * p.Cin; p.mode = mode; p.shift = shift; p.lgOW = host_ilog2_tc(OW); p.lgOHW = host_ilog2_tc(OH * OW);
  return launch_by_bn(p, M, Cout, 1, st);
}

int tc_fc_splits(int M, int K, int N) {
  const int tiles = ((M + tc::kBM - 1) / tc::kBM) * ((N + 127) / 128);
  int s = (148 + tiles - 1) / tiles;
  const int max_s = K / 128 > 0 ? K / 128 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

cudaError_t tc_fc_fwd_partials(const float* x, int ldx, const float* W, float* part, int M, int K, int N, int splits,
                               cudaStream_t st) {
  TcFcFwd p{};
  p.x = x; p.W = W; p.part = part; p.M = M; p.K = K; p.N = N; p.ldx = ldx;
  int kps = (K + splits - 1) / splits;
  kps = ((kps + tc::kBK - 1) / tc::kBK) * tc::kBK;
  p.k_per_split = kps;
  if (N <= 32) return tc::launch_tc_gemm<TcFcFwd, 32>(p, M, N, splits, st);
  if (N <= 64) return tc::launch_tc_gemm<TcFcFwd, 64>(p, M, N, splits, st);
  return tc::launch_tc_gemm<TcFcFwd, 128>(p, M, N, splits, st);
}

}  // namespace expo
