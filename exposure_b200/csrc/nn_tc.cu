// tcgen05 instantiations of the network primitives (see tc_engine.cuh): conv fprop (forward and
// tangent), conv dgrad, conv wgrad, FC forward.  Called from nn.cu when the GEMM backend allows
// it and the shape fits the 128 x BN x 32 tensor-core tile; otherwise nn.cu's exact-fp32 CUDA-core
// engine runs.  Both backends implement the same C-ABI entry points (include/exposure_b200.h).
#include "nn_tc.h"
#include "tc_engine.cuh"
#include "tc_engine_ws.cuh"

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

// ------------------------------------------------------------------------------------------
// conv dgrad, one parity class per blockIdx.z.  K index = t * Cout + co (t = 0..3 taps of the
// class), Cout % 32 == 0.  A = delta_out (co contiguous), B[n=ci][k] = W[tap][ci][co] (K-major).
// ------------------------------------------------------------------------------------------
struct TcConvDgrad {
  const float* dy; const float* W; const float* a_in; float* dx;
  int B, IH, IW, Cin, OH, OW, Cout, py, px, lgHW2, lgW2;
  struct RowA { int b, a, c; };
  struct KS { int oy_off, ox_off, tap, co0; };
  __device__ void init(int z) { py = z >> 1; px = z & 1; }
  __device__ int k_iters() const { return 4 * Cout / tc::kBK; }
  __device__ RowA row_a(int m) const {
    RowA r;
    if (m >= B * (IH / 2) * (IW / 2)) { r.b = -1; r.a = r.c = 0; return r; }
    r.b = m >> lgHW2;
    const int rem = m & ((1 << lgHW2) - 1);
    r.a = rem >> lgW2;
    r.c = rem & ((IW / 2) - 1);
    return r;
  }
  __device__ KS kstate(int ki) const {
    KS s;
    const int k0 = ki * tc::kBK;
    const int t = k0 / Cout;
    s.co0 = k0 - t * Cout;
    const int j = t >> 1, l = t & 1;
    const int ky = py == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
    const int kx = px == 0 ? (l == 0 ? 1 : 3) : (l == 0 ? 0 : 2);
    s.oy_off = py == 0 ? (j == 0 ? 0 : -1) : (j == 0 ? 1 : 0);
    s.ox_off = px == 0 ? (l == 0 ? 0 : -1) : (l == 0 ? 1 : 0);
    s.tap = ky * 4 + kx;
    return s;
  }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    const int oy = r.a + s.oy_off, ox = r.c + s.ox_off;
    if (r.b < 0 || (unsigned)oy >= (unsigned)OH || (unsigned)ox >= (unsigned)OW) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(dy + ((size_t)(r.b * OH + oy) * OW + ox) * Cout + s.co0 + 4 * c));
  }
  __device__ float4 load_b4(const KS& s, int n, int c) const {
    if (n >= Cin) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(W + ((size_t)s.tap * Cin + n) * Cout + s.co0 + 4 * c));
  }
  __device__ void store16(int m, int n0, const float (&v)[16]) const {
    if (m >= B * (IH / 2) * (IW / 2)) return;
    const int b = m >> lgHW2;
    const int rem = m & ((1 << lgHW2) - 1);
    const int iy = 2 * (rem >> lgW2) + py, ix = 2 * (rem & ((IW / 2) - 1)) + px;
    const size_t base = ((size_t)(b * IH + iy) * IW + ix) * Cin;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      if (n >= Cin) break;
      float o = v[i];
      if (a_in) o *= tc_dlrelu(__ldg(a_in + base + n));
      dx[base + n] = o;
    }
  }
};

// ------------------------------------------------------------------------------------------
// conv wgrad: part[z][(tap,ci)][co] = sum over the pixels of split z.  A rows = (tap,ci)
// (gathered row-fast: ci is the contiguous dimension), K = pixels.
// ------------------------------------------------------------------------------------------
struct TcConvWgrad {
  const float* x; const float* vec; const float* dy; float* part;
  int B, IH, IW, Cx, Cv, Cin, Cout, OH, OW, lgOW, lgOHW, pix_per_split, p_begin;
  float shift;
  struct RowA { int ky, kx, ci; };
  struct KS { int p0; };
  __device__ void init(int z) { p_begin = z * pix_per_split; part += (size_t)z * 16 * Cin * Cout; }
  __device__ int k_iters() const { return pix_per_split / tc::kBK; }
  __device__ RowA row_a(int m) const {
    RowA r;
    if (m >= 16 * Cin) { r.ci = -1; r.ky = r.kx = 0; return r; }
    const int tap = m / Cin;
    r.ci = m - tap * Cin;
    r.ky = tap >> 2; r.kx = tap & 3;
    return r;
  }
  __device__ KS kstate(int ki) const { KS s; s.p0 = p_begin + ki * tc::kBK; return s; }
  __device__ float a_elem(const RowA& r, int p) const {
    if (p >= B * OH * OW) return 0.f;
    const int b = p >> lgOHW;
    const int rem = p & ((1 << lgOHW) - 1);
    const int iy = 2 * (rem >> lgOW) - 1 + r.ky, ix = 2 * (rem & (OW - 1)) - 1 + r.kx;
    if ((unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return 0.f;
    const float v = r.ci < Cx ? __ldg(x + ((size_t)(b * IH + iy) * IW + ix) * Cx + r.ci)
                              : __ldg(vec + (size_t)b * Cv + (r.ci - Cx));
    return v - shift;
  }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.ci < 0) return v;
    const int p = s.p0 + 4 * c;
    v.x = a_elem(r, p); v.y = a_elem(r, p + 1); v.z = a_elem(r, p + 2); v.w = a_elem(r, p + 3);
    return v;
  }
  __device__ float4 load_b4(const KS& s, int n, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= Cout) return v;
    const int p = s.p0 + 4 * c, P = B * OH * OW;
    if (p + 0 < P) v.x = __ldg(dy + (size_t)(p + 0) * Cout + n);
    if (p + 1 < P) v.y = __ldg(dy + (size_t)(p + 1) * Cout + n);
    if (p + 2 < P) v.z = __ldg(dy + (size_t)(p + 2) * Cout + n);
    if (p + 3 < P) v.w = __ldg(dy + (size_t)(p + 3) * Cout + n);
    return v;
  }
  __device__ void store16(int m, int n0, const float (&v)[16]) const {
    if (m >= 16 * Cin) return;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (n0 + i < Cout) part[(size_t)m * Cout + n0 + i] = v[i];
  }
};

// dx[M,K] = (acc ? dx : 0) + dy[M,N] W^T * lrelu'(mul_act) * mul_plain   (B[n'=k][k'=n] = W[k][n], K-major)
struct TcFcDgrad {
  const float* dy; const float* W; const float* mul_act; const float* mul_plain; float* dx;
  int M, K, N, ldy, lddx, ldmul, accumulate;
  struct RowA { int m; };
  struct KS { int n0; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return (N + tc::kBK - 1) / tc::kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.m = m < M ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.n0 = ki * tc::kBK; return s; }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.m < 0) return v;
    const int n = s.n0 + 4 * c;
    const float* p = dy + (size_t)r.m * ldy + n;
    if (n + 0 < N) v.x = __ldg(p + 0);
    if (n + 1 < N) v.y = __ldg(p + 1);
    if (n + 2 < N) v.z = __ldg(p + 2);
    if (n + 3 < N) v.w = __ldg(p + 3);
    return v;
  }
  __device__ float4 load_b4(const KS& s, int col, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col >= K) return v;
    const int n = s.n0 + 4 * c;
    const float* p = W + (size_t)col * N + n;
    if (n + 0 < N) v.x = __ldg(p + 0);
    if (n + 1 < N) v.y = __ldg(p + 1);
    if (n + 2 < N) v.z = __ldg(p + 2);
    if (n + 3 < N) v.w = __ldg(p + 3);
    return v;
  }
  __device__ void store16(int m, int k0, const float (&v)[16]) const {
    if (m >= M) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = k0 + i;
      if (k >= K) break;
      float o = v[i];
      const size_t im = (size_t)m * ldmul + k;
      if (mul_act) o *= tc_dlrelu(__ldg(mul_act + im));
      if (mul_plain) o *= __ldg(mul_plain + im);
      const size_t idx = (size_t)m * lddx + k;
      dx[idx] = accumulate ? dx[idx] + o : o;
    }
  }
};

// gW[K,N] = (acc ? gW : 0) + x^T dy   (A rows = k gathered row-fast, reduction over the batch)
struct TcFcWgrad {
  const float* x; const float* dy; float* gW;
  int M, K, N, ldx, ldy, accumulate;
  struct RowA { int k; };
  struct KS { int s0; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return (M + tc::kBK - 1) / tc::kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.k = m < K ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.s0 = ki * tc::kBK; return s; }
  __device__ float4 load_a4(const RowA& r, const KS& s, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r.k < 0) return v;
    const int sm = s.s0 + 4 * c;
    if (sm + 0 < M) v.x = __ldg(x + (size_t)(sm + 0) * ldx + r.k);
    if (sm + 1 < M) v.y = __ldg(x + (size_t)(sm + 1) * ldx + r.k);
    if (sm + 2 < M) v.z = __ldg(x + (size_t)(sm + 2) * ldx + r.k);
    if (sm + 3 < M) v.w = __ldg(x + (size_t)(sm + 3) * ldx + r.k);
    return v;
  }
  __device__ float4 load_b4(const KS& s, int n, int c) const {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n >= N) return v;
    const int sm = s.s0 + 4 * c;
    if (sm + 0 < M) v.x = __ldg(dy + (size_t)(sm + 0) * ldy + n);
    if (sm + 1 < M) v.y = __ldg(dy + (size_t)(sm + 1) * ldy + n);
    if (sm + 2 < M) v.z = __ldg(dy + (size_t)(sm + 2) * ldy + n);
    if (sm + 3 < M) v.w = __ldg(dy + (size_t)(sm + 3) * ldy + n);
    return v;
  }
  __device__ void store16(int k, int n0, const float (&v)[16]) const {
    if (k >= K) return;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (n0 + i >= N) break;
      const size_t idx = (size_t)k * N + n0 + i;
      gW[idx] = accumulate ? gW[idx] + v[i] : v[i];
    }
  }
};

static int host_ilog2_tc(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Picks the N tile and the engine generation: backend 3 = warp-specialised engine (BN <= 128),
// backend 2 = first-generation engine (BN up to 256).
template <class P, bool ROWFAST = false, int MAXBN = 256>
static cudaError_t launch_auto(const P& p, int M, int N, int Z, cudaStream_t st) {
  if (gemm_backend() == kBackendTcgen05Ws) {
    if (N <= 32) return tc::launch_tc_gemm_ws<P, 32, ROWFAST>(p, M, N, Z, st);
    if (N <= 64) return tc::launch_tc_gemm_ws<P, 64, ROWFAST>(p, M, N, Z, st);
    return tc::launch_tc_gemm_ws<P, 128, ROWFAST>(p, M, N, Z, st);
  }
  if (N <= 32) return tc::launch_tc_gemm<P, 32, ROWFAST>(p, M, N, Z, st);
  if (N <= 64) return tc::launch_tc_gemm<P, 64, ROWFAST>(p, M, N, Z, st);
  if (N <= 128 || MAXBN < 256) return tc::launch_tc_gemm<P, 128, ROWFAST>(p, M, N, Z, st);
  return tc::launch_tc_gemm<P, 256, ROWFAST>(p, M, N, Z, st);
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
    return launch_auto(p, M, Cout, 1, st);
  }
  TcConvFprop<false> p{};
  p.x = x; p.vec = vec; p.W = W; p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
  p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cx + Cv; p.Cout = Cout; p.OH = OH; p.OW = OW;
  p.Ktot = 16 * p.Cin; p.mode = mode; p.shift = shift; p.lgOW = host_ilog2_tc(OW); p.lgOHW = host_ilog2_tc(OH * OW);
  return launch_auto(p, M, Cout, 1, st);
}

bool tc_conv_dgrad_supported(int Cout) { return Cout % 32 == 0; }

cudaError_t tc_conv_dgrad(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                          int Cout, cudaStream_t st) {
  TcConvDgrad p{};
  p.dy = dy; p.W = W; p.a_in = a_in; p.dx = dx; p.B = B; p.IH = IH; p.IW = IW; p.Cin = Cin; p.OH = IH / 2; p.OW = IW / 2;
  p.Cout = Cout; p.lgW2 = host_ilog2_tc(IW / 2); p.lgHW2 = host_ilog2_tc((IH / 2) * (IW / 2));
  const int M = B * (IH / 2) * (IW / 2);
  return launch_auto(p, M, Cin, 4, st);
}

int tc_wgrad_splits(int B, int OH, int OW, int Cin, int Cout) {
  const int pixels = B * OH * OW;
  const int tiles = ((16 * Cin + tc::kBM - 1) / tc::kBM) * ((Cout + 255) / 256);
  int s = (296 + tiles - 1) / tiles;
  const int max_s = pixels / 128 > 0 ? pixels / 128 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

cudaError_t tc_conv_wgrad_partials(const float* x, int Cx, const float* vec, int Cv, float shift, const float* dy,
                                   float* part, int B, int IH, int IW, int Cout, int splits, cudaStream_t st) {
  TcConvWgrad p{};
  const int OH = IH / 2, OW = IW / 2;
  p.x = x; p.vec = vec; p.dy = dy; p.part = part; p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cx + Cv;
  p.Cout = Cout; p.OH = OH; p.OW = OW; p.lgOW = host_ilog2_tc(OW); p.lgOHW = host_ilog2_tc(OH * OW); p.shift = shift;
  const int pixels = B * OH * OW;
  int pps = (pixels + splits - 1) / splits;
  pps = ((pps + tc::kBK - 1) / tc::kBK) * tc::kBK;
  p.pix_per_split = pps;
  const int M = 16 * p.Cin;
  return launch_auto<TcConvWgrad, true>(p, M, Cout, splits, st);
}

cudaError_t tc_fc_dgrad(const float* dy, int ldy, const float* W, const float* mul_act, const float* mul_plain, int ldmul,
                        float* dx, int lddx, int M, int K, int N, int accumulate, cudaStream_t st) {
  TcFcDgrad p{};
  p.dy = dy; p.W = W; p.mul_act = mul_act; p.mul_plain = mul_plain; p.dx = dx; p.M = M; p.K = K; p.N = N;
  p.ldy = ldy; p.lddx = lddx; p.ldmul = ldmul; p.accumulate = accumulate;
  return launch_auto<TcFcDgrad, false, 128>(p, M, K, 1, st);
}

cudaError_t tc_fc_wgrad(const float* x, int ldx, const float* dy, int ldy, float* gW, int M, int K, int N, int accumulate,
                        cudaStream_t st) {
  TcFcWgrad p{};
  p.x = x; p.dy = dy; p.gW = gW; p.M = M; p.K = K; p.N = N; p.ldx = ldx; p.ldy = ldy; p.accumulate = accumulate;
  return launch_auto<TcFcWgrad, true, 128>(p, K, N, 1, st);
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
  return launch_auto<TcFcFwd, false, 128>(p, M, N, splits, st);
}

}  // namespace expo
