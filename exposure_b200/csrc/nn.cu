// Policy / critic / value network primitives: 4x4 stride-2 SAME convolutions (fprop, dgrad,
// wgrad), fully connected layers (fwd, dgrad, wgrad), their forward-mode "tangent" variants
// used by the WGAN-GP second-order term, and column sums for bias gradients.
//
// Replaces the cuDNN / cuBLAS calls TF-1.6 issues for ly.conv2d / ly.fully_connected in
// agent.py:11-37,87-99, critics.py:6-38,94-97, filters.py:28-44 and their tf.gradients
// (Conv2DBackpropInput / Conv2DBackpropFilter), including the gradient-penalty double
// backward of net.py:174-194 (see DESIGN.md section 6 for the JVP formulation).
//
// Layouts follow the reference checkpoint: activations NHWC, conv weights HWIO
// [4,4,Cin,Cout], FC weights [in,out], flatten order (h*4+w)*256+c.
#include "gemm_engine.cuh"
#include "gemm_engine_v4.cuh"
#include "nn_internal.h"

namespace expo {

static int g_gemm_backend = kBackendAuto;
int gemm_backend() { return g_gemm_backend; }
// 0 (default): the TMA-fed tcgen05 engine for every contraction it supports (channel counts that are
// multiples of 32; the first layer through its staging copy), the exact-fp32 CUDA-core engine for the rest.
// 1: CUDA-core engine everywhere -- the A/B and bring-up switch of the tests, not a second product path.
bool use_tma() { return g_gemm_backend == kBackendAuto; }

__device__ __forceinline__ int ilog2(int v) { return 31 - __clz(v); }

// ------------------------------------------------------------------------------------------
// conv fprop:  y[b,oy,ox,co] = act( sum_{ky,kx,ci} in[b,2oy-1+ky,2ox-1+kx,ci] W[ky,kx,ci,co] + bias )
// in = concat(x[B,IH,IW,Cx], tile(vec[B,Cv])) - shift  (zero padding applied after the shift)
// ------------------------------------------------------------------------------------------
struct ConvFprop {
  const float* x; const float* vec; const float* W; const float* bias; const float* mask_ref;
  const float* post_mul; float* y; float* y2;
  int B, IH, IW, Cx, Cv, Cin, Cout, OH, OW, chunks, mode, lgOW, lgOHW;
  float shift;
  struct RowA { int b, iy0, ix0; };
  struct KS { int ky, kx, ci0, tap; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return 16 * chunks; }
  __device__ RowA row_a(int m) const {
    RowA r;
    if (m >= B * OH * OW) { r.b = -1; r.iy0 = r.ix0 = 0; return r; }
    r.b = m >> lgOHW;
    const int rem = m & ((1 << lgOHW) - 1);
    r.iy0 = 2 * (rem >> lgOW) - 1;
    r.ix0 = 2 * (rem & (OW - 1)) - 1;
    return r;
  }
  __device__ KS kstate(int ki) const {
    KS s;
    s.tap = ki / chunks;
    s.ci0 = (ki - s.tap * chunks) * kBK;
    s.ky = s.tap >> 2; s.kx = s.tap & 3;
    return s;
  }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int ci = s.ci0 + kk, iy = r.iy0 + s.ky, ix = r.ix0 + s.kx;
    if (r.b < 0 || ci >= Cin || (unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return 0.f;
    const float v = ci < Cx ? __ldg(x + ((size_t)(r.b * IH + iy) * IW + ix) * Cx + ci)
                            : __ldg(vec + (size_t)r.b * Cv + (ci - Cx));
    return v - shift;
  }
  __device__ float load_b(const KS& s, int kk, int n) const {
    const int ci = s.ci0 + kk;
    if (ci >= Cin || n >= Cout) return 0.f;
    return __ldg(W + ((size_t)s.tap * Cin + ci) * Cout + n);
  }
  // ---- vectorised engine (Cv == 0, Cx % 16 == 0, Cout % 4 == 0): K steps of 16 inside one tap
  __device__ int k_iters16() const { return 16 * (Cin / kBK4); }
  __device__ KS kstate16(int ki) const {
    KS s;
    const int c16 = Cin / kBK4;
    s.tap = ki / c16;
    s.ci0 = (ki - s.tap * c16) * kBK4;
    s.ky = s.tap >> 2; s.kx = s.tap & 3;
    return s;
  }
  __device__ float4 load_a_k4(const RowA& r, const KS& s, int c) const {
    const int iy = r.iy0 + s.ky, ix = r.ix0 + s.kx;
    if (r.b < 0 || (unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)(r.b * IH + iy) * IW + ix) * Cx + s.ci0 + 4 * c));
    v.x -= shift; v.y -= shift; v.z -= shift; v.w -= shift;
    return v;
  }
  __device__ float4 load_b_n4(const KS& s, int kk, int n) const {
    if (n >= Cout) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(W + ((size_t)s.tap * Cin + s.ci0 + kk) * Cout + n));
  }
  __device__ void store(int m, int n, float v) const {
    if (m >= B * OH * OW || n >= Cout) return;
    const size_t idx = (size_t)m * Cout + n;
    if (mode == 0) { v = lrelu_f(v + (bias ? __ldg(bias + n) : 0.f)); }
    else { v *= dlrelu_from_out(__ldg(mask_ref + idx)); }
    y[idx] = v;
    if (y2) y2[idx] = v * __ldg(post_mul + idx);
  }
};

// ------------------------------------------------------------------------------------------
// conv dgrad (transposed conv), one parity class (py,px) of the input grid per blockIdx.z:
// dx[b,2a+py,2c+px,ci] = sum_{j,l,co} dy[b,a+dy_j,c+dx_l,co] W[ky_j,kx_l,ci,co]
// ------------------------------------------------------------------------------------------
struct ConvDgrad {
  const float* dy; const float* W; const float* a_in; float* dx;
  int B, IH, IW, Cin, OH, OW, Cout, chunks, py, px, lgHW2, lgW2;
  struct RowA { int b, a, c; };
  struct KS { int oy_off, ox_off, tap, co0; };
  __device__ void init(int z) { py = z >> 1; px = z & 1; }
  __device__ int k_iters() const { return 4 * chunks; }
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
    const int t = ki / chunks;
    s.co0 = (ki - t * chunks) * kBK;
    const int j = t >> 1, l = t & 1;
    // iy = 2a+py: valid ky have parity (py+1)&1;  oy = (iy + 1 - ky) / 2
    const int ky = py == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
    const int kx = px == 0 ? (l == 0 ? 1 : 3) : (l == 0 ? 0 : 2);
    s.oy_off = py == 0 ? (j == 0 ? 0 : -1) : (j == 0 ? 1 : 0);
    s.ox_off = px == 0 ? (l == 0 ? 0 : -1) : (l == 0 ? 1 : 0);
    s.tap = ky * 4 + kx;
    return s;
  }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int oy = r.a + s.oy_off, ox = r.c + s.ox_off;
    if (r.b < 0 || (unsigned)oy >= (unsigned)OH || (unsigned)ox >= (unsigned)OW) return 0.f;
    return __ldg(dy + ((size_t)(r.b * OH + oy) * OW + ox) * Cout + s.co0 + kk);
  }
  __device__ float load_b(const KS& s, int kk, int n) const {
    if (n >= Cin) return 0.f;
    return __ldg(W + ((size_t)s.tap * Cin + n) * Cout + s.co0 + kk);
  }
  // ---- vectorised engine (Cout % 16 == 0)
  __device__ int k_iters16() const { return 4 * (Cout / kBK4); }
  __device__ KS kstate16(int ki) const {
    KS s;
    const int c16 = Cout / kBK4;
    const int t = ki / c16;
    s.co0 = (ki - t * c16) * kBK4;
    const int j = t >> 1, l = t & 1;
    const int ky = py == 0 ? (j == 0 ? 1 : 3) : (j == 0 ? 0 : 2);
    const int kx = px == 0 ? (l == 0 ? 1 : 3) : (l == 0 ? 0 : 2);
    s.oy_off = py == 0 ? (j == 0 ? 0 : -1) : (j == 0 ? 1 : 0);
    s.ox_off = px == 0 ? (l == 0 ? 0 : -1) : (l == 0 ? 1 : 0);
    s.tap = ky * 4 + kx;
    return s;
  }
  __device__ float4 load_a_k4(const RowA& r, const KS& s, int c) const {
    const int oy = r.a + s.oy_off, ox = r.c + s.ox_off;
    if (r.b < 0 || (unsigned)oy >= (unsigned)OH || (unsigned)ox >= (unsigned)OW) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(dy + ((size_t)(r.b * OH + oy) * OW + ox) * Cout + s.co0 + 4 * c));
  }
  __device__ float4 load_b_k4(const KS& s, int c, int n) const {
    if (n >= Cin) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(W + ((size_t)s.tap * Cin + n) * Cout + s.co0 + 4 * c));
  }
  __device__ void store(int m, int n, float v) const {
    if (m >= B * (IH / 2) * (IW / 2) || n >= Cin) return;
    const int b = m >> lgHW2;
    const int rem = m & ((1 << lgHW2) - 1);
    const int iy = 2 * (rem >> lgW2) + py, ix = 2 * (rem & ((IW / 2) - 1)) + px;
    const size_t idx = ((size_t)(b * IH + iy) * IW + ix) * Cin + n;
    if (a_in) v *= dlrelu_from_out(__ldg(a_in + idx));
    dx[idx] = v;
  }
};

// ------------------------------------------------------------------------------------------
// conv wgrad: gW[(tap,ci),co] = sum_{pixels} in[pixel shifted by tap, ci] * dy[pixel, co]
// split over blockIdx.z pixel ranges; partial sums go to part[z][16*Cin][Cout]
// ------------------------------------------------------------------------------------------
struct ConvWgrad {
  const float* x; const float* vec; const float* dy; float* part;
  int B, IH, IW, Cx, Cv, Cin, Cout, OH, OW, lgOW, lgOHW, pix_per_split, p_begin;
  float shift;
  struct RowA { int tap_ky, tap_kx, ci; };
  struct KS { int p0; };
  __device__ void init(int z) { p_begin = z * pix_per_split; part += (size_t)z * 16 * Cin * Cout; }
  __device__ int k_iters() const { return pix_per_split / kBK; }
  __device__ RowA row_a(int m) const {
    RowA r;
    if (m >= 16 * Cin) { r.ci = -1; r.tap_ky = r.tap_kx = 0; return r; }
    const int tap = m / Cin;
    r.ci = m - tap * Cin;
    r.tap_ky = tap >> 2; r.tap_kx = tap & 3;
    return r;
  }
  __device__ KS kstate(int ki) const { KS s; s.p0 = p_begin + ki * kBK; return s; }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int p = s.p0 + kk;
    if (r.ci < 0 || p >= B * OH * OW) return 0.f;
    const int b = p >> lgOHW;
    const int rem = p & ((1 << lgOHW) - 1);
    const int iy = 2 * (rem >> lgOW) - 1 + r.tap_ky, ix = 2 * (rem & (OW - 1)) - 1 + r.tap_kx;
    if ((unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return 0.f;
    const float v = r.ci < Cx ? __ldg(x + ((size_t)(b * IH + iy) * IW + ix) * Cx + r.ci)
                              : __ldg(vec + (size_t)b * Cv + (r.ci - Cx));
    return v - shift;
  }
  __device__ float load_b(const KS& s, int kk, int n) const {
    const int p = s.p0 + kk;
    if (p >= B * OH * OW || n >= Cout) return 0.f;
    return __ldg(dy + (size_t)p * Cout + n);
  }
  // ---- vectorised engine (Cv == 0, Cx % 4 == 0, Cout % 4 == 0): 4 consecutive ci per load
  __device__ int k_iters16() const { return pix_per_split / kBK4; }
  __device__ KS kstate16(int ki) const { KS s; s.p0 = p_begin + ki * kBK4; return s; }
  __device__ float4 load_a_m4(const KS& s, int kk, int m) const {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int p = s.p0 + kk;
    if (m >= 16 * Cin || p >= B * OH * OW) return z;
    const int tap = m / Cin, ci = m - tap * Cin;
    const int b = p >> lgOHW;
    const int rem = p & ((1 << lgOHW) - 1);
    const int iy = 2 * (rem >> lgOW) - 1 + (tap >> 2), ix = 2 * (rem & (OW - 1)) - 1 + (tap & 3);
    if ((unsigned)iy >= (unsigned)IH || (unsigned)ix >= (unsigned)IW) return z;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + ((size_t)(b * IH + iy) * IW + ix) * Cx + ci));
    v.x -= shift; v.y -= shift; v.z -= shift; v.w -= shift;
    return v;
  }
  __device__ float4 load_b_n4(const KS& s, int kk, int n) const {
    const int p = s.p0 + kk;
    if (p >= B * OH * OW || n >= Cout) return make_float4(0.f, 0.f, 0.f, 0.f);
    return __ldg(reinterpret_cast<const float4*>(dy + (size_t)p * Cout + n));
  }
  __device__ void store(int m, int n, float v) const {
    if (m >= 16 * Cin || n >= Cout) return;
    part[(size_t)m * Cout + n] = v;
  }
};

// ------------------------------------------------------------------------------------------
// fully connected: y = x[M,K] W[K,N] (split-K over blockIdx.z -> part[z][M][N])
// ------------------------------------------------------------------------------------------
struct FcFwd {
  const float* x; const float* W; float* part;
  int M, K, N, ldx, k_per_split, k_begin;
  struct RowA { int m; };
  struct KS { int k0; };
  __device__ void init(int z) { k_begin = z * k_per_split; part += (size_t)z * M * N; }
  __device__ int k_iters() const { return k_per_split / kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.m = m < M ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.k0 = k_begin + ki * kBK; return s; }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int k = s.k0 + kk;
    return (r.m < 0 || k >= K) ? 0.f : __ldg(x + (size_t)r.m * ldx + k);
  }
  __device__ float load_b(const KS& s, int kk, int n) const {
    const int k = s.k0 + kk;
    return (k >= K || n >= N) ? 0.f : __ldg(W + (size_t)k * N + n);
  }
  __device__ void store(int m, int n, float v) const {
    if (m < M && n < N) part[(size_t)m * N + n] = v;
  }
};

// dx[M,K] = (accumulate ? dx : 0) + dy[M,N] W^T * [lrelu'(mul_act)] * [mul_plain]
struct FcDgrad {
  const float* dy; const float* W; const float* mul_act; const float* mul_plain; float* dx;
  int M, K, N, ldy, lddx, ldmul, accumulate;
  struct RowA { int m; };
  struct KS { int n0; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return (N + kBK - 1) / kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.m = m < M ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.n0 = ki * kBK; return s; }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int n = s.n0 + kk;
    return (r.m < 0 || n >= N) ? 0.f : __ldg(dy + (size_t)r.m * ldy + n);
  }
  __device__ float load_b(const KS& s, int kk, int col) const {   // B[k'=n][col=k] = W[k][n]
    const int n = s.n0 + kk;
    return (n >= N || col >= K) ? 0.f : __ldg(W + (size_t)col * N + n);
  }
  __device__ void store(int m, int k, float v) const {
    if (m >= M || k >= K) return;
    const size_t im = (size_t)m * ldmul + k;
    if (mul_act) v *= dlrelu_from_out(__ldg(mul_act + im));
    if (mul_plain) v *= __ldg(mul_plain + im);
    const size_t idx = (size_t)m * lddx + k;
    dx[idx] = accumulate ? dx[idx] + v : v;
  }
};

// gW[K,N] = x^T[K,M] dy[M,N]
struct FcWgrad {
  const float* x; const float* dy; float* gW;
  int M, K, N, ldx, ldy, accumulate;
  struct RowA { int k; };
  struct KS { int s0; };
  __device__ void init(int) {}
  __device__ int k_iters() const { return (M + kBK - 1) / kBK; }
  __device__ RowA row_a(int m) const { RowA r; r.k = m < K ? m : -1; return r; }
  __device__ KS kstate(int ki) const { KS s; s.s0 = ki * kBK; return s; }
  __device__ float load_a(const RowA& r, const KS& s, int kk) const {
    const int smp = s.s0 + kk;
    return (r.k < 0 || smp >= M) ? 0.f : __ldg(x + (size_t)smp * ldx + r.k);
  }
  __device__ float load_b(const KS& s, int kk, int n) const {
    const int smp = s.s0 + kk;
    return (smp >= M || n >= N) ? 0.f : __ldg(dy + (size_t)smp * ldy + n);
  }
  __device__ void store(int k, int n, float v) const {
    if (k < K && n < N) { const size_t i = (size_t)k * N + n; gW[i] = accumulate ? gW[i] + v : v; }
  }
};

// out[row*ldo + col] = epi( sum_s part[s][row*ncols + col] ) (+ out when accumulate), fixed order.
// mode 0: lrelu(v + bias[col]); 1: v * dlrelu(mask_ref[row*ldmask+col]); 2: v + bias[col]; 3: v
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t count, int ncols,
                                     const float* __restrict__ bias, const float* __restrict__ mask_ref, int ldmask,
                                     int mode, float* __restrict__ out, int ldo, int accumulate) {
  EXP_PDL_ENTRY();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float v = 0.f;
  for (int s = 0; s < splits; ++s) v += part[(size_t)s * count + i];
  const int col = (int)(i % ncols);
  const size_t row = i / ncols;
  if (mode == 0) v = lrelu_f(v + (bias ? bias[col] : 0.f));
  else if (mode == 1) v *= dlrelu_from_out(mask_ref[row * ldmask + col]);
  else if (mode == 2) v += (bias ? bias[col] : 0.f);
  const size_t o = row * ldo + col;
  out[o] = accumulate ? out[o] + v : v;
}

// out[batch, cols] = column sums of a[batch, rows, cols] (bias gradients, per-image channel sums).
// Rows are split into gridDim.z chunks; chunk z of batch y writes out[(y * gridDim.z + z) * cols + col].
__global__ void colsum_kernel(const float* __restrict__ a, int rows, int cols, int rows_per_chunk,
                              float* __restrict__ out) {
  EXP_PDL_ENTRY();
  __shared__ float red[8][33];
  a += (size_t)blockIdx.y * rows * cols;
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const int r0 = blockIdx.z * rows_per_chunk;
  const int r1 = min(rows, r0 + rows_per_chunk);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (col < cols) {
    int r = r0 + ry;
    for (; r + 56 < r1; r += 64) {       // 8 independent loads in flight per thread
      const float v0 = a[(size_t)r * cols + col], v1 = a[(size_t)(r + 8) * cols + col];
      const float v2 = a[(size_t)(r + 16) * cols + col], v3 = a[(size_t)(r + 24) * cols + col];
      const float v4 = a[(size_t)(r + 32) * cols + col], v5 = a[(size_t)(r + 40) * cols + col];
      const float v6 = a[(size_t)(r + 48) * cols + col], v7 = a[(size_t)(r + 56) * cols + col];
      s0 += v0; s1 += v1; s2 += v2; s3 += v3;
      s0 += v4; s1 += v5; s2 += v6; s3 += v7;
    }
    for (; r < r1; r += 8) s0 += a[(size_t)r * cols + col];
  }
  red[ry][threadIdx.x & 31] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ry == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    out[((size_t)blockIdx.y * gridDim.z + blockIdx.z) * cols + col] = t;
  }
}
// out[b, col] = sum_z part[(b * chunks + z) * cols + col]   (fixed order)
__global__ void colsum_finish_kernel(const float* __restrict__ part, int chunks, int cols, int batch,
                                     float* __restrict__ out) {
  EXP_PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * cols) return;
  const int b = i / cols, col = i - b * cols;
  float s = 0.f;
  for (int z = 0; z < chunks; ++z) s += part[((size_t)b * chunks + z) * cols + col];
  out[i] = s;
}
// 256-row chunks: a thread then walks 32 rows (4 rounds of 8 loads); the partials [chunks][cols] are
// summed by the same kernel in a second pass (the serial 1024-row walk used to cost ~24 us a call)
static int colsum_chunks(int rows) {
  int c = (rows + 255) / 256;
  return c < 1 ? 1 : (c > 512 ? 512 : c);
}

// ------------------------------------------------------------------------------------------
// dgrad into a FEW input channels (first layer: Cin = 6 critic / 17 value network; the gradient
// w.r.t. the image, critics.py:48-87 through tf.gradients).  A 64 x 32 GEMM tile wastes most of
// its columns on N = 6 and the op is tiny (0.4 GFLOP, 15 MB).  One thread per INPUT pixel: a CTA
// covers a tile of th x tw positions (a, c) of the stride-2 grid, thread (class, a, c) computes input
// pixel (2a + py, 2c + px) of parity class (py, px) -- the class is uniform per warp, so its 4 taps
// and their weights are warp-uniform.  The (th + 2) x (tw + 2) neighbourhood of deltas is staged in
// shared memory by coalesced float4 loads (pixel pitch Cout + 4 floats: conflict-free float4 reads);
// the 16 x Cin x Cout weights are broadcast from shared memory.
// History: round 1 read the deltas from global memory at a 128-byte stride per lane (61 us for the
// critic's layer at batch 64, 124 us for the value network's, all L1 tag look-ups); one thread per
// 2 x 2 block of pixels with staged deltas: 40 / 132 us (14 warps per SM cannot hide the FMA chains).
// ------------------------------------------------------------------------------------------
// POS positions (a, c) per tile x 4 parity classes x CSPLIT slices of the input channels = 1024 threads per CTA:
// 256 positions for Cin <= 8; 128 positions x 2 channel slices of 10 for Cin <= 20 (20 accumulators per thread
// do not fit the 64 registers of a 1024-thread CTA)
struct DgsTile { int tw, th; };
static DgsTile dgs_tile(int OH, int OW, int pos) {
  DgsTile t;
  t.tw = OW < 32 ? OW : 32;
  t.th = pos / t.tw;
  if (t.th > OH) t.th = OH;
  return t;
}
static int dgs_pos(int Cin) { return Cin <= 8 ? 256 : 128; }
static size_t dgs_smem_bytes(int OH, int OW, int Cin, int Cout) {
  const DgsTile t = dgs_tile(OH, OW, dgs_pos(Cin));
  const size_t stage_in = (size_t)16 * Cin * Cout + (size_t)(t.th + 2) * (t.tw + 2) * (Cout + 4);
  const size_t stage_out = (size_t)4 * t.th * t.tw * Cin;           // the dx tile reuses the same memory
  return (stage_in > stage_out ? stage_in : stage_out) * sizeof(float);
}
template <int CMAX, int POS, int CSPLIT>
__global__ void __launch_bounds__(4 * POS * CSPLIT, 1) conv_dgrad_small_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                                       const float* __restrict__ a_in, float* __restrict__ dx,
                                                                       int IH, int IW, int Cin, int CinW, int Cout, int tw, int th,
                                                                       int tiles_x, int tiles_y) {
  EXP_PDL_ENTRY();
  extern __shared__ float4 dgs_smem4[];
  float4* const w_s4 = dgs_smem4;                                   // [ky][kx][Cin][Cout / 4]
  const int co4n = Cout >> 2, pitch4 = co4n + 1;
  float4* const d_s4 = dgs_smem4 + 16 * Cin * co4n;                 // [(th + 2)][(tw + 2)][pitch4]
  const int OH = IH >> 1, OW = IW >> 1;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int b = tile / tiles_y;
  const int a0 = ty * th, c0 = tx * tw;
  {
    // the first Cin of the CinW input channels of every tap (CinW > Cin: exp_conv_first_dgrad, image channels only)
    const float4* src = reinterpret_cast<const float4*>(W);
    const int per = Cin * co4n;
    for (int i = threadIdx.x; i < 16 * per; i += 4 * POS * CSPLIT) {
      const int tap = i / per, rem = i - tap * per;
      w_s4[i] = __ldg(src + (size_t)tap * CinW * co4n + rem);
    }
    const int hw = tw + 2, n4 = (th + 2) * hw * co4n;
    for (int i = threadIdx.x; i < n4; i += 4 * POS * CSPLIT) {
      const int q = i % co4n, pix = i / co4n;
      const int hc = pix % hw, hr = pix / hw;
      const int oy = a0 + hr - 1, ox = c0 + hc - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)oy < (unsigned)OH && (unsigned)ox < (unsigned)OW)
        v = __ldg(reinterpret_cast<const float4*>(dy + ((size_t)(b * OH + oy) * OW + ox) * Cout) + q);
      d_s4[pix * pitch4 + q] = v;
    }
  }
  __syncthreads();
  const int grp = threadIdx.x / POS, pos = threadIdx.x - grp * POS;      // grp is uniform per warp
  const int cls = grp & 3, ci0 = (grp >> 2) * CMAX;                      // parity class, first input channel of the slice
  const int py = cls >> 1, px = cls & 1;
  const int r = pos / tw, cc = pos - r * tw;
  const bool active = r < th;                       // OH, OW, th, tw are powers of two: every tile is full
  float acc[CMAX];
#pragma unroll
  for (int ci = 0; ci < CMAX; ++ci) acc[ci] = 0.f;
  // input row 2a + py takes output row a + dyy through tap ky = py - 2 dyy + 1 (stride 2, pad 1): dyy = 0 and the
  // neighbour on the side of the parity
  if (active) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int dyy = j == 0 ? 0 : (py ? 1 : -1);
      const int ky = py - 2 * dyy + 1;
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const int dxx = l == 0 ? 0 : (px ? 1 : -1);
        const int kx = px - 2 * dxx + 1;
        const float4* dp = d_s4 + ((r + 1 + dyy) * (tw + 2) + (cc + 1 + dxx)) * pitch4;
        const float4* wp = w_s4 + (size_t)(ky * 4 + kx) * Cin * co4n;
#pragma unroll 2
        for (int q = 0; q < co4n; ++q) {
          const float4 d = dp[q];
#pragma unroll
          for (int ci = 0; ci < CMAX; ++ci)
            if (ci0 + ci < Cin) {
              const float4 w = wp[(ci0 + ci) * co4n + q];
              acc[ci] = fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, fmaf(d.w, w.w, acc[ci]))));
            }
        }
      }
    }
  }
  // The tile of dx is 2 th rows of 2 tw x Cin contiguous floats: staged in shared memory (over the deltas and weights,
  // nobody needs them any more) and written out with full 128-byte lines -- a thread storing its own Cin floats
  // touches a sector per float.
  __syncthreads();
  float* const o_s = reinterpret_cast<float*>(dgs_smem4);
  const int rowlen = 2 * tw * Cin;
  if (active) {
    float* o = o_s + (2 * r + py) * rowlen + (2 * cc + px) * Cin + ci0;
#pragma unroll
    for (int ci = 0; ci < CMAX; ++ci)
      if (ci0 + ci < Cin) o[ci] = acc[ci];
  }
  __syncthreads();
  const int total = 2 * th * rowlen;
  for (int i = threadIdx.x; i < total; i += 4 * POS * CSPLIT) {
    const int row = i / rowlen, jj = i - row * rowlen;
    const size_t g = ((size_t)(b * IH + 2 * a0 + row) * IW + 2 * c0) * Cin + jj;
    float v = o_s[i];
    if (a_in) v *= dlrelu_from_out(__ldg(a_in + g));
    dx[g] = v;
  }
}
template <int CMAX, int POS, int CSPLIT>
static cudaError_t launch_dgrad_small(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW,
                                      int Cin, int CinW, int Cout, cudaStream_t stream) {
  const int OH = IH / 2, OW = IW / 2;
  const DgsTile t = dgs_tile(OH, OW, POS);
  const int tiles_x = (OW + t.tw - 1) / t.tw, tiles_y = (OH + t.th - 1) / t.th;
  const size_t smem = dgs_smem_bytes(OH, OW, Cin, Cout);
  static size_t opted = 0;                                       // per instantiation
  if (smem > 48 * 1024 && smem > opted) {
    const cudaError_t e = cudaFuncSetAttribute(conv_dgrad_small_kernel<CMAX, POS, CSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    opted = smem;
  }
  return launch_pdl(conv_dgrad_small_kernel<CMAX, POS, CSPLIT>, dim3((unsigned)(B * tiles_x * tiles_y)), dim3(4 * POS * CSPLIT), smem, stream, dy, W,
                    a_in, dx, IH, IW, Cin, CinW, Cout, t.tw, t.th, tiles_x, tiles_y);
}

// Cin = 3 (the image gradient of the first layer, exp_conv_first_dgrad) on 8 x 32 tiles: warp = row of the tile, lane =
// (parity class, l8), a thread owns positions l8 + 8 p (p < 4) of its row for its class: 4 x 3 accumulators, per tap and
// 4 output channels 4 delta loads + 3 weight loads for 48 FMAs (the generic kernel above: 1 + 3 loads for 12 FMAs, and
// at Cin = 3 it is bound by exactly those shared-memory loads: 27 us at batch 64).
constexpr int kDg3Threads = 256, kDg3TH = 8, kDg3TW = 32;
static size_t dg3_smem_bytes(int Cout) {
  const size_t stage_in = (size_t)16 * 3 * Cout + (size_t)(kDg3TH + 2) * (kDg3TW + 2) * (Cout + 4);
  const size_t stage_out = (size_t)4 * kDg3TH * kDg3TW * 3;
  return (stage_in > stage_out ? stage_in : stage_out) * sizeof(float);
}
__global__ void __launch_bounds__(kDg3Threads) conv_dgrad_img3_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                                      float* __restrict__ dx, int IH, int IW, int CinW, int Cout,
                                                                      int tiles_x, int tiles_y) {
  EXP_PDL_ENTRY();
  extern __shared__ float4 dg3_smem4[];
  constexpr int Cin = 3, tw = kDg3TW, th = kDg3TH;
  float4* const w_s4 = dg3_smem4;                                   // [ky][kx][3][Cout / 4]
  const int co4n = Cout >> 2, pitch4 = co4n + 1;
  float4* const d_s4 = dg3_smem4 + 16 * Cin * co4n;                 // [(th + 2)][(tw + 2)][pitch4]
  const int OH = IH >> 1, OW = IW >> 1;
  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int b = tile / tiles_y;
  const int a0 = ty * th, c0 = tx * tw;
  {
    const float4* src = reinterpret_cast<const float4*>(W);
    const int per = Cin * co4n;
    for (int i = threadIdx.x; i < 16 * per; i += kDg3Threads) {
      const int tap = i / per, rem = i - tap * per;
      w_s4[i] = __ldg(src + (size_t)tap * CinW * co4n + rem);
    }
    // (forcing 6 loads in flight per thread here measured SLOWER, 12.5 -> 14.2 us at batch 64: the co-resident CTAs
    // already hide each other's staging, and the extra registers cost occupancy)
    const int hw = tw + 2, n4 = (th + 2) * hw * co4n;
    for (int i = threadIdx.x; i < n4; i += kDg3Threads) {
      const int q = i % co4n, pix = i / co4n;
      const int hc = pix % hw, hr = pix / hw;
      const int oy = a0 + hr - 1, ox = c0 + hc - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((unsigned)oy < (unsigned)OH && (unsigned)ox < (unsigned)OW)
        v = __ldg(reinterpret_cast<const float4*>(dy + ((size_t)(b * OH + oy) * OW + ox) * Cout) + q);
      d_s4[pix * pitch4 + q] = v;
    }
  }
  __syncthreads();
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31, cls = lane >> 3, l8 = lane & 7;
  const int py = cls >> 1, px = cls & 1;
  float acc[4][3];
#pragma unroll
  for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = acc[p][2] = 0.f;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int dyy = j == 0 ? 0 : (py ? 1 : -1);
    const int ky = py - 2 * dyy + 1;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const int dxx = l == 0 ? 0 : (px ? 1 : -1);
      const int kx = px - 2 * dxx + 1;
      const float4* dp = d_s4 + ((r + 1 + dyy) * (tw + 2) + (l8 + 1 + dxx)) * pitch4;
      const float4* wp = w_s4 + (size_t)(ky * 4 + kx) * Cin * co4n;
#pragma unroll 2
      for (int q = 0; q < co4n; ++q) {
        const float4 w0 = wp[q], w1 = wp[co4n + q], w2 = wp[2 * co4n + q];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float4 d = dp[8 * p * pitch4 + q];
          acc[p][0] = fmaf(d.x, w0.x, fmaf(d.y, w0.y, fmaf(d.z, w0.z, fmaf(d.w, w0.w, acc[p][0]))));
          acc[p][1] = fmaf(d.x, w1.x, fmaf(d.y, w1.y, fmaf(d.z, w1.z, fmaf(d.w, w1.w, acc[p][1]))));
          acc[p][2] = fmaf(d.x, w2.x, fmaf(d.y, w2.y, fmaf(d.z, w2.z, fmaf(d.w, w2.w, acc[p][2]))));
        }
      }
    }
  }
  __syncthreads();                                   // deltas and weights are dead: the dx tile takes their place
  float* const o_s = reinterpret_cast<float*>(dg3_smem4);
  constexpr int rowlen = 2 * tw * Cin;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float* o = o_s + (2 * r + py) * rowlen + (2 * (l8 + 8 * p) + px) * Cin;
    o[0] = acc[p][0]; o[1] = acc[p][1]; o[2] = acc[p][2];
  }
  __syncthreads();
  constexpr int total = 2 * th * rowlen;
  for (int i = threadIdx.x; i < total; i += kDg3Threads) {
    const int row = i / rowlen, jj = i - row * rowlen;
    dx[((size_t)(b * IH + 2 * a0 + row) * IW + 2 * c0) * Cin + jj] = o_s[i];
  }
}

// dx[B,IH,IW,Cin] from the first Cin of CinW weight channels; the caller has checked dgs_smem_bytes / alignment / pow2 sizes
cudaError_t conv_dgrad_small_launch(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                                    int CinW, int Cout, cudaStream_t stream) {
  const int OH = IH / 2, OW = IW / 2;
  if (Cin == 3 && !a_in && OW % kDg3TW == 0 && OH % kDg3TH == 0 && dg3_smem_bytes(Cout) <= 100 * 1024) {
    const size_t smem = dg3_smem_bytes(Cout);
    static size_t opted = 0;
    if (smem > 48 * 1024 && smem > opted) {
      const cudaError_t e = cudaFuncSetAttribute(conv_dgrad_img3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      opted = smem;
    }
    const int tiles_x = OW / kDg3TW, tiles_y = OH / kDg3TH;
    return launch_pdl(conv_dgrad_img3_kernel, dim3((unsigned)(B * tiles_x * tiles_y)), dim3(kDg3Threads), smem, stream, dy, W, dx, IH, IW,
                      CinW, Cout, tiles_x, tiles_y);
  }
  return Cin <= 8 ? launch_dgrad_small<8, 256, 1>(dy, W, a_in, dx, B, IH, IW, Cin, CinW, Cout, stream)
                  : launch_dgrad_small<10, 128, 2>(dy, W, a_in, dx, B, IH, IW, Cin, CinW, Cout, stream);
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static int host_ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int wgrad_splits(int B, int OH, int OW, int Cin, int Cout) {
  const int pixels = B * OH * OW;
  const int tiles = ((16 * Cin + kBM - 1) / kBM) * ((Cout + 63) / 64);
  int s = (592 + tiles - 1) / tiles;            // aim at ~4 CTAs per SM
  const int max_s = pixels / 64 > 0 ? pixels / 64 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}
constexpr int kFcKU = 4;      // K-steps gathered per global-memory round trip by the fully connected GEMMs (gemm_engine.cuh)
static int fc_splits(int M, int K, int N) {
  const int tiles = ((M + kBM - 1) / kBM) * ((N + 63) / 64);
  int s = (296 + tiles - 1) / tiles;
  const int max_s = K / 64 > 0 ? K / 64 : 1;
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  return s;
}

}  // namespace expo

using namespace expo;

extern "C" {

int exp_conv_fwd(const float* x, int Cx, const float* vec, int Cv, float shift, const float* W, const float* bias,
                 const float* mask_ref, const float* post_mul, float* y, float* y2, int B, int IH, int IW, int Cout,
                 int mode, void* stream) {
  EXP_CHECK_ARG(x && W && y, "null pointer");
  EXP_CHECK_ARG(B > 0 && is_pow2(IH) && is_pow2(IW) && IH >= 2 && IW >= 2, "IH/IW must be powers of two >= 2 (got %dx%d)", IH, IW);
  EXP_CHECK_ARG(Cx > 0 && Cv >= 0 && (Cv == 0 || vec) && Cout > 0, "bad channel counts");
  EXP_CHECK_ARG(mode == 0 || (mode == 1 && mask_ref), "mode 1 needs mask_ref");
  EXP_CHECK_ARG(!y2 || post_mul, "y2 needs post_mul");
  ConvFprop p{};
  p.x = x; p.vec = vec; p.W = W; p.bias = bias; p.mask_ref = mask_ref; p.post_mul = post_mul; p.y = y; p.y2 = y2;
  p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cx + Cv; p.Cout = Cout; p.OH = IH / 2; p.OW = IW / 2;
  p.chunks = (p.Cin + kBK - 1) / kBK; p.mode = mode; p.shift = shift;
  p.lgOW = host_ilog2(p.OW); p.lgOHW = host_ilog2(p.OH * p.OW);
  const int M = B * p.OH * p.OW;
  if (use_tma() && tma_conv_fwd_supported(x, Cx, Cv, shift, W, bias, mask_ref, post_mul, y, y2, Cout)) {
    const cudaError_t e = tma_conv_fwd(x, Cx, W, bias, mask_ref, post_mul, y, y2, B, IH, IW, Cout, mode, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_fwd[tma]: %s", cudaGetErrorString(e));
    return EXP_OK;
  }
  if (Cv == 0 && Cx % kBK4 == 0 && Cout % 4 == 0 && aligned16(x) && aligned16(W)) {   // 16-byte gathers
    // the deep layers have few output rows: prefer the narrow N tile while the 64-wide grid
    // would leave SMs idle (2 x 148 CTAs), A is simply re-gathered from L2 per N tile
    const bool narrow = Cout <= 32 || ((M + kBM - 1) / kBM) * ((Cout + 63) / 64) < 296;
    if (narrow) launch_gemm_v4<ConvFprop, 32, false, false>(p, M, Cout, 1, (cudaStream_t)stream);
    else launch_gemm_v4<ConvFprop, 64, false, false>(p, M, Cout, 1, (cudaStream_t)stream);
  } else if (Cout <= 32) launch_gemm<ConvFprop, 32, false, false>(p, M, Cout, 1, (cudaStream_t)stream);
  else launch_gemm<ConvFprop, 64, false, false>(p, M, Cout, 1, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_conv_fwd");
  return EXP_OK;
}

int exp_conv_dgrad(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                   int Cout, void* stream) {
  EXP_CHECK_ARG(dy && W && dx, "null pointer");
  EXP_CHECK_ARG(B > 0 && is_pow2(IH) && is_pow2(IW) && IH >= 2 && IW >= 2, "IH/IW must be powers of two >= 2");
  EXP_CHECK_ARG(Cin > 0 && Cout > 0 && Cout % kBK == 0, "Cout must be a multiple of %d", kBK);
  ConvDgrad p{};
  p.dy = dy; p.W = W; p.a_in = a_in; p.dx = dx; p.B = B; p.IH = IH; p.IW = IW; p.Cin = Cin; p.OH = IH / 2;
  p.OW = IW / 2; p.Cout = Cout; p.chunks = Cout / kBK;
  p.lgW2 = host_ilog2(IW / 2); p.lgHW2 = host_ilog2((IH / 2) * (IW / 2));
  const int M = B * (IH / 2) * (IW / 2);
  if (use_tma() && Cin <= 20 && Cout % 4 == 0 && dgs_smem_bytes(IH / 2, IW / 2, Cin, Cout) <= 110 * 1024 && aligned16(dy) &&
      aligned16(W) && (long long)B * (IH / 2) * (IW / 2) < (1ll << 31)) {
    const cudaError_t e = conv_dgrad_small_launch(dy, W, a_in, dx, B, IH, IW, Cin, Cin, Cout, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_dgrad[small]: %s", cudaGetErrorString(e));
    EXP_CHECK_LAUNCH("exp_conv_dgrad[small]");
    return EXP_OK;
  }
  if (use_tma() && tma_conv_dgrad_supported(dy, W, a_in, dx, Cin, Cout)) {
    const cudaError_t e = tma_conv_dgrad(dy, W, a_in, dx, B, IH, IW, Cin, Cout, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_dgrad[tma]: %s", cudaGetErrorString(e));
    return EXP_OK;
  }
  if (Cout % kBK4 == 0 && aligned16(dy) && aligned16(W)) {                             // 16-byte gathers
    const bool narrow = Cin <= 32 || 4 * ((M + kBM - 1) / kBM) * ((Cin + 63) / 64) < 296;
    if (narrow) launch_gemm_v4<ConvDgrad, 32, false, true>(p, M, Cin, 4, (cudaStream_t)stream);
    else launch_gemm_v4<ConvDgrad, 64, false, true>(p, M, Cin, 4, (cudaStream_t)stream);
  } else if (Cin <= 32) launch_gemm<ConvDgrad, 32, false, true>(p, M, Cin, 4, (cudaStream_t)stream);
  else launch_gemm<ConvDgrad, 64, false, true>(p, M, Cin, 4, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_conv_dgrad");
  return EXP_OK;
}

size_t exp_conv_wgrad_workspace_bytes(int B, int IH, int IW, int Cin, int Cout) {
  if (B <= 0 || IH < 2 || IW < 2 || Cin <= 0 || Cout <= 0) return 0;
  const int a = wgrad_splits(B, IH / 2, IW / 2, Cin, Cout), c = tma_wgrad_splits(B, IH / 2, IW / 2, Cin, Cout);
  const int m = a > c ? a : c;
  return (size_t)m * 16 * Cin * Cout * sizeof(float);
}

int exp_conv_wgrad(const float* x, int Cx, const float* vec, int Cv, float shift, const float* dy, float* gW, int B,
                   int IH, int IW, int Cout, int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  EXP_CHECK_ARG(x && dy && gW && workspace, "null pointer");
  EXP_CHECK_ARG(B > 0 && is_pow2(IH) && is_pow2(IW) && IH >= 2 && IW >= 2, "IH/IW must be powers of two >= 2");
  EXP_CHECK_ARG(Cx > 0 && Cv >= 0 && (Cv == 0 || vec) && Cout > 0, "bad channel counts");
  const int Cin = Cx + Cv, OH = IH / 2, OW = IW / 2;
  const bool tmab = use_tma() && tma_conv_wgrad_supported(x, Cx, Cv, shift, dy, Cout) && aligned16(workspace);
  const int splits = tmab ? tma_wgrad_splits(B, OH, OW, Cin, Cout) : wgrad_splits(B, OH, OW, Cin, Cout);
  const size_t need = (size_t)splits * 16 * Cin * Cout * sizeof(float);
  if (workspace_bytes < need) return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (tmab) {
    const cudaError_t e = tma_conv_wgrad_partials(x, Cx, dy, reinterpret_cast<float*>(workspace), B, IH, IW, Cout, splits,
                                                  (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_wgrad[tma]: %s", cudaGetErrorString(e));
    const size_t cnt = (size_t)16 * Cin * Cout;
    launch_pdl(splitk_reduce_kernel, dim3((unsigned)((cnt + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<float*>(workspace), splits, cnt, Cout, nullptr, nullptr, 0, 3, gW, Cout, accumulate);
    EXP_CHECK_LAUNCH("exp_conv_wgrad[reduce]");
    return EXP_OK;
  }
  ConvWgrad p{};
  p.x = x; p.vec = vec; p.dy = dy; p.part = reinterpret_cast<float*>(workspace);
  p.B = B; p.IH = IH; p.IW = IW; p.Cx = Cx; p.Cv = Cv; p.Cin = Cin; p.Cout = Cout; p.OH = OH; p.OW = OW;
  p.lgOW = host_ilog2(OW); p.lgOHW = host_ilog2(OH * OW); p.shift = shift;
  const int pixels = B * OH * OW;
  int pps = (pixels + splits - 1) / splits;
  pps = ((pps + kBK4 - 1) / kBK4) * kBK4;
  p.pix_per_split = pps;
  if (Cv == 0 && Cx % 4 == 0 && Cout % 4 == 0 && aligned16(x) && aligned16(dy)) {         // 16-byte gathers
    if (Cout <= 32) launch_gemm_v4<ConvWgrad, 32, true, false>(p, 16 * Cin, Cout, splits, (cudaStream_t)stream);
    else launch_gemm_v4<ConvWgrad, 64, true, false>(p, 16 * Cin, Cout, splits, (cudaStream_t)stream);
  } else if (Cout <= 32) launch_gemm<ConvWgrad, 32, true, false>(p, 16 * Cin, Cout, splits, (cudaStream_t)stream);
  else launch_gemm<ConvWgrad, 64, true, false>(p, 16 * Cin, Cout, splits, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_conv_wgrad");
  const size_t count = (size_t)16 * Cin * Cout;
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)((count + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      p.part, splits, count, Cout, nullptr, nullptr, 0, 3, gW, Cout, accumulate);
  EXP_CHECK_LAUNCH("exp_conv_wgrad[reduce]");
  return EXP_OK;
}

size_t exp_fc_workspace_bytes(int M, int K, int N) {
  if (M <= 0 || K <= 0 || N <= 0) return 0;
  return (size_t)fc_splits(M, K, N) * M * N * sizeof(float);
}

int exp_set_gemm_backend(int backend) {
  EXP_CHECK_ARG(backend == kBackendAuto || backend == kBackendSimt,
                "backend must be 0 (TMA-fed tcgen05 where supported, default) or 1 (exact-fp32 CUDA cores everywhere)");
  g_gemm_backend = backend;
  return EXP_OK;
}

int exp_fc_fwd(const float* x, int ldx, const float* W, const float* bias, const float* mask_ref, int ldmask, float* y,
               int ldy, int M, int K, int N, int mode, void* workspace, size_t workspace_bytes, void* stream) {
  EXP_CHECK_ARG(x && W && y && workspace, "null pointer");
  EXP_CHECK_ARG(M > 0 && K > 0 && N > 0 && ldx >= K && ldy >= N, "bad shape");
  EXP_CHECK_ARG(mode >= 0 && mode <= 3 && (mode != 1 || (mask_ref && ldmask >= N)), "bad mode");
  const int splits = fc_splits(M, K, N);
  const size_t need = (size_t)splits * M * N * sizeof(float);
  if (workspace_bytes < need) return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  FcFwd p{};
  p.x = x; p.W = W; p.part = reinterpret_cast<float*>(workspace); p.M = M; p.K = K; p.N = N; p.ldx = ldx;
  int kps = (K + splits - 1) / splits;
  kps = ((kps + kBK - 1) / kBK) * kBK;
  p.k_per_split = kps;
  if (N <= 32) launch_gemm<FcFwd, 32, false, false, kFcKU>(p, M, N, splits, (cudaStream_t)stream);
  else launch_gemm<FcFwd, 64, false, false, kFcKU>(p, M, N, splits, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_fc_fwd");
  const size_t count = (size_t)M * N;
  launch_pdl(splitk_reduce_kernel, dim3((unsigned)((count + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      p.part, splits, count, N, bias, mask_ref, ldmask, mode, y, ldy, 0);
  EXP_CHECK_LAUNCH("exp_fc_fwd[reduce]");
  return EXP_OK;
}

int exp_fc_dgrad(const float* dy, int ldy, const float* W, const float* mul_act, const float* mul_plain, int ldmul,
                 float* dx, int lddx, int M, int K, int N, int accumulate, void* stream) {
  EXP_CHECK_ARG(dy && W && dx, "null pointer");
  EXP_CHECK_ARG(M > 0 && K > 0 && N > 0 && ldy >= N && lddx >= K, "bad shape");
  EXP_CHECK_ARG((!mul_act && !mul_plain) || ldmul >= K, "bad ldmul");
  FcDgrad p{};
  p.dy = dy; p.W = W; p.mul_act = mul_act; p.mul_plain = mul_plain; p.dx = dx; p.M = M; p.K = K; p.N = N;
  p.ldy = ldy; p.lddx = lddx; p.ldmul = ldmul; p.accumulate = accumulate;
  // the 4096 x 128 heads at batch 64: 64-wide tiles give 64 CTAs on 148 SMs -> halve the tile until the grid fills them
  if (((M + kBM - 1) / kBM) * ((K + 63) / 64) < 148) launch_gemm<FcDgrad, 32, false, true, kFcKU>(p, M, K, 1, (cudaStream_t)stream);
  else launch_gemm<FcDgrad, 64, false, true, kFcKU>(p, M, K, 1, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_fc_dgrad");
  return EXP_OK;
}

int exp_fc_wgrad(const float* x, int ldx, const float* dy, int ldy, float* gW, int M, int K, int N, int accumulate,
                 void* stream) {
  EXP_CHECK_ARG(x && dy && gW, "null pointer");
  EXP_CHECK_ARG(M > 0 && K > 0 && N > 0 && ldx >= K && ldy >= N, "bad shape");
  FcWgrad p{};
  p.x = x; p.dy = dy; p.gW = gW; p.M = M; p.K = K; p.N = N; p.ldx = ldx; p.ldy = ldy; p.accumulate = accumulate;
  if (N <= 32) launch_gemm<FcWgrad, 32, true, false, kFcKU>(p, K, N, 1, (cudaStream_t)stream);
  else launch_gemm<FcWgrad, 64, true, false, kFcKU>(p, K, N, 1, (cudaStream_t)stream);
  EXP_CHECK_LAUNCH("exp_fc_wgrad");
  return EXP_OK;
}

size_t exp_colsum_workspace_bytes(int batch, int rows, int cols) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return 0;
  const int chunks = colsum_chunks(rows);
  return chunks > 1 ? (size_t)batch * chunks * cols * sizeof(float) : 0;
}

int exp_colsum(const float* a, int batch, int rows, int cols, float* out, void* workspace, size_t workspace_bytes,
               void* stream) {
  EXP_CHECK_ARG(a && out && batch > 0 && batch <= 65535 && rows > 0 && cols > 0, "bad args");
  int chunks = colsum_chunks(rows);
  const size_t need = (size_t)batch * chunks * cols * sizeof(float);
  if (chunks > 1 && (!workspace || workspace_bytes < need)) chunks = 1;   // single pass (slower, still exact)
  const int rpc = (rows + chunks - 1) / chunks;
  dim3 grid((cols + 31) / 32, batch, chunks);
  if (chunks == 1) {
    launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, rows, cols, rpc, out);
    EXP_CHECK_LAUNCH("exp_colsum");
    return EXP_OK;
  }
  launch_pdl(colsum_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, a, rows, cols, rpc, reinterpret_cast<float*>(workspace));
  EXP_CHECK_LAUNCH("exp_colsum");
  if (chunks <= 16) {
    launch_pdl(colsum_finish_kernel, dim3((batch * cols + 255) / 256), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const float*>(workspace), chunks, cols, batch, out);
  } else {
    launch_pdl(colsum_kernel, dim3(dim3((cols + 31) / 32, batch, 1)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const float*>(workspace), chunks, cols, chunks, out);
  }
  EXP_CHECK_LAUNCH("exp_colsum[finish]");
  return EXP_OK;
}

}  // extern "C"
