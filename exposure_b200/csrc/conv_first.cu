// First convolution of the policy / critic / value CNNs (agent.py:11-37, critics.py:6-38, 42-98): a 4x4 / stride 2
// SAME convolution over  concat(image[B,IH,IW,3], tile(vec[B,Cv])) - shift  into 32 channels.
//
// Only the 3 image channels vary over the pixels.  The other Cv channels (11 agent states, 3 statistics, or both) are
// constants of the image, so their share of every output is one of at most 16 per-image values per output channel:
// which of the 4 x 4 taps fall inside the image depends only on whether the output row / column is the first and / or
// the last one ("border class").  That splits the layer into
//
//   forward   y[b,oy,ox,:] = act( T[b][class(oy,ox)][:] + sum_{taps, c<3} (x - shift)[tap, c] W[tap][c][:] )
//             T[b][cls][:] = bias + sum_{taps inside for cls} V[b][tap][:],   V[b][tap][:] = sum_c (vec[b][c] - shift) W[tap][3+c][:]
//   wgrad     c <  3: gW[tap][c][:]   = sum_{b,oy,ox} (x - shift)[tap, c] dy[b,oy,ox,:]
//             c >= 3: gW[tap][3+c][:] = sum_b (vec[b][c] - shift) E[b][tap][:],  E[b][tap][:] = sum_{(oy,ox): tap inside} dy[b,oy,ox,:]
//   dgrad     image channels only (conv_dgrad_small_kernel, nn.cu) + the per-image SUM over the pixels of the constant
//             channels' gradient,  gvec[b][c] = sum_{tap} <W[tap][3+c][:], E[b][tap][:]>  -- all the critic needs of
//             them (the statistic channels are pulled back through J_stats^T per image, critics.py:48-87).
//
// i.e. K = 48 instead of 16 * (3 + Cv) = 96 ... 272: exact-fp32 CUDA-core kernels that read the image itself.  Round 1
// ran the layer on the TMA-fed tcgen05 engine over a 16-channel zero-bordered staging copy (K = 256): 12 us staging +
// 42 us GEMM per forward at batch 64, shared-memory-bandwidth bound on N = 32 tiles (tools/tma_trace.py).
#include "common.cuh"

namespace expo {
namespace first {

constexpr int kTH = 8, kTW = 32;                  // output tile: one warp per row, 8 lanes x 4 pixels per row
constexpr int kThreads = 256;
constexpr int kCout = 32, kCx = 3, kK = 16 * kCx;  // 48
constexpr int kInRows = 2 * kTH + 2;               // 18
constexpr int kInCols = 2 * kTW + 2;               // 66
constexpr int kPitch = kInCols * kCx + 2;          // 200 floats: rows stay 8-byte aligned
constexpr int kMaxCv = 32;

__device__ __forceinline__ float lrelu(float v) { return 0.6f * v + 0.4f * fabsf(v); }                       // util.py:225-229
__device__ __forceinline__ float dlrelu_from_out(float a) { return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f); }

// border class of an output row / column: bit 0 = first, bit 1 = last; tap k (0..3) reads input 2 o + k - 1
__device__ __forceinline__ int border_class(int o, int n) { return (o == 0 ? 1 : 0) | (o == n - 1 ? 2 : 0); }
__device__ __forceinline__ bool tap_inside(int k, int cls) { return !((k == 0 && (cls & 1)) || (k == 3 && (cls & 2))); }

struct Tile { int b, oy0, ox0; };
__device__ __forceinline__ Tile decode_tile(int t, int tiles_x, int tiles_y) {
  Tile r;
  r.ox0 = (t % tiles_x) * kTW; t /= tiles_x;
  r.oy0 = (t % tiles_y) * kTH;
  r.b = t / tiles_y;
  return r;
}

// (x - shift) of the tile's receptive field, zero outside the image: xs[row 2 (oy - oy0) + ky][col 2 (ox - ox0) + kx][c]
__device__ __forceinline__ void load_input_tile(float* xs, const float* __restrict__ x, const Tile& t, int IH, int IW, float shift) {
  const int iy0 = 2 * t.oy0 - 1, ix0 = 2 * t.ox0 - 1;
  batched_fill<7>(kInRows * kInCols * kCx, kThreads,
                  [&](int i) {
                    const int r = i / (kInCols * kCx), j = i - r * (kInCols * kCx);
                    const int iy = iy0 + r, ix = ix0 + j / kCx;
                    float v = 0.f;
                    if ((unsigned)iy < (unsigned)IH && (unsigned)ix < (unsigned)IW)
                      v = __ldg(x + ((size_t)(t.b * IH + iy) * IW + ix0) * kCx + j) - shift;
                    return v;
                  },
                  [&](int i, float v) {
                    const int r = i / (kInCols * kCx), j = i - r * (kInCols * kCx);
                    xs[r * kPitch + j] = v;
                  });
}

// ------------------------------------------------------------------------------------------------ forward
struct FwdArgs {
  const float *x, *vec, *W, *bias, *mask_ref, *post_mul;
  float *y, *y2;
  int IH, IW, Cv, mode, tiles_x, tiles_y;
  float shift;
};

__global__ void __launch_bounds__(kThreads) conv_first_fwd_kernel(const FwdArgs A) {
  EXP_PDL_ENTRY();
  __shared__ __align__(16) float xs[kInRows * kPitch];
  __shared__ __align__(16) float ws[kK * kCout];
  __shared__ float V[16 * kCout], T[16 * kCout], vs[kMaxCv];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int OH = A.IH >> 1, OW = A.IW >> 1, Cin = kCx + A.Cv;
  const Tile t = decode_tile(blockIdx.x, A.tiles_x, A.tiles_y);
  batched_fill<6>(kK * kCout, kThreads,
                  [&](int i) {
                    const int co = i & 31, k = i >> 5, tap = k / kCx, c = k - tap * kCx;
                    return __ldg(A.W + ((size_t)tap * Cin + c) * kCout + co);
                  },
                  [&](int i, float v) { ws[i] = v; });
  if (tid < A.Cv) vs[tid] = __ldg(A.vec + (size_t)t.b * A.Cv + tid) - A.shift;
  load_input_tile(xs, A.x, t, A.IH, A.IW, A.shift);
  __syncthreads();
  for (int i = tid; i < 16 * kCout; i += kThreads) {                    // V[tap][co], 8 weight loads in flight
    const int co = i & 31, tap = i >> 5;
    const float* w = A.W + ((size_t)tap * Cin + kCx) * kCout + co;
    float s = 0.f;
    for (int c0 = 0; c0 < A.Cv; c0 += 8) {
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = c0 + u < A.Cv ? __ldg(w + (size_t)(c0 + u) * kCout) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < A.Cv) s = fmaf(vs[c0 + u], wv[u], s);
    }
    V[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < 16 * kCout; i += kThreads) {                    // T[class][co]
    const int co = i & 31, cls = i >> 5, rc = cls >> 2, cc = cls & 3;
    float s = A.bias ? __ldg(A.bias + co) : 0.f;
#pragma unroll
    for (int ky = 0; ky < 4; ++ky)
#pragma unroll
      for (int kx = 0; kx < 4; ++kx)
        if (tap_inside(ky, rc) && tap_inside(kx, cc)) s += V[(ky * 4 + kx) * kCout + co];
    T[i] = s;
  }
  __syncthreads();

  // warp = output row of the tile; lane = (group of 4 adjacent pixels, 4 + 4 output channels cg*4.. and 16 + cg*4..)
  const int cg = lane & 3, pg = lane >> 2, r = warp;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4* ws4 = reinterpret_cast<const float4*>(ws);
#pragma unroll
  for (int ky = 0; ky < 4; ++ky) {
    float xv[30];                                                       // 10 input columns x 3 channels of this tap row
    const float2* xr = reinterpret_cast<const float2*>(xs + (2 * r + ky) * kPitch + 24 * pg);
#pragma unroll
    for (int i = 0; i < 15; ++i) { const float2 v = xr[i]; xv[2 * i] = v.x; xv[2 * i + 1] = v.y; }
#pragma unroll
    for (int kx = 0; kx < 4; ++kx)
#pragma unroll
      for (int c = 0; c < kCx; ++c) {
        const int k = (ky * 4 + kx) * kCx + c;
        const float4 w0 = ws4[k * 8 + cg], w1 = ws4[k * 8 + 4 + cg];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xval = xv[(2 * i + kx) * kCx + c];
          acc[i][0] = fmaf(xval, w0.x, acc[i][0]); acc[i][1] = fmaf(xval, w0.y, acc[i][1]);
          acc[i][2] = fmaf(xval, w0.z, acc[i][2]); acc[i][3] = fmaf(xval, w0.w, acc[i][3]);
          acc[i][4] = fmaf(xval, w1.x, acc[i][4]); acc[i][5] = fmaf(xval, w1.y, acc[i][5]);
          acc[i][6] = fmaf(xval, w1.z, acc[i][6]); acc[i][7] = fmaf(xval, w1.w, acc[i][7]);
        }
      }
  }
  const int oy = t.oy0 + r;
  if (oy >= OH) return;
  const int rc = border_class(oy, OH);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = t.ox0 + pg * 4 + i;
    if (ox >= OW) continue;
    const float* Tc = T + (rc * 4 + border_class(ox, OW)) * kCout;
    const size_t o = ((size_t)(t.b * OH + oy) * OW + ox) * kCout;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int co = h * 16 + cg * 4;
      float4 v = make_float4(acc[i][4 * h] + Tc[co], acc[i][4 * h + 1] + Tc[co + 1], acc[i][4 * h + 2] + Tc[co + 2],
                             acc[i][4 * h + 3] + Tc[co + 3]);
      if (A.mode == 0) {
        v.x = lrelu(v.x); v.y = lrelu(v.y); v.z = lrelu(v.z); v.w = lrelu(v.w);
      } else {
        const float4 m = __ldg(reinterpret_cast<const float4*>(A.mask_ref + o + co));
        v.x *= dlrelu_from_out(m.x); v.y *= dlrelu_from_out(m.y); v.z *= dlrelu_from_out(m.z); v.w *= dlrelu_from_out(m.w);
      }
      *reinterpret_cast<float4*>(A.y + o + co) = v;
      if (A.y2) {
        const float4 pm = __ldg(reinterpret_cast<const float4*>(A.post_mul + o + co));
        *reinterpret_cast<float4*>(A.y2 + o + co) = make_float4(v.x * pm.x, v.y * pm.y, v.z * pm.z, v.w * pm.w);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ class sums of dy
// bins[cls][slice][co] (cls = row class * 4 + column class): every thread (slice, co) owns its 16 bins, so the
// dynamically indexed accumulation needs neither atomics nor local memory and its order is fixed.
__device__ __forceinline__ void reduce_bins_to_E(const float* bins, float* R, float* E, int tid) {
  for (int i = tid; i < 16 * kCout; i += kThreads) {                    // R[cls][co] = sum over the 8 slices, in order
    const int co = i & 31, cls = i >> 5;
    float s = 0.f;
#pragma unroll
    for (int sl = 0; sl < 8; ++sl) s += bins[(cls * 8 + sl) * kCout + co];
    R[i] = s;
  }
  __syncthreads();
  for (int i = tid; i < 16 * kCout; i += kThreads) {                    // E[tap][co] = sum of the classes the tap is inside of
    const int co = i & 31, tap = i >> 5, ky = tap >> 2, kx = tap & 3;
    float s = 0.f;
#pragma unroll
    for (int cls = 0; cls < 16; ++cls)
      if (tap_inside(ky, cls >> 2) && tap_inside(kx, cls & 3)) s += R[cls * kCout + co];
    E[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct WgradArgs {
  const float *x, *dy;
  float* part;             // [tiles][kK * 32 + 16 * 32]: image-channel partial, then E[tap][co] of the tile
  int IH, IW, tiles_x, tiles_y;
  float shift;
};
constexpr int kPartFloats = kK * kCout + 16 * kCout;                    // 2048 per tile
constexpr size_t kWgradSmem = (size_t)(kInRows * kPitch + kTH * kTW * kCout + kK * kCout + 16 * 8 * kCout + 16 * kCout) * sizeof(float);

__global__ void __launch_bounds__(kThreads) conv_first_wgrad_kernel(const WgradArgs A) {
  EXP_PDL_ENTRY();
  extern __shared__ __align__(16) float wg_smem[];
  float* xs = wg_smem;                                                  // [18][200]
  float* dys = xs + kInRows * kPitch;                                   // [8 rows][32 px][32 co]
  float* red = dys + kTH * kTW * kCout;                                 // [48][32]
  float* bins = red + kK * kCout;                                       // [16][8][32]
  float* R = bins + 16 * 8 * kCout;                                     // [16][32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int OH = A.IH >> 1, OW = A.IW >> 1;
  const Tile t = decode_tile(blockIdx.x, A.tiles_x, A.tiles_y);
  load_input_tile(xs, A.x, t, A.IH, A.IW, A.shift);
  batched_fill<8>(kTH * kTW * (kCout / 4), kThreads,
                  [&](int i) {
                    const int q = i & 7, px = i >> 3, r = px / kTW, c = px - r * kTW;
                    const int oy = t.oy0 + r, ox = t.ox0 + c;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (oy < OH && ox < OW)
                      v = __ldg(reinterpret_cast<const float4*>(A.dy + ((size_t)(t.b * OH + oy) * OW + ox) * kCout) + q);
                    return v;
                  },
                  [&](int i, float4 v) { reinterpret_cast<float4*>(dys)[i] = v; });
  for (int i = tid; i < 16 * 8 * kCout; i += kThreads) bins[i] = 0.f;
  __syncthreads();

  // class sums: thread (slice = row, co) walks its row
  {
    const int co = lane, r = warp, oy = t.oy0 + r;
    if (oy < OH) {
      const int rc = border_class(oy, OH);
      for (int c = 0; c < kTW; ++c) {
        const int ox = t.ox0 + c;
        if (ox >= OW) break;
        bins[((rc * 4 + border_class(ox, OW)) * 8 + r) * kCout + co] += dys[(r * kTW + c) * kCout + co];
      }
    }
  }
  // image channels: warp = row of the tile, lane = (6 consecutive k = (ky, kx pair, c), 4 + 4 output channels)
  const int cg = lane & 3, kg = lane >> 2, ky = kg >> 1, kxp = kg & 1, r = warp;
  float acc[6][8];
#pragma unroll
  for (int j = 0; j < 6; ++j)
#pragma unroll
    for (int h = 0; h < 8; ++h) acc[j][h] = 0.f;
  const float2* xr = reinterpret_cast<const float2*>(xs + (2 * r + ky) * kPitch + 6 * kxp);
  const float4* d4 = reinterpret_cast<const float4*>(dys) + (size_t)r * kTW * 8;
#pragma unroll 4
  for (int px = 0; px < kTW; ++px) {
    const float2 a = xr[3 * px], b2 = xr[3 * px + 1], c2 = xr[3 * px + 2];
    const float xv[6] = {a.x, a.y, b2.x, b2.y, c2.x, c2.y};
    const float4 d0 = d4[px * 8 + cg], d1 = d4[px * 8 + 4 + cg];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      acc[j][0] = fmaf(xv[j], d0.x, acc[j][0]); acc[j][1] = fmaf(xv[j], d0.y, acc[j][1]);
      acc[j][2] = fmaf(xv[j], d0.z, acc[j][2]); acc[j][3] = fmaf(xv[j], d0.w, acc[j][3]);
      acc[j][4] = fmaf(xv[j], d1.x, acc[j][4]); acc[j][5] = fmaf(xv[j], d1.y, acc[j][5]);
      acc[j][6] = fmaf(xv[j], d1.z, acc[j][6]); acc[j][7] = fmaf(xv[j], d1.w, acc[j][7]);
    }
  }
  // rows are added in row order (fixed summation order)
  for (int w = 0; w < kTH; ++w) {
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int k = (ky * 4 + 2 * kxp) * kCx + j;                      // (kx, c) = (2 kxp + j / 3, j % 3)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4* p = reinterpret_cast<float4*>(red + k * kCout + h * 16 + cg * 4);
          float4 v = make_float4(acc[j][4 * h], acc[j][4 * h + 1], acc[j][4 * h + 2], acc[j][4 * h + 3]);
          if (w > 0) { const float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
          *p = v;
        }
      }
    }
  }
  __syncthreads();
  float* part = A.part + (size_t)blockIdx.x * kPartFloats;
  for (int i = tid; i < kK * kCout; i += kThreads) part[i] = red[i];
  reduce_bins_to_E(bins, R, red /* reused as E[16][32] after the copy */, tid);   // (the copy above only READS red)
  __syncthreads();
  for (int i = tid; i < 16 * kCout; i += kThreads) part[kK * kCout + i] = red[i];
}

// gW[tap][c][co] (=|+=) over the tiles, in tile order.  grid (16 taps, Cin), 256 threads = 8 lanes x 32 co
__global__ void __launch_bounds__(kThreads) conv_first_wgrad_finish_kernel(const float* __restrict__ part, const float* __restrict__ vec,
                                                                           int Cv, float shift, int tiles, int tiles_per_image,
                                                                           float* __restrict__ gW, int accumulate) {
  EXP_PDL_ENTRY();
  __shared__ float red[8][kCout];
  const int tap = blockIdx.x, c = blockIdx.y, co = threadIdx.x & 31, ln = threadIdx.x >> 5, Cin = kCx + Cv;
  float s0 = 0.f, s1 = 0.f;
  if (c < kCx) {
    const float* p = part + (size_t)(tap * kCx + c) * kCout + co;
    int i = ln;
    for (; i + 8 < tiles; i += 16) { s0 += p[(size_t)i * kPartFloats]; s1 += p[(size_t)(i + 8) * kPartFloats]; }
    for (; i < tiles; i += 8) s0 += p[(size_t)i * kPartFloats];
  } else {
    const float* p = part + kK * kCout + (size_t)tap * kCout + co;
    const float* v = vec + (c - kCx);
    int i = ln;
    for (; i + 8 < tiles; i += 16) {
      s0 = fmaf(__ldg(v + (size_t)(i / tiles_per_image) * Cv) - shift, p[(size_t)i * kPartFloats], s0);
      s1 = fmaf(__ldg(v + (size_t)((i + 8) / tiles_per_image) * Cv) - shift, p[(size_t)(i + 8) * kPartFloats], s1);
    }
    for (; i < tiles; i += 8) s0 = fmaf(__ldg(v + (size_t)(i / tiles_per_image) * Cv) - shift, p[(size_t)i * kPartFloats], s0);
  }
  red[ln][co] = s0 + s1;
  __syncthreads();
  if (ln == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][co];
    float* o = gW + ((size_t)tap * Cin + c) * kCout + co;
    *o = accumulate ? *o + s : s;
  }
}

// ------------------------------------------------------------------------------------------------ dgrad, constant channels
// gvec[b][c] = sum over the pixels of d/d(channel 3 + c) = sum_tap <W[tap][3 + c][:], E[b][tap][:]>      one CTA per image
__global__ void __launch_bounds__(kThreads) conv_first_dgrad_vec_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                                        int Cv, int OH, int OW, float* __restrict__ gvec) {
  EXP_PDL_ENTRY();
  __shared__ float bins[16 * 8 * kCout], R[16 * kCout], E[16 * kCout];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b = blockIdx.x, Cin = kCx + Cv;
  for (int i = tid; i < 16 * 8 * kCout; i += kThreads) bins[i] = 0.f;
  __syncthreads();
  const float* d = dy + (size_t)b * OH * OW * kCout + lane;
  for (int oy = warp; oy < OH; oy += 8) {                               // slice = warp, co = lane
    const int rc = border_class(oy, OH);
    const float* dr = d + (size_t)oy * OW * kCout;
    // interior columns: a plain register sum, 8 independent loads in flight (a load per iteration next to a
    // shared-memory update is one L2 round trip per pixel: 13 us for a 32 x 32 map)
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
    int ox = 1;
    for (; ox + 8 <= OW - 1; ox += 8) {
      const float v0 = __ldg(dr + (size_t)(ox + 0) * kCout), v1 = __ldg(dr + (size_t)(ox + 1) * kCout);
      const float v2 = __ldg(dr + (size_t)(ox + 2) * kCout), v3 = __ldg(dr + (size_t)(ox + 3) * kCout);
      const float v4 = __ldg(dr + (size_t)(ox + 4) * kCout), v5 = __ldg(dr + (size_t)(ox + 5) * kCout);
      const float v6 = __ldg(dr + (size_t)(ox + 6) * kCout), v7 = __ldg(dr + (size_t)(ox + 7) * kCout);
      m0 += v0; m1 += v1; m2 += v2; m3 += v3;
      m0 += v4; m1 += v5; m2 += v6; m3 += v7;
    }
    for (; ox < OW - 1; ++ox) m0 += __ldg(dr + (size_t)ox * kCout);
    if (OW > 2) bins[((rc * 4) * 8 + warp) * kCout + lane] += (m0 + m1) + (m2 + m3);
    bins[((rc * 4 + border_class(0, OW)) * 8 + warp) * kCout + lane] += __ldg(dr);
    if (OW > 1) bins[((rc * 4 + border_class(OW - 1, OW)) * 8 + warp) * kCout + lane] += __ldg(dr + (size_t)(OW - 1) * kCout);
  }
  __syncthreads();
  reduce_bins_to_E(bins, R, E, tid);
  __syncthreads();
  for (int c = warp; c < Cv; c += 8) {                                  // warp per constant channel, lane = co
    float s = 0.f;
    for (int tap = 0; tap < 16; ++tap) s = fmaf(__ldg(W + ((size_t)tap * Cin + kCx + c) * kCout + lane), E[tap * kCout + lane], s);
    s = warp_sum(s);
    if (lane == 0) gvec[(size_t)b * Cv + c] = s;
  }
}

}  // namespace first

// nn.cu: transposed convolution into the first `Cin` of `CinW` weight channels (dx has Cin channels)
cudaError_t conv_dgrad_small_launch(const float* dy, const float* W, const float* a_in, float* dx, int B, int IH, int IW, int Cin,
                                    int CinW, int Cout, cudaStream_t stream);
}  // namespace expo

using namespace expo;
using namespace expo::first;

extern "C" {

int exp_conv_first_supported(int Cx, int Cv, int Cout) { return Cx == kCx && Cv >= 0 && Cv <= kMaxCv && Cout == kCout; }

static int first_shape_ok(int B, int IH, int IW, int* tiles_x, int* tiles_y) {
  if (B <= 0 || IH < 2 || IW < 2 || (IH & 1) || (IW & 1)) return 0;
  *tiles_x = (IW / 2 + kTW - 1) / kTW;
  *tiles_y = (IH / 2 + kTH - 1) / kTH;
  return (long long)B * *tiles_x * *tiles_y < (1ll << 31) && (long long)B * IH * IW * kCout < (1ll << 40);
}

int exp_conv_first_fwd(const float* x, const float* vec, int Cv, float shift, const float* W, const float* bias,
                       const float* mask_ref, const float* post_mul, float* y, float* y2, int B, int IH, int IW, int mode,
                       void* stream) {
  EXP_CHECK_ARG(x && W && y && (Cv == 0 || vec) && Cv >= 0 && Cv <= kMaxCv, "null pointer / Cv out of [0, %d]", kMaxCv);
  EXP_CHECK_ARG(mode == 0 || (mode == 1 && mask_ref), "mode 1 needs mask_ref");
  EXP_CHECK_ARG(!y2 || post_mul, "y2 needs post_mul");
  FwdArgs A{};
  EXP_CHECK_ARG(first_shape_ok(B, IH, IW, &A.tiles_x, &A.tiles_y), "bad shape (even IH, IW >= 2)");
  if (!aligned16(y) || !aligned16(y2) || !aligned16(mask_ref) || !aligned16(post_mul))
    return set_error(EXP_ERR_ALIGNMENT, "y / y2 / mask_ref / post_mul must be 16-byte aligned");
  A.x = x; A.vec = vec; A.W = W; A.bias = bias; A.mask_ref = mask_ref; A.post_mul = post_mul; A.y = y; A.y2 = y2;
  A.IH = IH; A.IW = IW; A.Cv = Cv; A.mode = mode; A.shift = shift;
  launch_pdl(conv_first_fwd_kernel, dim3((unsigned)(B * A.tiles_x * A.tiles_y)), dim3(kThreads), 0, (cudaStream_t)stream, A);
  EXP_CHECK_LAUNCH("exp_conv_first_fwd");
  return EXP_OK;
}

size_t exp_conv_first_wgrad_workspace_bytes(int B, int IH, int IW) {
  int tx, ty;
  if (!first_shape_ok(B, IH, IW, &tx, &ty)) return 0;
  return (size_t)B * tx * ty * kPartFloats * sizeof(float);
}

int exp_conv_first_wgrad(const float* x, const float* vec, int Cv, float shift, const float* dy, float* gW, int B, int IH, int IW,
                         int accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  EXP_CHECK_ARG(x && dy && gW && workspace && (Cv == 0 || vec) && Cv >= 0 && Cv <= kMaxCv, "null pointer / Cv out of [0, %d]", kMaxCv);
  WgradArgs A{};
  EXP_CHECK_ARG(first_shape_ok(B, IH, IW, &A.tiles_x, &A.tiles_y), "bad shape (even IH, IW >= 2)");
  const size_t need = exp_conv_first_wgrad_workspace_bytes(B, IH, IW);
  if (workspace_bytes < need) return set_error(EXP_ERR_WORKSPACE, "workspace %zu B < required %zu B", workspace_bytes, need);
  if (!aligned16(dy) || !aligned16(workspace)) return set_error(EXP_ERR_ALIGNMENT, "dy / workspace must be 16-byte aligned");
  static bool opted = false;
  if (!opted) {
    const cudaError_t e = cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgradSmem);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_first_wgrad: %s", cudaGetErrorString(e));
    opted = true;
  }
  A.x = x; A.dy = dy; A.part = reinterpret_cast<float*>(workspace); A.IH = IH; A.IW = IW; A.shift = shift;
  const int per_image = A.tiles_x * A.tiles_y, tiles = B * per_image;
  launch_pdl(conv_first_wgrad_kernel, dim3((unsigned)tiles), dim3(kThreads), kWgradSmem, (cudaStream_t)stream, A);
  EXP_CHECK_LAUNCH("exp_conv_first_wgrad");
  launch_pdl(conv_first_wgrad_finish_kernel, dim3(16, kCx + Cv), dim3(kThreads), 0, (cudaStream_t)stream,
             reinterpret_cast<const float*>(workspace), vec, Cv, shift, tiles, per_image, gW, accumulate);
  EXP_CHECK_LAUNCH("exp_conv_first_wgrad[finish]");
  return EXP_OK;
}

int exp_conv_first_dgrad(const float* dy, const float* W, int Cv, float* dx_img, float* gvec, int B, int IH, int IW, void* stream) {
  EXP_CHECK_ARG(dy && W && (dx_img || gvec) && Cv >= 0 && Cv <= kMaxCv, "null pointer / Cv out of [0, %d]", kMaxCv);
  int tx, ty;
  EXP_CHECK_ARG(first_shape_ok(B, IH, IW, &tx, &ty), "bad shape (even IH, IW >= 2)");
  EXP_CHECK_ARG(!dx_img || ((IH & (IH - 1)) == 0 && (IW & (IW - 1)) == 0), "the image gradient needs power-of-two IH, IW");
  if (!aligned16(dy) || !aligned16(W)) return set_error(EXP_ERR_ALIGNMENT, "dy / W must be 16-byte aligned");
  if (dx_img) {
    const cudaError_t e = conv_dgrad_small_launch(dy, W, nullptr, dx_img, B, IH, IW, kCx, kCx + Cv, kCout, (cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(EXP_ERR_CUDA, "exp_conv_first_dgrad: %s", cudaGetErrorString(e));
    EXP_CHECK_LAUNCH("exp_conv_first_dgrad[image]");
  }
  if (gvec && Cv > 0) {
    launch_pdl(conv_first_dgrad_vec_kernel, dim3(B), dim3(kThreads), 0, (cudaStream_t)stream, dy, W, Cv, IH / 2, IW / 2, gvec);
    EXP_CHECK_LAUNCH("exp_conv_first_dgrad[vec]");
  }
  return EXP_OK;
}

}  // extern "C"
