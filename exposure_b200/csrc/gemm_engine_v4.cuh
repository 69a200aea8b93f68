// Vectorised variant of gemm_engine.cuh: 64 x BN x 16 CTA tile, every operand gather is ONE
// 16-byte load per thread per K step (4x fewer load instructions and index computations per FMA
// than the scalar engine), for the operands whose contiguous dimension is a multiple of 4:
// conv fprop with Cin % 16 == 0, conv dgrad (Cout % 16 == 0), conv wgrad with Cin % 4 == 0.
//
// Operand modes
//   A_KVEC: float4 = A[m, k0+4c .. +3]   (K contiguous)  -> 4 scalar STS into As[k][m]
//   A_MVEC: float4 = A[m .. m+3, k0+kk]  (M contiguous)  -> one STS.128 into As[kk][m..]
//   B_NVEC: float4 = B[k0+kk, n .. n+3]  (N contiguous)  -> one STS.128 into Bs[kk][n..]
//   B_KVEC: float4 = B[k0+4c .. +3, n]   (K contiguous)  -> 4 scalar STS into Bs[k][n]
// Functor P (in addition to init / k_iters16 / kstate16 / store):
//   A_KVEC: RowA row_a(m);  float4 load_a_k4(RowA, KS, c)
//   A_MVEC:                 float4 load_a_m4(KS, kk, m)      (m multiple of 4)
//   B_NVEC:                 float4 load_b_n4(KS, kk, n)      (n multiple of 4)
//   B_KVEC:                 float4 load_b_k4(KS, c, n)
#pragma once
#include "gemm_engine.cuh"

namespace expo {

constexpr int kBK4 = 16;

template <class P, int BN, bool A_MVEC, bool B_KVEC>
__global__ void __launch_bounds__(kGemmThreads) gemm_v4_kernel(const P p_in) {
  EXP_PDL_ENTRY();
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][kBK4][kBM + 4];
  __shared__ __align__(16) float Bs[2][kBK4][BN + 4];

  P p = p_in;
  p.init(blockIdx.z);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int KI = p.k_iters16();

  // A: 64 x 16 = 256 float4 (one per thread); B: 16 x BN = 4*BN float4 (threads < 4*BN)
  const int a_m = A_MVEC ? (tid & 15) * 4 : (tid >> 2);
  const int a_k = A_MVEC ? (tid >> 4) : (tid & 3);          // kk (MVEC) or 4-chunk index c (KVEC)
  constexpr bool kBActiveAll = (4 * BN >= kGemmThreads);
  const bool b_active = kBActiveAll || tid < 4 * BN;
  const int b_n = B_KVEC ? (tid >> 2) : (tid % (BN / 4)) * 4;
  const int b_k = B_KVEC ? (tid & 3) : (tid / (BN / 4));
  typename P::RowA ra;
  if constexpr (!A_MVEC) ra = p.row_a(m0 + a_m);

  float4 areg, breg;
  auto gather = [&](int ki) {
    const typename P::KS ks = p.kstate16(ki);
    if constexpr (A_MVEC) areg = p.load_a_m4(ks, a_k, m0 + a_m);
    else areg = p.load_a_k4(ra, ks, a_k);
    if (b_active) {
      if constexpr (B_KVEC) breg = p.load_b_k4(ks, b_k, n0 + b_n);
      else breg = p.load_b_n4(ks, b_k, n0 + b_n);
    }
  };
  auto stash = [&](int buf) {
    if constexpr (A_MVEC) {
      *reinterpret_cast<float4*>(&As[buf][a_k][a_m]) = areg;
    } else {
      As[buf][4 * a_k + 0][a_m] = areg.x; As[buf][4 * a_k + 1][a_m] = areg.y;
      As[buf][4 * a_k + 2][a_m] = areg.z; As[buf][4 * a_k + 3][a_m] = areg.w;
    }
    if (b_active) {
      if constexpr (B_KVEC) {
        Bs[buf][4 * b_k + 0][b_n] = breg.x; Bs[buf][4 * b_k + 1][b_n] = breg.y;
        Bs[buf][4 * b_k + 2][b_n] = breg.z; Bs[buf][4 * b_k + 3][b_n] = breg.w;
      } else {
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = breg;
      }
    }
  };

  const int tx = tid % 16, ty = tid / 16;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  if (KI > 0) {
    gather(0);
    stash(0);
  }
  __syncthreads();
  for (int ki = 0; ki < KI; ++ki) {
    const int buf = ki & 1;
    if (ki + 1 < KI) gather(ki + 1);
#pragma unroll
    for (int kk = 0; kk < kBK4; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float bv[TN];
      if constexpr (TN == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        bv[0] = t.x; bv[1] = t.y; bv[2] = t.z; bv[3] = t.w;
      } else {
        const float2 t = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * 2]);
        bv[0] = t.x; bv[1] = t.y;
      }
      const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a4[i], bv[j], acc[i][j]);
    }
    if (ki + 1 < KI) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) p.store(m0 + ty * 4 + i, n0 + tx * TN + j, acc[i][j]);
}

template <class P, int BN, bool A_MVEC, bool B_KVEC>
inline void launch_gemm_v4(const P& p, int M, int N, int Z, cudaStream_t st) {
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  launch_pdl(gemm_v4_kernel<P, BN, A_MVEC, B_KVEC>, dim3(grid), dim3(kGemmThreads), 0, st, p);
}

}  // namespace expo
