// Tiled fp32 GEMM engine with functor-defined operand gathers (implicit GEMM).
//
// C[M,N] = sum_k A[m,k] * B[k,n] where A and B are *computed addresses*: the 4x4/stride-2
// convolutions of the policy / critic / value CNNs (agent.py:11-37, critics.py:6-38), their
// dgrad / wgrad and the fully connected heads all run through this one engine by supplying a
// problem functor P:
//
//   void  init(int z)                     sub-problem selected by blockIdx.z
//   int   M(), N(), k_iters()             k_iters counts BK(=8)-wide K steps
//   RowA  row_a(int m)                    per-row state hoisted out of the K loop
//   KS    kstate(int ki)                  per-K-step state (CTA uniform)
//   float load_a(RowA, KS, int kk)        A[m, ki*8+kk]   (0 outside the problem)
//   float load_b(KS, int kk, int n)       B[ki*8+kk, n]
//   void  store(int m, int n, float acc)  epilogue
//
// 64 x BN x 8 CTA tile, 256 threads, 4 x (BN/16) register micro-tile, double-buffered shared
// memory with register prefetch (one __syncthreads per stage of KU K-steps).  KU = 1 keeps one
// gather in flight, enough when many CTAs share an SM; the fully connected heads at batch 64 are a
// handful of CTAs with 8-16 K-steps each, every step a full global-memory round trip (~0.7 us):
// they run with KU = 4, i.e. 4 K-steps gathered per round trip (same summation order, same bits).  A_MFAST / B_KFAST pick which
// index runs across the lanes of a warp when gathering, so that the global reads follow the
// contiguous dimension of the operand.
//
// This is the exact-fp32 (CUDA core) path; the tcgen05 path for the same entry points is
// tracked in DESIGN.md section 7.
#pragma once
#include "common.cuh"

namespace expo {

constexpr int kGemmThreads = 256;
constexpr int kBM = 64;
constexpr int kBK = 8;

template <class P, int BN, bool A_MFAST, bool B_KFAST, int KU = 1>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const P p_in) {
  EXP_PDL_ENTRY();
  constexpr int TN = BN / 16;
  constexpr int NB = (kBK * BN) / kGemmThreads;   // B elements per thread per K step (2 or 1)
  __shared__ __align__(16) float As[2][KU * kBK][kBM + 4];
  __shared__ __align__(16) float Bs[2][KU * kBK][BN + 4];

  P p = p_in;
  p.init(blockIdx.z);
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  const int KI = p.k_iters();
  const int KG = (KI + KU - 1) / KU;              // stages of KU K-steps

  // gather mappings
  int a_m[2], a_k[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (A_MFAST) { a_m[j] = tid % kBM; a_k[j] = tid / kBM + 4 * j; }
    else         { a_k[j] = tid % kBK; a_m[j] = tid / kBK + 32 * j; }
  }
  int b_n[NB], b_k[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (B_KFAST) { b_k[j] = tid % kBK; b_n[j] = tid / kBK + 32 * j; }
    else         { b_n[j] = tid % BN; b_k[j] = tid / BN + (kGemmThreads / BN) * j; }
  }
  typename P::RowA ra[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) ra[j] = p.row_a(m0 + a_m[j]);

  float areg[KU][2], breg[KU][NB];
  auto gather = [&](int kg) {
#pragma unroll
    for (int u = 0; u < KU; ++u) {
      const int ki = kg * KU + u;
      if (KU == 1 || ki < KI) {
        const typename P::KS ks = p.kstate(ki);
#pragma unroll
        for (int j = 0; j < 2; ++j) areg[u][j] = p.load_a(ra[j], ks, a_k[j]);
#pragma unroll
        for (int j = 0; j < NB; ++j) breg[u][j] = p.load_b(ks, b_k[j], n0 + b_n[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) areg[u][j] = 0.f;
#pragma unroll
        for (int j = 0; j < NB; ++j) breg[u][j] = 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int u = 0; u < KU; ++u) {
#pragma unroll
      for (int j = 0; j < 2; ++j) As[buf][u * kBK + a_k[j]][a_m[j]] = areg[u][j];
#pragma unroll
      for (int j = 0; j < NB; ++j) Bs[buf][u * kBK + b_k[j]][b_n[j]] = breg[u][j];
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
  for (int kg = 0; kg < KG; ++kg) {
    const int buf = kg & 1;
    if (kg + 1 < KG) gather(kg + 1);
#pragma unroll
    for (int kk = 0; kk < KU * kBK; ++kk) {
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
    if (kg + 1 < KG) stash(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) p.store(m0 + ty * 4 + i, n0 + tx * TN + j, acc[i][j]);
}

template <class P, int BN, bool A_MFAST, bool B_KFAST, int KU = 1>
inline void launch_gemm(const P& p, int M, int N, int Z, cudaStream_t st) {
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  launch_pdl(gemm_kernel<P, BN, A_MFAST, B_KFAST, KU>, dim3(grid), dim3(kGemmThreads), 0, st, p);
}

__device__ __forceinline__ float lrelu_f(float v) { return 0.6f * v + 0.4f * fabsf(v); }   // util.py:225-229
// derivative of lrelu expressed through its (sign preserving) output: 1, 0.2, or 0.6 at 0
__device__ __forceinline__ float dlrelu_from_out(float a) { return a > 0.f ? 1.0f : (a < 0.f ? 0.2f : 0.6f); }

}  // namespace expo
