// tcgen05 (5th-gen tensor core) GEMM engine with functor-defined operand gathers.
//
// Same role as gemm_engine.cuh, but the 128 x BN x 32 CTA tile is multiplied by
// tcgen05.mma.kind::tf32 with the accumulator in TMEM:
//
//   all threads : gather A (128 rows x 32 k) and B (BN rows x 32 k) from global memory with
//                 computed addresses (implicit GEMM), split every value v into
//                 hi = v with the low 13 mantissa bits cleared (exact in TF32) and lo = v - hi,
//                 and store both, K-major, in the canonical 128-byte-swizzled layout
//   one thread  : per 32-wide K step, 4 x 3 MMAs (K = 8 each):  hi*hi + hi*lo + lo*hi
//                 ("3xTF32": the dropped lo*lo term and the TF32 rounding of lo are ~2^-22
//                 relative, i.e. fp32-level accuracy -- the 1e-3 logit bound of the north star
//                 is kept with three orders of magnitude to spare), tcgen05.commit -> mbarrier
//   all threads : while the tensor core works on stage s they gather stage s^1
//   epilogue    : tcgen05.ld 32x32b from TMEM, functor store (bias/lrelu/mask or split-K partial)
//
// Operand descriptors follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor version 1,
// SWIZZLE_128B, K-major: LBO = 1, SBO = 1024 B; InstrDescriptor: c_format F32, a/b TF32).
#pragma once
#include "common.cuh"
#include "async_ptx.cuh"     // mbarrier helpers

namespace expo {
namespace tc {

constexpr int kThreads = 256;
constexpr int kBM = 128;
constexpr int kBK = 32;                  // floats per K step = one 128-byte swizzle row
constexpr int kTileABytes = kBM * 128;   // 16 KiB

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  // start address [0,14) (>>4), LBO [16,30) = 1, SBO [32,46) = 1024>>4, version [46,48) = 1,
  // layout_type [61,64) = 2 (SWIZZLE_128B)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  // c_format F32 (1) @4, a_format TF32 (2) @7, b_format TF32 (2) @10, K-major both, n_dim @17, m_dim @24
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // one full warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of 16-byte chunk c (0..7) of row r inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}
// round-to-nearest fp32 -> tf32 (result has the low 13 mantissa bits cleared, so whatever
// conversion the tensor core applies to it is exact)
__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_store(unsigned char* hi_tile, unsigned char* lo_tile, uint32_t off, float4 v) {
  // v = hi + lo exactly; hi and lo are both TF32-representable up to an UNBIASED 2^-22 |v|
  // rounding of lo (truncation instead would bias every product the same way: measured 1e-5)
  float4 h, l;
  h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
  h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
  h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
  h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// Problem functor P:
//   void   init(int z);  int k_iters();                       (K steps of 32)
//   RowA   row_a(int m);  KS kstate(int ki);
//   float4 load_a4(RowA, KS, int c)        A[m, ki*32 + 4c .. +3]
//   float4 load_b4(KS, int n, int c)       B[n, ki*32 + 4c .. +3]     (B is N x K, "K-major")
//   void   store16(int m, int n0, const float (&v)[16])   C[m, n0..n0+15]
// A_ROWFAST: consecutive lanes gather consecutive A rows (operands whose contiguous dimension is
// M, e.g. wgrad) instead of the 8 chunks of one row.
template <class P, int BN, bool A_ROWFAST = false>
__global__ void __launch_bounds__(kThreads, (BN <= 64) ? 2 : 1) tc_gemm_kernel(const P p_in) {   // 2 CTAs/SM fit for BN <= 64
  static_assert(BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN must be a power of two in [32,256]");
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  constexpr int kTileBBytes = BN * 128;
  constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t tmem_base_s;

  P p = p_in;
  p.init(blockIdx.z);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  // dynamic smem is only guaranteed 16-byte aligned: round up to 1024 (the host adds the slack)
  unsigned char* base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);

  if (warp == 0) tmem_alloc(&tmem_base_s, BN);
  if (tid == 32) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_acc = tmem_base_s;

  // gather mappings: A: 8 lanes cover one 128-byte row; B: lanes run along n
  typename P::RowA ra[4];
  int a_row[4], a_c[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int idx = tid + kThreads * j;
    if (A_ROWFAST) { a_row[j] = idx & (kBM - 1); a_c[j] = idx >> 7; }
    else { a_row[j] = idx >> 3; a_c[j] = idx & 7; }
    ra[j] = p.row_a(m0 + a_row[j]);
  }
  constexpr int NBV = (BN * 8) / kThreads;   // float4 of B per thread per stage
  const int KI = p.k_iters();
  constexpr uint32_t idesc = idesc_tf32(kBM, BN);

  // register prefetch: the global gathers of K step ki+1 are in flight while step ki is split,
  // stored, fenced and handed to the tensor core (the gather latency was fully exposed before)
  float4 areg[4], breg[NBV];
  auto gather = [&](int ki) {
    const typename P::KS ks = p.kstate(ki);
#pragma unroll
    for (int j = 0; j < 4; ++j) areg[j] = p.load_a4(ra[j], ks, a_c[j]);
#pragma unroll
    for (int j = 0; j < NBV; ++j) {
      const int idx = tid + kThreads * j;
      breg[j] = p.load_b4(ks, n0 + idx % BN, idx / BN);
    }
  };
  if (KI > 0) gather(0);

  for (int ki = 0; ki < KI; ++ki) {
    const int s = ki & 1;
    unsigned char* a_hi = base + (size_t)s * kStageBytes;
    unsigned char* a_lo = a_hi + kTileABytes;
    unsigned char* b_hi = a_lo + kTileABytes;
    unsigned char* b_lo = b_hi + kTileBBytes;
    if (ki >= 2) mbar_wait(&mbar[s], (uint32_t)(((ki >> 1) - 1) & 1));   // MMAs of step ki-2 done with this stage
#pragma unroll
    for (int j = 0; j < 4; ++j) split_store(a_hi, a_lo, sw128_off(a_row[j], a_c[j]), areg[j]);
#pragma unroll
    for (int j = 0; j < NBV; ++j) {
      const int idx = tid + kThreads * j;
      split_store(b_hi, b_lo, sw128_off(idx % BN, idx / BN), breg[j]);
    }
    if (ki + 1 < KI) gather(ki + 1);
    fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
#pragma unroll
      for (int k = 0; k < 4; ++k) {        // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle row
        const uint64_t dah = smem_desc_sw128(sa_hi + k * 32), dal = smem_desc_sw128(sa_lo + k * 32);
        const uint64_t dbh = smem_desc_sw128(sb_hi + k * 32), dbl = smem_desc_sw128(sb_lo + k * 32);
        mma_tf32(tmem_acc, dah, dbh, idesc, (ki > 0 || k > 0) ? 1u : 0u);
        mma_tf32(tmem_acc, dah, dbl, idesc, 1u);
        mma_tf32(tmem_acc, dal, dbh, idesc, 1u);
      }
      mma_commit(&mbar[s]);
    }
  }
  if (KI > 0) mbar_wait(&mbar[(KI - 1) & 1], (uint32_t)(((KI - 1) >> 1) & 1));
  fence_after_sync();

  // epilogue: warp w reads TMEM lanes 32*(w%4).. (its row quarter), column half w/4
  const int q = warp & 3, half = warp >> 2;
  const int m = m0 + q * 32 + lane;
#pragma unroll 1
  for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
    float v[16];
    if (KI > 0) tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    p.store16(m, n0 + c0, v);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, BN);
}

template <class P, int BN, bool A_ROWFAST = false>
inline cudaError_t launch_tc_gemm(const P& p, int M, int N, int Z, cudaStream_t st) {
  constexpr size_t smem = 2 * (2 * (size_t)kTileABytes + 2 * (size_t)BN * 128) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<P, BN, A_ROWFAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  tc_gemm_kernel<P, BN, A_ROWFAST><<<grid, kThreads, smem, st>>>(p);
  return cudaSuccess;
}

}  // namespace tc
}  // namespace expo
