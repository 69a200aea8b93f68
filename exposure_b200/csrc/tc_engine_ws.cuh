// Warp-specialised tcgen05 GEMM engine (second generation of tc_engine.cuh).
//
// The first engine let all threads gather a K step, __syncthreads, and then one thread issued the
// MMAs: the tensor pipe sat at 3-6 % because every K step paid a full global-load round trip
// plus a CTA barrier, and operands whose contiguous dimension is M or N were gathered with four
// scalar loads per 16 bytes.  Here
//
//   * roles are split and decoupled by mbarriers (no __syncthreads in the main loop):
//       warps 0-5 (192 threads)  PRODUCERS: gather A (128 x 32) and B (BN x 32), split into TF32
//                                hi/lo, store SWIZZLE_128B tiles into an NS-stage ring,
//                                fence.proxy.async, arrive on full[s]; the gathers of step k+1
//                                are issued (register prefetch) before step k is stored
//       warp 6, one lane         MMA ISSUER: waits full[s], 4 x 3 tcgen05.mma.kind::tf32
//                                (hi*hi + hi*lo + lo*hi), tcgen05.commit -> empty[s]
//       all 8 warps              EPILOGUE: tcgen05.ld 32x32b from TMEM, functor store
//   * each operand is staged in the layout its MEMORY layout makes cheap: K-major tiles for
//     operands contiguous along K (im2col rows, dgrad weights), MN-major tiles (a_major / b_major
//     = 1 in the instruction descriptor) for operands contiguous along M or N (weights [K][N],
//     wgrad activations): every gather is one 16-byte load.
//
// Problem functor P: init / k_iters / kstate / row_a / store16 as in tc_engine.cuh, plus
//   A K-major : float4 load_a4(RowA, KS, c)        A[m, k0+4c..+3]
//   A MN-major: float4 load_a_mn4(KS, kk, m)       A[m..m+3, k0+kk]     (m % 4 == 0)
//   B K-major : float4 load_b4(KS, n, c)           B[n, k0+4c..+3]
//   B MN-major: float4 load_b_mn4(KS, kk, n)       B[n..n+3, k0+kk]     (n % 4 == 0)
#pragma once
#include "tc_engine.cuh"

namespace expo {
namespace tc {

constexpr int kProducers = 192;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// MN-major SWIZZLE_128B tile of R (M or N extent) x 32 (K): canonical layout
// ((4,8,R/32),(8,4)) : ((1,4,LBO),(32,SBO)) in elements, LBO = 1024 B between 32-wide MN blocks,
// SBO = (R/32) * 1024 B between groups of 8 K rows; 16-byte chunk index XOR (k % 8).
__device__ __forceinline__ uint32_t mn128_off(int R, int mn, int kk) {
  return (uint32_t)((kk >> 3) * (R >> 5) * 1024 + (mn >> 5) * 1024 + (kk & 7) * 128 + ((((mn >> 2) & 7) ^ (kk & 7)) << 4));
}
__device__ __forceinline__ uint64_t smem_desc_sw128_mn(uint32_t saddr, int R) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (64ull << 16) | ((uint64_t)((R >> 5) * 64) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32_major(int M, int N, bool a_mn, bool b_mn) {
  return idesc_tf32(M, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}

template <int BN>
struct WsCfg {
  static constexpr int kStages = BN >= 128 ? 3 : (BN == 64 ? 4 : 2);
  static constexpr int kTileB = BN * 128;
  static constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileB;
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024;
  static constexpr int kTotalV = 1024 + BN * 8;                         // float4 per stage (A then B)
  static constexpr int kSlots = (kTotalV + kProducers - 1) / kProducers;
};

template <class P, int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, (BN <= 32) ? 2 : 1) tc_gemm_ws_kernel(const P p_in) {
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN must be 32, 64 or 128");
  using C = WsCfg<BN>;
  constexpr int NS = C::kStages;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ __align__(8) uint64_t full[NS];
  __shared__ __align__(8) uint64_t empty[NS];
  __shared__ __align__(8) uint64_t accum;
  __shared__ uint32_t tmem_base_s;

  P p = p_in;
  p.init(blockIdx.z);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  unsigned char* base = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);

  if (warp == 0) tmem_alloc(&tmem_base_s, BN);
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], kProducers);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&accum, 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_acc = tmem_base_s;
  const int KI = p.k_iters();
  constexpr uint32_t idesc = idesc_tf32_major(kBM, BN, A_MN, B_MN);

  if (tid < kProducers) {
    // ================================ producers ================================
    float4 reg[C::kSlots];
    // slot j <-> flat float4 index i = tid + 192 j: A first (1024), then B (BN * 8).
    // K-major: (row = MN index, c = 16-byte K chunk); MN-major: (mn = 4 * (i % (R/4)), kk = i / (R/4))
    auto gather = [&](int ki) {
      const typename P::KS ks = p.kstate(ki);
#pragma unroll
      for (int j = 0; j < C::kSlots; ++j) {
        const int i = tid + kProducers * j;
        if (i < 1024) {
          if constexpr (A_MN) reg[j] = p.load_a_mn4(ks, i >> 5, m0 + ((i & 31) << 2));
          else reg[j] = p.load_a4(p.row_a(m0 + (i >> 3)), ks, i & 7);
        } else if (i < C::kTotalV) {
          const int k = i - 1024;
          if constexpr (B_MN) reg[j] = p.load_b_mn4(ks, k / (BN / 4), n0 + ((k % (BN / 4)) << 2));
          else reg[j] = p.load_b4(ks, n0 + k % BN, k / BN);
        }
      }
    };
    if (KI > 0) gather(0);
    for (int ki = 0; ki < KI; ++ki) {
      const int s = ki % NS, use = ki / NS;
      unsigned char* a_hi = base + (size_t)s * C::kStageBytes;
      unsigned char* a_lo = a_hi + kTileABytes;
      unsigned char* b_hi = a_lo + kTileABytes;
      unsigned char* b_lo = b_hi + C::kTileB;
      if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));     // MMAs of the previous use are done
#pragma unroll
      for (int j = 0; j < C::kSlots; ++j) {
        const int i = tid + kProducers * j;
        if (i < 1024) {
          const uint32_t off = A_MN ? mn128_off(kBM, (i & 31) << 2, i >> 5) : sw128_off(i >> 3, i & 7);
          split_store(a_hi, a_lo, off, reg[j]);
        } else if (i < C::kTotalV) {
          const int k = i - 1024;
          const uint32_t off = B_MN ? mn128_off(BN, (k % (BN / 4)) << 2, k / (BN / 4)) : sw128_off(k % BN, k / BN);
          split_store(b_hi, b_lo, off, reg[j]);
        }
      }
      if (ki + 1 < KI) gather(ki + 1);                                  // in flight while the next wait runs
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }
  } else if (warp == 6) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      for (int ki = 0; ki < KI; ++ki) {
        const int s = ki % NS, use = ki / NS;
        mbar_wait(&full[s], (uint32_t)(use & 1));
        fence_after_sync();
        const uint32_t sa_hi = smem_u32(base + (size_t)s * C::kStageBytes), sa_lo = sa_hi + kTileABytes,
                       sb_hi = sa_lo + kTileABytes, sb_lo = sb_hi + C::kTileB;
#pragma unroll
        for (int k = 0; k < 4; ++k) {           // 4 x (K = 8) per stage
          const uint32_t ao = A_MN ? k * (kBM / 32) * 1024 : k * 32;
          const uint32_t bo = B_MN ? k * (BN / 32) * 1024 : k * 32;
          const uint64_t dah = A_MN ? smem_desc_sw128_mn(sa_hi + ao, kBM) : smem_desc_sw128(sa_hi + ao);
          const uint64_t dal = A_MN ? smem_desc_sw128_mn(sa_lo + ao, kBM) : smem_desc_sw128(sa_lo + ao);
          const uint64_t dbh = B_MN ? smem_desc_sw128_mn(sb_hi + bo, BN) : smem_desc_sw128(sb_hi + bo);
          const uint64_t dbl = B_MN ? smem_desc_sw128_mn(sb_lo + bo, BN) : smem_desc_sw128(sb_lo + bo);
          mma_tf32(tmem_acc, dah, dbh, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          mma_tf32(tmem_acc, dah, dbl, idesc, 1u);
          mma_tf32(tmem_acc, dal, dbh, idesc, 1u);
        }
        mma_commit(&empty[s]);                    // frees the stage when these MMAs have read it
      }
      mma_commit(&accum);                         // accumulator complete
    }
    __syncwarp();
  }

  // ================================ epilogue (all warps) ================================
  if (KI > 0) mbar_wait(&accum, 0u);
  fence_after_sync();
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

template <class P, int BN, bool A_MN, bool B_MN>
inline cudaError_t launch_tc_gemm_ws(const P& p, int M, int N, int Z, cudaStream_t st) {
  constexpr size_t smem = WsCfg<BN>::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_ws_kernel<P, BN, A_MN, B_MN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  tc_gemm_ws_kernel<P, BN, A_MN, B_MN><<<grid, kThreads, smem, st>>>(p);
  return cudaSuccess;
}

}  // namespace tc
}  // namespace expo
