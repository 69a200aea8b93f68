// Warp-specialised tcgen05 GEMM engine (second generation of tc_engine.cuh).
//
// The first engine let all threads gather a K step, __syncthreads, and then one thread issued the
// MMAs: the tensor pipe sat at 3-6 % because every K step paid a full global-load round trip
// plus a CTA barrier.  Here the roles are split and decoupled by mbarriers:
//
//   warps 0-5 (192 threads)  PRODUCERS: gather A (128 x 32) and B (BN x 32) with computed
//                            addresses, split into TF32 hi/lo, store K-major SWIZZLE_128B tiles
//                            into an NS-stage shared-memory ring, fence.proxy.async, arrive on
//                            full[s].  The gathers of step k+1 are issued (register prefetch)
//                            before step k is stored, and producers run up to NS steps ahead.
//   warp 6, one lane         MMA ISSUER: waits full[s], issues 4 x 3 tcgen05.mma.kind::tf32
//                            (hi*hi + hi*lo + lo*hi), tcgen05.commit -> empty[s]; after the last
//                            K step commits to the accumulator barrier.
//   all 8 warps              EPILOGUE: wait for the accumulator, tcgen05.ld 32x32b, functor store.
//
// No __syncthreads in the main loop.  Same problem-functor interface as tc_engine.cuh.
#pragma once
#include "tc_engine.cuh"

namespace expo {
namespace tc {

constexpr int kProducers = 192;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

template <class P, int BN, bool A_ROWFAST = false>
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
  constexpr uint32_t idesc = idesc_tf32(kBM, BN);

  if (tid < kProducers) {
    // ================================ producers ================================
    float4 reg[C::kSlots];
    auto slot_coords = [&](int j, bool& is_a, int& row, int& c) {
      const int i = tid + kProducers * j;
      is_a = i < 1024;
      if (is_a) {
        if (A_ROWFAST) { row = i & (kBM - 1); c = i >> 7; }
        else { row = i >> 3; c = i & 7; }
      } else {
        const int k = i - 1024;
        row = k % BN; c = k / BN;
      }
    };
    auto gather = [&](int ki) {
      const typename P::KS ks = p.kstate(ki);
#pragma unroll
      for (int j = 0; j < C::kSlots; ++j) {
        bool is_a; int row, c;
        slot_coords(j, is_a, row, c);
        if (tid + kProducers * j < C::kTotalV) {
          if (is_a) reg[j] = p.load_a4(p.row_a(m0 + row), ks, c);
          else reg[j] = p.load_b4(ks, n0 + row, c);
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
        bool is_a; int row, c;
        slot_coords(j, is_a, row, c);
        if (tid + kProducers * j < C::kTotalV) {
          if (is_a) split_store(a_hi, a_lo, sw128_off(row, c), reg[j]);
          else split_store(b_hi, b_lo, sw128_off(row, c), reg[j]);
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
        unsigned char* a_hi = base + (size_t)s * C::kStageBytes;
        const uint32_t sa_hi = smem_u32(a_hi), sa_lo = sa_hi + kTileABytes, sb_hi = sa_lo + kTileABytes,
                       sb_lo = sb_hi + C::kTileB;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t dah = smem_desc_sw128(sa_hi + k * 32), dal = smem_desc_sw128(sa_lo + k * 32);
          const uint64_t dbh = smem_desc_sw128(sb_hi + k * 32), dbl = smem_desc_sw128(sb_lo + k * 32);
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

template <class P, int BN, bool A_ROWFAST = false>
inline cudaError_t launch_tc_gemm_ws(const P& p, int M, int N, int Z, cudaStream_t st) {
  constexpr size_t smem = WsCfg<BN>::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_ws_kernel<P, BN, A_ROWFAST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  tc_gemm_ws_kernel<P, BN, A_ROWFAST><<<grid, kThreads, smem, st>>>(p);
  return cudaSuccess;
}

}  // namespace tc
}  // namespace expo
