// Persistent variant of the TMA-fed tcgen05 engine (tc_engine_tma.cuh) for grids of MANY SHORT tiles
// (layer 1: 512-1536 tiles of 8 K steps; layer-2 dgrad).  A one-tile CTA spends ~10 us outside its K
// loop (launch, TMEM allocation, barrier setup, pipeline fill, epilogue, teardown) -- more than the
// loop itself when K is 256.  Here one CTA per SM walks tiles t = blockIdx.x, += gridDim.x and
//
//   * the stage ring never drains: the producer runs ahead into the next tile's K steps while the
//     current tile is still being multiplied (one global K-step counter across tiles);
//   * the accumulator is double-buffered in TMEM (2 x BN columns): the MMA issuer starts tile j+1 in
//     buffer (j+1)&1 as soon as its operands are converted, while four DEDICATED epilogue warps drain
//     buffer j&1 (tcgen05.ld -> functor store) -- acc_full[b] / acc_empty[b] mbarriers;
//   * TMEM allocation, barrier initialisation and descriptor prefetch happen once per SM.
//
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 converters (3xTF32 lo
// tiles), warps 6-9 epilogue (warp w reads TMEM lanes 32*(w%4)..).  Same problem functor as
// tc_engine_tma.cuh; cluster split-K is not combined with it (splits == 1).
#pragma once
#include "tc_engine_tma.cuh"

namespace expo {
namespace tma {

constexpr int kPersistThreads = 320;


template <int BN>
struct PersistCfg {
  static constexpr int kStages = BN >= 128 ? 3 : (BN == 64 ? 4 : 5);
  static constexpr int kTileB = BN * 128;
  static constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024;
  static constexpr int kVecPerThread = (kTileA + kTileB) / 16 / kConverters;
};

template <class P, int BN>
__global__ void __launch_bounds__(kPersistThreads, 1) tma_gemm_persistent_kernel(const __grid_constant__ P p, int MX, int NY, int tiles) {
  using C = PersistCfg<BN>;
  constexpr int NS = C::kStages;
  extern __shared__ __align__(1024) unsigned char tma_smem_p[];
  __shared__ __align__(8) uint64_t raw_full[NS];
  __shared__ __align__(8) uint64_t conv_full[NS];
  __shared__ __align__(8) uint64_t empty[NS];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  pdl_trigger();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* base = tma_smem_p + ((1024u - (smem_u32(tma_smem_p) & 1023u)) & 1023u);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 2 * BN);
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], kConverters);
      mbar_init(&empty[s], 1);
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 128);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.ta);
    tma_prefetch_desc(&p.tb);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();                                                 // the predecessor's writes are visible from here
  const uint32_t tmem_acc = tmem_base_s;

  // tile t -> (mx, ny, z): M fastest so that neighbouring CTAs share the B (weight) tile in L2
  auto decode = [&](int t, int& m0, int& n0, int& z) {
    const int mx = t % MX, r = t / MX;
    m0 = mx * kBM; n0 = (r % NY) * BN; z = r / NY;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int it = 0;                                              // global K-step counter (ring position)
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        int m0, n0, z;
        decode(t, m0, n0, z);
        const int KI = p.k_iters(z);
        for (int ki = 0; ki < KI; ++ki, ++it) {
          const int s = it % NS, use = it / NS;
          if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
          EXPO_TRACE(it, 0);
          unsigned char* a_raw = base + (size_t)s * C::kStageBytes;
          unsigned char* b_raw = a_raw + 2 * kTileA;
          mbar_expect_tx(&raw_full[s], (uint32_t)(kTileA + C::kTileB));
          p.template load<BN>(ki, z, m0, n0, a_raw, b_raw, &raw_full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32_major(kBM, BN, P::kAMn, P::kBMn);
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++j) {
        int m0, n0, z;
        decode(t, m0, n0, z);
        const int KI = p.k_iters(z);
        const int b = j & 1, useb = j >> 1;
        if (useb > 0) mbar_wait(&acc_empty[b], (uint32_t)((useb - 1) & 1));   // epilogue has drained this buffer
        tc::fence_after_sync();
        const uint32_t acc = tmem_acc + (uint32_t)(b * BN);
        for (int ki = 0; ki < KI; ++ki, ++it) {
          const int s = it % NS, use = it / NS;
          mbar_wait(&conv_full[s], (uint32_t)(use & 1));
          EXPO_TRACE(it, 3);
          tc::fence_after_sync();
          const uint32_t sa_hi = smem_u32(base + (size_t)s * C::kStageBytes), sa_lo = sa_hi + kTileA,
                         sb_hi = sa_lo + kTileA, sb_lo = sb_hi + C::kTileB;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ao = P::kAMn ? k * 1024 : k * 32, bo = P::kBMn ? k * 1024 : k * 32;
            const uint64_t dah = P::kAMn ? smem_desc_mn32(sa_hi + ao) : tc::smem_desc_sw128(sa_hi + ao);
            const uint64_t dal = P::kAMn ? smem_desc_mn32(sa_lo + ao) : tc::smem_desc_sw128(sa_lo + ao);
            const uint64_t dbh = P::kBMn ? smem_desc_mn32(sb_hi + bo) : tc::smem_desc_sw128(sb_hi + bo);
            const uint64_t dbl = P::kBMn ? smem_desc_mn32(sb_lo + bo) : tc::smem_desc_sw128(sb_lo + bo);
            tc::mma_tf32(acc, dah, dbh, idesc, (ki > 0 || k > 0) ? 1u : 0u);
            tc::mma_tf32(acc, dah, dbl, idesc, 1u);
            tc::mma_tf32(acc, dal, dbh, idesc, 1u);
          }
          tc::mma_commit(&empty[s]);
          EXPO_TRACE(it, 4);
        }
        tc::mma_commit(&acc_full[b]);                          // this tile's accumulator is complete
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // ================================ converters ================================
    const int tc_ = tid - 64;
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      int m0, n0, z;
      decode(t, m0, n0, z);
      const int KI = p.k_iters(z);
      for (int ki = 0; ki < KI; ++ki, ++it) {
        const int s = it % NS, use = it / NS;
        unsigned char* a_raw = base + (size_t)s * C::kStageBytes;
        mbar_wait(&raw_full[s], (uint32_t)(use & 1));
        if (tc_ == 0) EXPO_TRACE(it, 1);
#pragma unroll
        for (int q = 0; q < C::kVecPerThread; ++q) {
          const int i = tc_ + kConverters * q;
          const bool is_a = i < kTileA / 16;
          unsigned char* src = is_a ? a_raw + (size_t)i * 16 : a_raw + 2 * kTileA + (size_t)(i - kTileA / 16) * 16;
          const float4 v = *reinterpret_cast<const float4*>(src);
          float4 l;
          l.x = tc::rn_tf32(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
          l.y = tc::rn_tf32(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
          l.z = tc::rn_tf32(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
          l.w = tc::rn_tf32(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
          *reinterpret_cast<float4*>(src + (is_a ? kTileA : C::kTileB)) = l;
        }
        fence_proxy_async();
        mbar_arrive(&conv_full[s]);
        if (tc_ == 0) EXPO_TRACE(it, 2);
      }
    }
  } else {
    // ================================ epilogue ================================
    const int q = warp & 3;                                    // warps 6..9 -> TMEM lane quarters 2,3,0,1
    const int row = q * 32 + lane;
    int j = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++j) {
      int m0, n0, z;
      decode(t, m0, n0, z);
      const int KI = p.k_iters(z);
      const int b = j & 1, useb = j >> 1;
      if (tid == 192) EXPO_TRACE_EPI(j, 0);
      mbar_wait(&acc_full[b], (uint32_t)(useb & 1));
      if (tid == 192) EXPO_TRACE_EPI(j, 1);
      tc::fence_after_sync();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        if (KI > 0) tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + c0), v);
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        p.store16(z, m0 + row, n0 + c0, v);
      }
      tc::fence_before_sync();
      mbar_arrive(&acc_empty[b]);                              // buffer b may be overwritten
      if (tid == 192) EXPO_TRACE_EPI(j, 2);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_acc, 2 * BN);
}

template <class P, int BN>
inline cudaError_t launch_tma_gemm_persistent(const P& p, int M, int N, int Z, cudaStream_t st) {
  constexpr size_t smem = PersistCfg<BN>::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tma_gemm_persistent_kernel<P, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int MX = (M + kBM - 1) / kBM, NY = (N + BN - 1) / BN;
  const int tiles = MX * NY * Z;
  const int grid = tiles < 148 ? tiles : 148;
  launch_pdl(tma_gemm_persistent_kernel<P, BN>, dim3(grid), dim3(kPersistThreads), smem, st, p, MX, NY, tiles);
  return cudaGetLastError();
}

// Picks the persistent kernel for grids of at least two waves of one-CTA-per-SM tiles (no cluster
// split-K), the one-tile-per-CTA kernel otherwise.  EXPOSURE_TMA_PERSIST=0/1 forces the choice.
template <class P, int BN>
inline cudaError_t launch_tma_auto(const P& p, int M, int N, int Z, cudaStream_t st) {
  static const int forced = [] { const char* e = getenv("EXPOSURE_TMA_PERSIST"); return e ? atoi(e) : -1; }();   // tuning aid
  const long tiles = (long)((M + kBM - 1) / kBM) * ((N + BN - 1) / BN) * Z;
  const bool persist = p.splits <= 1 && (forced >= 0 ? forced != 0 : tiles >= 296);
  if (persist) return launch_tma_gemm_persistent<P, BN>(p, M, N, Z, st);
  return launch_tma_gemm<P, BN>(p, M, N, Z, st);
}

}  // namespace tma
}  // namespace expo
