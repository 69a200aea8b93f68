// TMA-fed tcgen05 GEMM engine (third generation; backend 4).
//
// The first two engines gather operands through registers: even warp-specialised, the producers
// spend ~50 instructions per 16 bytes on im2col address arithmetic and the tensor pipe idles
// (profiles/r1_layer_bench_*.md).  Here NO thread computes an operand address:
//
//   warp 0, one lane   TMA PRODUCER   one cp.async.bulk.tensor per operand block and K step:
//                                     * K-major operands (im2col rows, delta rows, dgrad weights):
//                                       4-D boxes {32 ch, w, h, b} with SWIZZLE_128B -- for the
//                                       4x4 stride-2 convolution the box walks the input with
//                                       element strides {1,2,2,1} and the -1 padding, the image
//                                       border and the batch tail are the TMA's out-of-bounds
//                                       zero fill;
//                                     * MN-major operands (weights [K][N], wgrad activations and
//                                       deltas [pixel][channel]): boxes {32 ch, 32 k rows} with
//                                       SWIZZLE_128B_ATOM_32B, which is the only layout
//                                       tcgen05.mma accepts for MN-major TF32
//                                       (UMMA layout_type 1, "128B_BASE32B"; verified with
//                                       tools/probe_umma.cu, profiles/r1_probe_umma.txt)
//   warps 2-5          CONVERTERS     3xTF32 split in shared memory: kind::tf32 TRUNCATES fp32
//                                     inputs (probe), so the raw TMA tile already is the "hi"
//                                     operand; the converters only write lo = rna(v - trunc(v))
//                                     at the same byte offset of a second tile (layout-agnostic)
//   warp 1, one lane   MMA ISSUER     4 x 3 tcgen05.mma.kind::tf32 per K step
//                                     (hi*hi + hi*lo + lo*hi), tcgen05.commit frees the stage
//   warps 2-5          EPILOGUE       tcgen05.ld 32x32b (warp w owns TMEM lanes 32*(w%4)..)
//
// mbarriers per stage: raw_full (TMA transaction bytes) -> conv_full (128 converter arrivals)
// -> empty (tcgen05.commit).
#pragma once
#include <cuda.h>

#include <cstdlib>

#include "tcgen05_ptx.cuh"

namespace expo {
namespace tma {

constexpr int kThreads = 192;
constexpr int kConverters = 128;
constexpr int kBM = 128;
constexpr int kBK = 32;
constexpr int kTileA = kBM * 128;        // bytes of one A tile (128 x 32 fp32)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// MN-major TF32 tile: blocks of 32 (MN) x 32 (K rows of 128 B) = 4 KiB; LBO = 4096 B between
// MN blocks, SBO = 512 B between groups of 4 K rows, layout_type 1 (SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (256ull << 16) | (32ull << 32) | (1ull << 46) | (1ull << 61);
}
__host__ __device__ constexpr uint32_t idesc_tf32_major(int M, int N, bool a_mn, bool b_mn) {
  return tc::idesc_tf32(M, N) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
}

// NS = 0: deepest ring that fits one CTA per SM (3 stages at BN = 128, else 4).  NS = 2: a two-stage ring
// small enough for two CTAs per SM -- used for multi-wave grids of short CTAs, where the second
// resident CTA hides the prologue / epilogue of the first (BN <= 64 only).
template <int BN, int NS_ = 0>
struct Cfg {
  static constexpr int kStages = NS_ > 0 ? NS_ : (BN >= 128 ? 3 : 4);
  static constexpr int kCtasPerSm = (NS_ == 2 && BN <= 64) ? 2 : 1;
  static constexpr int kTileB = BN * 128;
  static constexpr int kStageBytes = 2 * kTileA + 2 * kTileB;       // raw + lo for A and B
  static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024;
  static constexpr int kVecPerThread = (kTileA + kTileB) / 16 / kConverters;
};

// thread-block cluster helpers (split-K reduction through distributed shared memory)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t local_addr, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}

// sum of the float4 at local shared address `a` over ranks 0..S-1, in rank order; the DSMEM loads
// are issued back to back before the first add (a remote load costs ~1 us)
template <int S>
__device__ __forceinline__ float4 dsmem_sum4(uint32_t a) {
  float4 t[S];
#pragma unroll
  for (int s = 0; s < S; ++s) t[s] = ld_dsmem_v4(a, (uint32_t)s);
  float4 v = t[0];
#pragma unroll
  for (int s = 1; s < S; ++s) { v.x += t[s].x; v.y += t[s].y; v.z += t[s].z; v.w += t[s].w; }
  return v;
}

// Development aid (-DEXPO_TMA_TRACE, tools/tma_trace.py): SM-clock stamps of the ring hand-overs of the first CTAs.
#ifdef EXPO_TMA_TRACE
constexpr int kTraceCtas = 4, kTraceSteps = 96, kTraceTiles = 16;
__device__ long long g_tma_trace[kTraceCtas][kTraceSteps][5];
__device__ long long g_tma_trace_epi[kTraceCtas][kTraceTiles][3];
__device__ long long g_tma_trace_cta[kTraceCtas][8];      // one-tile kernel: entry, set-up done, accumulator ready, tile stored, exit
#define EXPO_TRACE(it, slot)                                                                          \
  do {                                                                                                \
    if (blockIdx.x < kTraceCtas && (it) < kTraceSteps) g_tma_trace[blockIdx.x][(it)][(slot)] = clock64(); \
  } while (0)
#define EXPO_TRACE_EPI(j, slot)                                                                          \
  do {                                                                                                   \
    if (blockIdx.x < kTraceCtas && (j) < kTraceTiles) g_tma_trace_epi[blockIdx.x][(j)][(slot)] = clock64(); \
  } while (0)
#define EXPO_TRACE_CTA(slot)                                                                      \
  do {                                                                                            \
    const int cta__ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);             \
    if (cta__ < kTraceCtas) g_tma_trace_cta[cta__][(slot)] = clock64();                           \
  } while (0)
#else
#define EXPO_TRACE_CTA(slot) do {} while (0)
#define EXPO_TRACE(it, slot) do {} while (0)
#define EXPO_TRACE_EPI(j, slot) do {} while (0)
#endif

// Problem functor P (passed as a __grid_constant__ parameter: the CUtensorMaps inside it must stay
// in parameter space):
//   static constexpr bool kAMn, kBMn            operand majorness
//   int  splits                                  cluster split-K factor S (1, 2, 4 or 8; BN / S >= 16):
//                                                blockIdx.z = slice * S + rank, the S CTAs of a cluster
//                                                each take 1/S of the K steps, park their accumulators in
//                                                their own shared memory, and after a cluster barrier CTA
//                                                `rank` sums column chunk `rank` over all S CTAs in rank
//                                                order (deterministic) through DSMEM and stores it
//   int  k_iters(int slice)                      K steps of 32 for grid slice `slice`
//   void load<BN>(ki, slice, m0, n0, a_dst, b_dst, bar)   issue the TMA loads of K step ki
//                                                (exactly kTileA + BN * 128 bytes in total)
//   void store16(slice, m, n0, v[16])            C[m, n0..n0+15]
//   void store4(slice, m, n, float4)             C[m, n..n+3]
//   Aux  epi_load(slice, m, n); void epi_store(slice, m, n, float4, Aux)   the same in two halves (store pass)
template <class P, int BN, int NS_ = 0>
__global__ void __launch_bounds__(kThreads, Cfg<BN, NS_>::kCtasPerSm) tma_gemm_kernel(const __grid_constant__ P p) {
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN must be 32, 64 or 128");
  using C = Cfg<BN, NS_>;
  constexpr int NS = C::kStages;
  extern __shared__ __align__(1024) unsigned char tma_smem[];
  __shared__ __align__(8) uint64_t raw_full[NS];
  __shared__ __align__(8) uint64_t conv_full[NS];
  __shared__ __align__(8) uint64_t empty[NS];
  __shared__ __align__(8) uint64_t accum;
  __shared__ uint32_t tmem_base_s;

  pdl_trigger();                                              // dependents may start their prologue now
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) EXPO_TRACE_CTA(0);
  const int S = p.splits;
  const int z = blockIdx.z / S;                               // problem slice
  const int rank = S > 1 ? (int)cluster_ctarank() : 0;       // == blockIdx.z % S
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
  unsigned char* base = tma_smem + ((1024u - (smem_u32(tma_smem) & 1023u)) & 1023u);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, BN);
  if (tid == 32) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      mbar_init(&raw_full[s], 1);
      mbar_init(&conv_full[s], kConverters);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&accum, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.ta);
    tma_prefetch_desc(&p.tb);
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();                                                 // prologue done; the predecessor's writes are visible from here
  if (tid == 0) EXPO_TRACE_CTA(1);
  const uint32_t tmem_acc = tmem_base_s;
  const int KI_all = p.k_iters(z);
  const int per = (KI_all + S - 1) / S;
  const int kb = rank * per;                                  // first K step of this CTA
  const int KI = KI_all - kb < per ? (KI_all - kb > 0 ? KI_all - kb : 0) : per;
  float* park = reinterpret_cast<float*>(base);               // [128][BN + 4] accumulators (S > 1), reuses the stage ring
  constexpr int kParkLd = BN + 4;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      for (int ki = 0; ki < KI; ++ki) {
        const int s = ki % NS, use = ki / NS;
        if (use > 0) mbar_wait(&empty[s], (uint32_t)((use - 1) & 1));
        if (blockIdx.y == 0 && blockIdx.z == 0) EXPO_TRACE(ki, 0);
        unsigned char* a_raw = base + (size_t)s * C::kStageBytes;
        unsigned char* b_raw = a_raw + 2 * kTileA;
        mbar_expect_tx(&raw_full[s], (uint32_t)(kTileA + C::kTileB));
        p.template load<BN>(kb + ki, z, m0, n0, a_raw, b_raw, &raw_full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_tf32_major(kBM, BN, P::kAMn, P::kBMn);
      for (int ki = 0; ki < KI; ++ki) {
        const int s = ki % NS, use = ki / NS;
        mbar_wait(&conv_full[s], (uint32_t)(use & 1));
        if (blockIdx.y == 0 && blockIdx.z == 0) EXPO_TRACE(ki, 3);
        tc::fence_after_sync();
        const uint32_t sa_hi = smem_u32(base + (size_t)s * C::kStageBytes), sa_lo = sa_hi + kTileA,
                       sb_hi = sa_lo + kTileA, sb_lo = sb_hi + C::kTileB;
#pragma unroll
        for (int k = 0; k < 4; ++k) {             // 4 x (K = 8) per stage
          const uint32_t ao = P::kAMn ? k * 1024 : k * 32, bo = P::kBMn ? k * 1024 : k * 32;
          const uint64_t dah = P::kAMn ? smem_desc_mn32(sa_hi + ao) : tc::smem_desc_sw128(sa_hi + ao);
          const uint64_t dal = P::kAMn ? smem_desc_mn32(sa_lo + ao) : tc::smem_desc_sw128(sa_lo + ao);
          const uint64_t dbh = P::kBMn ? smem_desc_mn32(sb_hi + bo) : tc::smem_desc_sw128(sb_hi + bo);
          const uint64_t dbl = P::kBMn ? smem_desc_mn32(sb_lo + bo) : tc::smem_desc_sw128(sb_lo + bo);
          tc::mma_tf32(tmem_acc, dah, dbh, idesc, (ki > 0 || k > 0) ? 1u : 0u);
          tc::mma_tf32(tmem_acc, dah, dbl, idesc, 1u);
          tc::mma_tf32(tmem_acc, dal, dbh, idesc, 1u);
        }
        tc::mma_commit(&empty[s]);
        if (blockIdx.y == 0 && blockIdx.z == 0) EXPO_TRACE(ki, 4);
      }
      tc::mma_commit(&accum);
    }
    __syncwarp();
  } else {
    // ================================ converters ================================
    const int t = tid - 64;
    for (int ki = 0; ki < KI; ++ki) {
      const int s = ki % NS, use = ki / NS;
      unsigned char* a_raw = base + (size_t)s * C::kStageBytes;
      mbar_wait(&raw_full[s], (uint32_t)(use & 1));
      if (t == 0 && blockIdx.y == 0 && blockIdx.z == 0) EXPO_TRACE(ki, 1);
#pragma unroll
      for (int j = 0; j < C::kVecPerThread; ++j) {
        // flat 16-byte index over [A raw | A lo | B raw | B lo]: raw at off, lo at off + tile size
        const int i = t + kConverters * j;
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
      if (t == 0 && blockIdx.y == 0 && blockIdx.z == 0) EXPO_TRACE(ki, 2);
    }
    // ================================ epilogue ================================
    // Every tile is first parked in shared memory (the stage ring is idle by now), row per thread as tcgen05.ld
    // delivers it; the store pass below then walks it with consecutive threads on consecutive float4 of a row, so
    // the functor's global loads (lrelu' mask, dropout multiplier) and stores are whole 128-byte lines, four
    // independent float4 in flight per thread.  (Storing straight from the tcgen05.ld registers -- a thread per row,
    // 16 bytes per lane at a row stride -- cost 4-8 k clocks per tile, a third of the kernel: profiles/r2t_tma_trace.md.)
    if (KI > 0) mbar_wait(&accum, 0u);
    if (tid == 64) EXPO_TRACE_CTA(2);
    tc::fence_after_sync();
    const int q = warp & 3;
    const int row = q * 32 + lane;
#pragma unroll 2
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      if (KI > 0) tc::tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(park + (size_t)row * kParkLd + c0);
#pragma unroll
      for (int g = 0; g < 4; ++g) dst[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
    if constexpr (P::kPlainRows) {
      // a functor that stores the tile as it is: every thread hands ITS parked row to the bulk-copy engine
      // (one cp.async.bulk of BN * 4 bytes), no second pass and no barrier
      if (S == 1) {
        float* g = p.row_ptr(z, m0 + row, n0);
        fence_proxy_async();
        if (g) bulk_s2g(g, park + (size_t)row * kParkLd, BN * 4);
        bulk_commit();
        bulk_wait_read<0>();
      }
    }
    if (S == 1) asm volatile("bar.sync 1, %0;" ::"n"(kConverters) : "memory");   // the four epilogue warps only
  }
  if (tid == 64) EXPO_TRACE_CTA(3);
  if (S > 1) cluster_sync_all();                      // every CTA of the cluster has parked its partial tile
  if (warp >= 2 && !(P::kPlainRows && S == 1)) {
    // CTA `rank` owns columns [rank*W, (rank+1)*W) (W = BN when the tile is not split): consecutive threads take
    // consecutive float4 of a row chunk, so the (D)SMEM reads and the global accesses are contiguous runs of W*4 bytes
    const int W = BN / S, lgW4 = 31 - __clz(W >> 2);
    const int total = kBM << lgW4;
#pragma unroll 1
    for (int idx0 = tid - 64; idx0 < total; idx0 += 4 * kConverters) {
      float4 v[4];
      typename P::Aux aux[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                   // the functor's global operands (bias / mask / multiplier) first
        const int idx = idx0 + u * kConverters;
        if (idx < total) aux[u] = p.epi_load(z, m0 + (idx >> lgW4), n0 + rank * W + ((idx & ((1 << lgW4) - 1)) << 2));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * kConverters;
        if (idx < total) {
          const int r = idx >> lgW4, col = rank * W + ((idx & ((1 << lgW4) - 1)) << 2);
          const float* src = park + (size_t)r * kParkLd + col;
          if (S == 1) v[u] = *reinterpret_cast<const float4*>(src);
          else {
            const uint32_t a = smem_u32(src);
            v[u] = S == 2 ? dsmem_sum4<2>(a) : (S == 4 ? dsmem_sum4<4>(a) : dsmem_sum4<8>(a));   // fixed order: reproducible
          }
        }
      }
      if (tid == 64 && idx0 == 0) EXPO_TRACE_CTA(6);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * kConverters;
        if (idx < total) p.epi_store(z, m0 + (idx >> lgW4), n0 + rank * W + ((idx & ((1 << lgW4) - 1)) << 2), v[u], aux[u]);
      }
      if (tid == 64 && idx0 == 0) EXPO_TRACE_CTA(7);
    }
  }
  if (S > 1) cluster_sync_all();                      // nobody leaves while its shared memory is still being read
  if (tid == 64) EXPO_TRACE_CTA(4);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_acc, BN);
  if (tid == 0) EXPO_TRACE_CTA(5);
}

template <class P, int BN, int NS_>
inline cudaError_t launch_tma_gemm_ns(const P& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = Cfg<BN, NS_>::kSmem;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tma_gemm_kernel<P, BN, NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (p.splits <= 1) return launch_pdl(tma_gemm_kernel<P, BN, NS_>, grid, dim3(kThreads), smem, st, p);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)p.splits;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cfg.attrs = attr;
  return cudaLaunchKernelEx(&cfg, tma_gemm_kernel<P, BN, NS_>, p);
}

// Z = slices * p.splits; the p.splits CTAs that share a tile form one cluster along z
template <class P, int BN>
inline cudaError_t launch_tma_gemm(const P& p, int M, int N, int Z, cudaStream_t st) {
  const dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN, Z);
  if constexpr (BN <= 64) {
    static const int forced = [] { const char* e = getenv("EXPOSURE_TMA_2CTA"); return e ? atoi(e) : -1; }();   // tuning aid
    const long ctas = (long)grid.x * grid.y * grid.z;
    // measured (profiles/r1_layer_bench_tma.md): pays for more than 1.5 waves of CTAs with <= 8 K steps each
    const bool two = forced >= 0 ? forced != 0 : (ctas > 222 && p.k_iters(0) <= 8 * p.splits);
    if (two && p.splits <= 1) return launch_tma_gemm_ns<P, BN, 2>(p, grid, st);
  }
  return launch_tma_gemm_ns<P, BN, 0>(p, grid, st);
}

// cluster split-K factor: double while the doubled grid fits the 148 SMs, every CTA keeps >= 4 K
// steps and every rank keeps >= 16 output columns
inline int pick_splits(int base_ctas, int k_iters, int BN) {
  static const int forced = [] { const char* e = getenv("EXPOSURE_TMA_SPLITS"); return e ? atoi(e) : 0; }();   // tuning aid
  if (forced > 0) {
    int f = 1;
    while (f < forced && f < 8 && BN / (2 * f) >= 16 && k_iters / (2 * f) >= 1) f *= 2;
    return f;
  }
  int s = 1;
  // measured (profiles/r1_layer_bench_tma.md): a split pays only while the doubled grid still fits one
  // wave, and clusters of 8 schedule poorly with 1 CTA per SM
  while (s < 4 && 2 * base_ctas * s <= 148 && k_iters / (2 * s) >= 4 && BN / (2 * s) >= 16) s *= 2;
  return s;
}

}  // namespace tma
}  // namespace expo
