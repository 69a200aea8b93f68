// Persistent TMA-staged variant of the fused filter step (EXP_VARIANT_TMA).
//
// The batch is one flat sequence of 1024-pixel tiles (12 KiB of interleaved RGB fp32).  The
// grid is sized to the number of CTAs that are co-resident (SMs x occupancy); CTA i owns the
// contiguous tile range [i*q, (i+1)*q) -- no wave quantisation, no tail.  Per tile:
//
//   elected thread: cp.async.bulk global -> shared (x tile [+ gy tile]) on an mbarrier ring
//   all threads   : wait(mbarrier) -> 3 x LDS.128 per operand (48 B stride: conflict free)
//                   -> per-pixel math in registers -> 3 x STS.128 in place
//   elected thread: fence.proxy.async + cp.async.bulk shared -> global (bulk_group)
//
// so every HBM transaction is a full, aligned, coalesced bulk copy and the number of bytes
// in flight is set by the ring depth, not by register occupancy.  Parameter-gradient
// accumulators live in registers across tiles and are flushed when the CTA crosses an image
// boundary: record index = cta + image (unique, ordered), finished in fixed order by the
// last CTA to arrive for that image (deterministic).
#pragma once
#include "async_ptx.cuh"
#include "filter_math.cuh"

namespace expo {

constexpr int kTilePx = 1024;                       // 256 threads x 4 pixels
constexpr int kTileBytes = kTilePx * 12;            // 12 KiB
constexpr int kTmaThreads = 256;

struct TmaArgs {
  const float* x;
  const float* gy;
  float* out;
  const float* params;
  int pstride;
  int P;                  // pixels per image (multiple of 4)
  int tiles_per_image;    // ceil(P / 1024)
  int total_tiles;        // B * tiles_per_image
  int B;
  float* partials;        // [(grid + B)][kAccStride]
  unsigned* counters;     // [B]
  float* gparams;
  int logits;             // params are raw regressor logits (EXP_OPT_LOGITS)
  FilterRanges rg;        // cfg-driven regressor ranges (exp_set_filter_ranges)
};

// Balanced static partition of the flat tile sequence: CTA i owns [floor(i*total/grid),
// floor((i+1)*total/grid)); sizes differ by at most one tile.
__device__ __forceinline__ int cta_first_tile(int i, int grid, int total) {
  return (int)(((long long)i * total) / grid);
}
__device__ __forceinline__ int cta_of_tile(long long n, int grid, int total) {
  return (int)(((n + 1) * grid - 1) / total);
}

// Flush the CTA's accumulators for image b into record (cta + b) and, if this CTA is the last
// of the image to arrive, finish dL/dparams[b] in fixed record order.
template <int FID>
__device__ __forceinline__ void tma_flush(float* acc, const TmaArgs& A, const FilterConsts& sc, float (*red)[kAccStride],
                                          double* tot, unsigned* ticket, int b) {
  constexpr int NACC = num_acc(FID);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NACC; ++a) {
    const float v = warp_sum(acc[a]);
    if (lane == 0) red[warp][a] = v;
    acc[a] = 0.f;
  }
  __syncthreads();
  const int first = cta_of_tile((long long)b * A.tiles_per_image, gridDim.x, A.total_tiles);
  const int last = cta_of_tile((long long)(b + 1) * A.tiles_per_image - 1, gridDim.x, A.total_tiles);
  if (threadIdx.x < NACC) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kTmaThreads / 32; ++w) s += red[w][threadIdx.x];
    A.partials[((size_t)blockIdx.x + b) * kAccStride + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *ticket = atomicAdd(A.counters + b, 1u);
  __syncthreads();
  if (*ticket == (unsigned)(last - first)) {
    __threadfence();
    if (threadIdx.x < NACC) {
      double s = 0.0;
      for (int i = first; i <= last; ++i)
        s += (double)__ldcg(A.partials + ((size_t)i + b) * kAccStride + threadIdx.x);
      tot[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      finalize_grads(FID, tot, sc, A.logits, A.gparams + (size_t)b * A.pstride);
      A.counters[b] = 0u;
    }
  }
  __syncthreads();   // sc / red / tot may be rewritten by the caller after this
}

template <int FID, bool BWD, bool HAS_GX, int STAGES>
__global__ void __launch_bounds__(kTmaThreads) filter_step_tma_kernel(const TmaArgs A) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int kOperands = BWD ? 2 : 1;
  float* bufx = reinterpret_cast<float*>(smem_raw);                              // [STAGES][3072]
  float* bufg = reinterpret_cast<float*>(smem_raw + (size_t)STAGES * kTileBytes); // [STAGES][3072] (BWD)
  __shared__ __align__(8) uint64_t full[STAGES];
  __shared__ FilterConsts sc;
  __shared__ float red[BWD ? kTmaThreads / 32 : 1][kAccStride];
  __shared__ double tot[BWD ? kAccStride : 1];
  __shared__ unsigned ticket;

  pdl_trigger();             // the next step's CTAs may be scheduled as ours drain
  const int t0 = cta_first_tile(blockIdx.x, gridDim.x, A.total_tiles);
  const int t1 = cta_first_tile(blockIdx.x + 1, gridDim.x, A.total_tiles);
  if (t0 >= t1) return;      // only when grid > total_tiles (the host never launches that)
  const int ntiles = t1 - t0;
  const int tid = threadIdx.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();                // x / gy / params of this step are the previous kernel's outputs

  // tile n (global index) -> image, pixel offset, byte count
  auto tile_info = [&](int n, int& b, size_t& off_floats, uint32_t& bytes) {
    b = n / A.tiles_per_image;
    const int j = n - b * A.tiles_per_image;
    const int px0 = j * kTilePx;
    const int npx = min(kTilePx, A.P - px0);
    off_floats = ((size_t)b * A.P + px0) * 3;
    bytes = (uint32_t)npx * 12u;
  };
  auto issue_load = [&](int i) {   // i = local tile index; elected thread only
    int b; size_t off; uint32_t bytes;
    tile_info(t0 + i, b, off, bytes);
    const int s = i % STAGES;
    mbar_expect_tx(&full[s], bytes * kOperands);
    bulk_g2s(bufx + (size_t)s * (kTileBytes / 4), A.x + off, bytes, &full[s]);
    if constexpr (BWD) bulk_g2s(bufg + (size_t)s * (kTileBytes / 4), A.gy + off, bytes, &full[s]);
  };

  if (tid == 0) {
    const int pre = min(STAGES - 1, ntiles);
    for (int i = 0; i < pre; ++i) issue_load(i);
  }

  constexpr int NACC = num_acc(FID);
  float acc[NACC];
#pragma unroll
  for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
  int cur_b = -1;
  constexpr bool kStores = !BWD || HAS_GX;

  for (int i = 0; i < ntiles; ++i) {
    int b; size_t off; uint32_t bytes;
    tile_info(t0 + i, b, off, bytes);
    if (b != cur_b) {                                  // CTA-uniform
      if constexpr (BWD) {
        if (cur_b >= 0) tma_flush<FID>(acc, A, sc, red, tot, &ticket, cur_b);
      }
      __syncthreads();                                 // everyone done with the old constants
      if (tid < 32) setup_consts(sc, A.params + (size_t)b * A.pstride, FID, A.logits, A.rg);
      __syncthreads();
      cur_b = b;
    }
    const int s = i % STAGES;
    mbar_wait(&full[s], (uint32_t)((i / STAGES) & 1));

    float4* sx = reinterpret_cast<float4*>(bufx + (size_t)s * (kTileBytes / 4)) + tid * 3;
    float4* sg = reinterpret_cast<float4*>(bufg + (size_t)s * (kTileBytes / 4)) + tid * 3;
    if ((uint32_t)tid * 48u < bytes) {                 // partial tiles: bytes is a multiple of 48
      Px4 vx;
      vx.a = sx[0]; vx.b = sx[1]; vx.c = sx[2];
      float px[4][3], py[4][3];
      unpack(vx, px);
      if constexpr (BWD) {
        Px4 vg;
        vg.a = sg[0]; vg.b = sg[1]; vg.c = sg[2];
        float pg[4][3];
        unpack(vg, pg);
#pragma unroll
        for (int k = 0; k < 4; ++k) px_bwd<FID, HAS_GX>(px[k], pg[k], py[k], acc, sc);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) px_fwd<FID>(px[k], py[k], sc);
      }
      if constexpr (kStores) {
        const Px4 vy = pack(py);
        sx[0] = vy.a; sx[1] = vy.b; sx[2] = vy.c;      // in place: same thread, same 48 bytes
      }
    }
    if constexpr (kStores) fence_proxy_async();         // generic-proxy writes -> visible to TMA
    __syncthreads();
    if (tid == 0) {
      if constexpr (kStores) {
        bulk_s2g(A.out + off, bufx + (size_t)s * (kTileBytes / 4), bytes);
        bulk_commit();
      }
      // refill the stage of tile i-1 (its store, issued last iteration, must have been read)
      const int nxt = i + STAGES - 1;
      if (nxt < ntiles) {
        if constexpr (kStores) bulk_wait_read<1>();
        issue_load(nxt);
      }
    }
  }
  if constexpr (BWD) tma_flush<FID>(acc, A, sc, red, tot, &ticket, cur_b);
  if constexpr (kStores) {
    if (tid == 0) bulk_wait_read<0>();
  }
}

}  // namespace expo
