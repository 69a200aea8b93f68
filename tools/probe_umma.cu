// Hardware probe (development tool, not part of the product library): pins down
//   (1) the shared-memory layout tcgen05.mma reads for MN-major SWIZZLE_128B TF32 operands,
//   (2) how kind::tf32 converts fp32 bit patterns (truncate / round),
//   (3) what a 4-D tiled TMA load with element strides (the 4x4 stride-2 im2col box) writes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/probe_umma tools/probe_umma.cu
// Run on a B200: tools/bin/probe_umma > gpurun_out/probe_umma.txt
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// (1)+(2): one tcgen05.mma (M = 128, N = 32, K = 8, kind::tf32) on caller-filled smem images
// ---------------------------------------------------------------------------------------------
struct MmaProbe {
  uint64_t adesc_hi, bdesc_hi;   // descriptor bits above the start address
  uint32_t idesc, a_off, b_off;  // byte offsets of the descriptor start inside the regions
};

__global__ void __launch_bounds__(128) mma_probe_kernel(const float* a_img, const float* b_img, MmaProbe pr, float* C) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* As = reinterpret_cast<float*>(base);            // 16 KiB
  float* Bs = reinterpret_cast<float*>(base + 16384);    // 16 KiB
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4096; i += 128) { As[i] = a_img[i]; Bs[i] = b_img[i]; }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tmem_base_s;
  if (tid == 0) {
    const uint64_t ad = pr.adesc_hi | (uint64_t)(((smem_u32(As) + pr.a_off) >> 4) & 0x3FFF);
    const uint64_t bd = pr.bdesc_hi | (uint64_t)(((smem_u32(Bs) + pr.b_off) >> 4) & 0x3FFF);
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tacc),
        "l"(ad), "l"(bd), "r"(pr.idesc), "r"(0u)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  mbar_wait(&bar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tacc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) C[(warp * 32 + lane) * 32 + c0 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc), "r"(32u) : "memory");
}

static uint32_t k_off(int r, int c) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)); }   // K-major SW128
static uint64_t desc_hi(uint32_t lbo, uint32_t sbo) { return ((uint64_t)lbo << 16) | ((uint64_t)sbo << 32) | (1ull << 46) | (2ull << 61); }
static uint32_t idesc(int M, int N, bool amn, bool bmn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (amn ? 1u << 15 : 0u) | (bmn ? 1u << 16 : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

static std::vector<float> run_mma(const std::vector<float>& a, const std::vector<float>& b, MmaProbe pr) {
  float *da, *db, *dc;
  CK(cudaMalloc(&da, 16384)); CK(cudaMalloc(&db, 16384)); CK(cudaMalloc(&dc, 128 * 32 * 4));
  CK(cudaMemcpy(da, a.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), 16384, cudaMemcpyHostToDevice));
  CK(cudaMemset(dc, 0xff, 128 * 32 * 4));
  CK(cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024));
  mma_probe_kernel<<<1, 128, 34 * 1024>>>(da, db, pr, dc);
  CK(cudaDeviceSynchronize());
  std::vector<float> c(128 * 32);
  CK(cudaMemcpy(c.data(), dc, 128 * 32 * 4, cudaMemcpyDeviceToHost));
  cudaFree(da); cudaFree(db); cudaFree(dc);
  return c;
}

static void probe_layouts() {
  std::vector<float> ident_k(4096, 0.f);            // K-major tile X[r][k] = (r == k), k < 8
  for (int r = 0; r < 8; ++r) ident_k[k_off(r, r >> 2) / 4 + (r & 3)] = 1.f;
  // sanity: K-major x K-major, A[m][k] = m + 1 for k == 0 -> C[m][0] = m + 1
  {
    std::vector<float> a(4096, 0.f);
    for (int m = 0; m < 128; ++m) a[k_off(m, 0) / 4] = (float)(m + 1);
    MmaProbe pr{desc_hi(1, 64), desc_hi(1, 64), idesc(128, 32, false, false), 0, 0};
    auto c = run_mma(a, ident_k, pr);
    int bad = 0;
    for (int m = 0; m < 128; ++m) bad += c[m * 32] != (float)(m + 1);
    printf("[sanity K-major] mismatches %d (C[5][0]=%g C[5][1]=%g)\n", bad, c[5 * 32], c[5 * 32 + 1]);
  }
  // conversion of fp32 bit patterns by kind::tf32
  {
    std::vector<float> a(4096, 0.f);
    const uint32_t pats[4] = {0x3F801000u, 0x3F801800u, 0x3F800FFFu, 0x3F803000u};   // 1+2^-11, 1+2^-11+2^-12, 1+(2^-11-ulp), 1+2^-10+2^-11
    for (int m = 0; m < 4; ++m) memcpy(&a[k_off(m, 0) / 4], &pats[m], 4);
    MmaProbe pr{desc_hi(1, 64), desc_hi(1, 64), idesc(128, 32, false, false), 0, 0};
    auto c = run_mma(a, ident_k, pr);
    for (int m = 0; m < 4; ++m) {
      uint32_t o;
      memcpy(&o, &c[m * 32], 4);
      printf("[tf32 conversion] in 0x%08x -> out 0x%08x\n", pats[m], o);
    }
  }
  // MN-major TF32 operands: the only legal layout is SWIZZLE_128B_BASE32B (layout_type 1): atoms of
  // 32 MN elements x 4 K rows, 32-byte chunk index XOR (k & 3); LBO between 32-wide MN blocks, SBO
  // between groups of 4 K rows.  Discovery: fill the image with the float index, other operand = identity.
  auto mn_off = [](int mn, int k, int lbo_b, int sbo_b) {
    return ((mn >> 5) * lbo_b + (k >> 2) * sbo_b + (k & 3) * 128 + ((((mn >> 3) & 3) ^ (k & 3)) << 5) + (mn & 7) * 4) / 4;
  };
  auto desc_hi_l1 = [](uint32_t lbo, uint32_t sbo) { return ((uint64_t)lbo << 16) | ((uint64_t)sbo << 32) | (1ull << 46) | (1ull << 61); };
  {
    std::vector<float> lo(4096), hi(4096);
    for (int p = 0; p < 4096; ++p) { lo[p] = (float)(p & 1023); hi[p] = (float)(p >> 10); }
    MmaProbe pr{desc_hi_l1(64, 32), desc_hi(1, 64), idesc(128, 32, true, false), 0, 0};
    auto c1 = run_mma(lo, ident_k, pr), c2 = run_mma(hi, ident_k, pr);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int k = 0; k < 8; ++k) bad += ((int)c2[m * 32 + k] * 1024 + (int)c1[m * 32 + k]) != mn_off(m, k, 1024, 512);
    printf("[A MN-major BASE32B, LBO=1024B SBO=512B] float index read for (m,k):\n");
    for (int m : {0, 1, 2, 7, 8, 9, 16, 24, 31, 32, 33, 64, 127}) {
      printf("  m=%3d:", m);
      for (int k = 0; k < 8; ++k) printf(" %5d", (int)c2[m * 32 + k] * 1024 + (int)c1[m * 32 + k]);
      printf("\n");
    }
    printf("  mismatches vs assumed layout: %d of 1024\n", bad);
  }
  {
    std::vector<float> lo(4096), hi(4096);
    for (int p = 0; p < 4096; ++p) { lo[p] = (float)(p & 1023); hi[p] = (float)(p >> 10); }
    MmaProbe pr{desc_hi(1, 64), desc_hi_l1(64, 32), idesc(128, 32, false, true), 0, 0};
    auto c1 = run_mma(ident_k, lo, pr), c2 = run_mma(ident_k, hi, pr);
    int bad = 0;
    for (int k = 0; k < 8; ++k)
      for (int n = 0; n < 32; ++n) bad += ((int)c2[k * 32 + n] * 1024 + (int)c1[k * 32 + n]) != mn_off(n, k, 1024, 512);
    printf("[B MN-major BASE32B N=32, SBO=512B] float index read for (k, n=0..11):\n");
    for (int k = 0; k < 8; ++k) {
      printf("  k=%d:", k);
      for (int n = 0; n < 12; ++n) printf(" %5d", (int)c2[k * 32 + n] * 1024 + (int)c1[k * 32 + n]);
      printf("\n");
    }
    printf("  mismatches vs assumed layout: %d of 256\n", bad);
  }
}

// ---------------------------------------------------------------------------------------------
// (3) 4-D tiled TMA with element strides: dump the smem box
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3,
                                                        uint32_t bytes, float* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* S = reinterpret_cast<float*>(base);
  for (int i = threadIdx.x; i < (int)(bytes / 4); i += 128) S[i] = -7.f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(S)),
        "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(&bar))
        : "memory");
  }
  mbar_wait(&bar, 0u);
  for (int i = threadIdx.x; i < (int)(bytes / 4); i += 128) out[i] = S[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


static void probe_tma_mn(EncodeTiledFn encode) {
  // weights-like matrix W[K=64][N=96] (N contiguous): box {32 n, 32 k} with SWIZZLE_128B_ATOM_32B
  const int K = 64, N = 96;
  std::vector<float> w((size_t)K * N);
  for (size_t i = 0; i < w.size(); ++i) w[i] = (float)i;
  float* dw;
  CK(cudaMalloc(&dw, w.size() * 4));
  CK(cudaMemcpy(dw, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)K, 1, 1};
  cuuint64_t strides[3] = {(cuuint64_t)N * 4, (cuuint64_t)N * K * 4, (cuuint64_t)N * K * 4};
  cuuint32_t box[4] = {32, 32, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("[tma mn] encode SWIZZLE_128B_ATOM_32B -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return;
  const uint32_t bytes = 32 * 32 * 4;
  float* dout;
  CK(cudaMalloc(&dout, bytes));
  tma_probe_kernel<<<1, 128, 20 * 1024>>>(tm, 64, 32, 0, 0, bytes, dout);    // n0 = 64, k0 = 32
  CK(cudaDeviceSynchronize());
  std::vector<float> s(bytes / 4);
  CK(cudaMemcpy(s.data(), dout, bytes, cudaMemcpyDeviceToHost));
  int bad = 0, shown = 0;
  for (int k = 0; k < 32; ++k)
    for (int n = 0; n < 32; ++n) {
      const float expect = w[(size_t)(32 + k) * N + 64 + n];
      const int off = ((k >> 2) * 512 + (k & 3) * 128 + ((((n >> 3) & 3) ^ (k & 3)) << 5) + (n & 7) * 4) / 4;
      if (s[off] != expect) {
        ++bad;
        if (shown++ < 8) printf("    k %d n %d: got %g expect %g\n", k, n, s[off], expect);
      }
    }
  printf("[tma mn box 32x32] mismatches vs BASE32B layout (k-groups 512 B apart): %d of 1024\n", bad);
  cudaFree(dw); cudaFree(dout);
}

static void probe_tma() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("[tma] no cuTensorMapEncodeTiled\n"); return; }
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const int B = 3, IH = 8, IW = 8, C = 64;
  std::vector<float> x((size_t)B * IH * IW * C);
  for (size_t i = 0; i < x.size(); ++i) x[i] = (float)i;
  float* dx;
  CK(cudaMalloc(&dx, x.size() * 4));
  CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
  const int tw = 4, th = 4, tb = 2;        // 32 output pixels per box
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)IW * C * 4, (cuuint64_t)IH * IW * C * 4};
  cuuint32_t box[4] = {32, 2 * tw, 2 * th, (cuuint32_t)tb};
  cuuint32_t estr[4] = {1, 2, 2, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("[tma] encode (box 32,%d,%d,%d strides 1,2,2,1) -> %d\n", 2 * tw, 2 * th, tb, (int)r);
  if (r != CUDA_SUCCESS) return;
  const uint32_t bytes = 32 * 4 * tw * th * tb;
  float* dout;
  CK(cudaMalloc(&dout, bytes));
  CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 1024));
  for (int trial = 0; trial < 3; ++trial) {
    const int ky = trial == 0 ? 0 : (trial == 1 ? 3 : 1), kx = trial == 0 ? 0 : (trial == 1 ? 3 : 2);
    const int ci0 = trial == 2 ? 32 : 0, b0 = trial == 1 ? 2 : 0;      // trial 1: second image of the box is past the batch
    tma_probe_kernel<<<1, 128, 20 * 1024>>>(tm, ci0, kx - 1, ky - 1, b0, bytes, dout);
    CK(cudaDeviceSynchronize());
    std::vector<float> s(bytes / 4);
    CK(cudaMemcpy(s.data(), dout, bytes, cudaMemcpyDeviceToHost));
    int bad = 0, shown = 0;
    for (int b = 0; b < tb; ++b)
      for (int oy = 0; oy < th; ++oy)
        for (int ox = 0; ox < tw; ++ox)
          for (int c = 0; c < 32; ++c) {
            const int row = (b * th + oy) * tw + ox;
            const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx, bb = b0 + b;
            float expect = 0.f;
            if (iy >= 0 && iy < IH && ix >= 0 && ix < IW && bb < B) expect = x[(((size_t)bb * IH + iy) * IW + ix) * C + ci0 + c];
            const float got = s[k_off(row, c >> 2) / 4 + (c & 3)];
            if (got != expect) {
              ++bad;
              if (shown++ < 6) printf("    row %d c %d: got %g expect %g\n", row, c, got, expect);
            }
          }
    printf("[tma im2col ky=%d kx=%d ci0=%d b0=%d] mismatches %d of %d\n", ky, kx, ci0, b0, bad, 32 * tw * th * tb);
  }
  cudaFree(dx); cudaFree(dout);
  probe_tma_mn(encode);
}

int main() {
  probe_layouts();
  probe_tma();
  return 0;
}
