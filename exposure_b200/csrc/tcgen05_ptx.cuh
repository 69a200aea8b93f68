// tcgen05 (5th-gen tensor core) PTX wrappers shared by the TMA-fed GEMM engines (tc_engine_tma*.cuh):
// TMEM allocation, tcgen05.mma.kind::tf32 issue / commit, tcgen05.ld, shared-memory and instruction
// descriptors, and the round-to-nearest TF32 conversion of the "3xTF32" split (every operand value v is
// used as hi + lo; each K step issues hi*hi + hi*lo + lo*hi: the dropped lo*lo term and the TF32 rounding
// of lo are ~2^-22 relative, i.e. fp32-level accuracy -- the 1e-3 logit bound of the north star is kept
// with three orders of magnitude to spare).
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

}  // namespace tc
}  // namespace expo
