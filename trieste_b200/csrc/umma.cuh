// tcgen05 / TMEM plumbing and the operand-tile geometry shared by the digit GEMMs (ozaki.cuh, ozaki5.cuh).
#pragma once
#include "common.cuh"

namespace tb {
namespace oz {

constexpr int KST = 64;                      // K bytes (= k columns) per pipeline stage
constexpr uint32_t LBO = 128;                // core matrices adjacent in K (no-swizzle K-major layout)
constexpr uint32_t SBO = (KST / 16) * 128;   // 8-row groups

__host__ __device__ inline int64_t a_stage_offset(int I) {  // stages before row-block I: Σ 2(i+1)
  return (int64_t)I * (I + 1);
}

// insert byte SRC (0..3) of w into byte POS (0..3) of acc: one PRMT
template <int POS, int SRC>
__device__ __forceinline__ uint32_t put_byte(uint32_t acc, uint32_t w) {
  constexpr uint32_t sel = (POS == 0 ? (4u + SRC) : 0u) | ((POS == 1 ? (4u + SRC) : 1u) << 4) | ((POS == 2 ? (4u + SRC) : 2u) << 8) |
                           ((POS == 3 ? (4u + SRC) : 3u) << 12);
  return __byte_perm(acc, w, sel);
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace oz
}  // namespace tb
