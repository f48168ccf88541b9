// Bring-up of the CTA-pair (cta_group::2) variant of the int8 digit MMA — the building block of the next GEMM step
// (DESIGN.md §7.1).  A cluster of two CTAs computes C[256 x 128] (s32) = A[256 x K] (s8) * B[128 x K]^T (s8):
// CTA r stages A rows [128 r, 128 r + 128) and B rows [64 r, 64 r + 64) (half of the N operand), the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = 128) over both CTAs' shared memory, every CTA reads its own 128 accumulator rows
// from its own TMEM.  Same no-swizzle K-major core-matrix layout and 1-D bulk copies as tools/i8_umma_test.cu; the
// peer CTA reports "operands loaded" to the leader with a remote mbarrier arrive, MMA completion is multicast to both.
// Checks against a CPU reference and measures the smem-resident MMA rate.  Measured on B200 (round 1): exact (0 mismatches
// of 32768; both CTAs' allocating warps are handed the same TMEM address), 3218 TOP/s at 8000 repetitions against 3838 TOP/s
// for cta_group::1 with the same N = 128: pairing alone does not raise the rate at this tile shape.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// acquire at cluster scope: the waiter consumes data published by a thread of the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t cta) {  // arrive on the same barrier of CTA `cta`
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (sm100); SWIZZLE_NONE
  return d;
}
// s8 x s8 -> s32, K-major A and B, N = 128, M = 256 (128 rows per CTA of the pair)
constexpr uint32_t IDESC_S8_256x128 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {  // arrives on `bar` of both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

constexpr int KSTAGE = 64;                  // K bytes per smem stage
constexpr int A_TILE = 128 * KSTAGE;        // 8 KB: this CTA's 128 rows of A
constexpr int B_TILE = 64 * KSTAGE;         // 4 KB: this CTA's 64 rows (half of N) of B
constexpr uint32_t LBO = 128, SBO = (KSTAGE / 16) * 128;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
i8_gemm2_kernel(const int8_t* __restrict__ Ap, const int8_t* __restrict__ Bp, int ksteps, int reps, int32_t* __restrict__ C) {
  extern __shared__ __align__(1024) unsigned char smem[];
  int8_t* sA = reinterpret_cast<int8_t*>(smem);                   // [ksteps][A_TILE]
  int8_t* sB = sA + (size_t)ksteps * A_TILE;                      // [ksteps][B_TILE]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)ksteps * B_TILE);  // [0] own loads, [1] peer loaded (leader), [2] mma done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp of EACH CTA of the pair takes part in the pair-wide allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();  // barriers of both CTAs initialised, TMEM of both allocated
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_own = *tmem_slot;  // what this CTA's allocating warp was handed (freed by the same warp)
  uint32_t tmem;                         // accumulator address the pair computes into: the leader's
  {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(tmem_slot)), "r"(0u));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(tmem) : "r"(raddr) : "memory");
  }

  if (threadIdx.x == 0) {
    // operands of this CTA: A rows [128 rank, +128) and B rows [64 rank, +64) of the pair's tile
    const int8_t* a_src = Ap + ((size_t)rank * ksteps) * A_TILE;
    const int8_t* b_src = Bp + ((size_t)rank * ksteps) * B_TILE;
    mbar_expect_tx(&bars[0], (uint32_t)ksteps * (A_TILE + B_TILE));
    for (int k = 0; k < ksteps; ++k) {
      bulk_g2s(sA + (size_t)k * A_TILE, a_src + (size_t)k * A_TILE, A_TILE, &bars[0]);
      bulk_g2s(sB + (size_t)k * B_TILE, b_src + (size_t)k * B_TILE, B_TILE, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    if (rank != 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive_remote(&bars[1], 0);  // tell the leader that the peer's half is in shared memory
    } else {
      mbar_wait_cluster(&bars[1], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int r = 0; r < reps; ++r) {
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t a0 = smem_u32(sA + (size_t)k * A_TILE), b0 = smem_u32(sB + (size_t)k * B_TILE);
#pragma unroll
          for (int kk = 0; kk < KSTAGE / 32; ++kk) {
            uint64_t da = make_desc(a0 + kk * 2 * LBO, LBO, SBO);
            uint64_t db = make_desc(b0 + kk * 2 * LBO, LBO, SBO);
            umma2_i8(tmem, da, db, IDESC_S8_256x128, (r | k | kk) ? 1u : 0u);
          }
        }
      }
      umma2_commit_multicast(&bars[2]);
    }
  }
  __syncwarp();
  mbar_wait(&bars[2], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: this CTA's 128 rows of the 256 x 128 tile; warp w reads TMEM lanes 32w..32w+31
  if (pair == 0) {
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
            "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
            "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int row = (int)rank * 128 + warp * 32 + lane;
      for (int j = 0; j < 32; ++j) C[row * 128 + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();  // both CTAs are done with the pair's TMEM
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_own), "r"(128u) : "memory");
}

// host-side packing of `rows` rows starting at row0: [kstep][rowgroup][kchunk 4][row 8][16 bytes]
static void pack(const std::vector<int8_t>& M, int K, int row0, int rows, std::vector<int8_t>& P) {
  const size_t tile = (size_t)rows * KSTAGE;
  P.assign((size_t)rows * K, 0);
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) {
      int ks = k / KSTAGE, kin = k % KSTAGE;
      size_t off = (size_t)ks * tile + (size_t)(r / 8) * SBO + (size_t)(kin / 16) * LBO + (r % 8) * 16 + (kin % 16);
      P[off] = M[(size_t)(row0 + r) * K + k];
    }
}

int main() {
  const int K = 256, ksteps = K / KSTAGE;
  std::vector<int8_t> A((size_t)256 * K), B((size_t)128 * K);
  srand(1);
  for (auto& x : A) x = (int8_t)(rand() % 129 - 64);
  for (auto& x : B) x = (int8_t)(rand() % 129 - 64);
  // device layout: [rank][kstep][tile]
  std::vector<int8_t> Ap, Bp, t;
  for (int r = 0; r < 2; ++r) {
    pack(A, K, 128 * r, 128, t);
    Ap.insert(Ap.end(), t.begin(), t.end());
    pack(B, K, 64 * r, 64, t);
    Bp.insert(Bp.end(), t.begin(), t.end());
  }
  int8_t *dA, *dB;
  int32_t* dC;
  CK(cudaMalloc(&dA, Ap.size()));
  CK(cudaMalloc(&dB, Bp.size()));
  CK(cudaMalloc(&dC, 256 * 128 * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, 256 * 128 * 4));
  const size_t smem = (size_t)ksteps * (A_TILE + B_TILE) + 64;
  CK(cudaFuncSetAttribute(i8_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  i8_gemm2_kernel<<<2, 128, smem>>>(dA, dB, ksteps, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> C(256 * 128);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < 256; ++i)
    for (int j = 0; j < 128; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)i * K + k] * (int32_t)B[(size_t)j * K + k];
      if (ref != C[i * 128 + j]) {
        if (bad < 5) printf("mismatch C[%d][%d] = %d, ref %d\n", i, j, C[i * 128 + j], ref);
        ++bad;
      }
    }
  printf("cta_group::2 correctness: %ld mismatches of %d\n", bad, 256 * 128);

  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int ctas = (p.multiProcessorCount / 2) * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int reps : {2000, 8000}) {
    cudaEventRecord(e0);
    i8_gemm2_kernel<<<ctas, 128, smem>>>(dA, dB, ksteps, reps, dC);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 256 * 128 * K * (double)reps * (ctas / 2);
    printf("cta_group::2 throughput reps=%d: %.3f ms  %.1f TOPS on %d SMs (int8 dense nominal 4500; cta_group::1 measured 3838)\n", reps, ms,
           ops / ms * 1e-9, ctas);
  }
  return 0;
}
