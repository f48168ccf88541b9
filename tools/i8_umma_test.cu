// Stage-0 experiment for the fp64-emulation engine (Ozaki splitting on the int8 tensor cores):
// one CTA computes C[128 x 128] (s32) = A[128 x K] (s8, K-major) * B[128 x K]^T (s8, K-major) with
// tcgen05.mma kind::i8, operands pre-packed in the no-swizzle UMMA core-matrix layout and moved with
// 1-D bulk copies, accumulator in TMEM, read back with tcgen05.ld.  Checks against a CPU reference and
// measures the sustained MMA rate over smem-resident operands.
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 B contiguous); LBO = stride between core matrices
// along K, SBO = stride between 8-row groups
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (sm100)
  return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset = 0
}
constexpr uint32_t IDESC_S8_128x128 = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int KSTAGE = 64;                 // K bytes per smem stage
constexpr int TILE_BYTES = 128 * KSTAGE;   // 8 KB
constexpr uint32_t LBO = 128, SBO = (KSTAGE / 16) * 128;

// grid = 1 (correctness) or many (throughput: every CTA repeats the same tile `reps` times)
__global__ void __launch_bounds__(128, 1)
i8_gemm_kernel(const int8_t* __restrict__ Ap, const int8_t* __restrict__ Bp, int ksteps, int reps, int32_t* __restrict__ C) {
  extern __shared__ __align__(1024) unsigned char smem[];
  int8_t* sA = reinterpret_cast<int8_t*>(smem);                   // [ksteps][TILE_BYTES]
  int8_t* sB = sA + (size_t)ksteps * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)ksteps * TILE_BYTES);  // [0]=loaded, [1]=mma done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], 2u * ksteps * TILE_BYTES);
    for (int k = 0; k < ksteps; ++k) {
      bulk_g2s(sA + (size_t)k * TILE_BYTES, Ap + (size_t)k * TILE_BYTES, TILE_BYTES, &bars[0]);
      bulk_g2s(sB + (size_t)k * TILE_BYTES, Bp + (size_t)k * TILE_BYTES, TILE_BYTES, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int r = 0; r < reps; ++r) {
      for (int k = 0; k < ksteps; ++k) {
        const uint32_t a0 = smem_u32(sA + (size_t)k * TILE_BYTES), b0 = smem_u32(sB + (size_t)k * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < KSTAGE / 32; ++kk) {  // one MMA consumes K = 32 bytes = two 16-byte core columns
          uint64_t da = make_desc(a0 + kk * 2 * LBO, LBO, SBO);
          uint64_t db = make_desc(b0 + kk * 2 * LBO, LBO, SBO);
          umma_i8(tmem, da, db, IDESC_S8_128x128, (r | k | kk) ? 1u : 0u);
        }
      }
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (rows), 128 columns in 4 chunks of 32
  if (blockIdx.x == 0) {
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
      const int row = warp * 32 + lane;
      for (int j = 0; j < 32; ++j) C[row * 128 + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

// host-side packing: row-major [128][K] -> [kstep][rowgroup 16][kchunk 4][row 8][16 bytes]
static void pack(const std::vector<int8_t>& M, int K, std::vector<int8_t>& P) {
  P.assign((size_t)128 * K, 0);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < K; ++k) {
      int ks = k / KSTAGE, kin = k % KSTAGE;
      size_t off = (size_t)ks * TILE_BYTES + (size_t)(r / 8) * SBO + (size_t)(kin / 16) * LBO + (r % 8) * 16 + (kin % 16);
      P[off] = M[(size_t)r * K + k];
    }
}

int main() {
  const int K = 256, ksteps = K / KSTAGE;
  std::vector<int8_t> A((size_t)128 * K), B((size_t)128 * K), Ap, Bp;
  srand(1);
  for (auto& x : A) x = (int8_t)(rand() % 129 - 64);
  for (auto& x : B) x = (int8_t)(rand() % 129 - 64);
  pack(A, K, Ap);
  pack(B, K, Bp);
  int8_t *dA, *dB;
  int32_t* dC;
  CK(cudaMalloc(&dA, Ap.size()));
  CK(cudaMalloc(&dB, Bp.size()));
  CK(cudaMalloc(&dC, 128 * 128 * 4));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, 128 * 128 * 4));
  const size_t smem = (size_t)2 * ksteps * TILE_BYTES + 64;
  CK(cudaFuncSetAttribute(i8_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  i8_gemm_kernel<<<1, 128, smem>>>(dA, dB, ksteps, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> C(128 * 128);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)i * K + k] * (int32_t)B[(size_t)j * K + k];
      if (ref != C[i * 128 + j]) {
        if (bad < 5) printf("mismatch C[%d][%d] = %d, ref %d\n", i, j, C[i * 128 + j], ref);
        ++bad;
      }
    }
  printf("correctness: %ld mismatches of %d\n", bad, 128 * 128);

  // throughput: every SM repeats the K=256 tile product `reps` times from shared memory
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int reps : {2000, 8000}) {
    cudaEventRecord(e0);
    i8_gemm_kernel<<<p.multiProcessorCount, 128, smem>>>(dA, dB, ksteps, reps, dC);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = 2.0 * 128 * 128 * K * (double)reps * p.multiProcessorCount;
    printf("throughput reps=%d: %.3f ms  %.1f TOPS (int8 dense nominal 4500)\n", reps, ms, ops / ms * 1e-9);
  }
  return 0;
}
