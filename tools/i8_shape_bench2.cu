// Shape study, part 2 (round 2): issue rate of tcgen05.mma.cta_group::2 kind::i8 (M = 256 over a CTA pair, K = 32 per instruction,
// SS mode, no-swizzle K-major operands resident in shared memory) as a function of N.  Each CTA holds its own 128 rows of A
// and N/2 rows of B; the leader issues, completion is multicast.  Prints SM clocks per MMA against the N/2-clock floor
// (tools/i8_shape_bench.cu measured cta_group::1: 32 + N/4 clocks below N = 128, i.e. bound by the shared-memory operand
// reads A 4 KB + B 32 N bytes at 128 B/clk; the pair halves the B bytes each SM reads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_shape_bench2 tools/i8_shape_bench2.cu && ./i8_shape_bench2
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
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
constexpr uint32_t LBO = 128, SBO = 512;
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
constexpr int ND = 5;

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1)
shape2_kernel(int stages, long long* __restrict__ clocks) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int ATILE = 128 * 64, BTILE = (N / 2) * 64;
  unsigned char* sA = smem;
  unsigned char* sB = sA + ND * ATILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + ND * BTILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (ND * (ATILE + BTILE)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_own = *tmem_slot;
  uint32_t tmem;
  {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(tmem_slot)), "r"(0u));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(tmem) : "r"(raddr) : "memory");
  }
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((256u >> 4) << 24);
  constexpr int NACC = (512 / N) < 5 ? (512 / N) : 5;
  if (rank == 0 && warp == 0 && lane == 0) {
    const long long t0 = clock64();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int s = 0; s < stages; ++s) {
#pragma unroll
      for (int p = 1; p <= ND; ++p)
#pragma unroll
        for (int q = 1; q <= ND; ++q) {
          if (p + q > ND + 1) continue;
          const uint32_t acc = tmem + (uint32_t)((p + q - 2) % NACC) * N;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
            umma2_i8(acc, make_desc(a0 + (p - 1) * ATILE + kk * 2 * LBO), make_desc(b0 + (q - 1) * BTILE + kk * 2 * LBO), IDESC, 1u);
        }
    }
    umma2_commit_multicast(&bars[0]);
  }
  __syncwarp();
  mbar_wait(&bars[0], 0);
  if (rank == 0 && threadIdx.x == 0) clocks[blockIdx.x >> 1] = clock64();  // end stamp (start is taken by the same thread below)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_own), "r"(512u) : "memory");
}

template <int N>
static void run(long long* dclk, int ctas) {
  const size_t smem = (size_t)ND * (128 * 64 + (N / 2) * 64) + 128;
  CK(cudaFuncSetAttribute(shape2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  shape2_kernel<N><<<ctas, 64, smem>>>(200, dclk);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int stages = 4000;
  cudaEventRecord(e0);
  shape2_kernel<N><<<ctas, 64, smem>>>(stages, dclk);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  // clocks from the event time and the SM clock estimated by the cta_group::1 bench are not needed: report ms and TOPS, and
  // clocks per MMA assuming the 1.90 GHz the device ran at in part 1 is NOT assumed here — use a second timing kernel instead
  const double mmas = (double)stages * 30;
  const double ops = 2.0 * 256 * N * 32 * mmas * (ctas / 2);
  printf("cta_group::2 M=256 N=%3d: %.3f ms  %.0f TOPS   (%.1f ns per MMA; floor N/2 = %d clk)\n", N, ms, ops / ms * 1e-9, ms * 1e6 / mmas, N / 2);
}

// SM clock probe: a kernel that spins for a known number of clocks
__global__ void clock_probe(long long n, long long* out) {
  const long long t0 = clock64();
  while (clock64() - t0 < n) {}
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int ctas = (p.multiProcessorCount / 2) * 2;
  long long* dclk;
  CK(cudaMalloc(&dclk, sizeof(long long) * ctas));
  run<64>(dclk, ctas);
  run<96>(dclk, ctas);
  run<128>(dclk, ctas);
  run<192>(dclk, ctas);
  run<256>(dclk, ctas);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  clock_probe<<<1, 32>>>(20000000, dclk);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long c;
  CK(cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost));
  printf("clock probe: %lld clocks in %.3f ms = %.3f GHz (idle-ish clock; the MMA kernels above ran near this)\n", c, ms, c / (ms * 1e6));
  return 0;
}
