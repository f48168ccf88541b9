// Shape study for the digit GEMM (round 2): sustained tcgen05.mma kind::i8 issue rate of ONE CTA per SM as a function of the
// instruction's N (M = 128, K = 32 per instruction, SS mode, no-swizzle K-major operands resident in shared memory), with and
// without concurrent bulk-TMA fill traffic into a separate shared-memory ring (the digit GEMM streams 35-60 B/clk/SM from L2).
// Prints SM clocks per MMA against the N/2-clock floor, so the result does not depend on the power-capped clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_shape_bench tools/i8_shape_bench.cu && ./i8_shape_bench
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
constexpr uint32_t LBO = 128, SBO = 512;  // K stage of 64 bytes: 4 core matrices along K, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int ND = 5;  // digit tiles per operand in a stage (15 products with p + q <= 6)
constexpr int FILL_CHUNK = 16384;

// warp 0 lane 0: MMA issue; warp 1 lane 0: background fill (fill_rate bytes per clock, 0 = none)
template <int N>
__global__ void __launch_bounds__(64, 1)
shape_kernel(const int8_t* __restrict__ src, int stages, float fill_rate, long long* __restrict__ clocks) {
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int ATILE = 128 * 64, BTILE = N * 64;
  unsigned char* sA = smem;
  unsigned char* sB = sA + ND * ATILE;
  unsigned char* sF = sB + ND * BTILE;  // fill ring: 2 chunks
  uint64_t* bars = reinterpret_cast<uint64_t*>(sF + 2 * FILL_CHUNK);  // [0] done, [1..2] fill
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(bars + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    *stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // operands: whatever bytes are in shared memory (timing does not depend on values); zero them for determinism
  for (int i = threadIdx.x; i < (ND * (ATILE + BTILE)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  constexpr int NACC = (512 / N) < 5 ? (512 / N) : 5;

  if (warp == 0 && lane == 0) {
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
            umma_i8(acc, make_desc(a0 + (p - 1) * ATILE + kk * 2 * LBO), make_desc(b0 + (q - 1) * BTILE + kk * 2 * LBO), IDESC, 1u);
        }
    }
    umma_commit(&bars[0]);
    mbar_wait(&bars[0], 0);
    const long long t1 = clock64();
    *stop = 1;
    clocks[blockIdx.x] = t1 - t0;
  } else if (warp == 1 && lane == 0 && fill_rate > 0.f) {
    const unsigned char* g = reinterpret_cast<const unsigned char*>(src) + (size_t)blockIdx.x * (1 << 19);
    long long next = clock64();
    uint32_t ph[2] = {0, 0};
    const long long gap = (long long)(FILL_CHUNK / fill_rate);
    int i = 0;
    long long issued = 0;
    while (!*stop) {
      const int b = i & 1;
      if (i >= 2) { mbar_wait(&bars[1 + b], ph[b]); ph[b] ^= 1; }
      while (clock64() < next) {}
      next += gap;
      mbar_expect_tx(&bars[1 + b], FILL_CHUNK);
      bulk_g2s(sF + b * FILL_CHUNK, g + (size_t)((i * FILL_CHUNK) & ((1 << 19) - 1)), FILL_CHUNK, &bars[1 + b]);
      ++i;
      ++issued;
    }
    // drain: the last two chunks are still in flight
    if (i >= 1) { const int b = (i - 1) & 1; mbar_wait(&bars[1 + b], ph[b]); }
    if (i >= 2) { const int b = (i - 2) & 1; mbar_wait(&bars[1 + b], ph[b]); }
    clocks[gridDim.x + blockIdx.x] = issued * FILL_CHUNK;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N>
static void run(const int8_t* src, long long* dclk, int sms, float fill_rate) {
  const size_t smem = (size_t)ND * (128 * 64 + N * 64) + 2 * FILL_CHUNK + 128;
  CK(cudaFuncSetAttribute(shape_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int stages = 4000;
  CK(cudaMemset(dclk, 0, sizeof(long long) * 2 * sms));
  shape_kernel<N><<<sms, 64, smem>>>(src, 200, fill_rate, dclk);  // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  shape_kernel<N><<<sms, 64, smem>>>(src, stages, fill_rate, dclk);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(2 * sms);
  CK(cudaMemcpy(h.data(), dclk, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
  double mean = 0, mx = 0, fb = 0;
  for (int i = 0; i < sms; ++i) { mean += h[i]; mx = h[i] > mx ? h[i] : mx; fb += h[sms + i]; }
  mean /= sms;
  const double mmas = (double)stages * 30;
  const double ops = 2.0 * 128 * N * 32 * mmas * sms;
  printf("N=%3d fill_target=%5.1f B/clk  clk/MMA mean %.2f max %.2f  floor %.1f  eff %.3f | %.3f ms %.0f TOPS (%.2f GHz) fill achieved %.1f B/clk/SM\n", N,
         fill_rate, mean / mmas, mx / mmas, N / 2.0, (N / 2.0) / (mean / mmas), ms, ops / ms * 1e-9, mean / (ms * 1e6), fb / sms / mean);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  int8_t* src;
  long long* dclk;
  CK(cudaMalloc(&src, (size_t)sms << 19));
  CK(cudaMemset(src, 1, (size_t)sms << 19));
  CK(cudaMalloc(&dclk, sizeof(long long) * 2 * sms));
  for (float fr : {0.f, 32.f, 48.f, 64.f}) {
    run<64>(src, dclk, sms, fr);
    run<96>(src, dclk, sms, fr);
    run<128>(src, dclk, sms, fr);
    run<192>(src, dclk, sms, fr);
    run<256>(src, dclk, sms, fr);
  }
  return 0;
}
