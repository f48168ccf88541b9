// Shared helpers: error plumbing, PTX wrappers (mbarrier, 1-D bulk TMA copy, fp64 DMMA).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <atomic>

namespace tb {

// ---- thread-local error string (C-ABI convention, include/trieste_b200.h) -------------------
inline std::string& last_error() {
  thread_local std::string e;
  return e;
}
// status codes of the C-ABI (include/trieste_b200.h, enum tb_status)
constexpr int ERR_INVALID = 1;   // bad argument / precondition: the Python layer raises ValueError (the reference's InvalidArgumentError)
constexpr int ERR_RUNTIME = 2;   // CUDA / library failure: NativeLibraryError (there is no CPU fallback to retry on)
constexpr int ERR_NUMERIC = 3;   // a factorisation met a non-positive-definite matrix: ValueError, as tf.linalg.cholesky's InvalidArgumentError
inline int fail(const std::string& msg, int code = ERR_INVALID) {
  last_error() = msg;
  return code;
}
#define TB_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::tb::fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")", ::tb::ERR_RUNTIME);            \
  } while (0)
#define TB_CHECK(cond, msg)              \
  do {                                   \
    if (!(cond)) return ::tb::fail(msg); \
  } while (0)
#define TB_CHECK_CODE(cond, msg, code)         \
  do {                                         \
    if (!(cond)) return ::tb::fail(msg, code); \
  } while (0)
#define TB_TRY(expr)        \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

inline std::atomic<int64_t>& launch_counter() {
  static std::atomic<int64_t> c{0};
  return c;
}
#define TB_LAUNCHED() (::tb::launch_counter().fetch_add(1, std::memory_order_relaxed))

// ---- tile geometry of the packed panels -------------------------------------------------------
// A "panel" is a 128 x 16 block of an operand, stored in DMMA-fragment order so that one warp-wide
// 16-byte load yields the fragments of two consecutive m8n8k4 steps:
//   [g8 = 16 groups of 8 rows/cols][p = 2 k-pairs][lane = 32][s = 2]   (2048 elements)
//   lane l <-> (row-or-col within the group = l / 4, k within the k4 step = l % 4); k = (2p+s)*4 + l%4
constexpr int BM = 128;          // rows of Linv per row-block
constexpr int BT = 128;          // candidates per tile
constexpr int BK = 16;           // k depth of one panel
constexpr int PANEL = BM * BK;   // elements per panel (2048)

__host__ __device__ inline int64_t rowblock_panel_offset(int I) {  // panels before row-block I
  return (int64_t)(BM / BK / 2) * I * (I + 1);                      // sum_{i<I} 8 (i+1) = 4 I (I+1)
}
__host__ __device__ inline int panel_elem_index(int r, int k) {  // r in [0,128), k in [0,16)
  int g8 = r >> 3, rr = r & 7, k4 = k >> 2, kq = k & 3;
  int p = k4 >> 1, s = k4 & 1, lane = rr * 4 + kq;
  return (((g8 * 2 + p) * 32 + lane) << 1) + s;
}

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// D(8x8) += A(8x4) * B(4x8), fp64 tensor-core MMA (SASS DMMA).
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

}  // namespace tb
