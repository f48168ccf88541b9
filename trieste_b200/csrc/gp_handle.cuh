// The model handle: device-resident posterior cache + scratch, and its once-per-step precompute.
#pragma once
#include "common.cuh"
#include "../../include/trieste_b200.h"
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace tb {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    TB_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

inline bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// supported padded input dimensions of the distance loop (even, so rows load as double2)
inline int pick_dp(int D) {
  static const int opts[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32};
  for (int o : opts)
    if (D <= o) return o;
  return -1;
}

}  // namespace tb

struct tb_gp {
  int device = 0;
  int dtype = TB_F64;
  cudaStream_t stream = nullptr;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;

  // model (host copies of the small things)
  int64_t N = 0;
  int D = 0, DP = 0;
  int kernel = TB_MATERN52;
  double variance = 1.0, noise = 1.0, mean_const = 0.0;
  std::vector<double> ls;  // [D]
  bool have_data = false, have_hyper = false, cache_valid = false;

  // geometry of the packed cache
  int nkc = 0;  // k panels = ceil(N/16)
  int NB = 0;   // row-blocks = ceil(N/128)

  tb::DevBuf dX, dy;            // raw data [N,D], [N]
  tb::DevBuf dXs;               // [nkc*16][DP] scaled, zero padded
  tb::DevBuf dInvLs;            // [DP]
  tb::DevBuf dAlpha;            // [nkc*16]
  tb::DevBuf dL;                // [N,N] column-major lower Cholesky factor
  tb::DevBuf dLinv;             // [N,N] column-major Linv (kept: predict_joint / gradients reuse it)
  tb::DevBuf dLinvP;            // packed lower panels
  tb::DevBuf dLinvTP;           // packed upper panels of Linv^T (lazy; gradient path)
  bool upper_valid = false;
  // int8 (Ozaki) engine: digit tiles of Linv, per-row scales, K* scale
  int engine = 1;  // 0 = fp64 DMMA, 1 = int8 tensor cores (default; same stated tolerances, ~3x faster)
  tb::DevBuf dAS, dRowScale;
  tb::DevBuf dKinv, dKinvS, dKinvScale;  // gradient path of the int8 engine: digit tiles of K^-1 (full rows)
  bool kinv_valid = false;       // digit tiles of K^-1 current
  bool kinv_dense_valid = false; // dense K^-1 (dKinv, lower triangle, ld = kinv_dense_N) current: kept so that an append can
  int64_t kinv_dense_N = 0;      // update it by rank m (tb_gp_append_data) instead of rebuilding it in O(N^3)
  tb::DevBuf dKinvSpare;
  tb::DevBuf sKs2, sMean2, sPartial2;  // second scratch slot of the pipelined driver
  tb::DevBuf sMeanPart;                // per-split mean partials of the k-split K* generation (few candidate tiles)
  cudaStream_t stream2 = nullptr;      // K* digit generation stream (overlaps the digit GEMM)
  cudaEvent_t evK[2] = {nullptr, nullptr}, evDone[2] = {nullptr, nullptr};
  bool oz_valid = false;
  int nst = 0, oz_bscale_exp = 0;
  double oz_out_scale = 1.0;
  // single-pass digit engine (ozaki5.cuh): tight row scales + row sums + S-digit tiles of Linv; mode = digits per operand
  // (5: fp64 handles, 15 products; 3: fp32 handles, 6 products; 0: not eligible -> the 6-digit / 21-product kernels)
  tb::DevBuf dAS5, dRowScale5, dRowSum5, dX2;
  bool oz5_valid = false;
  bool oz_full = false;  // tb_gp_set_engine(2): always the 6-digit / 21-product kernels
  int oz5_mode = 0;
  int oz5_planes = 0;    // digit planes stored per operand stage (5: fp64 handles; 4: fp32 handles, whose variance GEMM computes
                         // with the 3 leading planes and whose store-A / V GEMMs use all 4)
  double oz5_est = 0.0;  // a-priori estimate of max |Δvar| / σ_f² in the chosen mode
  tb::DevBuf dKinvS5, dKinvScale5, dKinvSum5;  // tight digit tiles / row scales / row sums of the dense K^-1 (gradient path)
  bool kinv5_valid = false, kinv5_ok = false;  // kinv5_ok: the V GEMM's own error estimate admits the single-pass engine
  tb::DevBuf dWork, dInfo;      // cusolver workspace / info flag
  tb::DevBuf dDinv;             // inverses of the diagonal blocks of L (hand-written factorisation)
  bool factor_own = true;       // false (TB_FACTOR=cusolver): cuSOLVER / cuBLAS cross-check path

  // per-call scratch
  tb::DevBuf sKs, sPartial, sMean, sVals, sVar, sXc, sBlkBest, sBlkIdx, sRun;
  tb::DevBuf sA, sV, sGrad, sMisc;  // A / V panels (joint + gradient paths), misc staging
  tb::DevBuf dXspare, dyspare, dLspare, dLinvSpare;  // ping-pong partners of dX / dy / dL / dLinv (tb_gp_append_data)
  tb::DevBuf dMes;              // min-value samples of TB_ACQ_MES (tb_acq_set_min_value_samples)
  int mesS = 0;

  // profiling of the dominant kernel
  bool profile = false;
  double prof_ms = 0.0, prof_flops = 0.0;
  int64_t prof_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<double> prof_event_flops;
};
