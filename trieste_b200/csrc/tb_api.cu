// C-ABI entry points (include/trieste_b200.h).  Host orchestration only: every per-candidate flop
// runs in the kernels of kernels_f64.cuh / kernels_extra.cuh.
#include "gp_handle.cuh"
#include "kernels_extra.cuh"
#include "ozaki.cuh"
#include "oz5_api.h"
#include <chrono>
#include "factor.cuh"
#include "lbfgs.cuh"

using namespace tb;
namespace tb {
int kernels_init();
}

// =================================================================================================
// once-per-step precompute kernels (posterior cache)
// =================================================================================================
namespace tb {

// Xs[k][d] = X[k][d] / l_d, zero padded to [nkc*16][DP]
__global__ void scale_inputs_kernel(const double* __restrict__ X, const double* __restrict__ inv_ls,
                                    int64_t N, int D, int DP, int64_t rows, double* __restrict__ Xs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * DP) return;
  int64_t k = i / DP;
  int d = (int)(i % DP);
  Xs[i] = (k < N && d < D) ? X[k * D + d] * inv_ls[d] : 0.0;
}

// K(X,X) + noise I, full symmetric, column-major [N,N]
template <int KIND>
__global__ void gram_kernel(const double* __restrict__ Xs, int64_t N, int DP, double variance,
                            double noise, double* __restrict__ K) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i >= N) return;
  double r2 = 0.0;
  for (int d = 0; d < DP; ++d) {
    double df = Xs[i * DP + d] - Xs[j * DP + d];
    r2 = fma(df, df, r2);
  }
  double v = kernel_from_r2<KIND>(r2, variance);
  if (i == j) v = variance + noise;
  K[i + j * N] = v;
}

// ---- rank-m append of training points to an existing cache (tb_gp_append_data) ----
// W[:, j] (column-major [N, m]) = k(x_i, xnew_j) for every row i of the grown data set; the diagonal entry of the new
// block carries the likelihood noise
template <int KIND>
__global__ void append_cross_kernel(const double* __restrict__ Xs, int DP, int64_t N0, int64_t N, double variance,
                                    double noise, double* __restrict__ W) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t j = blockIdx.y;
  if (i >= N) return;
  double r2 = 0.0;
  for (int d = 0; d < DP; ++d) {
    const double df = Xs[i * DP + d] - Xs[(N0 + j) * DP + d];
    r2 = fma(df, df, r2);
  }
  double v = kernel_from_r2<KIND>(r2, variance);
  if (i == N0 + j) v = variance + noise;
  W[i + j * N] = v;
}

// Y[:, j] = T[0:n, 0:n] X[:, j] for a column-major lower-triangular T with leading dimension ld (one thread per row:
// coalesced along the rows of a column)
__global__ void trmv_lower_cols_kernel(const double* __restrict__ T, int64_t n, int64_t ld, const double* __restrict__ X,
                                       int64_t ldx, double* __restrict__ Y, int64_t ldy) {
  __shared__ double xs[128];
  const int64_t r = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t j = blockIdx.y;
  const int64_t kend = min(n, ((int64_t)blockIdx.x + 1) * 128);
  double s = 0.0;
  for (int64_t k0 = 0; k0 < kend; k0 += 128) {
    __syncthreads();
    xs[threadIdx.x] = (k0 + threadIdx.x < n) ? X[k0 + threadIdx.x + j * ldx] : 0.0;
    __syncthreads();
    if (r < n) {
      const int64_t kk_end = min((int64_t)128, r - k0 + 1);
      for (int64_t kk = 0; kk < kk_end; ++kk) s = fma(T[r + (k0 + kk) * ld], xs[kk], s);
    }
  }
  if (r < n) Y[r + j * ldy] = s;
}

// U[:, j] = T[0:n, 0:n]^T X[:, j] (one warp per column of T)
__global__ void trmv_lower_t_cols_kernel(const double* __restrict__ T, int64_t n, int64_t ld, const double* __restrict__ X,
                                         int64_t ldx, double* __restrict__ U, int64_t ldu) {
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int64_t j = blockIdx.y;
  if (k >= n) return;
  double s = 0.0;
  for (int64_t r = k + lane; r < n; r += 32) s = fma(T[r + k * ld], X[r + j * ldx], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) U[k + j * ldu] = s;
}

// S[i + j*m] = C[i][j] - l_i . l_j (Schur complement of the new block; l_i = Y[:, i], C = W[N0:, :])
__global__ void append_schur_kernel(const double* __restrict__ Y, const double* __restrict__ W, int64_t N0, int64_t N, int m,
                                    double* __restrict__ S) {
  __shared__ double red[8];
  const int i = blockIdx.x, j = blockIdx.y;
  double s = 0.0;
  for (int64_t k = threadIdx.x; k < N0; k += blockDim.x) s = fma(Y[k + (int64_t)i * N], Y[k + (int64_t)j * N], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    S[i + j * m] = W[N0 + i + (int64_t)j * N] - t;
  }
}

constexpr int APPEND_MAX = 64;  // larger appends refactorise from scratch

// one CTA of APPEND_MAX threads: L22 = chol(S) in shared memory, R = L22^-1; writes both into the new corner of L / Linv
// and R (row-major [m, m]) to Rout; info = failing leading minor (global index) if S is not positive definite
__global__ void append_chol_kernel(const double* __restrict__ S, int m, int64_t N0, int64_t N, double* __restrict__ L,
                                   double* __restrict__ Linv, double* __restrict__ Rout, int* __restrict__ info) {
  __shared__ double A[APPEND_MAX][APPEND_MAX + 1];
  __shared__ int bad;
  const int t = threadIdx.x;
  if (t == 0) bad = 0;
  for (int j = 0; j < m; ++j)
    if (t < m) A[t][j] = S[t + j * m];
  __syncthreads();
  for (int j = 0; j < m; ++j) {
    if (t == j) {
      const double d = A[j][j];
      if (!(d > 0.0)) {
        if (!bad) bad = (int)(N0 + j + 1);
        A[j][j] = 1.0;
      } else {
        A[j][j] = sqrt(d);
      }
    }
    __syncthreads();
    if (t > j && t < m) A[t][j] /= A[j][j];
    __syncthreads();
    if (t > j && t < m)
      for (int k = j + 1; k <= t; ++k) A[t][k] -= A[t][j] * A[k][j];
    __syncthreads();
  }
  // column t of R: forward substitution of L22 r = e_t
  if (t < m) {
    for (int i = 0; i < m; ++i) {
      double s = (i == t) ? 1.0 : 0.0;
      for (int k = t; k < i; ++k) s -= A[i][k] * Rout[k * m + t];
      Rout[i * m + t] = (i < t) ? 0.0 : s / A[i][i];  // column t is private to this thread
    }
  }
  __syncthreads();
  if (t == 0 && bad) *info = bad;
  if (t < m)
    for (int j = 0; j < m; ++j) {
      L[(N0 + t) + (N0 + j) * N] = (j <= t) ? A[t][j] : 0.0;
      Linv[(N0 + t) + (N0 + j) * N] = (j <= t) ? Rout[t * m + j] : 0.0;
    }
}

// new rows left of the corner: L[N0+i, k] = Y[k, i] (Y = Linv0 B, so L21 = Y^T) and Linv[N0+i, k] = -(R U^T)[i, k] with
// U = Linv0^T Y (Linv21 = -L22^-1 L21 Linv0)
__global__ void append_rows_kernel(const double* __restrict__ Y, const double* __restrict__ U, const double* __restrict__ R,
                                   int m, int64_t N0, int64_t N, double* __restrict__ L, double* __restrict__ Linv) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (k >= N0) return;
  double b = 0.0;
  for (int j = 0; j <= i; ++j) b = fma(R[i * m + j], U[k + (int64_t)j * N], b);
  L[(N0 + i) + k * N] = Y[k + (int64_t)i * N];
  Linv[(N0 + i) + k * N] = -b;
}

__global__ void identity_kernel(int64_t N, double* __restrict__ A) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * N) return;
  A[i] = (i % N == i / N) ? 1.0 : 0.0;
}

__global__ void zero_upper_kernel(int64_t N, double* __restrict__ A) {  // column-major: zero i < j
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i < N && i < j) A[i + j * N] = 0.0;
}

__global__ void residual_kernel(const double* __restrict__ y, int64_t N, int64_t rows, double mean_const,
                                double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) out[i] = (i < N) ? y[i] - mean_const : 0.0;
}

// one CTA per (row-block I, k-panel kc) of the lower triangle: Linv (column-major) -> packed panel
__global__ void pack_lower_panels_kernel(const double* __restrict__ Linv, int64_t N, int nkc,
                                         double* __restrict__ P) {
  const int I = blockIdx.y, kc = blockIdx.x;
  const int nk = min((I + 1) * (BM / BK), nkc);
  if (kc >= nk) return;
  double* dst = P + (rowblock_panel_offset(I) + kc) * PANEL;
  for (int e = threadIdx.x; e < PANEL; e += blockDim.x) {
    // iterate in source-friendly order: r fastest (column-major source), scatter into the panel
    int r = e % BM, k = e / BM;
    int64_t n = (int64_t)I * BM + r, kk = (int64_t)kc * BK + k;
    double v = (n < N && kk <= n) ? Linv[n + kk * N] : 0.0;
    dst[panel_elem_index(r, k)] = v;
  }
}

// column-major lower factor -> row-major dense (upper zeroed), for tb_gp_get_cholesky
__global__ void colmajor_lower_to_rowmajor_kernel(const double* __restrict__ A, int64_t N,
                                                  double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t j = blockIdx.y;
  if (i < N) out[i * N + j] = (j <= i) ? A[i + j * N] : 0.0;
}

}  // namespace tb

#define TB_CUSOLVER(expr)                                                                          \
  do {                                                                                             \
    cusolverStatus_t _s = (expr);                                                                  \
    if (_s != CUSOLVER_STATUS_SUCCESS) return tb::fail(std::string(#expr) + ": cusolver status " + \
                                                       std::to_string((int)_s), tb::ERR_RUNTIME);  \
  } while (0)
#define TB_CUBLAS(expr)                                                                        \
  do {                                                                                         \
    cublasStatus_t _s = (expr);                                                                \
    if (_s != CUBLAS_STATUS_SUCCESS) return tb::fail(std::string(#expr) + ": cublas status " + \
                                                     std::to_string((int)_s), tb::ERR_RUNTIME);\
  } while (0)

static const char* kVersion = "trieste_b200 0.1 (sm_100a; fp64 DMMA triangular GEMM)";

extern "C" {

const char* tb_last_error(void) { return tb::last_error().c_str(); }
const char* tb_version(void) { return kVersion; }
int tb_device_count(int* count) {
  TB_CHECK(count != nullptr, "tb_device_count: null output");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
    return tb::fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e), tb::ERR_RUNTIME);
  }
  return 0;
}
int64_t tb_launch_count(void) { return tb::launch_counter().load(); }
void tb_launch_count_reset(void) { tb::launch_counter().store(0); }

int tb_gp_create(tb_gp** out, int device, int dtype) {
  TB_CHECK(out != nullptr, "tb_gp_create: null output");
  TB_CHECK(dtype == TB_F64 || dtype == TB_F32, "tb_gp_create: dtype must be TB_F64 or TB_F32");
  int n = 0;
  TB_TRY(tb_device_count(&n));
  TB_CHECK_CODE(n > 0, "tb_gp_create: no CUDA device visible (this library has no CPU fallback)", tb::ERR_RUNTIME);
  TB_CHECK(device >= 0 && device < n, "tb_gp_create: device index out of range");
  TB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TB_CUDA(cudaGetDeviceProperties(&prop, device));
  TB_CHECK(prop.major == 10, "tb_gp_create: this library is built for sm_100a (B200) only; found sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor));
  tb_gp* gp = new tb_gp();
  gp->device = device;
  gp->dtype = dtype;
  if (const char* e = std::getenv("TB_FACTOR")) gp->factor_own = std::string(e) != "cusolver";
  if (const char* e = std::getenv("TB_ENGINE")) gp->engine = (std::string(e) == "fp64") ? 0 : 1;
  {
    // main stream at the highest priority: when the K*-generation stream (lowest) runs concurrently, GEMM CTAs are placed
    // first and the generation CTAs only fill the register / thread slots a GEMM CTA leaves free
    int least = 0, greatest = 0;
    TB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    TB_CUDA(cudaStreamCreateWithPriority(&gp->stream, cudaStreamNonBlocking, greatest));
  }
  TB_TRY(tb::kernels_init());
  *out = gp;
  return 0;
}

int tb_gp_destroy(tb_gp* gp) {
  if (!gp) return 0;
  cudaSetDevice(gp->device);
  cudaStreamSynchronize(gp->stream);
  for (tb::DevBuf* b : {&gp->dX, &gp->dy, &gp->dXs, &gp->dInvLs, &gp->dAlpha, &gp->dL, &gp->dLinv,
                        &gp->dLinvP, &gp->dLinvTP, &gp->dAS, &gp->dRowScale, &gp->dKinv, &gp->dKinvS, &gp->dKinvScale, &gp->dDinv, &gp->sKs2, &gp->sMean2, &gp->sPartial2, &gp->dWork, &gp->dInfo, &gp->sKs, &gp->sPartial, &gp->sMean,
                        &gp->sVals, &gp->sVar, &gp->sXc, &gp->sBlkBest, &gp->sBlkIdx, &gp->sRun,
                        &gp->sA, &gp->sV, &gp->sGrad, &gp->sMisc, &gp->dMes, &gp->dXspare, &gp->dyspare, &gp->dLspare, &gp->dLinvSpare,
                        &gp->dAS5, &gp->dRowScale5, &gp->dRowSum5, &gp->dX2, &gp->dKinvS5, &gp->dKinvScale5, &gp->dKinvSum5,
                        &gp->dKinvSpare, &gp->sMeanPart})
    b->release();
  for (auto& ev : gp->prof_events) {
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  if (gp->stream2) {
    cudaStreamSynchronize(gp->stream2);
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(gp->evK[i]);
      cudaEventDestroy(gp->evDone[i]);
    }
    cudaStreamDestroy(gp->stream2);
  }
  if (gp->cusolver) cusolverDnDestroy(gp->cusolver);
  if (gp->cublas) cublasDestroy(gp->cublas);
  if (gp->stream) cudaStreamDestroy(gp->stream);
  delete gp;
  return 0;
}

static int tb_gp_set_data_f64(tb_gp* gp, const void* X, const void* y, int64_t N, int D) {
  TB_CHECK(gp && X && y, "tb_gp_set_data: null argument");
  TB_CHECK(N > 0, "tb_gp_set_data: dataset must be populated (N > 0)");
  TB_CHECK(D > 0 && tb::pick_dp(D) > 0, "tb_gp_set_data: input dimension must be in [1, 32]");
  TB_CHECK(N <= 65535, "tb_gp_set_data: N > 65535 is not supported");  // grid.y of the per-column kernels
  TB_CUDA(cudaSetDevice(gp->device));
  gp->N = N;
  gp->D = D;
  gp->DP = tb::pick_dp(D);
  gp->nkc = (int)((N + BK - 1) / BK);
  gp->NB = (int)((N + BM - 1) / BM);
  TB_TRY(gp->dX.reserve(sizeof(double) * N * D));
  TB_TRY(gp->dy.reserve(sizeof(double) * N));
  TB_CUDA(cudaMemcpyAsync(gp->dX.p, X, sizeof(double) * N * D, cudaMemcpyDefault, gp->stream));
  TB_CUDA(cudaMemcpyAsync(gp->dy.p, y, sizeof(double) * N, cudaMemcpyDefault, gp->stream));
  TB_CUDA(cudaStreamSynchronize(gp->stream));
  gp->have_data = true;
  gp->cache_valid = false;
  if ((int)gp->ls.size() != D && gp->ls.size() == 1) gp->ls.assign(D, gp->ls[0]);
  return 0;
}

int tb_gp_set_hyper(tb_gp* gp, int kernel, double variance, const double* lengthscales, int n_ls,
                    double noise_variance, double mean_const) {
  TB_CHECK(gp && lengthscales, "tb_gp_set_hyper: null argument");
  TB_CHECK(kernel >= TB_RBF && kernel <= TB_MATERN52, "tb_gp_set_hyper: unknown kernel kind");
  TB_CHECK(variance > 0.0, "tb_gp_set_hyper: kernel variance must be positive");
  TB_CHECK(noise_variance > 0.0, "tb_gp_set_hyper: likelihood variance must be positive");
  TB_CHECK(n_ls >= 1, "tb_gp_set_hyper: need at least one lengthscale");
  for (int i = 0; i < n_ls; ++i)
    TB_CHECK(lengthscales[i] > 0.0, "tb_gp_set_hyper: lengthscales must be positive");
  TB_CHECK(!gp->have_data || n_ls == 1 || n_ls == gp->D,
           "tb_gp_set_hyper: lengthscales must have 1 or D entries");
  gp->kernel = kernel;
  gp->variance = variance;
  gp->noise = noise_variance;
  gp->mean_const = mean_const;
  gp->ls.assign(lengthscales, lengthscales + n_ls);
  if (gp->have_data && n_ls == 1) gp->ls.assign(gp->D, lengthscales[0]);
  gp->have_hyper = true;
  gp->cache_valid = false;
  return 0;
}

static int scale_inputs(tb_gp* gp) {
  const int D = gp->D, DP = gp->DP;
  const int64_t rows = (int64_t)gp->NB * BM;
  std::vector<double> inv_ls(DP, 0.0);
  for (int d = 0; d < D; ++d) inv_ls[d] = 1.0 / gp->ls[d];
  TB_TRY(gp->dInvLs.reserve(sizeof(double) * DP));
  TB_CUDA(cudaMemcpyAsync(gp->dInvLs.p, inv_ls.data(), sizeof(double) * DP, cudaMemcpyHostToDevice, gp->stream));
  TB_TRY(gp->dXs.reserve(sizeof(double) * rows * DP));
  int64_t tot = rows * DP;
  scale_inputs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, gp->stream>>>(gp->dX.as<double>(), gp->dInvLs.as<double>(), gp->N, D,
                                                                            DP, rows, gp->dXs.as<double>());
  TB_LAUNCHED();
  return 0;
}

// cuSOLVER / cuBLAS handles exist only for the TB_FACTOR=cusolver cross-check path: created on first use
static int ensure_library_handles(tb_gp* gp) {
  if (gp->cublas && gp->cusolver) return 0;
  if (!gp->cublas) {
    TB_CUBLAS(cublasCreate(&gp->cublas));
    TB_CUBLAS(cublasSetStream(gp->cublas, gp->stream));
  }
  if (!gp->cusolver) {
    TB_CUSOLVER(cusolverDnCreate(&gp->cusolver));
    TB_CUSOLVER(cusolverDnSetStream(gp->cusolver, gp->stream));
  }
  return 0;
}

// alpha = Linv^T (Linv (y - m)) by two triangular mat-vecs on the current Linv
static int alpha_from_linv(tb_gp* gp) {
  const int64_t N = gp->N, rows = (int64_t)gp->NB * BM;
  cudaStream_t st = gp->stream;
  TB_TRY(gp->sMisc.reserve(sizeof(double) * 2 * rows));
  double* err = gp->sMisc.as<double>();
  double* tmp = err + rows;
  residual_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(gp->dy.as<double>(), N, rows, gp->mean_const, err);
  TB_LAUNCHED();
  TB_CUDA(cudaMemsetAsync(gp->dAlpha.p, 0, sizeof(double) * rows, st));
  fac::trmv_lower_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(gp->dLinv.as<double>(), N, err, tmp);
  TB_LAUNCHED();
  fac::trmv_lower_t_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(gp->dLinv.as<double>(), N, tmp, gp->dAlpha.as<double>());
  TB_LAUNCHED();
  return 0;
}

// pack the lower triangle of Linv into DMMA-fragment-ordered panels and mark the derived operand sets stale
// appended_from > 0: the cache was extended from that many rows by tb_gp_append_data (the dense K^-1, when it exists, is grown
// by the same rank instead of being invalidated)
static int finish_cache(tb_gp* gp, int64_t appended_from = 0) {
  const int64_t N = gp->N;
  cudaStream_t st = gp->stream;
  int64_t npanels = rowblock_panel_offset(gp->NB);
  TB_TRY(gp->dLinvP.reserve(sizeof(double) * npanels * PANEL));
  TB_CUDA(cudaMemsetAsync(gp->dLinvP.p, 0, sizeof(double) * npanels * PANEL, st));
  dim3 grid((unsigned)std::min<int64_t>(gp->nkc, (int64_t)gp->NB * (BM / BK)), (unsigned)gp->NB);
  pack_lower_panels_kernel<<<grid, 256, 0, st>>>(gp->dLinv.as<double>(), N, gp->nkc, gp->dLinvP.as<double>());
  TB_LAUNCHED();
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  gp->cache_valid = true;
  gp->upper_valid = false;
  gp->oz_valid = false;
  gp->oz5_valid = false;
  gp->kinv_valid = false;
  gp->kinv5_valid = false;
  if (appended_from > 0 && gp->kinv_dense_valid && gp->kinv_dense_N == appended_from) {
    TB_TRY(gp->dKinvSpare.reserve(sizeof(double) * N * N));
    fac::kinv_grow_kernel<<<dim3((unsigned)((N + 127) / 128), (unsigned)N), 128, 0, st>>>(gp->dKinv.as<double>(), appended_from,
                                                                                       gp->dLinv.as<double>(), N, gp->dKinvSpare.as<double>());
    TB_LAUNCHED();
    TB_CUDA(cudaStreamSynchronize(st));
    TB_CUDA(cudaGetLastError());
    std::swap(gp->dKinv, gp->dKinvSpare);
    gp->kinv_dense_N = N;
  } else {
    gp->kinv_dense_valid = false;
  }
  return 0;
}

int tb_gp_update_posterior_cache(tb_gp* gp) {
  TB_CHECK(gp, "tb_gp_update_posterior_cache: null handle");
  TB_CHECK(gp->have_data && gp->have_hyper, "tb_gp_update_posterior_cache: set data and hyper-parameters first");
  TB_CHECK((int)gp->ls.size() == gp->D, "tb_gp_update_posterior_cache: lengthscales must have 1 or D entries");
  TB_CUDA(cudaSetDevice(gp->device));
  const int64_t N = gp->N;
  const int D = gp->D, DP = gp->DP;
  const int64_t rows = (int64_t)gp->NB * BM;  // >= nkc*16 and >= nst*64
  cudaStream_t st = gp->stream;

  TB_TRY(scale_inputs(gp));
  TB_TRY(gp->dL.reserve(sizeof(double) * N * N));
  TB_TRY(gp->dLinv.reserve(sizeof(double) * N * N));
  {
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
    double* K = gp->dL.as<double>();
    const double* Xs = gp->dXs.as<double>();
    switch (gp->kernel) {
      case TB_RBF: gram_kernel<TB_RBF><<<grid, 128, 0, st>>>(Xs, N, DP, gp->variance, gp->noise, K); break;
      case TB_MATERN12: gram_kernel<TB_MATERN12><<<grid, 128, 0, st>>>(Xs, N, DP, gp->variance, gp->noise, K); break;
      case TB_MATERN32: gram_kernel<TB_MATERN32><<<grid, 128, 0, st>>>(Xs, N, DP, gp->variance, gp->noise, K); break;
      default: gram_kernel<TB_MATERN52><<<grid, 128, 0, st>>>(Xs, N, DP, gp->variance, gp->noise, K); break;
    }
    TB_LAUNCHED();
  }
  TB_TRY(gp->dInfo.reserve(sizeof(int)));
  TB_TRY(gp->dAlpha.reserve(sizeof(double) * rows));
  if (gp->factor_own) {
    // ---- hand-written path (factor.cuh): blocked Cholesky, Linv, alpha on the DMMA pipe; no library call ----
    const int nbk = (int)((N + fac::FB - 1) / fac::FB);
    TB_TRY(gp->dDinv.reserve(sizeof(double) * (size_t)nbk * fac::FB * fac::FB));
    TB_CUDA(cudaMemsetAsync(gp->dInfo.p, 0, sizeof(int), st));
    double* A = gp->dL.as<double>();
    const size_t diag_smem = sizeof(double) * fac::FB * (fac::FB + 1);
    for (int jb = 0; jb < nbk; ++jb) {
      const int j0 = jb * fac::FB;
      fac::chol_diag_kernel<<<1, fac::THREADS, diag_smem, st>>>(A, N, j0, gp->dDinv.as<double>(), gp->dInfo.as<int>());
      TB_LAUNCHED();
      const int64_t below = N - (int64_t)(j0 + fac::FB);
      if (below > 0) {
        const unsigned t = (unsigned)((below + fac::FB - 1) / fac::FB);
        fac::chol_panel_kernel<<<t, fac::THREADS, fac::GEMM_SMEM, st>>>(A, N, j0, gp->dDinv.as<double>());
        TB_LAUNCHED();
        fac::chol_syrk_kernel<<<dim3(t, t), fac::THREADS, fac::GEMM_SMEM, st>>>(A, N, j0);
        TB_LAUNCHED();
      }
    }
    int info = 0;
    TB_CUDA(cudaMemcpyAsync(&info, gp->dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    TB_CUDA(cudaGetLastError());
    TB_CHECK_CODE(info == 0, "tb_gp_update_posterior_cache: Cholesky decomposition was not successful "
                        "(K + noise*I not positive definite at leading minor " + std::to_string(info) + ")", tb::ERR_NUMERIC);
    {
      dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
      zero_upper_kernel<<<grid, 128, 0, st>>>(N, A);
      TB_LAUNCHED();
    }
    TB_CUDA(cudaMemsetAsync(gp->dLinv.p, 0, sizeof(double) * N * N, st));
    // Linv by recursive doubling: diagonal 128-blocks first, then block sizes 128, 256, ... (two GEMM launches per level)
    fac::trinv_diag_kernel<<<nbk, 256, 0, st>>>(gp->dLinv.as<double>(), N, gp->dDinv.as<double>());
    TB_LAUNCHED();
    if (nbk > 1) {
      TB_TRY(gp->dKinv.reserve(sizeof(double) * N * N));  // scratch T (the buffer is reused later for K^-1)
      for (int64_t n = fac::FB; n < (int64_t)nbk * fac::FB; n *= 2) {
        const unsigned tiles = (unsigned)(n / fac::FB), pairs = (unsigned)((N + 2 * n - 1) / (2 * n));
        fac::trinv_level_kernel<1><<<dim3(tiles, tiles, pairs), fac::THREADS, fac::GEMM_SMEM, st>>>(A, gp->dLinv.as<double>(),
                                                                                                   gp->dKinv.as<double>(), N, (int)n);
        TB_LAUNCHED();
        fac::trinv_level_kernel<2><<<dim3(tiles, tiles, pairs), fac::THREADS, fac::GEMM_SMEM, st>>>(A, gp->dLinv.as<double>(),
                                                                                                   gp->dKinv.as<double>(), N, (int)n);
        TB_LAUNCHED();
      }
    }
    TB_TRY(alpha_from_linv(gp));
  } else {
    // ---- library path (TB_FACTOR=cusolver): cuSOLVER potrf / potrs + cuBLAS trsm; kept as a cross-check ----
    TB_TRY(ensure_library_handles(gp));
    int lwork = 0;
    TB_CUSOLVER(cusolverDnDpotrf_bufferSize(gp->cusolver, CUBLAS_FILL_MODE_LOWER, (int)N, gp->dL.as<double>(), (int)N, &lwork));
    TB_TRY(gp->dWork.reserve(sizeof(double) * (size_t)std::max(lwork, 1)));
    TB_CUSOLVER(cusolverDnDpotrf(gp->cusolver, CUBLAS_FILL_MODE_LOWER, (int)N, gp->dL.as<double>(), (int)N,
                                 gp->dWork.as<double>(), lwork, gp->dInfo.as<int>()));
    int info = 0;
    TB_CUDA(cudaMemcpyAsync(&info, gp->dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    TB_CHECK_CODE(info == 0, "tb_gp_update_posterior_cache: Cholesky decomposition was not successful "
                        "(K + noise*I not positive definite at leading minor " + std::to_string(info) + ")", tb::ERR_NUMERIC);
    {
      dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
      zero_upper_kernel<<<grid, 128, 0, st>>>(N, gp->dL.as<double>());
      TB_LAUNCHED();
    }
    residual_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(gp->dy.as<double>(), N, rows, gp->mean_const,
                                                                    gp->dAlpha.as<double>());
    TB_LAUNCHED();
    TB_CUSOLVER(cusolverDnDpotrs(gp->cusolver, CUBLAS_FILL_MODE_LOWER, (int)N, 1, gp->dL.as<double>(), (int)N,
                                 gp->dAlpha.as<double>(), (int)N, gp->dInfo.as<int>()));
    int64_t tot = N * N;
    identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(N, gp->dLinv.as<double>());
    TB_LAUNCHED();
    const double one = 1.0;
    TB_CUBLAS(cublasDtrsm(gp->cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)N,
                          (int)N, &one, gp->dL.as<double>(), (int)N, gp->dLinv.as<double>(), (int)N));
  }
  return finish_cache(gp);
}

// Rank-m append: the reference refactorises from scratch whenever the data change
// (models.py:171-186 -> interface.py:108-112); one BO step only appends rows, so L, Linv and alpha are extended in
// O(m N^2) instead of O(N^3).  The hyper-parameters must be unchanged since the cache was built.
static int tb_gp_append_data_f64(tb_gp* gp, const double* Xnew, const double* ynew, int64_t m) {
  TB_CHECK(gp && Xnew && ynew, "tb_gp_append_data: null argument");
  TB_CHECK(gp->cache_valid, "tb_gp_append_data: posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(m > 0 && m <= APPEND_MAX, "tb_gp_append_data: between 1 and " + std::to_string(APPEND_MAX) + " new points per call");
  const int64_t N0 = gp->N, N = N0 + m;
  TB_CHECK(N <= 65535, "tb_gp_append_data: N > 65535 is not supported");
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D, DP = gp->DP;
  // grow the raw data and the two triangular factors (leading dimension N0 -> N) into the handle's spare buffers, which are
  // sized with slack (next multiple of 256 rows + 256) and ping-pong with the live ones: after the second append of a run
  // no cudaMalloc / cudaFree (both device-synchronising) is left on this path
  const int64_t cap_rows = ((N + 255) / 256) * 256 + 256;
  tb::DevBuf &nX = gp->dXspare, &ny = gp->dyspare, &nL = gp->dLspare, &nLinv = gp->dLinvSpare;
  TB_TRY(nX.reserve(sizeof(double) * cap_rows * D));
  TB_TRY(ny.reserve(sizeof(double) * cap_rows));
  TB_TRY(nL.reserve(sizeof(double) * cap_rows * cap_rows));
  TB_TRY(nLinv.reserve(sizeof(double) * cap_rows * cap_rows));
  TB_CUDA(cudaMemcpyAsync(nX.p, gp->dX.p, sizeof(double) * N0 * D, cudaMemcpyDeviceToDevice, st));
  TB_CUDA(cudaMemcpyAsync(nX.as<double>() + N0 * D, Xnew, sizeof(double) * m * D, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(ny.p, gp->dy.p, sizeof(double) * N0, cudaMemcpyDeviceToDevice, st));
  TB_CUDA(cudaMemcpyAsync(ny.as<double>() + N0, ynew, sizeof(double) * m, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemsetAsync(nL.p, 0, sizeof(double) * N * N, st));
  TB_CUDA(cudaMemsetAsync(nLinv.p, 0, sizeof(double) * N * N, st));
  TB_CUDA(cudaMemcpy2DAsync(nL.p, sizeof(double) * N, gp->dL.p, sizeof(double) * N0, sizeof(double) * N0, N0,
                            cudaMemcpyDeviceToDevice, st));
  TB_CUDA(cudaMemcpy2DAsync(nLinv.p, sizeof(double) * N, gp->dLinv.p, sizeof(double) * N0, sizeof(double) * N0, N0,
                            cudaMemcpyDeviceToDevice, st));
  TB_CUDA(cudaStreamSynchronize(st));
  std::swap(gp->dX, nX);
  std::swap(gp->dy, ny);
  std::swap(gp->dL, nL);
  std::swap(gp->dLinv, nLinv);
  gp->N = N;
  gp->nkc = (int)((N + BK - 1) / BK);
  gp->NB = (int)((N + BM - 1) / BM);
  gp->cache_valid = false;  // until the append completes
  const int64_t rows = (int64_t)gp->NB * BM;
  TB_TRY(scale_inputs(gp));
  TB_TRY(gp->dAlpha.reserve(sizeof(double) * rows));
  // scratch: W (cross kernel block), Y = Linv0 B, U = Linv0^T Y as [N, m] column-major; S, R as [m, m]
  // sized for the largest append at the spare buffers' row capacity: no reallocation while the capacity lasts
  TB_TRY(gp->sA.reserve(sizeof(double) * (3 * cap_rows * APPEND_MAX + 2 * APPEND_MAX * APPEND_MAX)));
  double* W = gp->sA.as<double>();
  double* Y = W + N * m;
  double* U = Y + N * m;
  double* S = U + N * m;
  double* R = S + m * m;
  TB_TRY(gp->dInfo.reserve(sizeof(int)));
  TB_CUDA(cudaMemsetAsync(gp->dInfo.p, 0, sizeof(int), st));
  {
    dim3 grid((unsigned)((N + 127) / 128), (unsigned)m);
    const double* Xs = gp->dXs.as<double>();
    switch (gp->kernel) {
      case TB_RBF: append_cross_kernel<TB_RBF><<<grid, 128, 0, st>>>(Xs, DP, N0, N, gp->variance, gp->noise, W); break;
      case TB_MATERN12: append_cross_kernel<TB_MATERN12><<<grid, 128, 0, st>>>(Xs, DP, N0, N, gp->variance, gp->noise, W); break;
      case TB_MATERN32: append_cross_kernel<TB_MATERN32><<<grid, 128, 0, st>>>(Xs, DP, N0, N, gp->variance, gp->noise, W); break;
      default: append_cross_kernel<TB_MATERN52><<<grid, 128, 0, st>>>(Xs, DP, N0, N, gp->variance, gp->noise, W); break;
    }
    TB_LAUNCHED();
  }
  double* L = gp->dL.as<double>();
  double* Linv = gp->dLinv.as<double>();
  trmv_lower_cols_kernel<<<dim3((unsigned)((N0 + 127) / 128), (unsigned)m), 128, 0, st>>>(Linv, N0, N, W, N, Y, N);
  TB_LAUNCHED();
  append_schur_kernel<<<dim3((unsigned)m, (unsigned)m), 256, 0, st>>>(Y, W, N0, N, (int)m, S);
  TB_LAUNCHED();
  append_chol_kernel<<<1, APPEND_MAX, 0, st>>>(S, (int)m, N0, N, L, Linv, R, gp->dInfo.as<int>());
  TB_LAUNCHED();
  trmv_lower_t_cols_kernel<<<dim3((unsigned)((N0 + 7) / 8), (unsigned)m), 256, 0, st>>>(Linv, N0, N, Y, N, U, N);
  TB_LAUNCHED();
  append_rows_kernel<<<dim3((unsigned)((N0 + 127) / 128), (unsigned)m), 128, 0, st>>>(Y, U, R, (int)m, N0, N, L, Linv);
  TB_LAUNCHED();
  int info = 0;
  TB_CUDA(cudaMemcpyAsync(&info, gp->dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  TB_CHECK_CODE(info == 0, "tb_gp_append_data: Cholesky decomposition was not successful "
                      "(K + noise*I not positive definite at leading minor " + std::to_string(info) + ")", tb::ERR_NUMERIC);
  TB_TRY(alpha_from_linv(gp));
  return finish_cache(gp, N0);
}

static int tb_gp_get_cholesky_f64(tb_gp* gp, void* L_out) {
  TB_CHECK(gp && L_out, "tb_gp_get_cholesky: null argument");
  TB_CHECK(gp->cache_valid, "tb_gp_get_cholesky: posterior cache is not built");
  TB_CUDA(cudaSetDevice(gp->device));
  const int64_t N = gp->N;
  TB_TRY(gp->sMisc.reserve(sizeof(double) * N * N));
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)N);
  colmajor_lower_to_rowmajor_kernel<<<grid, 128, 0, gp->stream>>>(gp->dL.as<double>(), N, gp->sMisc.as<double>());
  TB_LAUNCHED();
  TB_CUDA(cudaMemcpyAsync(L_out, gp->sMisc.p, sizeof(double) * N * N, cudaMemcpyDefault, gp->stream));
  TB_CUDA(cudaStreamSynchronize(gp->stream));
  return 0;
}

}  // extern "C"

// =================================================================================================
// per-candidate path: chunked driver for predict / acquisition / argmax
// =================================================================================================
namespace tb {

struct EvalRequest {
  int acq = -1;  // -1: predict only
  double param = 0.0;
  const double* Xc = nullptr;  // host or device, [M, D]
  int64_t M = 0;
  double* out_vals = nullptr;  // host or device (nullable)
  double* out_mean = nullptr;
  double* out_var = nullptr;
  double* out_grad = nullptr;  // [M, D] (nullable)
  bool want_argmax = false;
  double best_value = 0.0;
  int64_t best_index = -1;
};

static int launch_kstar(tb_gp* gp, const double* Xc_dev, int64_t mc, int tiles, double* KsP, double* mean) {
  const double* Xs = gp->dXs.as<double>();
  const double* al = gp->dAlpha.as<double>();
  const double* il = gp->dInvLs.as<double>();
  const int N = (int)gp->N, nkc = gp->nkc, D = gp->D;
  const double var = gp->variance, mc0 = gp->mean_const;
  cudaStream_t st = gp->stream;
#define TB_KSTAR(KIND, DPV)                                                                               \
  kstar_panels_kernel<KIND, DPV><<<tiles, 512, 0, st>>>(Xs, al, Xc_dev, il, N, nkc, D, mc, 0, var, mc0, KsP, mean)
#define TB_KSTAR_DP(KIND)                                   \
  switch (gp->DP) {                                         \
    case 2: TB_KSTAR(KIND, 2); break;                       \
    case 4: TB_KSTAR(KIND, 4); break;                       \
    case 6: TB_KSTAR(KIND, 6); break;                       \
    case 8: TB_KSTAR(KIND, 8); break;                       \
    case 10: TB_KSTAR(KIND, 10); break;                     \
    case 12: TB_KSTAR(KIND, 12); break;                     \
    case 16: TB_KSTAR(KIND, 16); break;                     \
    case 20: TB_KSTAR(KIND, 20); break;                     \
    case 24: TB_KSTAR(KIND, 24); break;                     \
    default: TB_KSTAR(KIND, 32); break;                     \
  }
  switch (gp->kernel) {
    case TB_RBF: TB_KSTAR_DP(TB_RBF); break;
    case TB_MATERN12: TB_KSTAR_DP(TB_MATERN12); break;
    case TB_MATERN32: TB_KSTAR_DP(TB_MATERN32); break;
    default: TB_KSTAR_DP(TB_MATERN52); break;
  }
#undef TB_KSTAR_DP
#undef TB_KSTAR
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  return 0;
}

// candidates per chunk: bounded by the Ks scratch budget, a whole number of 148-CTA waves when possible
static int64_t chunk_tiles(const tb_gp* gp) {
  const size_t per_tile = (size_t)gp->nkc * PANEL * sizeof(double);
  const size_t budget = (size_t)1536 << 20;
  int64_t t = (int64_t)(budget / per_tile);
  t = std::max<int64_t>(t, 1);
  if (t >= 148) t = (t / 148) * 148;
  return std::min<int64_t>(t, 148 * 16);
}

static int pick_groups(const tb_gp* gp, int tiles) {
  if (tiles >= 148) return 1;
  int G = (2 * 148 + tiles - 1) / tiles;
  return std::max(1, std::min(G, gp->NB));
}

int kernels_init() {
  TB_CUDA(cudaFuncSetAttribute(fac::chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * fac::FB * (fac::FB + 1))));
  TB_CUDA(cudaFuncSetAttribute(fac::chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM));
  TB_CUDA(cudaFuncSetAttribute(fac::chol_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM));
  TB_CUDA(cudaFuncSetAttribute(fac::trinv_level_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM));
  TB_CUDA(cudaFuncSetAttribute(fac::trinv_level_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM));
  TB_CUDA(cudaFuncSetAttribute(fac::kinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM));
  TB_CUDA(cudaFuncSetAttribute(oz::trigemm_i8_kernel<oz::OZ_SUMSQ, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz::SMEM_BYTES));
  TB_CUDA(cudaFuncSetAttribute(oz::trigemm_i8_lowreg_kernel<oz::OZ_SUMSQ, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz::SMEM_BYTES));
  TB_CUDA(cudaFuncSetAttribute(oz::trigemm_i8_kernel<oz::OZ_STORE, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)oz::SMEM_BYTES));
  TB_TRY(oz5_init());
  TB_CUDA(cudaFuncSetAttribute(trigemm_kernel<false, EPI_SUMSQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
  TB_CUDA(cudaFuncSetAttribute(trigemm_kernel<false, EPI_SUMSQ_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
  TB_CUDA(cudaFuncSetAttribute(trigemm_kernel<false, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
  TB_CUDA(cudaFuncSetAttribute(trigemm_kernel<true, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM));
  return 0;
}

// Linv^T packed upper panels: built lazily, only the gradient path needs them
static int ensure_upper_panels(tb_gp* gp) {
  if (gp->upper_valid) return 0;
  const int nkB = gp->NB * (BM / BK);
  const int64_t np = upper_panel_count(gp->NB, nkB);
  TB_TRY(gp->dLinvTP.reserve(sizeof(double) * np * PANEL));
  dim3 grid((unsigned)nkB, (unsigned)gp->NB);
  pack_upper_panels_kernel<<<grid, 256, 0, gp->stream>>>(gp->dLinv.as<double>(), gp->N, gp->NB, nkB,
                                                         gp->dLinvTP.as<double>());
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  gp->upper_valid = true;
  return 0;
}

template <int KIND>
static void launch_grad_dp(tb_gp* gp, const double* xc, int64_t mc, double* grad_dev) {
  const bool wide = mc <= 2048;  // one CTA (8 warps) per candidate when one warp each would leave SMs idle
  const int blocks = wide ? (int)mc : (int)((mc + 7) / 8);
  const double* Xs = gp->dXs.as<double>();
  const double* al = gp->dAlpha.as<double>();
  const double* il = gp->dInvLs.as<double>();
  const double* V = gp->sV.as<double>();
  const int64_t ldv = (int64_t)gp->NB * BM;
  const double* cmu = gp->sMisc.as<double>();
  const double* cvar = cmu + mc;
#define TB_GRAD(DPV)                                                                                                                     \
  if (wide)                                                                                                                              \
    grad_kernel<KIND, DPV, 8><<<blocks, 256, 0, gp->stream>>>(Xs, al, xc, il, (int)gp->N, gp->D, mc, V, ldv, cmu, cvar, gp->variance,     \
                                                              fm::Consts(), grad_dev);                                                   \
  else                                                                                                                                   \
    grad_kernel<KIND, DPV, 1><<<blocks, 256, 0, gp->stream>>>(Xs, al, xc, il, (int)gp->N, gp->D, mc, V, ldv, cmu, cvar, gp->variance,     \
                                                              fm::Consts(), grad_dev)
  switch (gp->DP) {
    case 2: TB_GRAD(2); break;
    case 4: TB_GRAD(4); break;
    case 6: TB_GRAD(6); break;
    case 8: TB_GRAD(8); break;
    case 10: TB_GRAD(10); break;
    case 12: TB_GRAD(12); break;
    case 16: TB_GRAD(16); break;
    case 20: TB_GRAD(20); break;
    case 24: TB_GRAD(24); break;
    default: TB_GRAD(32); break;
  }
#undef TB_GRAD
}

// after the lower GEMM (A stored as packed panels in sA, sum-of-squares in sPartial):
// V = Linv^T A, then the gradient assembly
static int gradient_chunk(tb_gp* gp, int acq, double param, const double* xc, int64_t mc, int tiles, int G,
                          int64_t McPad, double* out_grad) {
  cudaStream_t st = gp->stream;
  double* cmu = gp->sMisc.as<double>();
  acq_partials_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, st>>>(gp->sPartial.as<double>(), G, McPad,
                                                                    gp->sMean.as<double>(), mc, gp->variance, acq,
                                                                    param, gp->noise, gp->dMes.as<double>(), gp->mesS, cmu, cmu + mc);
  TB_LAUNCHED();
  const int nkB = gp->NB * (BM / BK);
  trigemm_kernel<true, EPI_PLAIN><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
      gp->dLinvTP.as<double>(), gp->sA.as<double>(), gp->NB, nkB, G, McPad, nullptr, nullptr, gp->sV.as<double>(),
      (int64_t)gp->NB * BM);
  TB_LAUNCHED();
  const bool gdev = is_device_ptr(out_grad);
  double* gd = gdev ? out_grad : gp->sGrad.as<double>();
  switch (gp->kernel) {
    case TB_RBF: launch_grad_dp<TB_RBF>(gp, xc, mc, gd); break;
    case TB_MATERN12: launch_grad_dp<TB_MATERN12>(gp, xc, mc, gd); break;
    case TB_MATERN32: launch_grad_dp<TB_MATERN32>(gp, xc, mc, gd); break;
    default: launch_grad_dp<TB_MATERN52>(gp, xc, mc, gd); break;
  }
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  if (!gdev) TB_CUDA(cudaMemcpyAsync(out_grad, gd, sizeof(double) * mc * gp->D, cudaMemcpyDeviceToHost, st));
  return 0;
}

// fp32 models (TB_F32 handles) run the HI pass only: 10 digit products, error ~1e-7 sigma_f^2 << the fp32 tolerance
static inline int oz_npass(const tb_gp* gp) { return gp->dtype == TB_F32 ? 1 : 2; }

// Ozaki engine state: digit tiles of Linv + row scales, built lazily after each cache refresh
static int ensure_ozaki(tb_gp* gp) {
  if (gp->oz_valid) return 0;
  TB_CHECK(gp->N <= 16384, "the int8 engine supports N <= 16384 (int32 accumulator headroom)");
  cudaStream_t st = gp->stream;
  const int64_t rows = (int64_t)gp->NB * BM;
  gp->nst = (int)((gp->N + oz::KST - 1) / oz::KST);
  TB_TRY(gp->dRowScale.reserve(sizeof(double) * rows));
  oz::linv_rowscale_kernel<<<(unsigned)rows, 256, 0, st>>>(gp->dLinv.as<double>(), gp->N, rows, gp->dRowScale.as<double>());
  TB_LAUNCHED();
  TB_LAUNCHED();
  const int64_t nstages = oz::a_stage_offset(gp->NB);
  TB_TRY(gp->dAS.reserve((size_t)nstages * oz::S * oz::TILE));
  TB_CUDA(cudaMemsetAsync(gp->dAS.p, 0, (size_t)nstages * oz::S * oz::TILE, st));
  oz::linv_digits_kernel<<<dim3(2 * gp->NB, gp->NB), 256, 0, st>>>(gp->dLinv.as<double>(), gp->N, gp->dRowScale.as<double>(),
                                                                   gp->dAS.as<int8_t>());
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  int e = 0;
  std::frexp(gp->variance, &e);  // variance = m 2^e, m in [0.5, 1)  ->  K* / 2^(e+2) < 1/4
  gp->oz_bscale_exp = e + 2;
  gp->oz_out_scale = std::ldexp(1.0, gp->oz_bscale_exp);
  gp->oz_valid = true;
  return 0;
}

// K^-1 digit tiles for the gradient path of the int8 engine: V = K^-1 k* as one dense digit GEMM (same K* digits as the
// variance GEMM).  K^-1 from the cached factor with cuSOLVER potri (once per BO step, lazily).
// dense K^-1 = Linv^T Linv (lower triangle, ld = N), O(N^3) on the DMMA pipe: once per full cache refresh; appends grow it in
// O(m N^2) (finish_cache / fac::kinv_grow_kernel)
static int ensure_kinv_dense(tb_gp* gp) {
  const int64_t N = gp->N;
  if (gp->kinv_dense_valid && gp->kinv_dense_N == N) return 0;
  cudaStream_t st = gp->stream;
  TB_TRY(gp->dKinv.reserve(sizeof(double) * N * N));
  if (gp->factor_own) {
    const unsigned t = (unsigned)((N + fac::FB - 1) / fac::FB);
    fac::kinv_kernel<<<dim3(t, t), fac::THREADS, fac::GEMM_SMEM, st>>>(gp->dLinv.as<double>(), gp->dKinv.as<double>(), N);
    TB_LAUNCHED();
  } else {
    TB_TRY(ensure_library_handles(gp));
    TB_CUDA(cudaMemcpyAsync(gp->dKinv.p, gp->dL.p, sizeof(double) * N * N, cudaMemcpyDeviceToDevice, st));
    int lwork = 0;
    cusolverStatus_t cs = cusolverDnDpotri_bufferSize(gp->cusolver, CUBLAS_FILL_MODE_LOWER, (int)N, gp->dKinv.as<double>(), (int)N, &lwork);
    TB_CHECK_CODE(cs == CUSOLVER_STATUS_SUCCESS, "cusolverDnDpotri_bufferSize failed", tb::ERR_RUNTIME);
    TB_TRY(gp->dWork.reserve(sizeof(double) * (size_t)std::max(lwork, 1)));
    cs = cusolverDnDpotri(gp->cusolver, CUBLAS_FILL_MODE_LOWER, (int)N, gp->dKinv.as<double>(), (int)N, gp->dWork.as<double>(), lwork,
                          gp->dInfo.as<int>());
    TB_CHECK_CODE(cs == CUSOLVER_STATUS_SUCCESS, "cusolverDnDpotri failed", tb::ERR_RUNTIME);
  }
  gp->kinv_dense_valid = true;
  gp->kinv_dense_N = N;
  gp->kinv5_valid = false;
  return 0;
}

static int ensure_kinv_digits(tb_gp* gp) {
  if (gp->kinv_valid) return 0;
  TB_TRY(ensure_ozaki(gp));
  cudaStream_t st = gp->stream;
  const int64_t N = gp->N, rows = (int64_t)gp->NB * BM;
  TB_TRY(ensure_kinv_dense(gp));
  TB_TRY(gp->dKinvScale.reserve(sizeof(double) * rows));
  oz::sym_rowscale_kernel<<<(unsigned)rows, 256, 0, st>>>(gp->dKinv.as<double>(), N, rows, gp->dKinvScale.as<double>());
  TB_LAUNCHED();
  const size_t bytes = (size_t)gp->NB * gp->nst * oz::S * oz::TILE;
  TB_TRY(gp->dKinvS.reserve(bytes));
  oz::sym_digits_kernel<<<dim3(gp->nst, gp->NB), 256, 0, st>>>(gp->dKinv.as<double>(), N, gp->nst, gp->dKinvScale.as<double>(),
                                                              gp->dKinvS.as<int8_t>());
  TB_LAUNCHED();
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  gp->kinv_valid = true;
  return 0;
}

// int8 engine, gradient path: sum-of-squares is already in sPartial (variance GEMM); V = K^-1 k* on the tensor cores
static int gradient_chunk_oz(tb_gp* gp, int acq, double param, const double* xc, int64_t mc, int tiles, int G, int64_t McPad,
                             double* out_grad) {
  cudaStream_t st = gp->stream;
  double* cmu = gp->sMisc.as<double>();
  acq_partials_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, st>>>(gp->sPartial.as<double>(), G, McPad, gp->sMean.as<double>(), mc,
                                                                    gp->variance, acq, param, gp->noise, gp->dMes.as<double>(), gp->mesS, cmu, cmu + mc);
  TB_LAUNCHED();
  const int Gv = std::max(1, std::min(gp->NB, std::max((gp->NB + 7) / 8, (2 * 148 + tiles - 1) / tiles)));
  oz::trigemm_i8_kernel<oz::OZ_STORE, 8><<<dim3(Gv, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
      gp->dKinvS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dKinvScale.as<double>(), gp->NB, gp->nst, Gv, McPad, gp->oz_out_scale,
      oz_npass(gp), 1, nullptr, gp->sV.as<double>(), (int64_t)gp->NB * BM);
  TB_LAUNCHED();
  const bool gdev = is_device_ptr(out_grad);
  double* gd = gdev ? out_grad : gp->sGrad.as<double>();
  switch (gp->kernel) {
    case TB_RBF: launch_grad_dp<TB_RBF>(gp, xc, mc, gd); break;
    case TB_MATERN12: launch_grad_dp<TB_MATERN12>(gp, xc, mc, gd); break;
    case TB_MATERN32: launch_grad_dp<TB_MATERN32>(gp, xc, mc, gd); break;
    default: launch_grad_dp<TB_MATERN52>(gp, xc, mc, gd); break;
  }
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  if (!gdev) TB_CUDA(cudaMemcpyAsync(out_grad, gd, sizeof(double) * mc * gp->D, cudaMemcpyDeviceToHost, st));
  return 0;
}

static int launch_kstar_digits(tb_gp* gp, const double* Xc_dev, int64_t mc, int tiles, int8_t* BS, double* mean) {
  const double* Xs = gp->dXs.as<double>();
  const double* al = gp->dAlpha.as<double>();
  const double* il = gp->dInvLs.as<double>();
  const int N = (int)gp->N, nst = gp->nst, D = gp->D;
  const double var = gp->variance, mc0 = gp->mean_const;
  const double inv_b = std::ldexp(1.0, 48 - gp->oz_bscale_exp);
  cudaStream_t st = gp->stream;
#define TB_KD(KIND, DPV) \
  oz::kstar_digits_kernel<KIND, DPV><<<tiles, 512, 0, st>>>(Xs, al, Xc_dev, il, N, nst, D, mc, var, inv_b, mc0, BS, mean)
#define TB_KD_DP(KIND)                                   \
  switch (gp->DP) {                                      \
    case 2: TB_KD(KIND, 2); break;                       \
    case 4: TB_KD(KIND, 4); break;                       \
    case 6: TB_KD(KIND, 6); break;                       \
    case 8: TB_KD(KIND, 8); break;                       \
    case 10: TB_KD(KIND, 10); break;                     \
    case 12: TB_KD(KIND, 12); break;                     \
    case 16: TB_KD(KIND, 16); break;                     \
    case 20: TB_KD(KIND, 20); break;                     \
    case 24: TB_KD(KIND, 24); break;                     \
    default: TB_KD(KIND, 32); break;                     \
  }
  switch (gp->kernel) {
    case TB_RBF: TB_KD_DP(TB_RBF); break;
    case TB_MATERN12: TB_KD_DP(TB_MATERN12); break;
    case TB_MATERN32: TB_KD_DP(TB_MATERN32); break;
    default: TB_KD_DP(TB_MATERN52); break;
  }
#undef TB_KD_DP
#undef TB_KD
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  return 0;
}

// Pipelined driver of the int8 engine: K* digit generation of chunk c+1 (fp64 / integer pipes, stream B) overlaps the
// digit GEMM of chunk c (tensor pipe, stream A); scratch is double-buffered and nothing synchronises with the host until
// the end of the call.
static int run_eval_oz(tb_gp* gp, EvalRequest& rq) {
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t sa = gp->stream;
  if (!gp->stream2) {
    int lo = 0, hi = 0;
    TB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    TB_CUDA(cudaStreamCreateWithPriority(&gp->stream2, cudaStreamNonBlocking, lo));  // lowest priority
    for (int i = 0; i < 2; ++i) {
      TB_CUDA(cudaEventCreateWithFlags(&gp->evK[i], cudaEventDisableTiming));
      TB_CUDA(cudaEventCreateWithFlags(&gp->evDone[i], cudaEventDisableTiming));
    }
  }
  // K* generation on a second stream is an experiment switch (TB_OZ_OVERLAP=1): measured on B200 it does not pay, because the
  // generation kernel needs >= 8 resident warps per SM to keep pace and a resident GEMM CTA leaves room for fewer
  cudaStream_t sb = std::getenv("TB_OZ_OVERLAP") ? gp->stream2 : sa;
  const int D = gp->D;
  if (rq.want_argmax) {
    TB_CHECK(rq.M > 0, "argmax over an empty candidate set");
    TB_TRY(gp->sRun.reserve(16));
    double init_v = -INFINITY;  // a candidate worth -inf still beats "nothing seen" through the lower-index tie rule
    int64_t init_i = INT64_MAX;
    TB_CUDA(cudaMemcpyAsync(gp->sRun.p, &init_v, 8, cudaMemcpyHostToDevice, sa));
    TB_CUDA(cudaMemcpyAsync((char*)gp->sRun.p + 8, &init_i, 8, cudaMemcpyHostToDevice, sa));
  }
  if (rq.M == 0) return 0;
  // single-pass engine (15 / 6 digit products) when its a-priori error estimate clears the bar, else the 6-digit kernels
  TB_TRY(oz5_ensure(gp));
  const bool fast = gp->oz5_mode != 0;
  if (!fast) TB_TRY(ensure_ozaki(gp));
  TB_CUDA(cudaStreamSynchronize(sa));  // digit tiles of Linv are built on stream A; stream B reads model state too
  const int nt = fast ? oz5_tile_width(gp) : BT;  // candidates per tile

  const bool xc_dev = is_device_ptr(rq.Xc);
  const bool vals_dev = is_device_ptr(rq.out_vals), mean_dev = is_device_ptr(rq.out_mean), var_dev = is_device_ptr(rq.out_var);
  // half the usual scratch budget per slot (two slots are live)
  const size_t per_tile = fast ? oz5_tile_bytes(gp) : (size_t)gp->nst * oz::S * oz::TILE;
  int64_t max_tiles = std::max<int64_t>(1, (int64_t)(((size_t)1280 << 20) / per_tile));
  // chunks of whole GEMM rounds (148 CTAs, G items per tile).  The generation kernel of the single-pass engine runs 1.5 CTAs per
  // tile, 4 resident per SM, so 592 tiles are 1.5 of its waves; chunks of 789 / 1184 tiles (2 / 3 whole waves) were measured
  // on B200 and change the step by < 0.2 % (the step is power-bound), so the smaller scratch stays
  if (max_tiles >= 296) max_tiles = (max_tiles / 296) * 296;
  else if (max_tiles >= 148) max_tiles = 148;
  max_tiles = std::min<int64_t>(max_tiles, 296 * 4);
  if (const char* e = std::getenv("TB_OZ_TILES")) max_tiles = std::max(1, std::atoi(e));  // experiment knob
  const int64_t chunk_cap = std::min<int64_t>(max_tiles * nt, ((rq.M + nt - 1) / nt) * nt);
  const int64_t tiles_cap = chunk_cap / nt;
  // row-block groups per candidate tile: ~4 row-blocks per CTA amortise the CTA prologue while the co-resident CTAs still
  // share few enough candidate tiles for the K* digits to live in L2; small batches get more groups to fill the GPU
  int G = std::max(1, (gp->NB + 3) / 4);
  {
    const int64_t tiles_all = std::min<int64_t>(tiles_cap, (rq.M + nt - 1) / nt);
    if (tiles_all * G < 2 * 148) G = (int)std::min<int64_t>(gp->NB, (2 * 148 + tiles_all - 1) / tiles_all);
  }
  if (const char* e = std::getenv("TB_OZ_G")) G = std::max(1, std::min(std::atoi(e), gp->NB));  // experiment knob
  tb::DevBuf* ks[2] = {&gp->sKs, &gp->sKs2};
  tb::DevBuf* mean[2] = {&gp->sMean, &gp->sMean2};
  tb::DevBuf* part[2] = {&gp->sPartial, &gp->sPartial2};
  const int nslots = rq.M > chunk_cap ? 2 : 1;
  for (int i = 0; i < nslots; ++i) {
    TB_TRY(ks[i]->reserve((size_t)tiles_cap * per_tile));
    TB_TRY(mean[i]->reserve(sizeof(double) * chunk_cap));
    TB_TRY(part[i]->reserve(sizeof(double) * (size_t)G * chunk_cap));
  }
  if (!xc_dev) TB_TRY(gp->sXc.reserve(sizeof(double) * chunk_cap * D));
  if (rq.out_vals && !vals_dev) TB_TRY(gp->sVals.reserve(sizeof(double) * chunk_cap));
  if (rq.out_var && !var_dev) TB_TRY(gp->sVar.reserve(sizeof(double) * chunk_cap));
  const int tail_blocks_cap = (int)((chunk_cap + 255) / 256);
  if (rq.want_argmax) {
    TB_TRY(gp->sBlkBest.reserve(sizeof(double) * tail_blocks_cap));
    TB_TRY(gp->sBlkIdx.reserve(sizeof(int64_t) * tail_blocks_cap));
  }

  int64_t c = 0;
  for (int64_t c0 = 0; c0 < rq.M; c0 += chunk_cap, ++c) {
    const int slot = (int)(c & 1);
    const int64_t mc = std::min<int64_t>(chunk_cap, rq.M - c0);
    const int tiles = (int)((mc + nt - 1) / nt);
    const int64_t McPad = (int64_t)tiles * nt;
    // ---- stream B: candidates in, K* digits + mean out ----
    if (c >= 2) TB_CUDA(cudaStreamWaitEvent(sb, gp->evDone[slot], 0));
    const double* xc_chunk;
    if (xc_dev) {
      xc_chunk = rq.Xc + c0 * D;
    } else {
      TB_CUDA(cudaMemcpyAsync(gp->sXc.p, rq.Xc + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, sb));
      xc_chunk = gp->sXc.as<double>();
    }
    if (fast) {
      TB_TRY(oz5_launch_kstar(gp, sb, xc_chunk, mc, tiles, ks[slot]->as<int8_t>(), mean[slot]->as<double>()));
    } else {
      cudaStream_t keep = gp->stream;
      gp->stream = sb;  // launch_kstar_digits launches on gp->stream
      int rc = launch_kstar_digits(gp, xc_chunk, mc, tiles, ks[slot]->as<int8_t>(), mean[slot]->as<double>());
      gp->stream = keep;
      TB_TRY(rc);
    }
    TB_CUDA(cudaEventRecord(gp->evK[slot], sb));
    // ---- stream A: digit GEMM, tail, reductions, results out ----
    TB_CUDA(cudaStreamWaitEvent(sa, gp->evK[slot], 0));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (gp->profile) {
      TB_CUDA(cudaEventCreate(&e0));
      TB_CUDA(cudaEventCreate(&e1));
      TB_CUDA(cudaEventRecord(e0, sa));
    }
    if (fast)
      TB_TRY(oz5_launch_gemm(gp, sa, ks[slot]->as<int8_t>(), tiles, G, McPad, part[slot]->as<double>()));
    else if (sb != sa)  // overlapped K* generation: register-capped GEMM so that generation CTAs fit beside it
      oz::trigemm_i8_lowreg_kernel<oz::OZ_SUMSQ, 8><<<dim3(G, tiles), 10 * 32, oz::SMEM_BYTES, sa>>>(
          gp->dAS.as<int8_t>(), ks[slot]->as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, G, McPad, gp->oz_out_scale,
          oz_npass(gp), 0, part[slot]->as<double>(), nullptr, 0);
    else
      oz::trigemm_i8_kernel<oz::OZ_SUMSQ, 8><<<dim3(G, tiles), 10 * 32, oz::SMEM_BYTES, sa>>>(
          gp->dAS.as<int8_t>(), ks[slot]->as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, G, McPad, gp->oz_out_scale,
          oz_npass(gp), 0, part[slot]->as<double>(), nullptr, 0);
    if (!fast) TB_LAUNCHED();
    if (gp->profile) {
      TB_CUDA(cudaEventRecord(e1, sa));
      gp->prof_events.emplace_back(e0, e1);
      gp->prof_event_flops.push_back((double)McPad * (double)gp->N * (double)gp->N);
    }
    TB_CUDA(cudaGetLastError());
    double* d_vals = rq.out_vals ? (vals_dev ? rq.out_vals + c0 : gp->sVals.as<double>()) : nullptr;
    double* d_mean = rq.out_mean ? (mean_dev ? rq.out_mean + c0 : nullptr) : nullptr;
    double* d_var = rq.out_var ? (var_dev ? rq.out_var + c0 : gp->sVar.as<double>()) : nullptr;
    const int tb_blocks = (int)((mc + 255) / 256);
    tail_kernel<<<tb_blocks, 256, 0, sa>>>(part[slot]->as<double>(), G, McPad, mean[slot]->as<double>(), mc, c0, gp->variance,
                                           rq.acq, rq.param, gp->noise, gp->dMes.as<double>(), gp->mesS, d_vals, d_mean, d_var,
                                           rq.want_argmax ? gp->sBlkBest.as<double>() : nullptr,
                                           rq.want_argmax ? gp->sBlkIdx.as<int64_t>() : nullptr);
    TB_LAUNCHED();
    if (rq.want_argmax) {
      argmax_fold_kernel<<<1, 256, 0, sa>>>(gp->sBlkBest.as<double>(), gp->sBlkIdx.as<int64_t>(), tb_blocks,
                                            gp->sRun.as<double>(), reinterpret_cast<int64_t*>((char*)gp->sRun.p + 8));
      TB_LAUNCHED();
    }
    if (rq.out_vals && !vals_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_vals + c0, gp->sVals.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, sa));
    if (rq.out_mean && !mean_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_mean + c0, mean[slot]->p, sizeof(double) * mc, cudaMemcpyDeviceToHost, sa));
    if (rq.out_var && !var_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_var + c0, gp->sVar.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, sa));
    TB_CUDA(cudaEventRecord(gp->evDone[slot], sa));
    // the host staging of the candidates is single-buffered: the next H2D (stream B) must not overtake this chunk's kstar,
    // which stream order on B already guarantees
  }
  if (rq.want_argmax) {
    TB_CUDA(cudaMemcpyAsync(&rq.best_value, gp->sRun.p, 8, cudaMemcpyDeviceToHost, sa));
    TB_CUDA(cudaMemcpyAsync(&rq.best_index, (char*)gp->sRun.p + 8, 8, cudaMemcpyDeviceToHost, sa));
  }
  TB_CUDA(cudaStreamSynchronize(sb));
  TB_CUDA(cudaStreamSynchronize(sa));
  TB_CUDA(cudaGetLastError());
  if (gp->profile) {
    for (size_t i = 0; i < gp->prof_events.size(); ++i) {
      float ms = 0.f;
      TB_CUDA(cudaEventElapsedTime(&ms, gp->prof_events[i].first, gp->prof_events[i].second));
      gp->prof_ms += ms;
      gp->prof_flops += gp->prof_event_flops[i];
      gp->prof_launches += 1;
      cudaEventDestroy(gp->prof_events[i].first);
      cudaEventDestroy(gp->prof_events[i].second);
    }
    gp->prof_events.clear();
    gp->prof_event_flops.clear();
  }
  return 0;
}

// Can the joint / gradient paths of this handle run on the single-pass engine (ozaki5.cuh)?  Needs an admitted variance mode;
// the gradient path also needs the V GEMM's own admission (oz5_ensure_kinv).
static int oz5_store_ready(tb_gp* gp, bool need_kinv, bool* ok) {
  *ok = false;
  if (!(gp->engine == 1 && gp->N <= 16384)) return 0;
  TB_TRY(oz5_ensure(gp));
  if (gp->oz5_planes == 0) return 0;
  if (need_kinv) {
    TB_TRY(ensure_kinv_dense(gp));
    TB_TRY(oz5_ensure_kinv(gp));
    if (!gp->kinv5_ok) return 0;
  }
  *ok = true;
  return 0;
}
static inline int oz5_groups(const tb_gp* gp, int tiles) {  // row-block groups per candidate tile: >= 2 items per SM
  return std::max(std::max(1, (gp->NB + 3) / 4), std::min(gp->NB, (2 * 148 + tiles - 1) / tiles));
}

// value + gradient on the single-pass engine: K* digits -> variance GEMM (sum of squares) -> V = K^-1 k* (store GEMM over the
// same K* digits, dense left factor) -> gradient assembly -> tail
static int run_eval_grad_oz5(tb_gp* gp, EvalRequest& rq) {
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D;
  if (rq.want_argmax) {
    TB_CHECK(rq.M > 0, "argmax over an empty candidate set");
    TB_TRY(gp->sRun.reserve(16));
    double init_v = -INFINITY;
    int64_t init_i = INT64_MAX;
    TB_CUDA(cudaMemcpyAsync(gp->sRun.p, &init_v, 8, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync((char*)gp->sRun.p + 8, &init_i, 8, cudaMemcpyHostToDevice, st));
  }
  if (rq.M == 0) return 0;
  TB_CHECK(rq.acq >= 0, "gradients need an acquisition kind");
  const int nt = oz5_tile_width(gp);
  const size_t per_tile = oz5_tile_bytes(gp);
  const int64_t ldv = (int64_t)gp->NB * BM;
  // scratch per chunk: K* digits + V (ldv doubles per candidate): bound both at ~1.3 GB
  int64_t max_tiles = std::max<int64_t>(1, (int64_t)(((size_t)1280 << 20) / std::max(per_tile, (size_t)nt * ldv * sizeof(double))));
  if (max_tiles >= 148) max_tiles = (max_tiles / 148) * 148;
  const int64_t chunk_cap = std::min<int64_t>(max_tiles * nt, ((rq.M + nt - 1) / nt) * nt);
  const int64_t tiles_cap = chunk_cap / nt;
  const bool xc_dev = is_device_ptr(rq.Xc);
  const bool vals_dev = is_device_ptr(rq.out_vals), mean_dev = is_device_ptr(rq.out_mean), var_dev = is_device_ptr(rq.out_var);
  const bool gdev = is_device_ptr(rq.out_grad);
  const int Gcap = gp->NB;
  TB_TRY(gp->sKs.reserve((size_t)tiles_cap * per_tile));
  TB_TRY(gp->sPartial.reserve(sizeof(double) * (size_t)Gcap * chunk_cap));
  TB_TRY(gp->sMean.reserve(sizeof(double) * chunk_cap));
  TB_TRY(gp->sV.reserve((size_t)chunk_cap * ldv * sizeof(double)));
  TB_TRY(gp->sMisc.reserve(sizeof(double) * 2 * chunk_cap));
  if (!xc_dev) TB_TRY(gp->sXc.reserve(sizeof(double) * chunk_cap * D));
  if (rq.out_vals && !vals_dev) TB_TRY(gp->sVals.reserve(sizeof(double) * chunk_cap));
  if (rq.out_var && !var_dev) TB_TRY(gp->sVar.reserve(sizeof(double) * chunk_cap));
  if (!gdev) TB_TRY(gp->sGrad.reserve(sizeof(double) * chunk_cap * D));
  const int tail_blocks_cap = (int)((chunk_cap + 255) / 256);
  if (rq.want_argmax) {
    TB_TRY(gp->sBlkBest.reserve(sizeof(double) * tail_blocks_cap));
    TB_TRY(gp->sBlkIdx.reserve(sizeof(int64_t) * tail_blocks_cap));
  }
  for (int64_t c0 = 0; c0 < rq.M; c0 += chunk_cap) {
    const int64_t mc = std::min<int64_t>(chunk_cap, rq.M - c0);
    const int tiles = (int)((mc + nt - 1) / nt);
    const int64_t McPad = (int64_t)tiles * nt;
    const int G = oz5_groups(gp, tiles);
    const double* xc_chunk;
    if (xc_dev) {
      xc_chunk = rq.Xc + c0 * D;
    } else {
      TB_CUDA(cudaMemcpyAsync(gp->sXc.p, rq.Xc + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, st));
      xc_chunk = gp->sXc.as<double>();
    }
    TB_TRY(oz5_launch_kstar(gp, st, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
    TB_TRY(oz5_launch_gemm(gp, st, gp->sKs.as<int8_t>(), tiles, G, McPad, gp->sPartial.as<double>()));
    double* cmu = gp->sMisc.as<double>();
    acq_partials_kernel<<<(unsigned)((mc + 255) / 256), 256, 0, st>>>(gp->sPartial.as<double>(), G, McPad, gp->sMean.as<double>(), mc,
                                                                      gp->variance, rq.acq, rq.param, gp->noise, gp->dMes.as<double>(), gp->mesS,
                                                                      cmu, cmu + mc);
    TB_LAUNCHED();
    TB_TRY(oz5_launch_gemm_store(gp, st, 1, gp->sKs.as<int8_t>(), tiles, G, gp->sV.as<double>(), ldv));
    double* gd = gdev ? rq.out_grad + c0 * D : gp->sGrad.as<double>();
    switch (gp->kernel) {
      case TB_RBF: launch_grad_dp<TB_RBF>(gp, xc_chunk, mc, gd); break;
      case TB_MATERN12: launch_grad_dp<TB_MATERN12>(gp, xc_chunk, mc, gd); break;
      case TB_MATERN32: launch_grad_dp<TB_MATERN32>(gp, xc_chunk, mc, gd); break;
      default: launch_grad_dp<TB_MATERN52>(gp, xc_chunk, mc, gd); break;
    }
    TB_LAUNCHED();
    TB_CUDA(cudaGetLastError());
    if (!gdev) TB_CUDA(cudaMemcpyAsync(rq.out_grad + c0 * D, gd, sizeof(double) * mc * D, cudaMemcpyDeviceToHost, st));
    double* d_vals = rq.out_vals ? (vals_dev ? rq.out_vals + c0 : gp->sVals.as<double>()) : nullptr;
    double* d_mean = rq.out_mean ? (mean_dev ? rq.out_mean + c0 : nullptr) : nullptr;
    double* d_var = rq.out_var ? (var_dev ? rq.out_var + c0 : gp->sVar.as<double>()) : nullptr;
    const int tb_blocks = (int)((mc + 255) / 256);
    tail_kernel<<<tb_blocks, 256, 0, st>>>(gp->sPartial.as<double>(), G, McPad, gp->sMean.as<double>(), mc, c0, gp->variance, rq.acq,
                                           rq.param, gp->noise, gp->dMes.as<double>(), gp->mesS, d_vals, d_mean, d_var,
                                           rq.want_argmax ? gp->sBlkBest.as<double>() : nullptr,
                                           rq.want_argmax ? gp->sBlkIdx.as<int64_t>() : nullptr);
    TB_LAUNCHED();
    if (rq.want_argmax) {
      argmax_fold_kernel<<<1, 256, 0, st>>>(gp->sBlkBest.as<double>(), gp->sBlkIdx.as<int64_t>(), tb_blocks, gp->sRun.as<double>(),
                                            reinterpret_cast<int64_t*>((char*)gp->sRun.p + 8));
      TB_LAUNCHED();
    }
    if (rq.out_vals && !vals_dev) TB_CUDA(cudaMemcpyAsync(rq.out_vals + c0, gp->sVals.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    if (rq.out_mean && !mean_dev) TB_CUDA(cudaMemcpyAsync(rq.out_mean + c0, gp->sMean.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    if (rq.out_var && !var_dev) TB_CUDA(cudaMemcpyAsync(rq.out_var + c0, gp->sVar.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    if (!xc_dev || (rq.out_vals && !vals_dev) || (rq.out_mean && !mean_dev) || (rq.out_var && !var_dev) || !gdev)
      TB_CUDA(cudaStreamSynchronize(st));  // scratch is reused by the next chunk: host-staged copies must drain first
  }
  if (rq.want_argmax) {
    TB_CUDA(cudaMemcpyAsync(&rq.best_value, gp->sRun.p, 8, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(&rq.best_index, (char*)gp->sRun.p + 8, 8, cudaMemcpyDeviceToHost, st));
  }
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}

static int run_eval(tb_gp* gp, EvalRequest& rq) {
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(rq.M >= 0, "negative candidate count");
  // int32 accumulators of the int8 engine are exact up to K = 16384; larger models use the native fp64 engine
  if (gp->engine == 1 && !rq.out_grad && gp->N <= 16384) return run_eval_oz(gp, rq);
  if (rq.out_grad && rq.acq >= 0) {
    bool fast = false;
    TB_CUDA(cudaSetDevice(gp->device));
    TB_TRY(oz5_store_ready(gp, true, &fast));
    if (fast) return run_eval_grad_oz5(gp, rq);
  }
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D;
  if (rq.want_argmax) {
    TB_CHECK(rq.M > 0, "argmax over an empty candidate set");
    TB_TRY(gp->sRun.reserve(16));
    double init_v = -INFINITY;  // a candidate worth -inf still beats "nothing seen" through the lower-index tie rule
    int64_t init_i = INT64_MAX;
    TB_CUDA(cudaMemcpyAsync(gp->sRun.p, &init_v, 8, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync((char*)gp->sRun.p + 8, &init_i, 8, cudaMemcpyHostToDevice, st));
  }
  if (rq.M == 0) return 0;

  const bool xc_dev = is_device_ptr(rq.Xc);
  const bool vals_dev = is_device_ptr(rq.out_vals), mean_dev = is_device_ptr(rq.out_mean),
             var_dev = is_device_ptr(rq.out_var);
  const int64_t max_tiles = chunk_tiles(gp);
  const int64_t Mc_max = max_tiles * BT;
  const int64_t chunk_cap = std::min<int64_t>(Mc_max, ((rq.M + BT - 1) / BT) * BT);
  const int64_t tiles_cap = chunk_cap / BT;
  const int Gmax = gp->NB;  // upper bound of every group choice below

  const bool ozaki = gp->engine == 1 && gp->N <= 16384;  // here: only reached with a gradient request
  if (ozaki) {
    TB_TRY(ensure_ozaki(gp));
    if (rq.out_grad) TB_TRY(ensure_kinv_digits(gp));
  }
  TB_TRY(gp->sKs.reserve(std::max((size_t)tiles_cap * gp->nkc * PANEL * sizeof(double),
                                  (size_t)tiles_cap * gp->nst * oz::S * oz::TILE)));
  TB_TRY(gp->sPartial.reserve(sizeof(double) * (size_t)Gmax * chunk_cap));
  TB_TRY(gp->sMean.reserve(sizeof(double) * chunk_cap));
  if (!xc_dev) TB_TRY(gp->sXc.reserve(sizeof(double) * chunk_cap * D));
  if (rq.out_vals && !vals_dev) TB_TRY(gp->sVals.reserve(sizeof(double) * chunk_cap));
  if (rq.out_var && !var_dev) TB_TRY(gp->sVar.reserve(sizeof(double) * chunk_cap));
  if (rq.out_grad) {
    TB_CHECK(rq.acq >= 0, "gradients need an acquisition kind");
    if (!ozaki) {
      TB_TRY(ensure_upper_panels(gp));
      TB_TRY(gp->sA.reserve((size_t)tiles_cap * gp->NB * (BM / BK) * PANEL * sizeof(double)));
    }
    TB_TRY(gp->sV.reserve((size_t)chunk_cap * gp->NB * BM * sizeof(double)));
    TB_TRY(gp->sMisc.reserve(sizeof(double) * 2 * chunk_cap));
    if (!is_device_ptr(rq.out_grad)) TB_TRY(gp->sGrad.reserve(sizeof(double) * chunk_cap * D));
  }
  const int tail_blocks_cap = (int)((chunk_cap + 255) / 256);
  if (rq.want_argmax) {
    TB_TRY(gp->sBlkBest.reserve(sizeof(double) * tail_blocks_cap));
    TB_TRY(gp->sBlkIdx.reserve(sizeof(int64_t) * tail_blocks_cap));
  }

  for (int64_t c0 = 0; c0 < rq.M; c0 += chunk_cap) {
    const int64_t mc = std::min<int64_t>(chunk_cap, rq.M - c0);
    const int tiles = (int)((mc + BT - 1) / BT);
    const int64_t McPad = (int64_t)tiles * BT;
    // int8 engine: one serpentine PAIR of row-blocks per CTA, so the ~150 co-resident CTAs touch only ~9 candidate tiles
    // and their K* digit tiles are re-read from L2 instead of HBM (ncu: 28 GB -> ~1 GB of DRAM reads per launch)
    const int G = ozaki ? std::max(std::max(1, (gp->NB + 3) / 4), std::min(gp->NB, (2 * 148 + tiles - 1) / tiles)) : pick_groups(gp, tiles);

    const double* xc_chunk;
    if (xc_dev) {
      xc_chunk = rq.Xc + c0 * D;
    } else {
      TB_CUDA(cudaMemcpyAsync(gp->sXc.p, rq.Xc + c0 * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, st));
      xc_chunk = gp->sXc.as<double>();
    }
    const bool use_oz = ozaki;
    if (use_oz)
      TB_TRY(launch_kstar_digits(gp, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
    else
      TB_TRY(launch_kstar(gp, xc_chunk, mc, tiles, gp->sKs.as<double>(), gp->sMean.as<double>()));

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (gp->profile) {
      TB_CUDA(cudaEventCreate(&e0));
      TB_CUDA(cudaEventCreate(&e1));
      TB_CUDA(cudaEventRecord(e0, st));
    }
    if (use_oz)
      oz::trigemm_i8_kernel<oz::OZ_SUMSQ, 8><<<dim3(G, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
          gp->dAS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, G, McPad,
          gp->oz_out_scale, oz_npass(gp), 0, gp->sPartial.as<double>(), nullptr, 0);
    else if (rq.out_grad)
      trigemm_kernel<false, EPI_SUMSQ_PACKED><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, gp->sPartial.as<double>(),
          gp->sA.as<double>(), nullptr, 0);
    else
      trigemm_kernel<false, EPI_SUMSQ><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, gp->sPartial.as<double>(),
          nullptr, nullptr, 0);
    TB_LAUNCHED();
    if (gp->profile) {
      TB_CUDA(cudaEventRecord(e1, st));
      gp->prof_events.emplace_back(e0, e1);
      gp->prof_event_flops.push_back((double)McPad * (double)gp->N * (double)gp->N);
    }
    TB_CUDA(cudaGetLastError());

    if (rq.out_grad) {
      if (use_oz)
        TB_TRY(gradient_chunk_oz(gp, rq.acq, rq.param, xc_chunk, mc, tiles, G, McPad, rq.out_grad + c0 * D));
      else
        TB_TRY(gradient_chunk(gp, rq.acq, rq.param, xc_chunk, mc, tiles, G, McPad, rq.out_grad + c0 * D));
    }

    double* d_vals = rq.out_vals ? (vals_dev ? rq.out_vals + c0 : gp->sVals.as<double>()) : nullptr;
    double* d_mean = rq.out_mean ? (mean_dev ? rq.out_mean + c0 : nullptr) : nullptr;  // sMean already holds it
    double* d_var = rq.out_var ? (var_dev ? rq.out_var + c0 : gp->sVar.as<double>()) : nullptr;
    const int tb_blocks = (int)((mc + 255) / 256);
    tail_kernel<<<tb_blocks, 256, 0, st>>>(gp->sPartial.as<double>(), G, McPad, gp->sMean.as<double>(), mc, c0,
                                           gp->variance, rq.acq, rq.param, gp->noise, gp->dMes.as<double>(), gp->mesS, d_vals, d_mean, d_var,
                                           rq.want_argmax ? gp->sBlkBest.as<double>() : nullptr,
                                           rq.want_argmax ? gp->sBlkIdx.as<int64_t>() : nullptr);
    TB_LAUNCHED();
    if (rq.want_argmax) {
      argmax_fold_kernel<<<1, 256, 0, st>>>(gp->sBlkBest.as<double>(), gp->sBlkIdx.as<int64_t>(), tb_blocks,
                                            gp->sRun.as<double>(), reinterpret_cast<int64_t*>((char*)gp->sRun.p + 8));
      TB_LAUNCHED();
    }
    if (rq.out_vals && !vals_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_vals + c0, gp->sVals.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    if (rq.out_mean && !mean_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_mean + c0, gp->sMean.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    if (rq.out_var && !var_dev)
      TB_CUDA(cudaMemcpyAsync(rq.out_var + c0, gp->sVar.p, sizeof(double) * mc, cudaMemcpyDeviceToHost, st));
    // scratch is reused by the next chunk: host-staged copies must drain first
    if (!xc_dev || (rq.out_vals && !vals_dev) || (rq.out_mean && !mean_dev) || (rq.out_var && !var_dev) ||
        (rq.out_grad && !is_device_ptr(rq.out_grad)))
      TB_CUDA(cudaStreamSynchronize(st));
  }
  if (rq.want_argmax) {
    TB_CUDA(cudaMemcpyAsync(&rq.best_value, gp->sRun.p, 8, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(&rq.best_index, (char*)gp->sRun.p + 8, 8, cudaMemcpyDeviceToHost, st));
  }
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  if (gp->profile) {
    for (size_t i = 0; i < gp->prof_events.size(); ++i) {
      float ms = 0.f;
      TB_CUDA(cudaEventElapsedTime(&ms, gp->prof_events[i].first, gp->prof_events[i].second));
      gp->prof_ms += ms;
      gp->prof_flops += gp->prof_event_flops[i];
      gp->prof_launches += 1;
      cudaEventDestroy(gp->prof_events[i].first);
      cudaEventDestroy(gp->prof_events[i].second);
    }
    gp->prof_events.clear();
    gp->prof_event_flops.clear();
  }
  return 0;
}

}  // namespace tb

extern "C" {

static int tb_gp_predict_f64(tb_gp* gp, const void* Xc, int64_t M, void* mean, void* var) {
  TB_CHECK(gp && (M == 0 || (Xc && mean && var)), "tb_gp_predict: null argument");
  tb::EvalRequest rq;
  rq.Xc = (const double*)Xc;
  rq.M = M;
  rq.out_mean = (double*)mean;
  rq.out_var = (double*)var;
  return tb::run_eval(gp, rq);
}

static int tb_acq_eval_f64(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out, void* grad) {
  TB_CHECK(gp && (M == 0 || (Xc && out)), "tb_acq_eval: null argument");
  TB_CHECK(acq >= TB_ACQ_EI && acq <= TB_ACQ_MES, "tb_acq_eval: unknown acquisition kind");
  if (acq == TB_ACQ_LCB || acq == TB_ACQ_NEG_LCB)
    TB_CHECK(param >= 0.0, "Standard deviation scaling parameter beta must not be negative");
  if (acq == TB_ACQ_MES) TB_CHECK(gp->mesS > 0, "min-value entropy search: set the min-value samples first (tb_acq_set_min_value_samples)");
  tb::EvalRequest rq;
  rq.acq = acq;
  rq.param = param;
  rq.Xc = (const double*)Xc;
  rq.M = M;
  rq.out_vals = (double*)out;
  rq.out_grad = (double*)grad;
  return tb::run_eval(gp, rq);
}

static int tb_acq_argmax_f64(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out, void* best_value,
                  int64_t* best_index) {
  TB_CHECK(gp && Xc && best_value && best_index, "tb_acq_argmax: null argument");
  TB_CHECK(acq >= TB_ACQ_EI && acq <= TB_ACQ_MES, "tb_acq_argmax: unknown acquisition kind");
  if (acq == TB_ACQ_LCB || acq == TB_ACQ_NEG_LCB)
    TB_CHECK(param >= 0.0, "Standard deviation scaling parameter beta must not be negative");
  if (acq == TB_ACQ_MES) TB_CHECK(gp->mesS > 0, "min-value entropy search: set the min-value samples first (tb_acq_set_min_value_samples)");
  tb::EvalRequest rq;
  rq.acq = acq;
  rq.param = param;
  rq.Xc = (const double*)Xc;
  rq.M = M;
  rq.out_vals = (double*)out;
  rq.want_argmax = true;
  TB_TRY(tb::run_eval(gp, rq));
  if (rq.best_index == INT64_MAX) {  // every value was NaN: tf.math.argmax still returns a valid index (optimizer.py:149)
    rq.best_index = 0;
    rq.best_value = std::nan("");
  }
  *(double*)best_value = rq.best_value;
  *best_index = rq.best_index;
  return 0;
}

int tb_acq_set_min_value_samples(tb_gp* gp, const double* samples, int S) {
  TB_CHECK(gp && samples, "tb_acq_set_min_value_samples: null argument");
  TB_CHECK(S > 0, "tb_acq_set_min_value_samples: need at least one sample");
  TB_CUDA(cudaSetDevice(gp->device));
  TB_TRY(gp->dMes.reserve(sizeof(double) * (size_t)S));
  TB_CUDA(cudaMemcpyAsync(gp->dMes.p, samples, sizeof(double) * (size_t)S, cudaMemcpyDefault, gp->stream));
  TB_CUDA(cudaStreamSynchronize(gp->stream));
  gp->mesS = S;
  return 0;
}

int tb_gp_profile(tb_gp* gp, int enable) {
  TB_CHECK(gp, "tb_gp_profile: null handle");
  gp->profile = enable != 0;
  gp->prof_ms = 0.0;
  gp->prof_flops = 0.0;
  gp->prof_launches = 0;
  return 0;
}
int tb_gp_set_engine(tb_gp* gp, int engine) {
  TB_CHECK(gp, "tb_gp_set_engine: null handle");
  TB_CHECK(engine >= 0 && engine <= 2, "tb_gp_set_engine: engine must be 0 (fp64 DMMA), 1 (int8 Ozaki) or 2 (int8, full 21 products)");
  gp->engine = engine == 0 ? 0 : 1;
  if (gp->oz_full != (engine == 2)) gp->oz5_valid = false;
  gp->oz_full = engine == 2;
  return 0;
}
int tb_gp_engine_info(tb_gp* gp, int* digit_products, double* error_estimate) {
  TB_CHECK(gp, "tb_gp_engine_info: null handle");
  TB_CHECK(gp->cache_valid, "tb_gp_engine_info: posterior cache is not built");
  int products = 0;
  double est = 0.0;
  if (gp->engine == 1 && gp->N <= 16384) {
    TB_CUDA(cudaSetDevice(gp->device));
    TB_TRY(oz5_ensure(gp));
    if (gp->oz5_mode == 5) products = 15;
    else if (gp->oz5_mode == 3) products = 6;
    else products = gp->dtype == TB_F32 ? 10 : 21;
    est = gp->oz5_est;
  }
  if (digit_products) *digit_products = products;
  if (error_estimate) *error_estimate = est;
  return 0;
}
int tb_gp_kinv_apply(tb_gp* gp, const double* B, int nrhs, double* out) {
  TB_CHECK(gp && B && out, "tb_gp_kinv_apply: null argument");
  TB_CHECK(gp->cache_valid, "tb_gp_kinv_apply: posterior cache is not built");
  TB_CHECK(nrhs >= 1, "tb_gp_kinv_apply: need at least one right-hand side");
  TB_CUDA(cudaSetDevice(gp->device));
  const int64_t N = gp->N;
  cudaStream_t st = gp->stream;
  TB_TRY(gp->sMisc.reserve(sizeof(double) * 2 * N * nrhs));
  double* rhs = gp->sMisc.as<double>();
  double* tmp = rhs + N * nrhs;
  TB_CUDA(cudaMemcpyAsync(rhs, B, sizeof(double) * N * nrhs, cudaMemcpyDefault, st));
  // (K + noise I)^-1 B = Linv^T (Linv B) through the cached triangular inverse: two batched triangular mat-vecs
  // (once per trajectory, off the candidate path)
  trmv_lower_cols_kernel<<<dim3((unsigned)((N + 127) / 128), (unsigned)nrhs), 128, 0, st>>>(gp->dLinv.as<double>(), N, N, rhs, N, tmp, N);
  TB_LAUNCHED();
  trmv_lower_t_cols_kernel<<<dim3((unsigned)((N + 7) / 8), (unsigned)nrhs), 256, 0, st>>>(gp->dLinv.as<double>(), N, N, tmp, N, rhs, N);
  TB_LAUNCHED();
  TB_CUDA(cudaMemcpyAsync(out, rhs, sizeof(double) * N * nrhs, cudaMemcpyDefault, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}
int tb_gp_stream(tb_gp* gp, void** stream) {
  TB_CHECK(gp && stream, "tb_gp_stream: null argument");
  *stream = (void*)gp->stream;
  return 0;
}
int tb_gp_profile_read(tb_gp* gp, double* trigemm_ms, int64_t* trigemm_launches, double* flops) {
  TB_CHECK(gp, "tb_gp_profile_read: null handle");
  if (trigemm_ms) *trigemm_ms = gp->prof_ms;
  if (trigemm_launches) *trigemm_launches = gp->prof_launches;
  if (flops) *flops = gp->prof_flops;
  return 0;
}

}  // extern "C"

// =================================================================================================
// joint posterior of q-batches: predict_joint / reparam samples / MC-qEI
// =================================================================================================
namespace tb {

struct JointRequest {
  int mode = JOINT_PREDICT;
  const double* Xc = nullptr;  // [B, q, D]
  int64_t B = 0;
  int q = 0;
  const double* eps = nullptr;  // [q, S] host or device
  int S = 0;
  double eta = 0.0, jitter = 0.0;
  double* out_mean = nullptr;     // [B, q]
  double* out_cov = nullptr;      // [B, q, q]
  double* out_samples = nullptr;  // [B, S, q]
  double* out_qei = nullptr;      // [B]
};

template <int KIND>
static int launch_joint(tb_gp* gp, int QT, int blocks, size_t smem, const double* A, int64_t lda, int Nrows,
                        const double* mean, const double* xc, int64_t nb, const JointRequest& rq, const double* eps_dev,
                        double* om, double* oc, double* os, double* oq, int* err) {
  const double* il = gp->dInvLs.as<double>();
#define TB_JOINT(QTV)                                                                                              \
  {                                                                                                                \
    TB_CUDA(cudaFuncSetAttribute(joint_kernel<KIND, QTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    joint_kernel<KIND, QTV><<<blocks, JOINT_WARPS * 32, smem, gp->stream>>>(A, lda, Nrows, mean, xc, il, gp->D, nb, rq.q, \
        gp->variance, rq.mode, eps_dev, rq.S, rq.eta, rq.jitter, om, oc, os, oq, err);                               \
  }
  switch (QT) {
    case 1: TB_JOINT(1); break;
    case 2: TB_JOINT(2); break;
    case 3: TB_JOINT(3); break;
    default: TB_JOINT(4); break;
  }
#undef TB_JOINT
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  return 0;
}

static int run_joint(tb_gp* gp, JointRequest& rq) {
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(rq.q >= 1 && rq.q <= 32, "batch size q must be in [1, 32]");
  TB_CHECK(rq.B >= 0, "negative batch count");
  if (rq.mode != JOINT_PREDICT) {
    TB_CHECK(rq.S >= 1 && rq.eps, "need S >= 1 base samples");
    TB_CHECK(rq.jitter >= 0.0, "jitter must be non-negative");
  }
  if (rq.B == 0) return 0;
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D, q = rq.q;
  const int QT = (q + 7) / 8, QP = QT * 8;
  const int64_t lda = (int64_t)gp->NB * BM;

  const int64_t max_tiles = chunk_tiles(gp);
  int64_t nbc_cap = std::max<int64_t>(1, (max_tiles * BT) / q);  // whole batches per chunk
  nbc_cap = std::min<int64_t>(nbc_cap, rq.B);
  const int64_t cand_cap = nbc_cap * q;
  const bool oz_joint = gp->engine == 1 && gp->N <= 16384;
  bool fast = false;  // single-pass engine (ozaki5.cuh) for A = Linv K*
  TB_TRY(oz5_store_ready(gp, false, &fast));
  const int nt = fast ? oz5_tile_width(gp) : BT;  // candidates per tile
  const int64_t tiles_cap = (cand_cap + nt - 1) / nt;
  if (oz_joint && !fast) {
    TB_TRY(ensure_ozaki(gp));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  TB_TRY(gp->sKs.reserve(fast ? (size_t)tiles_cap * oz5_tile_bytes(gp)
                              : std::max((size_t)tiles_cap * gp->nkc * PANEL * sizeof(double), (size_t)tiles_cap * gp->nst * oz::S * oz::TILE)));
  TB_TRY(gp->sV.reserve((size_t)tiles_cap * nt * lda * sizeof(double)));  // A plain
  TB_TRY(gp->sMean.reserve(sizeof(double) * tiles_cap * nt));
  const bool xc_dev = is_device_ptr(rq.Xc);
  if (!xc_dev) TB_TRY(gp->sXc.reserve(sizeof(double) * cand_cap * D));
  const double* eps_dev = nullptr;
  if (rq.mode != JOINT_PREDICT) {
    if (is_device_ptr(rq.eps)) {
      eps_dev = rq.eps;
    } else {
      TB_TRY(gp->sMisc.reserve(sizeof(double) * (size_t)q * rq.S + 64));
      TB_CUDA(cudaMemcpyAsync(gp->sMisc.p, rq.eps, sizeof(double) * (size_t)q * rq.S, cudaMemcpyHostToDevice, st));
      eps_dev = gp->sMisc.as<double>();
    }
  }
  TB_TRY(gp->sRun.reserve(16));
  int* err = reinterpret_cast<int*>(gp->sRun.p);
  TB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  // host-staged outputs
  struct Out { double* user; size_t per_batch; tb::DevBuf* buf; };
  Out outs[4] = {{rq.out_mean, (size_t)q, &gp->sVals}, {rq.out_cov, (size_t)q * q, &gp->sVar},
                 {rq.out_samples, (size_t)rq.S * q, &gp->sGrad}, {rq.out_qei, 1, &gp->sBlkBest}};
  bool any_host_out = false;
  for (auto& o : outs)
    if (o.user && !is_device_ptr(o.user)) {
      TB_TRY(o.buf->reserve(sizeof(double) * o.per_batch * nbc_cap));
      any_host_out = true;
    }
  const size_t smem = (size_t)JOINT_WARPS * (QP * QP + QP * D + QP) * sizeof(double);

  for (int64_t b0 = 0; b0 < rq.B; b0 += nbc_cap) {
    const int64_t nbc = std::min<int64_t>(nbc_cap, rq.B - b0);
    const int64_t mc = nbc * q;
    const int tiles = (int)((mc + nt - 1) / nt);
    const int64_t McPad = (int64_t)tiles * nt;
    const int G = pick_groups(gp, tiles);
    const double* xc_chunk;
    if (xc_dev) {
      xc_chunk = rq.Xc + b0 * q * D;
    } else {
      TB_CUDA(cudaMemcpyAsync(gp->sXc.p, rq.Xc + b0 * q * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, st));
      xc_chunk = gp->sXc.as<double>();
    }
    if (fast) {
      // A = Linv K* as 15 (fp32 models: 10) exact digit products in one pass, stored for the per-batch Gram kernel
      TB_TRY(oz5_launch_kstar(gp, st, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
      TB_TRY(oz5_launch_gemm_store(gp, st, 0, gp->sKs.as<int8_t>(), tiles, oz5_groups(gp, tiles), gp->sV.as<double>(), lda));
    } else if (oz_joint) {
      // A = Linv K* on the int8 tensor cores (fp64-accurate digit GEMM), stored for the per-batch Gram kernel
      TB_TRY(launch_kstar_digits(gp, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
      const int Goz = std::max(std::max(1, (gp->NB + 3) / 4), std::min(gp->NB, (2 * 148 + tiles - 1) / tiles));
      oz::trigemm_i8_kernel<oz::OZ_STORE, 8><<<dim3(Goz, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
          gp->dAS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, Goz, McPad,
          gp->oz_out_scale, oz_npass(gp), 0, nullptr, gp->sV.as<double>(), lda);
    } else {
      TB_TRY(launch_kstar(gp, xc_chunk, mc, tiles, gp->sKs.as<double>(), gp->sMean.as<double>()));
      trigemm_kernel<false, EPI_PLAIN><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, nullptr, nullptr,
          gp->sV.as<double>(), lda);
    }
    TB_LAUNCHED();
    TB_CUDA(cudaGetLastError());
    double* dptr[4];
    for (int i = 0; i < 4; ++i) {
      Out& o = outs[i];
      dptr[i] = !o.user ? nullptr : (is_device_ptr(o.user) ? o.user + b0 * o.per_batch : o.buf->as<double>());
    }
    const int blocks = (int)((nbc + JOINT_WARPS - 1) / JOINT_WARPS);
    const int Nrows = (int)lda;
    switch (gp->kernel) {
      case TB_RBF: TB_TRY(launch_joint<TB_RBF>(gp, QT, blocks, smem, gp->sV.as<double>(), lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, rq, eps_dev, dptr[0], dptr[1], dptr[2], dptr[3], err)); break;
      case TB_MATERN12: TB_TRY(launch_joint<TB_MATERN12>(gp, QT, blocks, smem, gp->sV.as<double>(), lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, rq, eps_dev, dptr[0], dptr[1], dptr[2], dptr[3], err)); break;
      case TB_MATERN32: TB_TRY(launch_joint<TB_MATERN32>(gp, QT, blocks, smem, gp->sV.as<double>(), lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, rq, eps_dev, dptr[0], dptr[1], dptr[2], dptr[3], err)); break;
      default: TB_TRY(launch_joint<TB_MATERN52>(gp, QT, blocks, smem, gp->sV.as<double>(), lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, rq, eps_dev, dptr[0], dptr[1], dptr[2], dptr[3], err)); break;
    }
    for (auto& o : outs)
      if (o.user && !is_device_ptr(o.user))
        TB_CUDA(cudaMemcpyAsync(o.user + b0 * o.per_batch, o.buf->p, sizeof(double) * o.per_batch * nbc,
                                cudaMemcpyDeviceToHost, st));
    if (!xc_dev || any_host_out) TB_CUDA(cudaStreamSynchronize(st));
  }
  int herr = 0;
  TB_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  TB_CHECK_CODE(herr == 0, "Cholesky decomposition was not successful. The input might not be valid "
                      "(covariance + jitter*I of a query batch is not positive definite)", tb::ERR_NUMERIC);
  return 0;
}

// A = Linv K(X, Xc) for M device-resident points, stored plain ([point][lda], lda = NB*128) in gp->sA; posterior means in
// gp->sMean.  One launch over all M points (callers bound M).
static int compute_a_plain(tb_gp* gp, const double* xc_dev, int64_t M) {
  cudaStream_t st = gp->stream;
  const int64_t lda = (int64_t)gp->NB * BM;
  const bool oz_path = gp->engine == 1 && gp->N <= 16384;
  bool fast = false;
  TB_TRY(oz5_store_ready(gp, false, &fast));
  const int nt = fast ? oz5_tile_width(gp) : BT;
  const int tiles = (int)((M + nt - 1) / nt);
  const int64_t McPad = (int64_t)tiles * nt;
  if (oz_path && !fast) TB_TRY(ensure_ozaki(gp));
  TB_TRY(gp->sKs.reserve(fast ? (size_t)tiles * oz5_tile_bytes(gp)
                              : std::max((size_t)tiles * gp->nkc * PANEL * sizeof(double), (size_t)tiles * gp->nst * oz::S * oz::TILE)));
  TB_TRY(gp->sA.reserve((size_t)McPad * lda * sizeof(double)));
  TB_TRY(gp->sMean.reserve(sizeof(double) * McPad));
  if (fast) {
    TB_TRY(oz5_launch_kstar(gp, st, xc_dev, M, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
    TB_TRY(oz5_launch_gemm_store(gp, st, 0, gp->sKs.as<int8_t>(), tiles, oz5_groups(gp, tiles), gp->sA.as<double>(), lda));
    return 0;
  }
  if (oz_path) {
    TB_TRY(launch_kstar_digits(gp, xc_dev, M, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
    const int Goz = std::max(std::max(1, (gp->NB + 3) / 4), std::min(gp->NB, (2 * 148 + tiles - 1) / tiles));
    oz::trigemm_i8_kernel<oz::OZ_STORE, 8><<<dim3(Goz, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
        gp->dAS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, Goz, McPad, gp->oz_out_scale,
        oz_npass(gp), 0, nullptr, gp->sA.as<double>(), lda);
  } else {
    const int G = pick_groups(gp, tiles);
    TB_TRY(launch_kstar(gp, xc_dev, M, tiles, gp->sKs.as<double>(), gp->sMean.as<double>()));
    trigemm_kernel<false, EPI_PLAIN><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
        gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, nullptr, nullptr, gp->sA.as<double>(), lda);
  }
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  return 0;
}

// out[i][j] = k(x1_i, x2_j) - sum_k A1[i][k] A2[j][k]: posterior covariance between two point sets, row-major [M1, M2]
// (covariance_between_points_encoded, models/gpflow/models.py:188-254: K12 - Kx1 (K + s^2 I)^-1 Kx2, no clipping)
template <int KIND>
__global__ void __launch_bounds__(fac::THREADS)
cross_cov_kernel(const double* __restrict__ A1, const double* __restrict__ A2, int64_t lda, int Nk, const double* __restrict__ X1,
                 const double* __restrict__ X2, const double* __restrict__ inv_ls, int D, int64_t M1, int64_t M2, double variance,
                 double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const int64_t r0 = (int64_t)blockIdx.y * fac::FB, c0 = (int64_t)blockIdx.x * fac::FB;
  double acc[8][4][2];
  fac::zero_acc(acc);
  fac::dmma_tile<true, true>([&](int m, int k) { return (r0 + m < M1) ? A1[(r0 + m) * lda + k] : 0.0; },
                             [&](int k, int n) { return (c0 + n < M2) ? A2[(c0 + n) * lda + k] : 0.0; }, 0, Nk, acc, sm);
  fac::for_each_acc(acc, [&](int m, int n, double& v) {
    const int64_t i = r0 + m, j = c0 + n;
    if (i >= M1 || j >= M2) return;
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) {
      const double df = (X1[i * D + d] - X2[j * D + d]) * inv_ls[d];
      r2 = fma(df, df, r2);
    }
    out[i * M2 + j] = kernel_from_r2<KIND>(r2, variance) - v;
  });
}

static int run_cross_cov(tb_gp* gp, const double* X1, int64_t M1, const double* X2, int64_t M2, double* out) {
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(M1 >= 1 && M2 >= 1 && M1 + M2 <= 16384, "tb_gp_covariance_between_points: between 1 and 16384 points in total");
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D;
  const int64_t lda = (int64_t)gp->NB * BM, M = M1 + M2;
  tb::DevBuf bx, bout;
  struct Release {
    std::vector<tb::DevBuf*> v;
    ~Release() { for (auto* b : v) b->release(); }
  } rel{{&bx, &bout}};
  TB_TRY(bx.reserve(sizeof(double) * M * D));  // [X1; X2] contiguous on the device
  TB_CUDA(cudaMemcpyAsync(bx.p, X1, sizeof(double) * M1 * D, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(bx.as<double>() + M1 * D, X2, sizeof(double) * M2 * D, cudaMemcpyDefault, st));
  TB_TRY(compute_a_plain(gp, bx.as<double>(), M));
  const bool out_dev = is_device_ptr(out);
  double* od = out;
  if (!out_dev) {
    TB_TRY(bout.reserve(sizeof(double) * (size_t)M1 * M2));
    od = bout.as<double>();
  }
  const double* A1 = gp->sA.as<double>();
  const double* A2 = A1 + M1 * lda;
  const double* x1 = bx.as<double>();
  const double* x2 = x1 + M1 * D;
  const double* il = gp->dInvLs.as<double>();
  const dim3 grid((unsigned)((M2 + fac::FB - 1) / fac::FB), (unsigned)((M1 + fac::FB - 1) / fac::FB));
#define TB_CCOV(KIND)                                                                                                        \
  TB_CUDA(cudaFuncSetAttribute(cross_cov_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM)); \
  cross_cov_kernel<KIND><<<grid, fac::THREADS, fac::GEMM_SMEM, st>>>(A1, A2, lda, (int)lda, x1, x2, il, D, M1, M2, gp->variance, od)
  switch (gp->kernel) {
    case TB_RBF: TB_CCOV(TB_RBF); break;
    case TB_MATERN12: TB_CCOV(TB_MATERN12); break;
    case TB_MATERN32: TB_CCOV(TB_MATERN32); break;
    default: TB_CCOV(TB_MATERN52); break;
  }
#undef TB_CCOV
  TB_LAUNCHED();
  if (!out_dev) TB_CUDA(cudaMemcpyAsync(out, od, sizeof(double) * (size_t)M1 * M2, cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}

// ---- joint samples over a LARGE point set (interface.py:135-138 -> gpflow predict_f_samples; the ExactThompsonSampler's
// model.sample, acquisition/sampler.py:85-123): full posterior covariance, blocked Cholesky, mean + L z ----
// cov[i][j] = k(x_i, x_j) - sum_k A[i][k] A[j][k]  (+ jitter on the clipped diagonal); lower 128-tiles, mirrored
template <int KIND>
__global__ void __launch_bounds__(fac::THREADS)
posterior_cov_kernel(const double* __restrict__ A, int64_t lda, int Nk, const double* __restrict__ Xc,
                     const double* __restrict__ inv_ls, int D, int64_t M, double variance, double jitter,
                     double* __restrict__ cov) {
  extern __shared__ __align__(16) double sm[];
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (ti < tj) return;
  const int64_t r0 = (int64_t)ti * fac::FB, c0 = (int64_t)tj * fac::FB;
  double acc[8][4][2];
  fac::zero_acc(acc);
  fac::dmma_tile<true, true>([&](int m, int k) { return (r0 + m < M) ? A[(r0 + m) * lda + k] : 0.0; },
                             [&](int k, int n) { return (c0 + n < M) ? A[(c0 + n) * lda + k] : 0.0; }, 0, Nk, acc, sm);
  fac::for_each_acc(acc, [&](int m, int n, double& v) {
    const int64_t i = r0 + m, j = c0 + n;
    if (i >= M || j >= M || i < j) return;
    double val;
    if (i == j) {
      val = fmax(variance - v, 1e-12) + jitter;
    } else {
      double r2 = 0.0;
      for (int d = 0; d < D; ++d) {
        const double df = (Xc[i * D + d] - Xc[j * D + d]) * inv_ls[d];
        r2 = fma(df, df, r2);
      }
      val = kernel_from_r2<KIND>(r2, variance) - v;
    }
    cov[i + j * M] = val;
    cov[j + i * M] = val;
  });
}

// out[s][i] += mean[i]
__global__ void add_mean_rows_kernel(double* __restrict__ out, const double* __restrict__ mean, int64_t M, int64_t total) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < total) out[e] += mean[e % M];
}

static int run_sample_joint(tb_gp* gp, const double* Xc, int64_t M, const double* z, int S, double jitter, double* out) {
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(M >= 1 && M <= 16384, "tb_gp_sample_joint: between 1 and 16384 points");
  TB_CHECK(S >= 1 && z && out && Xc, "tb_gp_sample_joint: need S >= 1 standard-normal draws per point");
  TB_CHECK(jitter >= 0.0, "jitter must be non-negative");
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D;
  const int64_t lda = (int64_t)gp->NB * BM;
  tb::DevBuf bx, bcov, bz, bout, bdinv;
  struct Release {
    std::vector<tb::DevBuf*> v;
    ~Release() { for (auto* b : v) b->release(); }
  } rel{{&bx, &bcov, &bz, &bout, &bdinv}};
  const double* xc = Xc;
  if (!is_device_ptr(Xc)) {
    TB_TRY(bx.reserve(sizeof(double) * M * D));
    TB_CUDA(cudaMemcpyAsync(bx.p, Xc, sizeof(double) * M * D, cudaMemcpyHostToDevice, st));
    xc = bx.as<double>();
  }
  TB_TRY(compute_a_plain(gp, xc, M));
  TB_TRY(bcov.reserve(sizeof(double) * M * M));
  {
    const unsigned t = (unsigned)((M + fac::FB - 1) / fac::FB);
    const double* A = gp->sA.as<double>();
    const double* il = gp->dInvLs.as<double>();
    double* cov = bcov.as<double>();
#define TB_PCOV(KIND)                                                                                                            \
  TB_CUDA(cudaFuncSetAttribute(posterior_cov_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fac::GEMM_SMEM)); \
  posterior_cov_kernel<KIND><<<dim3(t, t), fac::THREADS, fac::GEMM_SMEM, st>>>(A, lda, (int)lda, xc, il, D, M, gp->variance, jitter, cov)
    switch (gp->kernel) {
      case TB_RBF: TB_PCOV(TB_RBF); break;
      case TB_MATERN12: TB_PCOV(TB_MATERN12); break;
      case TB_MATERN32: TB_PCOV(TB_MATERN32); break;
      default: TB_PCOV(TB_MATERN52); break;
    }
#undef TB_PCOV
    TB_LAUNCHED();
  }
  // blocked Cholesky of the covariance with the cache-build kernels (factor.cuh)
  const int nbk = (int)((M + fac::FB - 1) / fac::FB);
  TB_TRY(bdinv.reserve(sizeof(double) * (size_t)nbk * fac::FB * fac::FB));
  TB_TRY(gp->dInfo.reserve(sizeof(int)));
  TB_CUDA(cudaMemsetAsync(gp->dInfo.p, 0, sizeof(int), st));
  {
    double* C = bcov.as<double>();
    const size_t diag_smem = sizeof(double) * fac::FB * (fac::FB + 1);
    for (int jb = 0; jb < nbk; ++jb) {
      const int j0 = jb * fac::FB;
      fac::chol_diag_kernel<<<1, fac::THREADS, diag_smem, st>>>(C, M, j0, bdinv.as<double>(), gp->dInfo.as<int>());
      TB_LAUNCHED();
      const int64_t below = M - (int64_t)(j0 + fac::FB);
      if (below > 0) {
        const unsigned t = (unsigned)((below + fac::FB - 1) / fac::FB);
        fac::chol_panel_kernel<<<t, fac::THREADS, fac::GEMM_SMEM, st>>>(C, M, j0, bdinv.as<double>());
        TB_LAUNCHED();
        fac::chol_syrk_kernel<<<dim3(t, t), fac::THREADS, fac::GEMM_SMEM, st>>>(C, M, j0);
        TB_LAUNCHED();
      }
    }
  }
  int info = 0;
  TB_CUDA(cudaMemcpyAsync(&info, gp->dInfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  TB_CHECK_CODE(info == 0, "Cholesky decomposition was not successful. The input might not be valid "
                      "(joint covariance + jitter*I is not positive definite at leading minor " + std::to_string(info) + ")", tb::ERR_NUMERIC);
  // samples = mean + L z
  const double* zd = z;
  if (!is_device_ptr(z)) {
    TB_TRY(bz.reserve(sizeof(double) * (size_t)S * M));
    TB_CUDA(cudaMemcpyAsync(bz.p, z, sizeof(double) * (size_t)S * M, cudaMemcpyHostToDevice, st));
    zd = bz.as<double>();
  }
  const bool out_dev = is_device_ptr(out);
  double* od = out;
  if (!out_dev) {
    TB_TRY(bout.reserve(sizeof(double) * (size_t)S * M));
    od = bout.as<double>();
  }
  trmv_lower_cols_kernel<<<dim3((unsigned)((M + 127) / 128), (unsigned)S), 128, 0, st>>>(bcov.as<double>(), M, M, zd, M, od, M);
  TB_LAUNCHED();
  add_mean_rows_kernel<<<(unsigned)(((int64_t)S * M + 255) / 256), 256, 0, st>>>(od, gp->sMean.as<double>(), M, (int64_t)S * M);
  TB_LAUNCHED();
  if (!out_dev) TB_CUDA(cudaMemcpyAsync(out, od, sizeof(double) * (size_t)S * M, cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}

// ---- value AND gradient of the batch Monte-Carlo EI (reverse pass of function.py:1181-1186) ----
// chunk pipeline: K* digits -> A = Linv K* (stored) -> per-batch mean / cov (joint_kernel) -> qei_backward_kernel
// (value, G_mu, Sigma_bar) -> V = K^-1 K* (dense digit GEMM over the same K* digits) -> per-batch mix V~ = Sigma_bar V
// -> grad_kernel (the training-point sums) -> qei_cross_kernel (the K(x_b, x_b) term).
template <int KIND>
static void launch_qei_cross(tb_gp* gp, const double* xc, int64_t npts, int q, const double* sbar, double* grad) {
  qei_cross_kernel<KIND><<<(unsigned)((npts + 127) / 128), 128, 0, gp->stream>>>(xc, gp->dInvLs.as<double>(), gp->D, npts, q, sbar,
                                                                                 gp->variance, grad);
}

static int run_qei_grad(tb_gp* gp, const double* Xc, int64_t B, int q, const double* eps, int S, double eta, double jitter,
                        double* out_val, double* out_grad) {
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  TB_CHECK(q >= 1 && q <= 32, "batch size q must be in [1, 32]");
  TB_CHECK(B >= 0, "negative batch count");
  TB_CHECK(S >= 1 && eps, "need S >= 1 base samples");
  TB_CHECK(jitter >= 0.0, "jitter must be non-negative");
  if (B == 0) return 0;
  const bool oz_path = gp->engine == 1 && gp->N <= 16384;  // else: native fp64 DMMA kernels
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D;
  const int QT = (q + 7) / 8, QP = QT * 8;
  const int64_t lda = (int64_t)gp->NB * BM;
  bool fast = false;  // single-pass engine for both store GEMMs (A = Linv K*, V = K^-1 K*)
  if (oz_path) TB_TRY(oz5_store_ready(gp, true, &fast));
  const int nt = fast ? oz5_tile_width(gp) : BT;
  if (fast) {
  } else if (oz_path) {
    TB_TRY(ensure_ozaki(gp));
    TB_TRY(ensure_kinv_digits(gp));
  } else {
    TB_TRY(ensure_upper_panels(gp));
  }
  const int64_t max_tiles = chunk_tiles(gp);
  int64_t nbc_cap = std::min<int64_t>(std::max<int64_t>(1, (max_tiles * BT) / q), B);
  const int64_t cand_cap = nbc_cap * q;
  const int64_t tiles_cap = (cand_cap + nt - 1) / nt;
  tb::DevBuf baplain;  // fp64 engine: plain copy of A for the Gram kernel (sA holds the packed panels the upper GEMM reads)
  struct ReleaseA {
    tb::DevBuf* b;
    ~ReleaseA() { b->release(); }
  } rel_a{&baplain};
  if (oz_path) {
    TB_TRY(gp->sKs.reserve(fast ? (size_t)tiles_cap * oz5_tile_bytes(gp) : (size_t)tiles_cap * gp->nst * oz::S * oz::TILE));
    TB_TRY(gp->sA.reserve((size_t)tiles_cap * nt * lda * sizeof(double)));  // A plain
  } else {
    TB_TRY(gp->sKs.reserve((size_t)tiles_cap * gp->nkc * PANEL * sizeof(double)));
    TB_TRY(gp->sA.reserve((size_t)tiles_cap * gp->NB * (BM / BK) * PANEL * sizeof(double)));  // A packed
    TB_TRY(baplain.reserve((size_t)tiles_cap * BT * lda * sizeof(double)));
    TB_TRY(gp->sPartial.reserve(sizeof(double) * (size_t)gp->NB * tiles_cap * BT));
  }
  TB_TRY(gp->sV.reserve((size_t)tiles_cap * nt * lda * sizeof(double)));  // V plain
  TB_TRY(gp->sMean.reserve(sizeof(double) * tiles_cap * nt));
  TB_TRY(gp->sMisc.reserve(sizeof(double) * 2 * cand_cap));  // c_mu, c_var
  const bool xc_dev = is_device_ptr(Xc), val_dev = is_device_ptr(out_val), grad_dev = is_device_ptr(out_grad);
  if (!xc_dev) TB_TRY(gp->sXc.reserve(sizeof(double) * cand_cap * D));
  if (!grad_dev) TB_TRY(gp->sGrad.reserve(sizeof(double) * cand_cap * D));
  tb::DevBuf beps, bcov, bmu, bsbar, bval;
  struct Release {
    std::vector<tb::DevBuf*> v;
    ~Release() { for (auto* b : v) b->release(); }
  } rel{{&beps, &bcov, &bmu, &bsbar, &bval}};
  const double* eps_dev = eps;
  if (!is_device_ptr(eps)) {
    TB_TRY(beps.reserve(sizeof(double) * (size_t)q * S));
    TB_CUDA(cudaMemcpyAsync(beps.p, eps, sizeof(double) * (size_t)q * S, cudaMemcpyHostToDevice, st));
    eps_dev = beps.as<double>();
  }
  TB_TRY(bcov.reserve(sizeof(double) * (size_t)nbc_cap * q * q));
  TB_TRY(bsbar.reserve(sizeof(double) * (size_t)nbc_cap * q * q));
  TB_TRY(bmu.reserve(sizeof(double) * (size_t)cand_cap));
  TB_TRY(bval.reserve(sizeof(double) * (size_t)nbc_cap));
  TB_TRY(gp->sRun.reserve(16));
  int* err = reinterpret_cast<int*>(gp->sRun.p);
  TB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  const size_t smem_joint = (size_t)JOINT_WARPS * (QP * QP + QP * D + QP) * sizeof(double);
  const size_t smem_back = (size_t)QEIG_WARPS * (3 * q * q + 2 * q) * sizeof(double);
  TB_CUDA(cudaFuncSetAttribute(qei_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_back));
  JointRequest jr;
  jr.mode = JOINT_PREDICT;
  jr.q = q;

  for (int64_t b0 = 0; b0 < B; b0 += nbc_cap) {
    const int64_t nbc = std::min<int64_t>(nbc_cap, B - b0);
    const int64_t mc = nbc * q;
    const int tiles = (int)((mc + nt - 1) / nt);
    const int64_t McPad = (int64_t)tiles * nt;
    const double* xc_chunk;
    if (xc_dev) {
      xc_chunk = Xc + b0 * q * D;
    } else {
      TB_CUDA(cudaMemcpyAsync(gp->sXc.p, Xc + b0 * q * D, sizeof(double) * mc * D, cudaMemcpyHostToDevice, st));
      xc_chunk = gp->sXc.as<double>();
    }
    const double* a_plain;
    if (fast) {
      TB_TRY(oz5_launch_kstar(gp, st, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
      const int Gs = oz5_groups(gp, tiles);
      TB_TRY(oz5_launch_gemm_store(gp, st, 0, gp->sKs.as<int8_t>(), tiles, Gs, gp->sA.as<double>(), lda));
      TB_TRY(oz5_launch_gemm_store(gp, st, 1, gp->sKs.as<int8_t>(), tiles, Gs, gp->sV.as<double>(), lda));
      a_plain = gp->sA.as<double>();
    } else if (oz_path) {
      TB_TRY(launch_kstar_digits(gp, xc_chunk, mc, tiles, gp->sKs.as<int8_t>(), gp->sMean.as<double>()));
      const int Goz = std::max(std::max(1, (gp->NB + 3) / 4), std::min(gp->NB, (2 * 148 + tiles - 1) / tiles));
      oz::trigemm_i8_kernel<oz::OZ_STORE, 8><<<dim3(Goz, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
          gp->dAS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dRowScale.as<double>(), gp->NB, gp->nst, Goz, McPad, gp->oz_out_scale,
          oz_npass(gp), 0, nullptr, gp->sA.as<double>(), lda);
      TB_LAUNCHED();
      const int Gv = std::max(1, std::min(gp->NB, std::max((gp->NB + 7) / 8, (2 * 148 + tiles - 1) / tiles)));
      oz::trigemm_i8_kernel<oz::OZ_STORE, 8><<<dim3(Gv, tiles), 10 * 32, oz::SMEM_BYTES, st>>>(
          gp->dKinvS.as<int8_t>(), gp->sKs.as<int8_t>(), gp->dKinvScale.as<double>(), gp->NB, gp->nst, Gv, McPad, gp->oz_out_scale,
          oz_npass(gp), 1, nullptr, gp->sV.as<double>(), lda);
      TB_LAUNCHED();
      a_plain = gp->sA.as<double>();
    } else {
      // native fp64 engine: A twice (plain for the Gram kernel, packed panels for the upper GEMM), then V = Linv^T A
      const int G = pick_groups(gp, tiles);
      TB_TRY(launch_kstar(gp, xc_chunk, mc, tiles, gp->sKs.as<double>(), gp->sMean.as<double>()));
      trigemm_kernel<false, EPI_PLAIN><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, nullptr, nullptr, baplain.as<double>(), lda);
      TB_LAUNCHED();
      trigemm_kernel<false, EPI_SUMSQ_PACKED><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvP.as<double>(), gp->sKs.as<double>(), gp->NB, gp->nkc, G, McPad, gp->sPartial.as<double>(), gp->sA.as<double>(),
          nullptr, 0);
      TB_LAUNCHED();
      const int nkB = gp->NB * (BM / BK);
      trigemm_kernel<true, EPI_PLAIN><<<dim3(tiles, G), TG_THREADS, TG_SMEM, st>>>(
          gp->dLinvTP.as<double>(), gp->sA.as<double>(), gp->NB, nkB, G, McPad, nullptr, nullptr, gp->sV.as<double>(), lda);
      TB_LAUNCHED();
      a_plain = baplain.as<double>();
    }
    TB_CUDA(cudaGetLastError());
    const int jblocks = (int)((nbc + JOINT_WARPS - 1) / JOINT_WARPS);
    const int Nrows = (int)lda;
    double* dmu = bmu.as<double>();
    double* dcov = bcov.as<double>();
    switch (gp->kernel) {
      case TB_RBF: TB_TRY(launch_joint<TB_RBF>(gp, QT, jblocks, smem_joint, a_plain, lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, jr, nullptr, dmu, dcov, nullptr, nullptr, err)); break;
      case TB_MATERN12: TB_TRY(launch_joint<TB_MATERN12>(gp, QT, jblocks, smem_joint, a_plain, lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, jr, nullptr, dmu, dcov, nullptr, nullptr, err)); break;
      case TB_MATERN32: TB_TRY(launch_joint<TB_MATERN32>(gp, QT, jblocks, smem_joint, a_plain, lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, jr, nullptr, dmu, dcov, nullptr, nullptr, err)); break;
      default: TB_TRY(launch_joint<TB_MATERN52>(gp, QT, jblocks, smem_joint, a_plain, lda, Nrows, gp->sMean.as<double>(), xc_chunk, nbc, jr, nullptr, dmu, dcov, nullptr, nullptr, err)); break;
    }
    double* cmu = gp->sMisc.as<double>();
    double* dval = val_dev ? out_val + b0 : bval.as<double>();
    qei_backward_kernel<<<(unsigned)((nbc + QEIG_WARPS - 1) / QEIG_WARPS), QEIG_WARPS * 32, smem_back, st>>>(
        dmu, dcov, nbc, q, eps_dev, S, eta, jitter, dval, cmu, cmu + mc, bsbar.as<double>(), err);
    TB_LAUNCHED();
    qei_mix_kernel<<<dim3((unsigned)nbc, (unsigned)((gp->N + 255) / 256)), 256, 0, st>>>(gp->sV.as<double>(), lda, (int)gp->N, q,
                                                                                        bsbar.as<double>());
    TB_LAUNCHED();
    double* gd = grad_dev ? out_grad + b0 * q * D : gp->sGrad.as<double>();
    switch (gp->kernel) {
      case TB_RBF: launch_grad_dp<TB_RBF>(gp, xc_chunk, mc, gd); break;
      case TB_MATERN12: launch_grad_dp<TB_MATERN12>(gp, xc_chunk, mc, gd); break;
      case TB_MATERN32: launch_grad_dp<TB_MATERN32>(gp, xc_chunk, mc, gd); break;
      default: launch_grad_dp<TB_MATERN52>(gp, xc_chunk, mc, gd); break;
    }
    TB_LAUNCHED();
    switch (gp->kernel) {
      case TB_RBF: launch_qei_cross<TB_RBF>(gp, xc_chunk, mc, q, bsbar.as<double>(), gd); break;
      case TB_MATERN12: launch_qei_cross<TB_MATERN12>(gp, xc_chunk, mc, q, bsbar.as<double>(), gd); break;
      case TB_MATERN32: launch_qei_cross<TB_MATERN32>(gp, xc_chunk, mc, q, bsbar.as<double>(), gd); break;
      default: launch_qei_cross<TB_MATERN52>(gp, xc_chunk, mc, q, bsbar.as<double>(), gd); break;
    }
    TB_LAUNCHED();
    TB_CUDA(cudaGetLastError());
    if (!val_dev) TB_CUDA(cudaMemcpyAsync(out_val + b0, bval.p, sizeof(double) * nbc, cudaMemcpyDeviceToHost, st));
    if (!grad_dev) TB_CUDA(cudaMemcpyAsync(out_grad + b0 * q * D, gd, sizeof(double) * mc * D, cudaMemcpyDeviceToHost, st));
    if (!xc_dev || !val_dev || !grad_dev) TB_CUDA(cudaStreamSynchronize(st));
  }
  int herr = 0;
  TB_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  TB_CHECK_CODE(herr == 0, "Cholesky decomposition was not successful. The input might not be valid "
                      "(covariance + jitter*I of a query batch is not positive definite)", tb::ERR_NUMERIC);
  return 0;
}

}  // namespace tb

extern "C" {

static int tb_gp_predict_joint_f64(tb_gp* gp, const void* Xc, int64_t B, int q, void* mean, void* cov) {
  TB_CHECK(gp && (B == 0 || (Xc && mean && cov)), "tb_gp_predict_joint: null argument");
  tb::JointRequest rq;
  rq.mode = JOINT_PREDICT;
  rq.Xc = (const double*)Xc;
  rq.B = B;
  rq.q = q;
  rq.out_mean = (double*)mean;
  rq.out_cov = (double*)cov;
  return tb::run_joint(gp, rq);
}

static int tb_acq_batch_mc_ei_f64(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double eta,
                       double jitter, void* out) {
  TB_CHECK(gp && (B == 0 || (Xc && eps && out)), "tb_acq_batch_mc_ei: null argument");
  tb::JointRequest rq;
  rq.mode = JOINT_QEI;
  rq.Xc = (const double*)Xc;
  rq.B = B;
  rq.q = q;
  rq.eps = (const double*)eps;
  rq.S = S;
  rq.eta = eta;
  rq.jitter = jitter;
  rq.out_qei = (double*)out;
  return tb::run_joint(gp, rq);
}

static int tb_gp_reparam_sample_f64(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double jitter,
                         void* samples) {
  TB_CHECK(gp && (B == 0 || (Xc && eps && samples)), "tb_gp_reparam_sample: null argument");
  tb::JointRequest rq;
  rq.mode = JOINT_SAMPLE;
  rq.Xc = (const double*)Xc;
  rq.B = B;
  rq.q = q;
  rq.eps = (const double*)eps;
  rq.S = S;
  rq.jitter = jitter;
  rq.out_samples = (double*)samples;
  return tb::run_joint(gp, rq);
}

// ---- top-k ---------------------------------------------------------------------------------------
int tb_topk(int device, int dtype, const void* values, int64_t M, int k, void* top_values, int64_t* top_indices) {
  TB_CHECK(dtype == TB_F64, "tb_topk: only TB_F64 is implemented in this build");
  TB_CHECK(values && top_values && top_indices, "tb_topk: null argument");
  TB_CHECK(M >= 1 && k >= 1 && k <= M, "tb_topk: need 1 <= k <= M");
  TB_CUDA(cudaSetDevice(device));
  int64_t P = BIT_TILE;
  while (P < M) P <<= 1;
  const bool vdev = is_device_ptr(values), tvdev = is_device_ptr(top_values), tidev = is_device_ptr(top_indices);
  tb::DevBuf a, vin, tv, ti;
  int rc = 0;
  auto body = [&]() -> int {
    TB_TRY(a.reserve(sizeof(VI) * P));
    const double* vd = (const double*)values;
    if (!vdev) {
      TB_TRY(vin.reserve(sizeof(double) * M));
      TB_CUDA(cudaMemcpy(vin.p, values, sizeof(double) * M, cudaMemcpyHostToDevice));
      vd = vin.as<double>();
    }
    topk_init_kernel<<<(unsigned)((P + 255) / 256), 256>>>(vd, M, P, a.as<VI>());
    TB_LAUNCHED();
    const unsigned nblk = (unsigned)(P / BIT_TILE);
    bitonic_local_kernel<<<nblk, 1024>>>(a.as<VI>(), 2, BIT_TILE);
    TB_LAUNCHED();
    for (int64_t kk = (int64_t)BIT_TILE * 2; kk <= P; kk <<= 1) {
      for (int64_t j = kk >> 1; j >= BIT_TILE; j >>= 1) {
        bitonic_global_kernel<<<(unsigned)((P / 2 + 255) / 256), 256>>>(a.as<VI>(), P, kk, j);
        TB_LAUNCHED();
      }
      bitonic_local_kernel<<<nblk, 1024>>>(a.as<VI>(), kk, kk);
      TB_LAUNCHED();
    }
    double* tvd = (double*)top_values;
    int64_t* tid = top_indices;
    if (!tvdev) { TB_TRY(tv.reserve(sizeof(double) * k)); tvd = tv.as<double>(); }
    if (!tidev) { TB_TRY(ti.reserve(sizeof(int64_t) * k)); tid = ti.as<int64_t>(); }
    topk_emit_kernel<<<(k + 255) / 256, 256>>>(a.as<VI>(), k, tvd, tid);
    TB_LAUNCHED();
    TB_CUDA(cudaGetLastError());
    if (!tvdev) TB_CUDA(cudaMemcpy(top_values, tvd, sizeof(double) * k, cudaMemcpyDeviceToHost));
    if (!tidev) TB_CUDA(cudaMemcpy(top_indices, tid, sizeof(int64_t) * k, cudaMemcpyDeviceToHost));
    TB_CUDA(cudaDeviceSynchronize());
    return 0;
  };
  rc = body();
  a.release(); vin.release(); tv.release(); ti.release();
  return rc;
}

}  // extern "C"

// =================================================================================================
// random-Fourier-feature trajectories
// =================================================================================================
struct tb_rff {
  int device = 0;
  cudaStream_t stream = nullptr;
  int F = 0, D = 0, DP = 0, nb = 0;
  double variance = 1.0, mean_const = 0.0;
  tb::DevBuf dW, dB, dTheta, dInvLs, sXc, sOut, sBlkBest, sBlkIdx, sRunV, sRunI;
  // canonical features of a decoupled trajectory: scaled training inputs + per-trajectory weights v [nbc, N]
  tb::DevBuf dXs, dV, sCanon;
  int kernel = TB_MATERN52, nbc = 0;
  int64_t N = 0;
};

extern "C" {

int tb_rff_create(tb_rff** out, int device) {
  TB_CHECK(out, "tb_rff_create: null output");
  int n = 0;
  TB_TRY(tb_device_count(&n));
  TB_CHECK_CODE(n > 0, "tb_rff_create: no CUDA device visible (this library has no CPU fallback)", tb::ERR_RUNTIME);
  TB_CHECK(device >= 0 && device < n, "tb_rff_create: device index out of range");
  TB_CUDA(cudaSetDevice(device));
  tb_rff* r = new tb_rff();
  r->device = device;
  TB_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  *out = r;
  return 0;
}

int tb_rff_destroy(tb_rff* r) {
  if (!r) return 0;
  cudaSetDevice(r->device);
  cudaStreamSynchronize(r->stream);
  for (tb::DevBuf* b : {&r->dW, &r->dB, &r->dTheta, &r->dInvLs, &r->sXc, &r->sOut, &r->sBlkBest, &r->sBlkIdx,
                        &r->sRunV, &r->sRunI, &r->dXs, &r->dV, &r->sCanon})
    b->release();
  cudaStreamDestroy(r->stream);
  delete r;
  return 0;
}

int tb_rff_set(tb_rff* r, const double* W, const double* b, int F, int D, const double* lengthscales,
               double variance, double mean_const) {
  TB_CHECK(r && W && b && lengthscales, "tb_rff_set: null argument");
  TB_CHECK(F >= 1, "tb_rff_set: need at least one feature");
  TB_CHECK(D >= 1 && tb::pick_dp(D) > 0, "tb_rff_set: input dimension must be in [1, 32]");
  TB_CHECK(variance > 0.0, "tb_rff_set: kernel variance must be positive");
  TB_CUDA(cudaSetDevice(r->device));
  const int DP = tb::pick_dp(D);
  std::vector<double> Wp((size_t)F * DP, 0.0), il(DP, 0.0);
  for (int f = 0; f < F; ++f)
    for (int d = 0; d < D; ++d) Wp[(size_t)f * DP + d] = W[(size_t)f * D + d];
  for (int d = 0; d < D; ++d) {
    TB_CHECK(lengthscales[d] > 0.0, "tb_rff_set: lengthscales must be positive");
    il[d] = 1.0 / lengthscales[d];
  }
  TB_TRY(r->dW.reserve(sizeof(double) * Wp.size()));
  TB_TRY(r->dB.reserve(sizeof(double) * F));
  TB_TRY(r->dInvLs.reserve(sizeof(double) * DP));
  TB_CUDA(cudaMemcpy(r->dW.p, Wp.data(), sizeof(double) * Wp.size(), cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemcpy(r->dB.p, b, sizeof(double) * F, cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemcpy(r->dInvLs.p, il.data(), sizeof(double) * DP, cudaMemcpyHostToDevice));
  r->F = F;
  r->D = D;
  r->DP = DP;
  r->variance = variance;
  r->mean_const = mean_const;
  r->nb = 0;
  r->nbc = 0;
  r->N = 0;
  return 0;
}

int tb_rff_set_theta(tb_rff* r, const double* theta, int nb) {
  TB_CHECK(r && theta, "tb_rff_set_theta: null argument");
  TB_CHECK(r->F > 0, "tb_rff_set_theta: call tb_rff_set first");
  TB_CHECK(nb >= 1, "tb_rff_set_theta: need at least one trajectory");
  TB_CUDA(cudaSetDevice(r->device));
  TB_TRY(r->dTheta.reserve(sizeof(double) * (size_t)nb * r->F));
  TB_CUDA(cudaMemcpy(r->dTheta.p, theta, sizeof(double) * (size_t)nb * r->F, cudaMemcpyDefault));
  r->nb = nb;
  return 0;
}

}  // extern "C"

namespace tb {
template <int DP>
static int launch_rff(tb_rff* r, int nbt, int blocks, const double* xc, int b0, int64_t mc, int64_t idx0, double scale,
                      const double* addend, double* out, double* bb, int64_t* bi) {
  const size_t smem = sizeof(double) * ((size_t)RFF_FCHUNK * DP + RFF_FCHUNK + (size_t)nbt * RFF_FCHUNK);
#define TB_RFF(NBT)                                                                                              \
  {                                                                                                              \
    TB_CUDA(cudaFuncSetAttribute(rff_eval_kernel<DP, NBT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    rff_eval_kernel<DP, NBT><<<blocks, RFF_THREADS, smem, r->stream>>>(r->dW.as<double>(), r->dB.as<double>(),      \
        r->dTheta.as<double>(), xc, r->dInvLs.as<double>(), r->D, r->F, r->nb, b0, mc, idx0, scale, r->mean_const,  \
        addend, fm::TrigConsts(), out, bb, bi);                                                                          \
  }
  switch (nbt) {
    case 1: TB_RFF(1); break;
    case 2: TB_RFF(2); break;
    case 4: TB_RFF(4); break;
    default: TB_RFF(8); break;
  }
#undef TB_RFF
  TB_LAUNCHED();
  TB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace tb

namespace tb {
template <int KIND, int DP>
static void launch_kdot_nbt(tb_rff* r, const double* xc, int64_t mc, double* out) {
  const int blocks = (int)((mc + 255) / 256);
  for (int b0 = 0; b0 < r->nbc; b0 += 4) {
    const int rem = r->nbc - b0;
    if (rem >= 3)
      kdot_kernel<KIND, DP, 4><<<blocks, 256, 0, r->stream>>>(r->dXs.as<double>(), r->dV.as<double>(), r->N, xc, r->dInvLs.as<double>(),
                                                            (int)r->N, r->D, r->nbc, b0, mc, r->variance, fm::Consts(), out);
    else if (rem == 2)
      kdot_kernel<KIND, DP, 2><<<blocks, 256, 0, r->stream>>>(r->dXs.as<double>(), r->dV.as<double>(), r->N, xc, r->dInvLs.as<double>(),
                                                            (int)r->N, r->D, r->nbc, b0, mc, r->variance, fm::Consts(), out);
    else
      kdot_kernel<KIND, DP, 1><<<blocks, 256, 0, r->stream>>>(r->dXs.as<double>(), r->dV.as<double>(), r->N, xc, r->dInvLs.as<double>(),
                                                            (int)r->N, r->D, r->nbc, b0, mc, r->variance, fm::Consts(), out);
    TB_LAUNCHED();
  }
}
template <int KIND>
static void launch_kdot_dp(tb_rff* r, const double* xc, int64_t mc, double* out) {
  switch (r->DP) {
    case 2: launch_kdot_nbt<KIND, 2>(r, xc, mc, out); break;
    case 4: launch_kdot_nbt<KIND, 4>(r, xc, mc, out); break;
    case 6: launch_kdot_nbt<KIND, 6>(r, xc, mc, out); break;
    case 8: launch_kdot_nbt<KIND, 8>(r, xc, mc, out); break;
    case 10: launch_kdot_nbt<KIND, 10>(r, xc, mc, out); break;
    case 12: launch_kdot_nbt<KIND, 12>(r, xc, mc, out); break;
    case 16: launch_kdot_nbt<KIND, 16>(r, xc, mc, out); break;
    case 20: launch_kdot_nbt<KIND, 20>(r, xc, mc, out); break;
    case 24: launch_kdot_nbt<KIND, 24>(r, xc, mc, out); break;
    default: launch_kdot_nbt<KIND, 32>(r, xc, mc, out); break;
  }
}
static int launch_kdot(tb_rff* r, const double* xc, int64_t mc, double* out) {
  switch (r->kernel) {
    case TB_RBF: launch_kdot_dp<TB_RBF>(r, xc, mc, out); break;
    case TB_MATERN12: launch_kdot_dp<TB_MATERN12>(r, xc, mc, out); break;
    case TB_MATERN32: launch_kdot_dp<TB_MATERN32>(r, xc, mc, out); break;
    default: launch_kdot_dp<TB_MATERN52>(r, xc, mc, out); break;
  }
  TB_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace tb

extern "C" {

int tb_rff_set_canonical(tb_rff* r, int kernel, const double* X, int64_t N, const double* v, int nb) {
  TB_CHECK(r, "tb_rff_set_canonical: null handle");
  TB_CHECK(r->F > 0, "tb_rff_set_canonical: call tb_rff_set first");
  if (N == 0) {  // switch the canonical part off
    r->N = 0;
    r->nbc = 0;
    return 0;
  }
  TB_CHECK(X && v, "tb_rff_set_canonical: null argument");
  TB_CHECK(kernel >= TB_RBF && kernel <= TB_MATERN52, "tb_rff_set_canonical: unknown kernel kind");
  TB_CHECK(N > 0 && nb >= 1, "tb_rff_set_canonical: need N >= 1 training points and nb >= 1 trajectories");
  TB_CUDA(cudaSetDevice(r->device));
  const int D = r->D, DP = r->DP;
  std::vector<double> il(DP, 0.0), Xs((size_t)N * DP, 0.0);
  TB_CUDA(cudaMemcpy(il.data(), r->dInvLs.p, sizeof(double) * DP, cudaMemcpyDeviceToHost));
  for (int64_t k = 0; k < N; ++k)
    for (int d = 0; d < D; ++d) Xs[(size_t)k * DP + d] = X[(size_t)k * D + d] * il[d];
  TB_TRY(r->dXs.reserve(sizeof(double) * Xs.size()));
  TB_TRY(r->dV.reserve(sizeof(double) * (size_t)nb * N));
  TB_CUDA(cudaMemcpy(r->dXs.p, Xs.data(), sizeof(double) * Xs.size(), cudaMemcpyHostToDevice));
  TB_CUDA(cudaMemcpy(r->dV.p, v, sizeof(double) * (size_t)nb * N, cudaMemcpyDefault));
  r->kernel = kernel;
  r->N = N;
  r->nbc = nb;
  return 0;
}

}  // extern "C"

extern "C" {

int tb_rff_eval(tb_rff* r, const void* Xc, int64_t M, void* out, double* min_value, int64_t* min_index) {
  TB_CHECK(r && (M == 0 || Xc), "tb_rff_eval: null argument");
  TB_CHECK(r->F > 0 && r->nb > 0, "tb_rff_eval: features and theta must be set first");
  TB_CHECK((min_value == nullptr) == (min_index == nullptr), "tb_rff_eval: min_value and min_index go together");
  TB_CHECK(M > 0 || !min_value, "tb_rff_eval: argmin over an empty candidate set");
  if (M == 0) return 0;
  TB_CUDA(cudaSetDevice(r->device));
  cudaStream_t st = r->stream;
  const int D = r->D, nb = r->nb;
  const double scale = std::sqrt(2.0 * r->variance / (double)r->F);
  const bool xdev = is_device_ptr(Xc), odev = is_device_ptr(out);
  const int64_t chunk = std::min<int64_t>(M, (int64_t)1 << 22);
  const int blocks_cap = (int)((chunk + RFF_THREADS - 1) / RFF_THREADS);
  if (!xdev) TB_TRY(r->sXc.reserve(sizeof(double) * chunk * D));
  if (out && !odev) TB_TRY(r->sOut.reserve(sizeof(double) * chunk * nb));
  const bool want_min = min_value != nullptr;
  if (want_min) {
    TB_TRY(r->sBlkBest.reserve(sizeof(double) * (size_t)nb * blocks_cap));
    TB_TRY(r->sBlkIdx.reserve(sizeof(int64_t) * (size_t)nb * blocks_cap));
    TB_TRY(r->sRunV.reserve(sizeof(double) * nb));
    TB_TRY(r->sRunI.reserve(sizeof(int64_t) * nb));
    std::vector<double> iv(nb, -DBL_MAX);
    std::vector<int64_t> ii(nb, INT64_MAX);
    TB_CUDA(cudaMemcpyAsync(r->sRunV.p, iv.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(r->sRunI.p, ii.data(), sizeof(int64_t) * nb, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaStreamSynchronize(st));
  }
  for (int64_t c0 = 0; c0 < M; c0 += chunk) {
    const int64_t mc = std::min<int64_t>(chunk, M - c0);
    const int blocks = (int)((mc + RFF_THREADS - 1) / RFF_THREADS);
    const double* xc = (const double*)Xc + c0 * D;
    if (!xdev) {
      TB_CUDA(cudaMemcpyAsync(r->sXc.p, xc, sizeof(double) * mc * D, cudaMemcpyHostToDevice, st));
      xc = r->sXc.as<double>();
    }
    double* od = out ? (odev ? (double*)out + c0 * nb : r->sOut.as<double>()) : nullptr;
    const double* addend = nullptr;
    if (r->N > 0) {
      TB_CHECK(r->nbc == nb, "tb_rff_eval: canonical weights and theta must have the same number of trajectories");
      TB_TRY(r->sCanon.reserve(sizeof(double) * (size_t)chunk * nb));
      TB_TRY(tb::launch_kdot(r, xc, mc, r->sCanon.as<double>()));
      addend = r->sCanon.as<double>();
    }
    for (int b0 = 0, nbt = 0; b0 < nb; b0 += nbt) {
      const int rem = nb - b0;
      nbt = rem >= 5 ? 8 : rem >= 3 ? 4 : rem;  // trajectories per pass: the cosine of a feature is shared by all of them
      double* bb = want_min ? r->sBlkBest.as<double>() : nullptr;
      int64_t* bi = want_min ? r->sBlkIdx.as<int64_t>() : nullptr;
#define TB_RFF_DP(DPV) TB_TRY((tb::launch_rff<DPV>(r, nbt, blocks, xc, b0, mc, c0, scale, addend, od, bb, bi)))
      switch (r->DP) {
        case 2: TB_RFF_DP(2); break;
        case 4: TB_RFF_DP(4); break;
        case 6: TB_RFF_DP(6); break;
        case 8: TB_RFF_DP(8); break;
        case 10: TB_RFF_DP(10); break;
        case 12: TB_RFF_DP(12); break;
        case 16: TB_RFF_DP(16); break;
        case 20: TB_RFF_DP(20); break;
        case 24: TB_RFF_DP(24); break;
        default: TB_RFF_DP(32); break;
      }
#undef TB_RFF_DP
    }
    if (want_min) {
      // blk arrays are laid out [nb][blocks of this launch]
      rff_fold_kernel<<<nb, 256, 0, st>>>(r->sBlkBest.as<double>(), r->sBlkIdx.as<int64_t>(), blocks,
                                          r->sRunV.as<double>(), r->sRunI.as<int64_t>());
      TB_LAUNCHED();
    }
    if (out && !odev)
      TB_CUDA(cudaMemcpyAsync((double*)out + c0 * nb, r->sOut.p, sizeof(double) * mc * nb, cudaMemcpyDeviceToHost, st));
    if (!xdev || (out && !odev)) TB_CUDA(cudaStreamSynchronize(st));
  }
  if (want_min) {
    std::vector<double> hv(nb);
    TB_CUDA(cudaMemcpyAsync(hv.data(), r->sRunV.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(min_index, r->sRunI.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < nb; ++b) min_value[b] = -hv[b];
  }
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"


// =================================================================================================
// dtype dispatch.  TB_F32 handles (fp32 models, e.g. BASELINE config 5) take and return float arrays;
// in this build the arithmetic still runs on the fp64 DMMA path (inputs widened on the device, outputs
// narrowed), which exceeds the fp32 tolerance; an fp32-native tensor path is listed as next in DESIGN.md.
// =================================================================================================
namespace tb {

__global__ void widen_kernel(const float* __restrict__ in, int64_t n, double* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (double)in[i];
}
__global__ void narrow_kernel(const double* __restrict__ in, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

// per-call staging of float arrays as device doubles
struct F32Bridge {
  tb_gp* gp;
  std::vector<void*> allocs;
  struct Pending { double* dev; void* user; int64_t n; };
  std::vector<Pending> outs;
  explicit F32Bridge(tb_gp* g) : gp(g) {}
  ~F32Bridge() { for (void* p : allocs) cudaFree(p); }
  int alloc(void** p, size_t bytes) {
    TB_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
    allocs.push_back(*p);
    return 0;
  }
  int in(const void* user, int64_t n, const double** out) {  // float (host or device) -> device double
    *out = nullptr;
    if (!user || n == 0) return 0;
    const float* src = (const float*)user;
    if (!is_device_ptr(user)) {
      void* tmp;
      TB_TRY(alloc(&tmp, sizeof(float) * n));
      TB_CUDA(cudaMemcpyAsync(tmp, user, sizeof(float) * n, cudaMemcpyHostToDevice, gp->stream));
      src = (const float*)tmp;
    }
    void* d;
    TB_TRY(alloc(&d, sizeof(double) * n));
    widen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, gp->stream>>>(src, n, (double*)d);
    TB_LAUNCHED();
    *out = (const double*)d;
    return 0;
  }
  int out(void* user, int64_t n, double** dev) {  // device double scratch, narrowed into `user` by finish()
    *dev = nullptr;
    if (!user || n == 0) return 0;
    void* d;
    TB_TRY(alloc(&d, sizeof(double) * n));
    *dev = (double*)d;
    outs.push_back({(double*)d, user, n});
    return 0;
  }
  int finish() {
    for (auto& o : outs) {
      float* dst = (float*)o.user;
      void* tmp = nullptr;
      const bool dev = is_device_ptr(o.user);
      if (!dev) {
        TB_TRY(alloc(&tmp, sizeof(float) * o.n));
        dst = (float*)tmp;
      }
      narrow_kernel<<<(unsigned)((o.n + 255) / 256), 256, 0, gp->stream>>>(o.dev, o.n, dst);
      TB_LAUNCHED();
      if (!dev) TB_CUDA(cudaMemcpyAsync(o.user, tmp, sizeof(float) * o.n, cudaMemcpyDeviceToHost, gp->stream));
    }
    TB_CUDA(cudaStreamSynchronize(gp->stream));
    TB_CUDA(cudaGetLastError());
    return 0;
  }
};

}  // namespace tb

extern "C" {

int tb_gp_set_data(tb_gp* gp, const void* X, const void* y, int64_t N, int D) {
  TB_CHECK(gp && X && y, "tb_gp_set_data: null argument");
  if (gp->dtype == TB_F64) return tb_gp_set_data_f64(gp, X, y, N, D);
  TB_CHECK(N > 0 && D > 0, "tb_gp_set_data: dataset must be populated (N > 0)");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *Xd, *yd;
  TB_TRY(br.in(X, N * D, &Xd));
  TB_TRY(br.in(y, N, &yd));
  return tb_gp_set_data_f64(gp, Xd, yd, N, D);
}

int tb_gp_append_data(tb_gp* gp, const void* Xnew, const void* ynew, int64_t m) {
  TB_CHECK(gp && Xnew && ynew, "tb_gp_append_data: null argument");
  if (gp->dtype == TB_F64) return tb_gp_append_data_f64(gp, (const double*)Xnew, (const double*)ynew, m);
  TB_CHECK(m > 0 && gp->have_data, "tb_gp_append_data: set the data first and append at least one point");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *Xd, *yd;
  TB_TRY(br.in(Xnew, m * gp->D, &Xd));
  TB_TRY(br.in(ynew, m, &yd));
  return tb_gp_append_data_f64(gp, Xd, yd, m);
}

int tb_gp_get_cholesky(tb_gp* gp, void* L_out) {
  TB_CHECK(gp && L_out, "tb_gp_get_cholesky: null argument");
  if (gp->dtype == TB_F64) return tb_gp_get_cholesky_f64(gp, L_out);
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  double* Ld;
  TB_TRY(br.out(L_out, gp->N * gp->N, &Ld));
  TB_TRY(tb_gp_get_cholesky_f64(gp, Ld));
  return br.finish();
}

int tb_gp_predict(tb_gp* gp, const void* Xc, int64_t M, void* mean, void* var) {
  TB_CHECK(gp && (M == 0 || (Xc && mean && var)), "tb_gp_predict: null argument");
  if (gp->dtype == TB_F64) return tb_gp_predict_f64(gp, Xc, M, mean, var);
  if (M == 0) return 0;
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double* xd;
  double *md, *vd;
  TB_TRY(br.in(Xc, M * gp->D, &xd));
  TB_TRY(br.out(mean, M, &md));
  TB_TRY(br.out(var, M, &vd));
  TB_TRY(tb_gp_predict_f64(gp, xd, M, md, vd));
  return br.finish();
}

int tb_acq_eval(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out, void* grad) {
  TB_CHECK(gp && (M == 0 || (Xc && out)), "tb_acq_eval: null argument");
  if (gp->dtype == TB_F64) return tb_acq_eval_f64(gp, acq, param, Xc, M, out, grad);
  if (M == 0) return 0;
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double* xd;
  double *od, *gd;
  TB_TRY(br.in(Xc, M * gp->D, &xd));
  TB_TRY(br.out(out, M, &od));
  TB_TRY(br.out(grad, M * gp->D, &gd));
  TB_TRY(tb_acq_eval_f64(gp, acq, param, xd, M, od, gd));
  return br.finish();
}

int tb_acq_argmax(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out, void* best_value,
                  int64_t* best_index) {
  TB_CHECK(gp && Xc && best_value && best_index, "tb_acq_argmax: null argument");
  if (gp->dtype == TB_F64) return tb_acq_argmax_f64(gp, acq, param, Xc, M, out, best_value, best_index);
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double* xd;
  double* od;
  double best = 0.0;
  TB_TRY(br.in(Xc, M * gp->D, &xd));
  TB_TRY(br.out(out, M, &od));
  TB_TRY(tb_acq_argmax_f64(gp, acq, param, xd, M, od, &best, best_index));
  *(float*)best_value = (float)best;
  return br.finish();
}

// ---- device-side multi-start L-BFGS (acquisition/optimizer.py:566-745) ----
__global__ void lbfgs_finish_kernel(tb::lb::State s, int64_t P, double* __restrict__ f_out, int32_t* __restrict__ success,
                                    int64_t* __restrict__ nfev) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  f_out[p] = -s.f[p];
  success[p] = s.status[p] == tb::lb::ST_SUCCESS ? 1 : 0;
  nfev[p] = (int64_t)s.nfev[p];
}

int tb_acq_maximize(tb_gp* gp, int acq, double param, const double* lower, const double* upper, const double* starts, int64_t P,
                    int maxcor, int maxiter, int maxls, double gtol, double ftol, double* x_out, double* f_out,
                    int32_t* success, int64_t* nfev) {
  TB_CHECK(gp && lower && upper, "tb_acq_maximize: null argument");
  TB_CHECK(P >= 0 && P < ((int64_t)1 << 31), "tb_acq_maximize: number of starts out of range");
  TB_CHECK(P == 0 || (starts && x_out && f_out && success && nfev), "tb_acq_maximize: null argument");
  TB_CHECK(acq >= TB_ACQ_EI && acq <= TB_ACQ_MES, "tb_acq_maximize: unknown acquisition kind");
  if (acq == TB_ACQ_LCB || acq == TB_ACQ_NEG_LCB)
    TB_CHECK(param >= 0.0, "Standard deviation scaling parameter beta must not be negative");
  if (acq == TB_ACQ_MES) TB_CHECK(gp->mesS > 0, "min-value entropy search: set the min-value samples first (tb_acq_set_min_value_samples)");
  TB_CHECK(maxcor >= 1 && maxcor <= tb::lb::MMAX, "tb_acq_maximize: maxcor must be in [1, " + std::to_string(tb::lb::MMAX) + "]");
  TB_CHECK(maxiter >= 1 && maxls >= 1, "tb_acq_maximize: maxiter and maxls must be positive");
  TB_CHECK(gtol >= 0.0 && ftol >= 0.0, "tb_acq_maximize: tolerances must be non-negative");
  TB_CHECK(gp->cache_valid, "posterior cache is not built: call tb_gp_update_posterior_cache first");
  if (P == 0) return 0;
  TB_CUDA(cudaSetDevice(gp->device));
  cudaStream_t st = gp->stream;
  const int D = gp->D, m = maxcor;
  const size_t PD = (size_t)P * D;
  // per-problem state + compact evaluation buffers (freed on return)
  tb::DevBuf bx, bf, bg, bd, bt, bS, bY, brho, bgam, bint, bnfev, btrial, bidx, bxt, bval, bgrad, bbox, bcount, bres;
  struct Release {
    std::vector<tb::DevBuf*> v;
    ~Release() { for (auto* b : v) b->release(); }
  } rel{{&bx, &bf, &bg, &bd, &bt, &bS, &bY, &brho, &bgam, &bint, &bnfev, &btrial, &bidx, &bxt, &bval, &bgrad, &bbox, &bcount, &bres}};
  TB_TRY(bx.reserve(8 * PD)); TB_TRY(bg.reserve(8 * PD)); TB_TRY(bd.reserve(8 * PD)); TB_TRY(btrial.reserve(8 * PD));
  TB_TRY(bf.reserve(8 * (size_t)P)); TB_TRY(bt.reserve(8 * (size_t)P)); TB_TRY(bgam.reserve(8 * (size_t)P));
  TB_TRY(bS.reserve(8 * PD * m)); TB_TRY(bY.reserve(8 * PD * m)); TB_TRY(brho.reserve(8 * (size_t)P * m));
  TB_TRY(bint.reserve(sizeof(int) * 6 * (size_t)P)); TB_TRY(bnfev.reserve(8 * (size_t)P));
  TB_TRY(bidx.reserve(sizeof(int) * (size_t)P)); TB_TRY(bxt.reserve(8 * PD)); TB_TRY(bval.reserve(8 * (size_t)P));
  TB_TRY(bgrad.reserve(8 * PD)); TB_TRY(bbox.reserve(8 * 2 * (size_t)D)); TB_TRY(bcount.reserve(sizeof(int)));
  tb::lb::State s;
  s.x = bx.as<double>(); s.f = bf.as<double>(); s.g = bg.as<double>(); s.d = bd.as<double>(); s.t = bt.as<double>();
  s.S = bS.as<double>(); s.Y = bY.as<double>(); s.rho = brho.as<double>(); s.gam = bgam.as<double>();
  int* ints = bint.as<int>();
  s.npairs = ints; s.head = ints + P; s.ls = ints + 2 * P; s.iters = ints + 3 * P; s.phase = ints + 4 * P; s.status = ints + 5 * P;
  s.nfev = bnfev.as<long long>();
  s.xtrial = btrial.as<double>();
  double* dlo = bbox.as<double>();
  double* dup = dlo + D;
  TB_CUDA(cudaMemcpyAsync(dlo, lower, 8 * (size_t)D, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(dup, upper, 8 * (size_t)D, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(bxt.p, starts, 8 * PD, cudaMemcpyDefault, st));  // staged through the trial buffer
  TB_CUDA(cudaMemsetAsync(bS.p, 0, 8 * PD * m, st));
  TB_CUDA(cudaMemsetAsync(bY.p, 0, 8 * PD * m, st));
  TB_CUDA(cudaMemsetAsync(brho.p, 0, 8 * (size_t)P * m, st));
  tb::lb::lbfgs_init_kernel<<<(unsigned)((PD + 255) / 256), 256, 0, st>>>(bxt.as<double>(), P, D, dlo, dup, s);
  TB_LAUNCHED();
  tb::lb::Options o{D, m, maxiter, maxls, gtol, ftol};
  int n_active = 0;
  auto compact = [&]() -> int {
    tb::lb::lbfgs_compact_kernel<<<1, 1024, 0, st>>>(s.status, P, bidx.as<int>(), bcount.as<int>());
    TB_LAUNCHED();
    TB_CUDA(cudaMemcpyAsync(&n_active, bcount.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (n_active > 0) {
      tb::lb::lbfgs_gather_kernel<<<(unsigned)(((size_t)n_active * D + 255) / 256), 256, 0, st>>>(s.xtrial, bidx.as<int>(), n_active, D,
                                                                                                 bxt.as<double>());
      TB_LAUNCHED();
    }
    return 0;
  };
  TB_TRY(compact());
  // every problem ends after at most maxiter accepted steps of at most maxls trials each
  const int64_t max_rounds = (int64_t)maxiter * (int64_t)maxls + 2;
  const bool trace = std::getenv("TB_LBFGS_TRACE") != nullptr;  // per-round (active starts, milliseconds) on stderr
  std::vector<std::pair<int, double>> trace_rows;
  for (int64_t round = 0; n_active > 0 && round < max_rounds; ++round) {
    const auto t_round = std::chrono::steady_clock::now();
    const int n_round = n_active;
    tb::EvalRequest rq;
    rq.acq = acq;
    rq.param = param;
    rq.Xc = bxt.as<double>();
    rq.M = n_active;
    rq.out_vals = bval.as<double>();
    rq.out_grad = bgrad.as<double>();
    TB_TRY(tb::run_eval(gp, rq));
    tb::lb::lbfgs_step_kernel<<<(unsigned)((n_active + 7) / 8), 256, 0, st>>>(s, o, n_active, bidx.as<int>(), bxt.as<double>(),
                                                                             bval.as<double>(), bgrad.as<double>(), dlo, dup);
    TB_LAUNCHED();
    TB_TRY(compact());
    if (trace) trace_rows.emplace_back(n_round, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_round).count());
  }
  if (trace) {
    double total = 0.0;
    for (auto& r : trace_rows) total += r.second;
    std::fprintf(stderr, "[tb_acq_maximize] P=%lld rounds=%zu total=%.2f ms:", (long long)P, trace_rows.size(), total);
    for (auto& r : trace_rows) std::fprintf(stderr, " %d:%.2f", r.first, r.second);
    std::fprintf(stderr, "\n");
  }
  TB_TRY(bres.reserve((8 + 4 + 8) * (size_t)P));
  double* rf = bres.as<double>();
  int64_t* rn = reinterpret_cast<int64_t*>(rf + P);
  int32_t* rs = reinterpret_cast<int32_t*>(rn + P);
  lbfgs_finish_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(s, P, rf, rs, rn);
  TB_LAUNCHED();
  TB_CUDA(cudaMemcpyAsync(x_out, s.x, 8 * PD, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(f_out, rf, 8 * (size_t)P, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(nfev, rn, 8 * (size_t)P, cudaMemcpyDefault, st));
  TB_CUDA(cudaMemcpyAsync(success, rs, 4 * (size_t)P, cudaMemcpyDefault, st));
  TB_CUDA(cudaStreamSynchronize(st));
  TB_CUDA(cudaGetLastError());
  return 0;
}

int tb_gp_predict_joint(tb_gp* gp, const void* Xc, int64_t B, int q, void* mean, void* cov) {
  TB_CHECK(gp && (B == 0 || (Xc && mean && cov)), "tb_gp_predict_joint: null argument");
  if (gp->dtype == TB_F64) return tb_gp_predict_joint_f64(gp, Xc, B, q, mean, cov);
  if (B == 0) return 0;
  TB_CHECK(q >= 1 && q <= 32, "batch size q must be in [1, 32]");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double* xd;
  double *md, *cd;
  TB_TRY(br.in(Xc, B * q * gp->D, &xd));
  TB_TRY(br.out(mean, B * q, &md));
  TB_TRY(br.out(cov, B * q * q, &cd));
  TB_TRY(tb_gp_predict_joint_f64(gp, xd, B, q, md, cd));
  return br.finish();
}

int tb_acq_batch_mc_ei(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double eta,
                       double jitter, void* out) {
  TB_CHECK(gp && (B == 0 || (Xc && eps && out)), "tb_acq_batch_mc_ei: null argument");
  if (gp->dtype == TB_F64) return tb_acq_batch_mc_ei_f64(gp, Xc, B, q, eps, S, eta, jitter, out);
  if (B == 0) return 0;
  TB_CHECK(q >= 1 && q <= 32 && S >= 1, "batch size q must be in [1, 32] and S >= 1");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *xd, *ed;
  double* od;
  TB_TRY(br.in(Xc, B * q * gp->D, &xd));
  TB_TRY(br.in(eps, (int64_t)q * S, &ed));
  TB_TRY(br.out(out, B, &od));
  TB_TRY(tb_acq_batch_mc_ei_f64(gp, xd, B, q, ed, S, eta, jitter, od));
  return br.finish();
}

int tb_gp_covariance_between_points(tb_gp* gp, const void* X1, int64_t M1, const void* X2, int64_t M2, void* out) {
  TB_CHECK(gp && X1 && X2 && out, "tb_gp_covariance_between_points: null argument");
  if (gp->dtype == TB_F64) return tb::run_cross_cov(gp, (const double*)X1, M1, (const double*)X2, M2, (double*)out);
  TB_CHECK(M1 >= 1 && M2 >= 1, "tb_gp_covariance_between_points: need at least one point in each set");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *x1, *x2;
  double* od;
  TB_TRY(br.in(X1, M1 * gp->D, &x1));
  TB_TRY(br.in(X2, M2 * gp->D, &x2));
  TB_TRY(br.out(out, M1 * M2, &od));
  TB_TRY(tb::run_cross_cov(gp, x1, M1, x2, M2, od));
  return br.finish();
}

int tb_gp_sample_joint(tb_gp* gp, const void* Xc, int64_t M, const double* z, int S, double jitter, void* out) {
  TB_CHECK(gp && Xc && z && out, "tb_gp_sample_joint: null argument");
  if (gp->dtype == TB_F64) return tb::run_sample_joint(gp, (const double*)Xc, M, z, S, jitter, (double*)out);
  TB_CHECK(M >= 1 && S >= 1, "tb_gp_sample_joint: need at least one point and one draw");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double* xd;
  double* od;
  TB_TRY(br.in(Xc, M * gp->D, &xd));
  TB_TRY(br.out(out, (int64_t)S * M, &od));
  TB_TRY(tb::run_sample_joint(gp, xd, M, z, S, jitter, od));
  return br.finish();
}

int tb_acq_batch_mc_ei_grad(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double eta, double jitter,
                            void* out, void* grad) {
  TB_CHECK(gp && (B == 0 || (Xc && eps && out && grad)), "tb_acq_batch_mc_ei_grad: null argument");
  if (gp->dtype == TB_F64)
    return tb::run_qei_grad(gp, (const double*)Xc, B, q, (const double*)eps, S, eta, jitter, (double*)out, (double*)grad);
  if (B == 0) return 0;
  TB_CHECK(q >= 1 && q <= 32 && S >= 1, "batch size q must be in [1, 32] and S >= 1");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *xd, *ed;
  double *od, *gd;
  TB_TRY(br.in(Xc, B * q * gp->D, &xd));
  TB_TRY(br.in(eps, (int64_t)q * S, &ed));
  TB_TRY(br.out(out, B, &od));
  TB_TRY(br.out(grad, B * q * gp->D, &gd));
  TB_TRY(tb::run_qei_grad(gp, xd, B, q, ed, S, eta, jitter, od, gd));
  return br.finish();
}

int tb_gp_reparam_sample(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double jitter,
                         void* samples) {
  TB_CHECK(gp && (B == 0 || (Xc && eps && samples)), "tb_gp_reparam_sample: null argument");
  if (gp->dtype == TB_F64) return tb_gp_reparam_sample_f64(gp, Xc, B, q, eps, S, jitter, samples);
  if (B == 0) return 0;
  TB_CHECK(q >= 1 && q <= 32 && S >= 1, "batch size q must be in [1, 32] and S >= 1");
  TB_CUDA(cudaSetDevice(gp->device));
  tb::F32Bridge br(gp);
  const double *xd, *ed;
  double* sd;
  TB_TRY(br.in(Xc, B * q * gp->D, &xd));
  TB_TRY(br.in(eps, (int64_t)q * S, &ed));
  TB_TRY(br.out(samples, B * S * q, &sd));
  TB_TRY(tb_gp_reparam_sample_f64(gp, xd, B, q, ed, S, jitter, sd));
  return br.finish();
}

}  // extern "C"
