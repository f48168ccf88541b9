// Hand-written posterior-cache precompute: blocked Cholesky of K + noise I, triangular
// inverse Linv = L^-1, alpha = K^-1 err and K^-1 = Linv^T Linv, all on the fp64 DMMA pipe — no cuSOLVER / cuBLAS.
// Column-major N x N matrices, block size 128.  Once per BO step (off the per-candidate path): written for clarity and
// determinism, not tuned (single-buffered shared-memory tiles).
#pragma once
#include "common.cuh"

namespace tb {
namespace fac {

constexpr int FB = 128;          // block size
constexpr int KS = 32;           // k depth of one shared-memory step
constexpr int LDS = KS + 4;      // padded row stride (doubles): conflict-free m8n8k4 fragment loads
constexpr int THREADS = 256;     // 8 warps: 2 (rows) x 4 (cols), warp tile 64 x 32
constexpr size_t GEMM_SMEM = 2 * FB * LDS * sizeof(double);  // 73,728 B

// acc[8][4][2] += A(128 x [k0,k1)) * B([k0,k1) x 128); a_at(m, k), b_at(k, n) return the operand elements (0 outside)
// A_KFAST / B_KFAST: the operand is contiguous in memory along k (else along its row / column index m, n): picks the
// thread -> element mapping of the staging loads so that global reads coalesce
template <bool A_KFAST, bool B_KFAST, class FA, class FB_>
__device__ __forceinline__ void dmma_tile(FA a_at, FB_ b_at, int k0, int k1, double (&acc)[8][4][2], double* sm) {
  double* As = sm;
  double* Bs = sm + FB * LDS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 2, wn = warp & 3;
  for (int kb = k0; kb < k1; kb += KS) {
    __syncthreads();
    for (int e = threadIdx.x; e < FB * KS; e += THREADS) {
      const int ma = A_KFAST ? e / KS : e % FB, ka = A_KFAST ? e % KS : e / FB;
      const int nb_ = B_KFAST ? e / KS : e % FB, kb_ = B_KFAST ? e % KS : e / FB;
      As[ma * LDS + ka] = (kb + ka < k1) ? a_at(ma, kb + ka) : 0.0;
      Bs[nb_ * LDS + kb_] = (kb + kb_ < k1) ? b_at(kb + kb_, nb_) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < KS / 4; ++k4) {
      double af[8], bf[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = As[(wm * 64 + i * 8 + (lane >> 2)) * LDS + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[(wn * 32 + j * 8 + (lane >> 2)) * LDS + k4 * 4 + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
}
// visit the accumulator elements of this thread: f(row m in [0,128), col n in [0,128), value&)
template <class F>
__device__ __forceinline__ void for_each_acc(double (&acc)[8][4][2], F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 2, wn = warp & 3;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) f(wm * 64 + i * 8 + (lane >> 2), wn * 32 + j * 8 + (lane & 3) * 2 + c, acc[i][j][c]);
}
__device__ __forceinline__ void zero_acc(double (&acc)[8][4][2]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
}

// ---- Cholesky ---------------------------------------------------------------------------------
// diagonal block j: left-looking column Cholesky in shared memory + its inverse; one CTA, thread r owns row r (threads
// >= 128 only take part in the barriers).  info = first failing pivot (1-based)
__global__ void __launch_bounds__(THREADS)
chol_diag_kernel(double* __restrict__ A, int64_t N, int j0, double* __restrict__ Dinv, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // [FB][FB + 1]
  __shared__ double xd[FB];
  const int nb = (int)min((int64_t)FB, N - j0);
  const int ld = FB + 1;
  for (int e = threadIdx.x; e < FB * FB; e += THREADS) {
    const int r = e % FB, c = e / FB;
    S[r * ld + c] = (r < nb && c < nb && r >= c) ? A[(j0 + r) + (int64_t)(j0 + c) * N] : 0.0;
  }
  __syncthreads();
  const int r = threadIdx.x;
  const double* Sr = S + r * ld;
  for (int k = 0; k < nb; ++k) {
    // column k: S[r][k] -= sum_{j<k} S[r][j] S[k][j]  (row r against the already-final row k)
    double s = 0.0;
    if (r >= k && r < nb) {
      const double* Sk = S + k * ld;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int j = 0;
      for (; j + 4 <= k; j += 4) {
        s0 = fma(Sr[j], Sk[j], s0);
        s1 = fma(Sr[j + 1], Sk[j + 1], s1);
        s2 = fma(Sr[j + 2], Sk[j + 2], s2);
        s3 = fma(Sr[j + 3], Sk[j + 3], s3);
      }
      for (; j < k; ++j) s0 = fma(Sr[j], Sk[j], s0);
      s = Sr[k] - ((s0 + s1) + (s2 + s3));
    }
    __syncthreads();  // every read of row k is done before its diagonal entry changes
    if (r == k) S[k * ld + k] = s;
    __syncthreads();
    const double piv = S[k * ld + k];
    if (!(piv > 0.0)) {  // uniform across the CTA
      if (threadIdx.x == 0) atomicCAS(info, 0, j0 + k + 1);
      return;
    }
    const double d = sqrt(piv);
    if (r > k && r < nb) S[r * ld + k] = s / d;
    __syncthreads();
    if (r == k) S[k * ld + k] = d;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * nb; e += THREADS) {
    const int rr = e % nb, c = e / nb;
    A[(j0 + rr) + (int64_t)(j0 + c) * N] = (rr >= c) ? S[rr * ld + c] : 0.0;
  }
  // inverse of the lower-triangular block by forward substitution, one column per thread.  X[r][c] (r > c) lives in the
  // unused strict upper triangle of S (at S[c][r]), the diagonal in xd: every operand of the FMA chains is in smem
  __syncthreads();
  for (int c = threadIdx.x; c < nb; c += THREADS) {
    xd[c] = 1.0 / S[c * ld + c];
    double* Xc = S + c * ld;  // Xc[r] = X[r][c] for r > c
    for (int rr = c + 1; rr < nb; ++rr) {
      const double* Lr = S + rr * ld;
      double s0 = -Lr[c] * xd[c], s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int k = c + 1;
      for (; k + 4 <= rr; k += 4) {
        s0 = fma(-Lr[k], Xc[k], s0);
        s1 = fma(-Lr[k + 1], Xc[k + 1], s1);
        s2 = fma(-Lr[k + 2], Xc[k + 2], s2);
        s3 = fma(-Lr[k + 3], Xc[k + 3], s3);
      }
      for (; k < rr; ++k) s0 = fma(-Lr[k], Xc[k], s0);
      Xc[rr] = ((s0 + s1) + (s2 + s3)) / Lr[rr];
    }
  }
  __syncthreads();
  double* X = Dinv + (int64_t)(j0 / FB) * FB * FB;  // column-major [FB][FB]
  for (int e = threadIdx.x; e < FB * FB; e += THREADS) {
    const int rr = e % FB, c = e / FB;
    double v = 0.0;
    if (rr < nb && c < nb) v = (rr == c) ? xd[c] : (rr > c ? S[c * ld + rr] : 0.0);
    X[rr + (int64_t)c * FB] = v;
  }
}

// panel below diagonal block j: L21 = A21 * invL11^T (in place); grid.x = 128-row tiles below the block
__global__ void __launch_bounds__(THREADS)
chol_panel_kernel(double* __restrict__ A, int64_t N, int j0, const double* __restrict__ Dinv) {
  extern __shared__ __align__(16) double sm[];
  const int64_t r0 = (int64_t)j0 + FB + (int64_t)blockIdx.x * FB;
  const double* X = Dinv + (int64_t)(j0 / FB) * FB * FB;
  double acc[8][4][2];
  zero_acc(acc);
  // out[m][n] = sum_k A21[m][k] * invL11[n][k]
  dmma_tile<false, false>([&](int m, int k) { return (r0 + m < N) ? A[(r0 + m) + (int64_t)(j0 + k) * N] : 0.0; },
                          [&](int k, int n) { return X[n + (int64_t)k * FB]; }, 0, FB, acc, sm);
  __syncthreads();
  for_each_acc(acc, [&](int m, int n, double& v) {
    if (r0 + m < N && j0 + n < N) A[(r0 + m) + (int64_t)(j0 + n) * N] = v;
  });
}

// trailing update after panel j: A22[I][J] -= L21[I] L21[J]^T for the lower tiles I >= J
__global__ void __launch_bounds__(THREADS)
chol_syrk_kernel(double* __restrict__ A, int64_t N, int j0) {
  extern __shared__ __align__(16) double sm[];
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (ti < tj) return;
  const int64_t r0 = (int64_t)j0 + FB + (int64_t)ti * FB, c0 = (int64_t)j0 + FB + (int64_t)tj * FB;
  if (r0 >= N || c0 >= N) return;
  double acc[8][4][2];
  zero_acc(acc);
  dmma_tile<false, false>([&](int m, int k) { return (r0 + m < N) ? A[(r0 + m) + (int64_t)(j0 + k) * N] : 0.0; },
                          [&](int k, int n) { return (c0 + n < N) ? A[(c0 + n) + (int64_t)(j0 + k) * N] : 0.0; }, 0, FB, acc, sm);
  for_each_acc(acc, [&](int m, int n, double& v) {
    if (r0 + m < N && c0 + n < N && r0 + m >= c0 + n) A[(r0 + m) + (int64_t)(c0 + n) * N] -= v;
  });
}

// ---- Linv by recursive doubling (fully parallel GEMMs): with L = [[A, 0], [B, C]] and A^-1, C^-1 known (size n),
//      Linv[B-block] = -C^-1 (B A^-1).  One level = two launches over all pairs; log2(N/128) levels.
//      STEP 1: T = B A^-1 into a scratch matrix (same coordinates);  STEP 2: Linv[B-block] = -C^-1 T.
template <int STEP>
__global__ void __launch_bounds__(THREADS)
trinv_level_kernel(const double* __restrict__ L, double* __restrict__ Linv, double* __restrict__ T, int64_t N, int n) {
  extern __shared__ __align__(16) double sm[];
  const int64_t c0 = (int64_t)blockIdx.z * 2 * n, r0 = c0 + n;  // A at (c0, c0), B at (r0, c0), C at (r0, r0)
  const int64_t m0 = r0 + (int64_t)blockIdx.y * FB, n0 = c0 + (int64_t)blockIdx.x * FB;
  if (m0 >= N) return;
  double acc[8][4][2];
  zero_acc(acc);
  if (STEP == 1) {
    // T[m][nn] = sum_{k >= nn} L[m][c0 + k] * Linv[c0 + k][nn]
    dmma_tile<false, true>([&](int m, int k) { return (m0 + m < N) ? L[(m0 + m) + (c0 + k) * N] : 0.0; },
                           [&](int k, int nn) { return (c0 + k >= n0 + nn) ? Linv[(c0 + k) + (n0 + nn) * N] : 0.0; },
                           (int)(n0 - c0), n, acc, sm);
    for_each_acc(acc, [&](int m, int nn, double& v) {
      if (m0 + m < N) T[(m0 + m) + (n0 + nn) * N] = v;
    });
  } else {
    // Linv[m][nn] = - sum_{k <= m} Linv[m][r0 + k] * T[r0 + k][nn]
    const int kend = (int)min((int64_t)n, m0 - r0 + FB);
    dmma_tile<false, true>([&](int m, int k) { return (m0 + m < N && r0 + k <= m0 + m) ? Linv[(m0 + m) + (r0 + k) * N] : 0.0; },
                           [&](int k, int nn) { return (r0 + k < N) ? T[(r0 + k) + (n0 + nn) * N] : 0.0; }, 0, kend, acc, sm);
    for_each_acc(acc, [&](int m, int nn, double& v) {
      if (m0 + m < N) Linv[(m0 + m) + (n0 + nn) * N] = -v;
    });
  }
}
// diagonal blocks of Linv = inverses of the diagonal blocks of L
__global__ void trinv_diag_kernel(double* __restrict__ Linv, int64_t N, const double* __restrict__ Dinv) {
  const int64_t b0 = (int64_t)blockIdx.x * FB;
  const double* X = Dinv + (int64_t)blockIdx.x * FB * FB;
  for (int e = threadIdx.x; e < FB * FB; e += blockDim.x) {
    const int m = e % FB, nn = e / FB;
    if (b0 + m < N && b0 + nn < N) Linv[(b0 + m) + (b0 + nn) * N] = X[m + (int64_t)nn * FB];
  }
}

// ---- K^-1 = Linv^T Linv, lower tiles (I >= J): sum over k >= I*128 of Linv[k][I-cols] * Linv[k][J-cols] ----
__global__ void __launch_bounds__(THREADS)
kinv_kernel(const double* __restrict__ Linv, double* __restrict__ Kinv, int64_t N) {
  extern __shared__ __align__(16) double sm[];
  const int tj = blockIdx.x, ti = blockIdx.y;
  if (ti < tj) return;
  const int64_t r0 = (int64_t)ti * FB, c0 = (int64_t)tj * FB;
  double acc[8][4][2];
  zero_acc(acc);
  dmma_tile<true, true>([&](int m, int k) { return (r0 + m < N && k >= r0 + m) ? Linv[k + (int64_t)(r0 + m) * N] : 0.0; },
                        [&](int k, int n) { return (c0 + n < N && k >= c0 + n) ? Linv[k + (int64_t)(c0 + n) * N] : 0.0; }, (int)r0, (int)N, acc, sm);
  for_each_acc(acc, [&](int m, int n, double& v) {
    if (r0 + m < N && c0 + n < N) Kinv[(r0 + m) + (int64_t)(c0 + n) * N] = v;
  });
}

// ---- rank-m growth of K^-1 = Linv^T Linv after m rows were appended to the factor (N0 -> N): K^-1[i,j] = Σ_{n >= max(i,j)}
// Linv[n,i] Linv[n,j], so the old block gains only the m new rows' contribution and the new rows are sums over <= m terms:
// O(m N^2) instead of the O(N^3) rebuild.  Lower triangle (i >= j), column-major, old ld = N0, new ld = N.
__global__ void kinv_grow_kernel(const double* __restrict__ Kold, int64_t N0, const double* __restrict__ Linv, int64_t N,
                                 double* __restrict__ Knew) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= N || j > i) return;
  double s = (i < N0) ? Kold[i + j * N0] : 0.0;
  for (int64_t n = (i > N0 ? i : N0); n < N; ++n) s = fma(Linv[n + i * N], Linv[n + j * N], s);
  Knew[i + j * N] = s;
}

// ---- alpha = Linv^T (Linv r): two triangular mat-vecs, one warp per output element ----
__global__ void trmv_lower_kernel(const double* __restrict__ Linv, int64_t N, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  double s = 0.0;
  for (int64_t k = lane; k <= n; k += 32) s = fma(Linv[n + k * N], x[k], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[n] = s;
}
__global__ void trmv_lower_t_kernel(const double* __restrict__ Linv, int64_t N, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= N) return;
  double s = 0.0;
  for (int64_t n = k + lane; n < N; n += 32) s = fma(Linv[n + k * N], x[n], s);  // column k: contiguous
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[k] = s;
}

}  // namespace fac
}  // namespace tb
