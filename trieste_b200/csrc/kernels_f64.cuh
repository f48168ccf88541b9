// fp64 hot-path kernels: K(X*,X) panel generation, triangular DMMA GEMM, acquisition tail.
//
// Data flow for one chunk of candidates (GPflow GPRPosterior.predict_f called at
// trieste/models/gpflow/interface.py:120):
//   kstar_panels_kernel : Ks = K(X, X*) written in DMMA-fragment-packed panels + mean = Ks^T alpha + m
//   trigemm_kernel      : A = Linv · Ks per (row-block, candidate-tile), epilogue sum_n A^2  (never stores A)
//   tail_kernel         : var = clip(k** - sum A^2), EI / log-EI / LCB, block argmax
#pragma once
#include "common.cuh"
#include "../../include/trieste_b200.h"
#include "kernel_fn.cuh"
#include <cfloat>

namespace tb {

// ------------------------------------------------------------------------------------------------
// K1a: cross-covariance panels.  grid.x = candidate tiles of the chunk; 512 threads = 16 warps,
// warp w owns candidates [8w, 8w+8) of the tile; lane l <-> (candidate l/4, k-within-k4 l%4).
// Every store is one 512-byte contiguous warp write into the packed panel.
// ------------------------------------------------------------------------------------------------
template <int KIND, int DP>
__global__ void __launch_bounds__(512)
kstar_panels_kernel(const double* __restrict__ Xs,      // [nkc*16][DP] training inputs / lengthscale
                    const double* __restrict__ alpha,   // [nkc*16]  K^-1 err (0 beyond N)
                    const double* __restrict__ Xc,      // [M][D] raw candidates
                    const double* __restrict__ inv_ls,  // [DP]
                    int N, int nkc, int D, int64_t M, int64_t cand0, double variance,
                    double mean_const, double* __restrict__ KsP, double* __restrict__ mean_out) {
  const int lane = threadIdx.x & 31, tj = threadIdx.x >> 5;
  const int tl = lane >> 2, kq = lane & 3;
  const int t_local = tj * 8 + tl;
  const int64_t t = cand0 + (int64_t)blockIdx.x * BT + t_local;
  const bool valid = t < M;

  double xc[DP];
#pragma unroll
  for (int d = 0; d < DP; ++d) xc[d] = (valid && d < D) ? Xc[t * D + d] * inv_ls[d] : 0.0;

  double* panel = KsP + (int64_t)blockIdx.x * nkc * PANEL;
  double macc = 0.0;
  for (int kc = 0; kc < nkc; ++kc) {
#pragma unroll
    for (int p = 0; p < 2; ++p) {
      double kv[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int k = kc * BK + (2 * p + s) * 4 + kq;
        const double* xr = Xs + (int64_t)k * DP;
        double r2 = 0.0;
#pragma unroll
        for (int d = 0; d < DP; d += 2) {
          double2 v = __ldg(reinterpret_cast<const double2*>(xr + d));
          double d0 = xc[d] - v.x, d1 = xc[d + 1] - v.y;
          r2 = fma(d0, d0, r2);
          r2 = fma(d1, d1, r2);
        }
        double kval = (valid && k < N) ? kernel_from_r2<KIND>(r2, variance) : 0.0;
        macc = fma(kval, __ldg(alpha + k), macc);
        kv[s] = kval;
      }
      reinterpret_cast<double2*>(panel + (int64_t)kc * PANEL)[(tj * 2 + p) * 32 + lane] =
          make_double2(kv[0], kv[1]);
    }
  }
  macc += __shfl_xor_sync(0xffffffffu, macc, 1);
  macc += __shfl_xor_sync(0xffffffffu, macc, 2);
  if (kq == 0) mean_out[(int64_t)blockIdx.x * BT + t_local] = macc + mean_const;
}

// ------------------------------------------------------------------------------------------------
// K1b: triangular DMMA GEMM  C = T · B  with T a packed triangular factor (Linv lower, or Linv^T upper)
//   grid = (candidate tiles, G row-block groups); 8 consumer warps (2 x 4, warp tile 64 rows x 32
//   candidates, 64 fp64 accumulators / thread) + 1 producer warp that streams packed 16 KB panels of
//   T and B with 1-D bulk TMA copies through a 4-stage mbarrier ring.
//   Epilogues (compile-time):
//     EPI_SUMSQ         partial[g][t] = sum over the group's rows n of C[n,t]^2            (variance)
//     EPI_SUMSQ_PACKED  + C stored as B-operand panels [tile][row/16][PANEL]   (feeds the Linv^T GEMM)
//     EPI_PLAIN         C stored plain, candidate-major: Cplain[t][ldc] (n contiguous)    (joint / V)
//     EPI_SUMSQ_PLAIN   both
// ------------------------------------------------------------------------------------------------
constexpr int TG_STAGES = 4;
constexpr int TG_CONSUMER_WARPS = 8;
constexpr int TG_THREADS = (TG_CONSUMER_WARPS + 1) * 32;
constexpr size_t TG_SMEM = (size_t)TG_STAGES * 2 * PANEL * sizeof(double) + 2 * TG_STAGES * 8 + 2 * BT * 8 + 64;

enum { EPI_SUMSQ = 0, EPI_SUMSQ_PACKED = 1, EPI_PLAIN = 2, EPI_SUMSQ_PLAIN = 3 };


// panel range [k0, k1) and storage offset of row-block I
template <bool UPPER>
__device__ __forceinline__ void rowblock_range(int I, int nkB, int& k0, int& k1, int64_t& off) {
  if (!UPPER) {
    k0 = 0;
    k1 = min((I + 1) * (BM / BK), nkB);
    off = rowblock_panel_offset(I);
  } else {
    k0 = I * (BM / BK);
    k1 = nkB;
    off = (int64_t)I * nkB - rowblock_panel_offset(I - 1) - (int64_t)k0;  // so that panel kc sits at off + kc
  }
}
__host__ __device__ inline int64_t upper_panel_count(int NB, int nkB) {
  return (int64_t)NB * nkB - rowblock_panel_offset(NB - 1);
}

template <bool UPPER, int EPI>
__global__ void __launch_bounds__(TG_THREADS, 1)
trigemm_kernel(const double* __restrict__ TP,   // packed triangular panels
               const double* __restrict__ BP,   // [tiles][nkB][PANEL]
               int NB, int nkB, int G, int64_t McPad,
               double* __restrict__ partial,    // [G][McPad]                (SUMSQ variants)
               double* __restrict__ Cpacked,    // [tiles][NB*8][PANEL]      (EPI_SUMSQ_PACKED)
               double* __restrict__ Cplain, int64_t ldc) {  // [McPad][ldc] (PLAIN variants)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* sA = reinterpret_cast<double*>(smem_raw);
  double* sB = sA + TG_STAGES * PANEL;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + TG_STAGES * PANEL);
  uint64_t* empty = full + TG_STAGES;
  double* red = reinterpret_cast<double*>(empty + TG_STAGES);  // [2][BT]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, g = blockIdx.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TG_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  const double* bTile = BP + (int64_t)tile * nkB * PANEL;

  if (warp == TG_CONSUMER_WARPS) {
    // ===== producer warp: one elected lane issues the bulk copies =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0;; ++i) {
        const int I = serpentine_rowblock(i, g, G);
        if (I >= NB) break;
        int k0, k1;
        int64_t off;
        rowblock_range<UPPER>(I, nkB, k0, k1, off);
        const double* aRow = TP + off * PANEL;
        for (int kc = k0; kc < k1; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], 2 * PANEL * sizeof(double));
          bulk_g2s(sA + stage * PANEL, aRow + (int64_t)kc * PANEL, PANEL * sizeof(double), &full[stage]);
          bulk_g2s(sB + stage * PANEL, bTile + (int64_t)kc * PANEL, PANEL * sizeof(double), &full[stage]);
          if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int wm = warp >> 2, wt = warp & 3;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  double colsum[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) colsum[j][0] = colsum[j][1] = 0.0;

  int stage = 0;
  uint32_t phase = 0;
  for (int i = 0;; ++i) {
    const int I = serpentine_rowblock(i, g, G);
    if (I >= NB) break;
    int k0, k1;
    int64_t off;
    rowblock_range<UPPER>(I, nkB, k0, k1, off);
    for (int kc = k0; kc < k1; ++kc) {
      mbar_wait(&full[stage], phase);
      const double2* a2 = reinterpret_cast<const double2*>(sA + stage * PANEL);
      const double2* b2 = reinterpret_cast<const double2*>(sB + stage * PANEL);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        double2 af[8], bf[4];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii) af[ii] = a2[((wm * 8 + ii) * 2 + p) * 32 + lane];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = b2[((wt * 4 + j) * 2 + p) * 32 + lane];
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[ii][j][0], acc[ii][j][1], af[ii].x, bf[j].x);
#pragma unroll
        for (int ii = 0; ii < 8; ++ii)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[ii][j][0], acc[ii][j][1], af[ii].y, bf[j].y);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == TG_STAGES) { stage = 0; phase ^= 1; }
    }
    // ---- row-block epilogue ----
    if (EPI == EPI_SUMSQ_PACKED) {
      // C[n,t] -> B-operand panel of the next GEMM: "k" = n % 16, "row" = t within the tile
      double* cTile = Cpacked + ((int64_t)tile * NB * (BM / BK) + (int64_t)I * (BM / BK)) * PANEL;
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        const int np = wm * 4 + (ii >> 1);  // 16-row panel within the row-block
        double* pn = cTile + (int64_t)np * PANEL;
        const int p = ii & 1, s = lane >> 4, kq = (lane >> 2) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int g8 = wt * 4 + j, rr = (lane & 3) * 2 + c;
            pn[(((g8 * 2 + p) * 32 + rr * 4 + kq) << 1) + s] = acc[ii][j][c];
          }
      }
    }
    if (EPI == EPI_PLAIN || EPI == EPI_SUMSQ_PLAIN) {
      const int64_t tbase = (int64_t)tile * BT + wt * 32 + (lane & 3) * 2;
      const int64_t nbase = (int64_t)I * BM + wm * 64 + (lane >> 2);
#pragma unroll
      for (int ii = 0; ii < 8; ++ii)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c)
            Cplain[(tbase + j * 8 + c) * ldc + nbase + ii * 8] = acc[ii][j][c];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int ii = 0; ii < 8; ++ii) {
        s0 = fma(acc[ii][j][0], acc[ii][j][0], s0);
        s1 = fma(acc[ii][j][1], acc[ii][j][1], s1);
        acc[ii][j][0] = 0.0;
        acc[ii][j][1] = 0.0;
      }
      colsum[j][0] += s0;
      colsum[j][1] += s1;
    }
  }

  if (EPI == EPI_PLAIN) return;
  // reduce over the 8 row-lanes (lane / 4) of the warp, then over the two row-warps
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      double v = colsum[j][c];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      colsum[j][c] = v;
    }
  if (lane < 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) red[wm * BT + wt * 32 + j * 8 + lane * 2 + c] = colsum[j][c];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(TG_CONSUMER_WARPS * 32));
  const int tid = threadIdx.x;
  if (tid < BT) partial[(int64_t)g * McPad + (int64_t)tile * BT + tid] = red[tid] + red[BT + tid];
}

// ------------------------------------------------------------------------------------------------
// acquisition tails (trieste/acquisition/function/function.py:221-223, 415-416; log-EI is ours)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ndtr_tfp(double x) {  // tfp special_math._ndtr piecewise form
  const double hs2 = 0.7071067811865476;
  double w = x * hs2, z = fabs(w);
  double y = (z < hs2) ? 1.0 + erf(w) : ((w > 0.0) ? 2.0 - erfc(z) : erfc(z));
  return 0.5 * y;
}
// aux: second parameter of the tail (the likelihood noise variance for AEI; unused otherwise)
__device__ __forceinline__ double acq_value(int acq, double param, double aux, double mean, double var) {
  const double sigma = sqrt(var);
  if (acq == TB_ACQ_LCB) return mean - param * sigma;
  if (acq == TB_ACQ_NEG_LCB) return -(mean - param * sigma);
  const double z = (param - mean) / sigma;
  if (acq == TB_ACQ_PBT) return ndtr_tfp(z);
  if (acq == TB_ACQ_EI || acq == TB_ACQ_AEI) {
    const double pdf_term = sigma * exp(-0.5 * z * z) * 0.3989422804014327;  // variance * N(eta; mean, sigma)
    const double ei = (param - mean) * ndtr_tfp(z) + pdf_term;
    if (acq == TB_ACQ_EI) return ei;
    return ei * (1.0 - sqrt(aux) / sqrt(aux + var));  // function.py:318-325
  }
  // log-EI: log(sigma) + log(phi(z) + z Phi(z))
  double lh;
  if (z > -1.0) {
    lh = log(z * ndtr_tfp(z) + exp(-0.5 * z * z) * 0.3989422804014327);
  } else {
    double t;
    if (z < -1e3) {
      double iz2 = 1.0 / (z * z);
      t = (1.0 - 3.0 * iz2) * iz2;
    } else {
      t = 1.0 + z * 1.2533141373155003 * erfcx(-z * 0.7071067811865476);
    }
    lh = -0.5 * z * z - 0.9189385332046727 + log(t);
  }
  return lh + log(sigma);
}

// d acq / d mean and d acq / d var (for the gradient path); clipped variance has zero gradient
__device__ __forceinline__ void acq_partials(int acq, double param, double aux, double mean, double var,
                                             bool clipped, double& dmu, double& dvar) {
  const double sigma = sqrt(var);
  if (acq == TB_ACQ_LCB || acq == TB_ACQ_NEG_LCB) {
    double sgn = (acq == TB_ACQ_LCB) ? 1.0 : -1.0;
    dmu = sgn;
    dvar = clipped ? 0.0 : -sgn * param / (2.0 * sigma);
    return;
  }
  const double z = (param - mean) / sigma;
  const double pdf = exp(-0.5 * z * z) * 0.3989422804014327;
  const double cdf = ndtr_tfp(z);
  if (acq == TB_ACQ_PBT) {
    dmu = -pdf / sigma;
    dvar = clipped ? 0.0 : -pdf * z / (2.0 * var);
    return;
  }
  if (acq == TB_ACQ_EI) {
    dmu = -cdf;
    dvar = clipped ? 0.0 : pdf / (2.0 * sigma);
    return;
  }
  if (acq == TB_ACQ_AEI) {
    const double ei = (param - mean) * cdf + sigma * pdf;
    const double tv = aux + var;
    const double aug = 1.0 - sqrt(aux) / sqrt(tv);
    dmu = -cdf * aug;
    dvar = clipped ? 0.0 : pdf / (2.0 * sigma) * aug + ei * 0.5 * sqrt(aux) / (tv * sqrt(tv));
    return;
  }
  // log-EI: d/dmu = -Phi/(sigma h), d/dvar = phi/(2 sigma^2 h)... with h = phi + z Phi; use the
  // erfcx ratio R = Phi/phi for stability: Phi/h = R/(1+zR), phi/h = 1/(1+zR)
  double R, one_zR;
  if (z > -1.0) {
    R = cdf / pdf;
    one_zR = 1.0 + z * R;
  } else if (z < -1e3) {
    double iz2 = 1.0 / (z * z);
    one_zR = (1.0 - 3.0 * iz2) * iz2;
    R = (one_zR - 1.0) / z;
  } else {
    R = 1.2533141373155003 * erfcx(-z * 0.7071067811865476);
    one_zR = 1.0 + z * R;
  }
  dmu = -R / (sigma * one_zR);
  dvar = clipped ? 0.0 : 1.0 / (2.0 * var * one_zR);
}

struct BestPair {
  double v;
  int64_t i;
};
__device__ __forceinline__ void best_merge(double& v, int64_t& i, double v2, int64_t i2) {
  // first-max semantics of tf.math.argmax (optimizer.py:149): larger value wins, ties -> lower index
  if (v2 > v || (v2 == v && i2 < i)) {
    v = v2;
    i = i2;
  }
}

// ---- min-value entropy search (entropy.py:193-213): mean over the S min-value samples of
//   -gamma r / 2 - log Phi(-gamma),  gamma = (y*_s - mean) / sd,  r = phi(gamma) / Phi(-gamma)
// log Phi(x) and r are evaluated through erfcx for x < -1 (no cancellation, no underflow)
__device__ __forceinline__ void mes_terms(double gamma, double& log_cdf_neg, double& ratio) {
  const double x = -gamma;
  if (x > -1.0) {
    const double c = ndtr_tfp(x);
    log_cdf_neg = (x > 8.0) ? -ndtr_tfp(-x) : log(c);
    ratio = exp(-0.5 * x * x) * 0.3989422804014327 / c;
  } else {
    const double e = 0.5 * erfcx(-x * 0.7071067811865476);  // Phi(x) = e * exp(-x^2/2)
    log_cdf_neg = log(e) - 0.5 * x * x;
    ratio = 0.3989422804014327 / e;
  }
}
constexpr double MES_CLAMP_LB = 1e-8;  // entropy.py:47
__device__ __forceinline__ double mes_value(const double* __restrict__ samp, int ns, double mean, double var) {
  const double sd = fmax(sqrt(var), MES_CLAMP_LB);
  double acc = 0.0;
  for (int s = 0; s < ns; ++s) {
    const double gamma = (samp[s] - mean) / sd;
    double lc, r;
    mes_terms(gamma, lc, r);
    acc += -0.5 * gamma * r - lc;
  }
  return acc / (double)ns;
}
// d/dmean and d/dvar of the above: df/dgamma = r/2 - gamma r (r - gamma) / 2, dgamma/dmean = -1/sd,
// dgamma/dvar = -gamma / (2 var)
__device__ __forceinline__ void mes_partials(const double* __restrict__ samp, int ns, double mean, double var, bool clipped,
                                             double& dmu, double& dvar) {
  const double sd = fmax(sqrt(var), MES_CLAMP_LB);
  double am = 0.0, av = 0.0;
  for (int s = 0; s < ns; ++s) {
    const double gamma = (samp[s] - mean) / sd;
    double lc, r;
    mes_terms(gamma, lc, r);
    const double dg = 0.5 * r - 0.5 * gamma * r * (r - gamma);
    am += dg;
    av += dg * gamma;
  }
  dmu = -am / (sd * (double)ns);
  dvar = clipped ? 0.0 : -av / (2.0 * var * (double)ns);
}

// one thread per candidate of the chunk; block-level first-max argmax
__global__ void __launch_bounds__(256)
tail_kernel(const double* __restrict__ partial, int G, int64_t McPad, const double* __restrict__ mean,
            int64_t Mc, int64_t idx0, double variance, int acq, double param, double aux,
            const double* __restrict__ samp, int nsamp, double* __restrict__ out_vals, double* __restrict__ out_mean, double* __restrict__ out_var,
            double* __restrict__ blk_best, int64_t* __restrict__ blk_idx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double bv = -INFINITY;
  int64_t bi = INT64_MAX;
  if (t < Mc) {
    double ss = 0.0;
    for (int g = 0; g < G; ++g) ss += partial[(int64_t)g * McPad + t];
    double var = fmax(variance - ss, 1e-12);
    double mu = mean[t];
    if (out_mean) out_mean[t] = mu;
    if (out_var) out_var[t] = var;
    if (acq >= 0) {
      double v = (acq == TB_ACQ_MES) ? mes_value(samp, nsamp, mu, var) : acq_value(acq, param, aux, mu, var);
      if (out_vals) out_vals[t] = v;
      if (v == v) { bv = v; bi = idx0 + t; }
    }
  }
  if (blk_best == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
    int64_t i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    best_merge(bv, bi, v2, i2);
  }
  __shared__ double sv[8];
  __shared__ int64_t si[8];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = bv;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best_merge(bv, bi, sv[w], si[w]);
    blk_best[blockIdx.x] = bv;
    blk_idx[blockIdx.x] = bi;
  }
}

// fold the per-block winners of one chunk into the running best (single block)
__global__ void __launch_bounds__(256)
argmax_fold_kernel(const double* __restrict__ blk_best, const int64_t* __restrict__ blk_idx, int nblk,
                   double* __restrict__ run_best, int64_t* __restrict__ run_idx) {
  double bv = -INFINITY;
  int64_t bi = INT64_MAX;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) best_merge(bv, bi, blk_best[i], blk_idx[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
    int64_t i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    best_merge(bv, bi, v2, i2);
  }
  __shared__ double sv[8];
  __shared__ int64_t si[8];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = bv;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) best_merge(bv, bi, sv[w], si[w]);
    double rv = *run_best;
    int64_t ri = *run_idx;
    best_merge(rv, ri, bv, bi);
    *run_best = rv;
    *run_idx = ri;
  }
}

}  // namespace tb
