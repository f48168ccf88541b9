// Device functions shared by every translation unit: the stationary kernels on the scaled squared distance and the
// serpentine row-block order of the triangular GEMMs.
#pragma once
#include "common.cuh"
#include "../../include/trieste_b200.h"
#include "fastmath.cuh"

namespace tb {

// ------------------------------------------------------------------------------------------------
// stationary kernels on the scaled squared distance (GPflow kernels/stationaries.py semantics,
// SURVEY.md Appendix A1; r = sqrt(max(r2, 1e-36)) for the Matern family)
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ double kernel_from_r2(double r2, double variance) {
  if (KIND == TB_RBF) return variance * exp(-0.5 * r2);
  double r = sqrt(fmax(r2, 1e-36));
  if (KIND == TB_MATERN12) return variance * exp(-r);
  if (KIND == TB_MATERN32) {
    double s = 1.7320508075688772 * r;
    return variance * (1.0 + s) * exp(-s);
  }
  double s = 2.23606797749979 * r;
  return variance * (1.0 + s + (5.0 / 3.0) * r * r) * exp(-s);
}

// The same kernels on the branch-free exp / sqrt of fastmath.cuh (T = the 64-entry 2^(j/64) table in shared memory).  r2 may come
// from the expansion form |a|^2 + |b|^2 - 2 a.b and be slightly negative, exactly as in GPflow's square_distance.
template <int KIND>
__device__ __forceinline__ double kernel_from_r2_fast(double r2, double variance, const double* T, const fm::Consts& c) {
  if (KIND == TB_RBF) return variance * fm::exp_neg(0.5 * fm::clamp_below(r2, 0.0), T, c);
  const double q = fm::clamp_below(r2, 1e-36);
  if (KIND == TB_MATERN12) return variance * fm::exp_neg(fm::sqrt_pos(q), T, c);
  if (KIND == TB_MATERN32) {
    const double s = fm::sqrt_pos(3.0 * q);
    return variance * (1.0 + s) * fm::exp_neg(s, T, c);
  }
  const double s2 = 5.0 * q;
  const double s = fm::sqrt_pos(s2);
  return variance * fma(s2, c.third, 1.0 + s) * fm::exp_neg(s, T, c);
}

template <int KIND>
__device__ __forceinline__ double kernel_dr2_fast(double r2, double variance, const double* T, const fm::Consts& c) {
  if (KIND == TB_RBF) return -0.5 * variance * fm::exp_neg(0.5 * fm::clamp_below(r2, 0.0), T, c);
  const double q = fm::clamp_below(r2, 1e-36);
  if (KIND == TB_MATERN12) {
    const double r = fm::sqrt_pos(q);
    return -variance * fm::exp_neg(r, T, c) / (2.0 * r);
  }
  if (KIND == TB_MATERN32) return -1.5 * variance * fm::exp_neg(fm::sqrt_pos(3.0 * q), T, c);
  const double s = fm::sqrt_pos(5.0 * q);
  return -(5.0 / 6.0) * variance * (1.0 + s) * fm::exp_neg(s, T, c);
}

// dk/d(r2) (for gradients w.r.t. x*: dk/dx*_d = dk/dr2 * 2 (x*_d - x_d) / l_d^2)
template <int KIND>
__device__ __forceinline__ double kernel_dr2(double r2, double variance) {
  if (KIND == TB_RBF) return -0.5 * variance * exp(-0.5 * r2);
  double r = sqrt(fmax(r2, 1e-36));
  if (KIND == TB_MATERN12) return -variance * exp(-r) / (2.0 * r);
  if (KIND == TB_MATERN32) return -1.5 * variance * exp(-1.7320508075688772 * r);
  double s = 2.23606797749979 * r;
  return -(5.0 / 6.0) * variance * (1.0 + s) * exp(-s);
}

__device__ __forceinline__ int serpentine_rowblock(int i, int g, int G) {
  // i-th row-block of group g (increasing in i); balances the triangular cost across groups
  int base = (i >> 1) * 2 * G;
  return (i & 1) ? base + 2 * G - 1 - g : base + g;
}

}  // namespace tb
