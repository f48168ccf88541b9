// Kernels beyond the marginal predict + acquisition path: gradients, predict_joint / MC-qEI,
// RFF trajectories, top-k.
#pragma once
#include "gp_handle.cuh"

namespace tb {

inline int extra_kernels_init() { return 0; }

inline int gradient_chunk(tb_gp*, int, double, const double*, int64_t, int, int, int64_t, double*) {
  return fail("gradients are not implemented in this build");
}

}  // namespace tb
