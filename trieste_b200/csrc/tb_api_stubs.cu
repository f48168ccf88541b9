// Entry points not implemented yet: exported so the ABI is complete; they fail loudly.
#include "common.cuh"
#include "../../include/trieste_b200.h"
extern "C" {
int tb_gp_predict_joint(tb_gp*, const void*, int64_t, int, void*, void*) { return tb::fail("tb_gp_predict_joint: not implemented in this build"); }
int tb_acq_batch_mc_ei(tb_gp*, const void*, int64_t, int, const void*, int, double, double, void*) { return tb::fail("tb_acq_batch_mc_ei: not implemented in this build"); }
int tb_gp_reparam_sample(tb_gp*, const void*, int64_t, int, const void*, int, double, void*) { return tb::fail("tb_gp_reparam_sample: not implemented in this build"); }
int tb_topk(int, int, const void*, int64_t, int, void*, int64_t*) { return tb::fail("tb_topk: not implemented in this build"); }
int tb_rff_create(tb_rff**, int) { return tb::fail("tb_rff_create: not implemented in this build"); }
int tb_rff_destroy(tb_rff*) { return 0; }
int tb_rff_set(tb_rff*, const double*, const double*, int, int, const double*, double, double) { return tb::fail("tb_rff_set: not implemented in this build"); }
int tb_rff_set_theta(tb_rff*, const double*, int) { return tb::fail("tb_rff_set_theta: not implemented in this build"); }
int tb_rff_eval(tb_rff*, const void*, int64_t, void*, double*, int64_t*) { return tb::fail("tb_rff_eval: not implemented in this build"); }
}
