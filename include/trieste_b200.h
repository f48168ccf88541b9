/*
 * trieste_b200 — C-ABI of the B200-native GP-posterior + acquisition engine.
 *
 * This is the drop-in boundary for the hot path named by BASELINE.json `north_star`
 * (SURVEY.md §8b).  The reference has no FFI: its boundary is a set of Python structural
 * protocols.  Each entry point below cites the reference interface it stands behind
 * (paths relative to /root/reference/).  The Python host layer (`trieste_b200/`) binds these
 * with ctypes and mirrors the reference's class/method names on top; INTEGRATION.md shows the
 * binding a trieste maintainer would add.
 *
 * Conventions
 *   - every function returns a tb_status: 0 on success, otherwise the class of the failure (below);
 *     tb_last_error() returns the thread-local message.  The Python layer maps the CODE (not the text) to the
 *     exception the reference raises for the same condition; nothing aborts the process (the BO loop records
 *     exceptions, bayesian_optimizer.py:855-875).
 *   - plain pointers + sizes only.  Every array pointer may be a HOST pointer or a DEVICE pointer
 *     on the handle's GPU (detected with cudaPointerGetAttributes); host buffers are staged
 *     through the handle's stream inside the call.
 *   - row-major, dense; fp64 unless the handle was created with TB_F32.
 *   - one caller per handle, one CUDA stream per handle, synchronous return
 *     (results are immediately `.numpy()`-ed by the reference, optimizer.py:665-666).
 */
#ifndef TRIESTE_B200_H
#define TRIESTE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb_gp tb_gp;   /* exact-GPR posterior: owns device copies of X, Linv, alpha, hyper-params */
typedef struct tb_rff tb_rff; /* random-Fourier-feature trajectory: owns W, b, theta */

enum tb_status {
  TB_OK = 0,
  TB_ERR_INVALID = 1, /* bad argument / unmet precondition: ValueError (tf.errors.InvalidArgumentError in the reference) */
  TB_ERR_RUNTIME = 2, /* CUDA or library failure, no GPU: RuntimeError (there is no CPU fallback) */
  TB_ERR_NUMERIC = 3  /* a Cholesky factorisation met a non-positive-definite matrix: ValueError, like tf.linalg.cholesky's
                         InvalidArgumentError ("Cholesky decomposition was not successful") */
};
enum tb_dtype { TB_F64 = 0, TB_F32 = 1 };
/* gpflow.kernels.{SquaredExponential,Matern12,Matern32,Matern52}; trieste default Matern52
 * (models/gpflow/builders.py:399) */
enum tb_kernel { TB_RBF = 0, TB_MATERN12 = 1, TB_MATERN32 = 2, TB_MATERN52 = 3 };
enum tb_acq {
  TB_ACQ_EI = 0,      /* expected_improvement.__call__, acquisition/function/function.py:215-223 */
  TB_ACQ_LOG_EI = 1,  /* log of the above; ABSENT in the reference */
  TB_ACQ_NEG_LCB = 2, /* NegativeLowerConfidenceBound, function.py:358-359 (−lower_confidence_bound :415-416) */
  TB_ACQ_LCB = 3,     /* lower_confidence_bound, function.py:389-418 */
  TB_ACQ_PBT = 4,     /* probability_below_threshold.__call__, function.py:501-509 (ProbabilityOfImprovement :47-93,
                         ProbabilityOfFeasibility :421-478): Normal(mean, sqrt(var)).cdf(param) */
  TB_ACQ_AEI = 5,     /* augmented_expected_improvement.__call__, function.py:311-325: EI(param = eta) times
                         1 − sqrt(noise)/sqrt(noise + var), noise = the handle's likelihood variance */
  TB_ACQ_MES = 6      /* min_value_entropy_search.__call__, entropy.py:193-213: mean over the min-value samples set by
                         tb_acq_set_min_value_samples of −γ·φ(γ)/(2Φ(−γ)) − log Φ(−γ), γ = (y* − mean)/sd; param unused */
};

/* ---- errors / build info ------------------------------------------------------------------ */
const char* tb_last_error(void);
const char* tb_version(void);
int tb_device_count(int* count);

/* ---- model handle ----------------------------------------------------------------------------
 * mirrors GaussianProcessRegression (models/gpflow/models.py:69-186) + GPflowPredictor's posterior
 * cache (models/gpflow/interface.py:89-112). */
int tb_gp_create(tb_gp** out, int device, int dtype);
int tb_gp_destroy(tb_gp* gp);

/* GPR data Variables assign, models.py:171-186 (update_encoded).  X [N,D], y [N] (E = 1). */
int tb_gp_set_data(tb_gp* gp, const void* X, const void* y, int64_t N, int D);

/* kernel / likelihood / mean-function hyper-parameters (gpflow Parameters read through
 * get_kernel / get_observation_noise / get_mean_function, models/interfaces.py:166-225).
 * lengthscales: n_ls = 1 (isotropic) or D (ARD), always double. */
int tb_gp_set_hyper(tb_gp* gp, int kernel, double variance, const double* lengthscales, int n_ls,
                    double noise_variance, double mean_const);

/* update_posterior_cache (interface.py:108-112): err = y − m(X), L = chol(K(X,X) + σ²I) (hand-written blocked
 * Cholesky on the DMMA pipe, once per BO step), then Linv and alpha = K⁻¹err packed for the kernels. */
int tb_gp_update_posterior_cache(tb_gp* gp);

/* Append m (1..64) observations to the data AND extend the cached L, Linv and alpha in O(m N²) — the incremental form of
 * update_encoded (models.py:171-186) followed by update_posterior_cache (interface.py:108-112), which in the reference
 * refactorise from scratch every BO step.  Requires a valid cache and unchanged hyper-parameters;
 * same error behaviour as tb_gp_update_posterior_cache (on failure the cache is left invalid).  Xnew [m,D], ynew [m]. */
int tb_gp_append_data(tb_gp* gp, const void* Xnew, const void* ynew, int64_t m);

/* copy out the cached Cholesky factor L [N,N] row-major lower (tests / diagnostics). */
int tb_gp_get_cholesky(tb_gp* gp, void* L_out);

/* predict_encoded (interface.py:119-124): Xc [M,D] → mean [M], var [M] (clipped to ≥ 1e-12). */
int tb_gp_predict(tb_gp* gp, const void* Xc, int64_t M, void* mean, void* var);

/* predict_joint_encoded (interface.py:126-133): Xc [B,q,D] → mean [B,q], cov [B,q,q]
 * (diagonal clipped to ≥ 1e-12). */
int tb_gp_predict_joint(tb_gp* gp, const void* Xc, int64_t B, int q, void* mean, void* cov);

/* ---- fused predict + acquisition tail --------------------------------------------------------
 * acq: tb_acq; param = eta (EI / log-EI) or beta (LCB).  Xc [M,D] → out [M].
 * grad (nullable): [M,D] = d out / d Xc (what tfp.math.value_and_gradient returns at
 * acquisition/optimizer.py:621-629). */
int tb_acq_eval(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out, void* grad);

/* generate_random_search_optimizer / _get_max_discrete_points (optimizer.py:124-150, 973-1011):
 * fused evaluation + first-max argmax.  best_value (1 scalar of the handle dtype, host),
 * best_index (host).  out may be NULL (values are then never written to HBM). */
int tb_acq_argmax(tb_gp* gp, int acq, double param, const void* Xc, int64_t M, void* out,
                  void* best_value, int64_t* best_index);

/* min_value_entropy_search.update / __init__ (entropy.py:166-191): the samples of the objective minimum y* are a
 * tf.Variable assigned in place; here they are a device array owned by the handle.  samples [S] (the reference holds them
 * as [S,1]), always double, S ≥ 1.  Required before any TB_ACQ_MES evaluation. */
int tb_acq_set_min_value_samples(tb_gp* gp, const double* samples, int S);

/* _perform_parallel_continuous_optimization (acquisition/optimizer.py:566-697) with one ScipyOptimizerGreenlet per start
 * (:700-745, L-BFGS-B): here P independent projected L-BFGS runs advance together ON THE DEVICE — one batched fused
 * value+gradient evaluation of all active trial points per round, then one warp per run updates its curvature history,
 * line search and convergence tests (SciPy's option names: maxcor ≤ 16, maxiter, maxls, gtol on the projected gradient,
 * ftol on the relative decrease).  Maximises the acquisition `acq` inside the box.  lower, upper [D]; starts [P,D];
 * x_out [P,D], f_out [P] (maximised values), success [P] (1 = converged), nfev [P].  All arrays double / as declared
 * (the reference's SciPy side is fp64 whatever the model dtype, optimizer.py:635-639), host or device pointers. */
int tb_acq_maximize(tb_gp* gp, int acq, double param, const double* lower, const double* upper, const double* starts,
                    int64_t P, int maxcor, int maxiter, int maxls, double gtol, double ftol, double* x_out, double* f_out,
                    int32_t* success, int64_t* nfev);

/* batch_monte_carlo_expected_improvement.__call__ (function.py:1181-1186) on top of
 * BatchReparametrizationSampler.sample (models/gpflow/sampler.py:208-287):
 * Xc [B,q,D], eps [q,S] (the sampler's fixed base samples, injected) → out [B]. */
int tb_acq_batch_mc_ei(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S,
                       double eta, double jitter, void* out);

/* The same value together with its gradient w.r.t. the query batches — what tfp.math.value_and_gradient
 * (acquisition/optimizer.py:621-629) differentiates when a batch function is maximised through batchify_joint
 * (:897-936): reduce_min routes to the arg-min sample, maximum(., 0) to the active ones, the Cholesky of the joint
 * covariance by its reverse-mode rule.  out [B], grad [B,q,D]. */
int tb_acq_batch_mc_ei_grad(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S, double eta,
                            double jitter, void* out, void* grad);

/* BatchReparametrizationSampler.sample (sampler.py:208-287): → samples [B,S,q]. */
int tb_gp_reparam_sample(tb_gp* gp, const void* Xc, int64_t B, int q, const void* eps, int S,
                         double jitter, void* samples);

/* model.sample over a large point set (GPflowPredictor.sample, interface.py:135-138 -> gpflow predict_f_samples: joint
 * mean / full covariance, Cholesky of cov + jitter I, mean + L z) — the call behind ExactThompsonSampler
 * (acquisition/sampler.py:85-123).  Xc [M,D] (handle dtype), z [S,M] standard-normal draws (double), out [S,M] (handle
 * dtype).  1 ≤ M ≤ 16384; the covariance is built and factorised on the device (DMMA Gram + the blocked Cholesky of the
 * cache build). */
int tb_gp_sample_joint(tb_gp* gp, const void* Xc, int64_t M, const double* z, int S, double jitter, void* out);

/* covariance_between_points_encoded (models/gpflow/models.py:188-254): posterior covariance between two point sets,
 * K12 − Kx1 (K + σ²I)⁻¹ Kx2, no clipping.  X1 [M1,D], X2 [M2,D], out [M1,M2] row-major (handle dtype); M1 + M2 ≤ 16384.
 * (The reference's leading dimensions of X1 are flattened into M1 by the caller.) */
int tb_gp_covariance_between_points(tb_gp* gp, const void* X1, int64_t M1, const void* X2, int64_t M2, void* out);

/* ---- streaming reductions over candidate scores ----------------------------------------------
 * tf.math.top_k as used by generate_initial_points (optimizer.py:321-335): values [M] →
 * top values [k] (descending, ties → lower index), indices [k].  Host or device pointers. */
int tb_topk(int device, int dtype, const void* values, int64_t M, int k, void* top_values,
            int64_t* top_indices);

/* ---- random-Fourier-feature trajectories -----------------------------------------------------
 * feature_decomposition_trajectory.__call__ (models/gpflow/sampler.py:901-936) with
 * ResampleableRandomFourierFeatureFunctions (:741-806): always fp64 (sampler.py:782). */
int tb_rff_create(tb_rff** out, int device);
int tb_rff_destroy(tb_rff* r);
/* W [F,D], b [F], lengthscales [D], theta [nb,F] (nb trajectories = batch size B). */
int tb_rff_set(tb_rff* r, const double* W, const double* b, int F, int D, const double* lengthscales,
               double variance, double mean_const);
int tb_rff_set_theta(tb_rff* r, const double* theta, int nb);
/* Xc [M,D] evaluated under all nb trajectories → out [M,nb]; ThompsonSamplerFromTrajectory
 * (acquisition/sampler.py:262-271): argmin per trajectory → min_value [nb], min_index [nb]
 * (either may be NULL). */
int tb_rff_eval(tb_rff* r, const void* Xc, int64_t M, void* out, double* min_value,
                int64_t* min_index);
/* DecoupledTrajectorySampler (models/gpflow/sampler.py:594-738) / ResampleableDecoupledFeatureFunctions (:809-855):
 * adds the canonical features  sum_j v[b][j] k(x, X_j)  to trajectory b.  X [N,D] raw training inputs (host),
 * v [nb,N] column layout [trajectory][training point] (host or device).  N = 0 switches the term off. */
int tb_rff_set_canonical(tb_rff* r, int kernel, const double* X, int64_t N, const double* v, int nb);
/* (K(X,X) + noise I)^-1 B = Linv^T (Linv B) through the cached triangular inverse: B, out [nrhs][N] (each right-hand side contiguous);
 * the v-weights of a decoupled trajectory (sampler.py:716, gpflux compute_A_inv_b).  fp64, host or device. */
int tb_gp_kinv_apply(tb_gp* gp, const double* B, int nrhs, double* out);

/* ---- instrumentation (bench / tests) ---------------------------------------------------------
 * kernels launched by this library in this process since the last reset; device time (ms) of the
 * dominant kernel (triangular DMMA GEMM) accumulated with CUDA events on the handle's stream when
 * profiling is enabled. */
int64_t tb_launch_count(void);
void tb_launch_count_reset(void);
/* engine of the variance GEMM: 0 = native fp64 (DMMA), 1 = fp64-accurate emulation on the INT8 tensor cores
 * (Ozaki splitting, tcgen05 kind::i8; same stated tolerances; N <= 16384, larger models fall back to engine 0), 2 = engine 1
 * with the number of digit products pinned to the full 21 (engine 1 drops to 15 — fp32 handles: 6 — when the a-priori error
 * estimate of the cache allows it; csrc/ozaki5.cuh).  The
 * engine serves every path: predict / acquisition values, gradients (V = K^-1 k* as a dense digit GEMM) and the joint
 * paths (predict_joint / reparam samples / MC-qEI through the store-A epilogue). */
int tb_gp_set_engine(tb_gp* gp, int engine);
/* what the variance GEMM of this handle runs right now: int8 digit products per k-step (15 or 21; fp32 handles 6 or 10;
 * 0 = native fp64 engine) and, for the reduced modes, the a-priori estimate of max |Δvar| / σ_f² that admitted them.
 * Needs a valid cache.  Either output may be NULL. */
int tb_gp_engine_info(tb_gp* gp, int* digit_products, double* error_estimate);
int tb_gp_profile(tb_gp* gp, int enable);
/* the handle's CUDA stream (cudaStream_t as void*), so callers can record CUDA events on the stream the
 * kernels are launched on (torch.cuda.ExternalStream in bench.py). */
int tb_gp_stream(tb_gp* gp, void** stream);
int tb_gp_profile_read(tb_gp* gp, double* trigemm_ms, int64_t* trigemm_launches, double* flops);

#ifdef __cplusplus
}
#endif
#endif /* TRIESTE_B200_H */
