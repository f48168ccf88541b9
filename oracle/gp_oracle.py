"""CPU oracle for the GP-posterior + acquisition hot path (TEST INFRASTRUCTURE — not product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product path (``trieste_b200``) never does.

PARITY STATUS: **parity unpinned by reference data**.  The arithmetic of this path lives in
un-vendored third parties (gpflow==2.9.2, gpflux==0.4.4, tensorflow-probability==0.24.0,
tensorflow==2.16.1; ``/root/reference/tests/latest/constraints.txt``) that are neither under
``/root/reference`` nor installed here, and the reference's tests hold no literal golden vectors for
``predict``.  This oracle restates the published algorithms at the
reference's own call sites, and is pinned by (i) the reference tests' closed-form / Monte-Carlo
known answers restated in ``tests/test_oracle.py`` and (ii) an independent third implementation
(scikit-learn ``GaussianProcessRegressor``) through the committed fixtures in ``tests/golden/``.

All citations are relative to ``/root/reference/``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
from scipy import linalg as sla
from scipy import special as ssp

JITTER = 1e-6  # trieste/utils/misc.py:183 (DEFAULTS.JITTER)
VAR_CLIP = 1e-12  # trieste/models/gpflow/interface.py:123

KERNEL_KINDS = ("rbf", "matern12", "matern32", "matern52")


# --------------------------------------------------------------------------------------------
# A1. Stationary kernels (EXT gpflow 2.9.2 kernels/stationaries.py; default Matern52 is built at
# trieste/models/gpflow/builders.py:399)
# --------------------------------------------------------------------------------------------
def scaled_square_dist(X1: np.ndarray, X2: np.ndarray, lengthscales: np.ndarray) -> np.ndarray:
    """r^2 via the expansion |a|^2 + |b|^2 - 2 a.b on inputs scaled by 1/lengthscale (GPflow's
    ``square_distance``); may be slightly negative."""
    A = X1 / lengthscales
    B = X2 / lengthscales
    return (A * A).sum(-1)[:, None] + (B * B).sum(-1)[None, :] - 2.0 * (A @ B.T)


def kernel_from_r2(kind: str, r2: np.ndarray, variance: float) -> np.ndarray:
    if kind == "rbf":
        return variance * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))  # GPflow IsotropicStationary.scaled_squared_euclid_dist
    if kind == "matern12":
        return variance * np.exp(-r)
    if kind == "matern32":
        s = math.sqrt(3.0) * r
        return variance * (1.0 + s) * np.exp(-s)
    if kind == "matern52":
        s = math.sqrt(5.0) * r
        return variance * (1.0 + s + (5.0 / 3.0) * np.square(r)) * np.exp(-s)
    raise ValueError(f"unknown kernel kind {kind!r}")


KERNEL_WORKERS = 1  # > 1: evaluate kernel matrices in column slabs on a thread pool (bench.py's CPU baseline; NumPy's
# elementwise loops release the GIL, so the Matern evaluation no longer runs on one core beside a multi-threaded dtrsm)


def kernel_matrix(kind, X1, X2, variance, lengthscales) -> np.ndarray:
    ls = np.broadcast_to(np.asarray(lengthscales, dtype=X1.dtype), (X1.shape[-1],))
    if KERNEL_WORKERS > 1 and X2.ndim == 2 and X2.shape[0] >= 4 * KERNEL_WORKERS:
        from concurrent.futures import ThreadPoolExecutor

        bounds = np.linspace(0, X2.shape[0], KERNEL_WORKERS + 1).astype(int)
        out = np.empty((X1.shape[0], X2.shape[0]), dtype=np.result_type(X1.dtype, X2.dtype))

        def slab(i):
            lo, hi = bounds[i], bounds[i + 1]
            out[:, lo:hi] = kernel_from_r2(kind, scaled_square_dist(X1, X2[lo:hi], ls), variance)

        with ThreadPoolExecutor(max_workers=KERNEL_WORKERS) as pool:
            list(pool.map(slab, range(KERNEL_WORKERS)))
        return out
    return kernel_from_r2(kind, scaled_square_dist(X1, X2, ls), variance)


# --------------------------------------------------------------------------------------------
# A2. GPR posterior with the once-per-step cache
# (trieste/models/gpflow/interface.py:89-133; algebra written out in-repo at models.py:208-238)
# --------------------------------------------------------------------------------------------
@dataclass
class GPRModel:
    kind: str
    X: np.ndarray  # [N, D]
    y: np.ndarray  # [N, 1]
    variance: float
    lengthscales: np.ndarray  # [D]
    noise: float
    mean_const: float
    L: Optional[np.ndarray] = None  # chol(K + noise I), lower
    err: Optional[np.ndarray] = None  # y - m(X)

    @property
    def dtype(self):
        return self.X.dtype


def build_model(kind, X, y, variance, lengthscales, noise, mean_const) -> GPRModel:
    X = np.ascontiguousarray(X)
    y = np.ascontiguousarray(y).reshape(-1, 1).astype(X.dtype)
    ls = np.broadcast_to(np.asarray(lengthscales, dtype=X.dtype), (X.shape[1],)).copy()
    m = GPRModel(kind, X, y, float(variance), ls, float(noise), float(mean_const))
    update_posterior_cache(m)
    return m


def update_posterior_cache(m: GPRModel) -> None:
    """interface.py:108-112 -> GPflow ``GPRPosterior._precompute``: err = y - m(X),
    L = chol(K(X,X) + noise I)."""
    K = kernel_matrix(m.kind, m.X, m.X, m.variance, m.lengthscales)
    K[np.diag_indices_from(K)] += m.noise
    m.L = np.linalg.cholesky(K).astype(m.X.dtype)
    m.err = m.y - m.mean_const


def predict_f(m: GPRModel, Xq: np.ndarray, full_cov: bool = False):
    """Unclipped GPflow ``GPRPosterior._conditional_with_precompute`` / ``base_conditional``:
    A = L^-1 Kmn; fvar = Knn - sum A^2 (or Knn - A^T A); A2 = L^-T A; fmean = A2^T err + m."""
    Xq = np.asarray(Xq, dtype=m.dtype)
    Kmn = kernel_matrix(m.kind, m.X, Xq, m.variance, m.lengthscales)  # [N, M]
    A = sla.solve_triangular(m.L, Kmn, lower=True, check_finite=False)
    if full_cov:
        Knn = kernel_matrix(m.kind, Xq, Xq, m.variance, m.lengthscales)
        fvar = Knn - A.T @ A
    else:
        fvar = m.variance - np.square(A).sum(0)  # K(x,x) diag = variance
    A2 = sla.solve_triangular(m.L.T, A, lower=False, check_finite=False)
    fmean = A2.T @ m.err + m.mean_const
    if full_cov:
        return fmean, fvar
    return fmean, fvar[:, None]


def predict(m: GPRModel, Xq: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``GPflowPredictor.predict_encoded`` (interface.py:119-124): [M,D] -> ([M,1],[M,1]),
    variance clipped to >= 1e-12."""
    mean, var = predict_f(m, Xq)
    return mean, np.clip(var, VAR_CLIP, np.finfo(var.dtype).max)


def predict_joint(m: GPRModel, Xq: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """``predict_joint_encoded`` (interface.py:126-133): [..., q, D] -> ([..., q, 1],
    [..., 1, q, q]); only the diagonal of cov is clipped."""
    Xq = np.asarray(Xq, dtype=m.dtype)
    lead = Xq.shape[:-2]
    q, D = Xq.shape[-2:]
    flat = Xq.reshape(-1, q, D)
    means = np.empty((flat.shape[0], q, 1), dtype=m.dtype)
    covs = np.empty((flat.shape[0], 1, q, q), dtype=m.dtype)
    for b in range(flat.shape[0]):
        mu, cov = predict_f(m, flat[b], full_cov=True)
        d = np.clip(np.diag(cov), VAR_CLIP, np.finfo(cov.dtype).max)
        cov = cov.copy()
        cov[np.diag_indices(q)] = d
        means[b] = mu
        covs[b, 0] = cov
    return means.reshape(*lead, q, 1), covs.reshape(*lead, 1, q, q)


def predict_batched(m: GPRModel, Xq: np.ndarray, chunk: int = 16384):
    """predict over many candidates in memory-bounded chunks — the reference's own scaling device
    is ``split_acquisition_function`` (trieste/acquisition/utils.py:31-80)."""
    means, vars_ = [], []
    for s in range(0, Xq.shape[0], chunk):
        mu, v = predict(m, Xq[s : s + chunk])
        means.append(mu)
        vars_.append(v)
    return np.concatenate(means), np.concatenate(vars_)


def covariance_between_points(m: GPRModel, X1: np.ndarray, X2: np.ndarray) -> np.ndarray:
    """trieste/models/gpflow/models.py:188-254: ``K12 - Kx1 (K + noise I)^-1 Kx2`` via two triangular solves against
    L = chol(K + noise I).  X1 [..., N, D], X2 [M, D] -> [..., 1, N, M] (no clipping)."""
    X1 = np.asarray(X1, dtype=m.dtype)
    X2 = np.asarray(X2, dtype=m.dtype)
    lead, n = X1.shape[:-2], X1.shape[-2]
    flat = X1.reshape(-1, X1.shape[-1])
    A1 = sla.solve_triangular(m.L, kernel_matrix(m.kind, m.X, flat, m.variance, m.lengthscales), lower=True, check_finite=False)
    A2 = sla.solve_triangular(m.L, kernel_matrix(m.kind, m.X, X2, m.variance, m.lengthscales), lower=True, check_finite=False)
    cov = kernel_matrix(m.kind, flat, X2, m.variance, m.lengthscales) - A1.T @ A2
    return cov.reshape(lead + (n, X2.shape[0]))[..., None, :, :]


def posterior_gradients(m: GPRModel, Xq: np.ndarray):
    """d mean / d x* and d var / d x* (what ``tfp.math.value_and_gradient`` differentiates at
    trieste/acquisition/optimizer.py:621-629).  Analytic: dmean = (dk*/dx)^T alpha,
    dvar = -2 (dk*/dx)^T V with V = K^-1 k*."""
    Xq = np.asarray(Xq, dtype=m.dtype)
    ls = m.lengthscales
    A_ = m.X / ls  # [N, D]
    B_ = Xq / ls  # [M, D]
    diff = B_[:, None, :] - A_[None, :, :]  # [M, N, D]  (x* - x)/l
    r2 = np.square(diff).sum(-1)
    if m.kind == "rbf":
        k = m.variance * np.exp(-0.5 * r2)
        dk_dr2 = -0.5 * k
    else:
        r = np.sqrt(np.maximum(r2, 1e-36))
        if m.kind == "matern12":
            dk_dr2 = -m.variance * np.exp(-r) / (2.0 * r)
        elif m.kind == "matern32":
            s3 = math.sqrt(3.0)
            dk_dr2 = -m.variance * 1.5 * np.exp(-s3 * r)
        else:
            s5 = math.sqrt(5.0)
            dk_dr2 = -m.variance * (5.0 / 6.0) * (1.0 + s5 * r) * np.exp(-s5 * r)
    dk = dk_dr2[:, :, None] * 2.0 * diff / ls  # [M, N, D]
    Kmn = kernel_matrix(m.kind, m.X, Xq, m.variance, m.lengthscales)
    alpha = sla.cho_solve((m.L, True), m.err, check_finite=False)[:, 0]  # [N]
    V = sla.cho_solve((m.L, True), Kmn, check_finite=False)  # [N, M]
    dmean = np.einsum("mnd,n->md", dk, alpha)
    dvar = -2.0 * np.einsum("mnd,nm->md", dk, V)
    return dmean, dvar


# --------------------------------------------------------------------------------------------
# A3. Elementwise acquisition tails
# --------------------------------------------------------------------------------------------
def ndtr(x: np.ndarray) -> np.ndarray:
    """EXT tfp 0.24 ``special_math._ndtr``: piecewise erf / erfc on w = x/sqrt(2)."""
    half_sqrt_2 = 0.5 * math.sqrt(2.0)
    w = x * half_sqrt_2
    z = np.abs(w)
    y = np.where(z < half_sqrt_2, 1.0 + ssp.erf(w), np.where(w > 0.0, 2.0 - ssp.erfc(z), ssp.erfc(z)))
    return 0.5 * y


def expected_improvement(mean, var, eta):
    """trieste/acquisition/function/function.py:221-223:
    ``(eta - mean) * normal.cdf(eta) + variance * normal.prob(eta)`` with Normal(mean, sqrt(var))."""
    sigma = np.sqrt(var)
    z = (eta - mean) / sigma
    cdf = ndtr(z)
    log_prob = -0.5 * z * z - np.log(sigma) - 0.5 * math.log(2.0 * math.pi)  # tfp Normal._log_prob
    return (eta - mean) * cdf + var * np.exp(log_prob)


def augmented_expected_improvement(mean, var, eta, noise):
    """function.py:318-325: EI times the augmentation ``1 - sqrt(noise) / sqrt(noise + variance)``."""
    return expected_improvement(mean, var, eta) * (1.0 - math.sqrt(noise) / np.sqrt(noise + var))


def aei_gradient(m: GPRModel, Xq: np.ndarray, eta: float):
    """Value and d AEI / d x* by the product rule on :func:`ei_gradient` and the augmentation factor."""
    mean, var = predict(m, Xq)
    _, dvar = posterior_gradients(m, Xq)
    ei, gei = ei_gradient(m, Xq, eta)
    aug = 1.0 - math.sqrt(m.noise) / np.sqrt(m.noise + var)
    daug = np.where(var <= VAR_CLIP, 0.0, 0.5 * math.sqrt(m.noise) * (m.noise + var) ** -1.5)
    return ei * aug, gei * aug + ei * daug * dvar


MES_CLAMP_LB = 1e-8  # entropy.py:47


def min_value_entropy_search(mean, var, samples):
    """entropy.py:193-213.  mean, var [M,1]; samples [S,1] -> [M,1]:
    gamma = (y* - mean) / clip(sd, 1e-8); mean_S( -gamma * exp(logpdf(gamma) - logcdf(-gamma)) / 2 - logcdf(-gamma) )."""
    from scipy.special import log_ndtr

    sd = np.maximum(np.sqrt(var), MES_CLAMP_LB)
    gamma = (np.asarray(samples).reshape(1, -1) - mean) / sd  # [M, S]
    log_minus_cdf = log_ndtr(-gamma)
    log_prob = -0.5 * gamma * gamma - 0.5 * math.log(2.0 * math.pi)
    ratio = np.exp(log_prob - log_minus_cdf)
    return (-gamma * ratio / 2.0 - log_minus_cdf).mean(axis=1, keepdims=True)


def gumbel_fit(fmean, fsd):
    """acquisition/sampler.py:186-204: Gumbel (a, b) matching the quartiles of Pr(y* < y) = 1 - prod Phi(-(y - mu)/sd),
    found by bisection on [min(mu - 5 sd), max(mu + 5 sd)]."""
    from scipy.optimize import bisect
    from scipy.special import log_ndtr

    fmean = np.asarray(fmean, dtype=np.float64).reshape(-1)
    fsd = np.asarray(fsd, dtype=np.float64).reshape(-1)

    def probf(y):
        return 1.0 - math.exp(float(np.sum(log_ndtr(-(y - fmean) / fsd))))

    left, right = float(np.min(fmean - 5 * fsd)), float(np.max(fmean + 5 * fsd))
    q1 = bisect(lambda y: probf(y) - 0.25, left, right, maxiter=10000)
    q2 = bisect(lambda y: probf(y) - 0.75, left, right, maxiter=10000)
    l1, l2 = math.log(math.log(4.0 / 3.0)), math.log(math.log(4.0))
    return (q2 * l1 - q1 * l2) / (l1 - l2), (q1 - q2) / (l1 - l2)


def gumbel_samples(a, b, uniform):
    """acquisition/sampler.py:206-211: inverse probability integral transform of uniform draws -> [S, 1]."""
    u = np.asarray(uniform, dtype=np.float64)
    return (np.log(-np.log(1.0 - u)) * b + a)[:, None]


def expected_improvement_at(m: GPRModel, Xq: np.ndarray, eta: float, chunk: int = 16384):
    mean, var = predict_batched(m, Xq, chunk)
    return expected_improvement(mean, var, eta)


def ei_eta(m: GPRModel) -> float:
    """``ExpectedImprovement.prepare_acquisition_function`` (function.py:145-149): eta = min over
    the training inputs of the posterior mean."""
    mean, _ = predict(m, m.X)
    return float(mean.min())


def lower_confidence_bound(mean, var, beta):
    """function.py:415-416: mean - beta * sqrt(var).  ``NegativeLowerConfidenceBound`` negates it
    (function.py:358-359)."""
    if beta < 0:
        raise ValueError("Standard deviation scaling parameter beta must not be negative")
    return mean - beta * np.sqrt(var)


def log_expected_improvement(mean, var, eta):
    """log of ``expected_improvement`` — ABSENT in the reference at this commit;
    defined by us as log(EI) with an erfcx-based branch for z < -1 so that it stays finite where EI
    underflows.  PARITY UNPINNED; tests compare with log(oracle EI) where EI > 1e-300."""
    sigma = np.sqrt(var)
    z = (eta - mean) / sigma
    out = np.empty_like(z)
    hi = z > -1.0
    zh = z[hi]
    out[hi] = np.log(zh * ndtr(zh) + np.exp(-0.5 * zh * zh) / math.sqrt(2.0 * math.pi))
    zl = z[~hi]
    # h(z) = phi(z) * (1 + z * Phi(z)/phi(z)),  Phi(z)/phi(z) = sqrt(pi/2) erfcx(-z/sqrt2)
    t = 1.0 + zl * math.sqrt(0.5 * math.pi) * ssp.erfcx(-zl / math.sqrt(2.0))
    far = zl < -1e3
    if np.any(far):  # asymptotic 1 + z R(z) ~ 1/z^2 - 3/z^4 (avoids cancellation)
        zf = zl[far]
        t[far] = (1.0 - 3.0 / (zf * zf)) / (zf * zf)
    out[~hi] = -0.5 * zl * zl - 0.5 * math.log(2.0 * math.pi) + np.log(t)
    return out + np.log(sigma)


def log_ei_gradient(m: GPRModel, Xq: np.ndarray, eta: float):
    """Value and gradient of :func:`log_expected_improvement` (ours; parity unpinned like the value).  With EI = sigma h(z),
    h = z Phi + phi:  d log EI / d mean = -(Phi / h) / sigma,  d log EI / d var = (phi / h) / (2 var); the two ratios are
    formed in log space so they stay finite where EI underflows."""
    mean, var = predict(m, Xq)
    dmean, dvar = posterior_gradients(m, Xq)
    sigma = np.sqrt(var)
    z = (eta - mean) / sigma
    val = log_expected_improvement(mean, var, eta)
    log_h = val - np.log(sigma)
    Phi_over_h = np.exp(ssp.log_ndtr(z) - log_h)
    phi_over_h = np.exp(-0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - log_h)
    clipped = var <= VAR_CLIP
    g = -(Phi_over_h / sigma) * dmean + np.where(clipped, 0.0, phi_over_h / (2.0 * var)) * dvar
    return val, g


def ei_gradient(m: GPRModel, Xq: np.ndarray, eta: float):
    """Value and d EI / d x*: dEI/dmean = -Phi(z), dEI/dvar = phi(z) / (2 sigma)."""
    mean, var = predict(m, Xq)
    dmean, dvar = posterior_gradients(m, Xq)
    sigma = np.sqrt(var)
    z = (eta - mean) / sigma
    pdf = np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
    clipped = var <= VAR_CLIP  # clip_by_value has zero gradient where it clips
    g = -ndtr(z) * dmean + np.where(clipped, 0.0, pdf / (2.0 * sigma)) * dvar
    return expected_improvement(mean, var, eta), g


# --------------------------------------------------------------------------------------------
# A4. Batch reparametrisation sampler + MC-qEI
# (trieste/models/gpflow/sampler.py:208-287, function.py:1181-1186)
# --------------------------------------------------------------------------------------------
def batch_reparam_sample(mean, cov, eps, jitter=JITTER):
    """mean [..., q, 1], cov [..., 1, q, q], eps [1, q, S] -> samples [..., S, q, 1]."""
    q = cov.shape[-1]
    chol = np.linalg.cholesky(cov + jitter * np.eye(q, dtype=cov.dtype))  # [..., 1, q, q]
    contrib = chol @ eps  # [..., 1, q, S]
    contrib = np.moveaxis(contrib, (-1, -2, -3), (-3, -2, -1))  # [..., S, q, 1]
    return mean[..., None, :, :] + contrib


def batch_monte_carlo_expected_improvement(m: GPRModel, Xq, eps, eta, jitter=JITTER):
    """Xq [..., q, D], eps [1, q, S] -> [..., 1]."""
    mean, cov = predict_joint(m, Xq)
    samples = batch_reparam_sample(mean, cov, eps, jitter)[..., 0]  # [..., S, q]
    min_per_batch = samples.min(-1)  # [..., S]
    improvement = np.maximum(eta - min_per_batch, 0.0)
    return improvement.mean(-1, keepdims=True)


def _kernel_dr2(kind: str, r2: np.ndarray, variance: float) -> np.ndarray:
    """d k / d r^2 of the stationary kernels (r^2 in lengthscale-scaled coordinates)."""
    if kind == "rbf":
        return -0.5 * variance * np.exp(-0.5 * r2)
    r = np.sqrt(np.maximum(r2, 1e-36))
    if kind == "matern12":
        return -variance * np.exp(-r) / (2.0 * r)
    if kind == "matern32":
        return -variance * 1.5 * np.exp(-math.sqrt(3.0) * r)
    s5 = math.sqrt(5.0)
    return -variance * (5.0 / 6.0) * (1.0 + s5 * r) * np.exp(-s5 * r)


def batch_mc_ei_gradient(m: GPRModel, Xb: np.ndarray, eps: np.ndarray, eta: float, jitter: float = JITTER):
    """Value and d/dX of :func:`batch_monte_carlo_expected_improvement` for ONE query batch Xb [q, D] with base samples
    eps [q, S] — what TensorFlow's autodiff returns for function.py:1181-1186 through sampler.py:262-287 (reduce_min
    passes the gradient to the arg-min sample, maximum(., 0) to the active ones, Cholesky by its reverse-mode rule
    Sigma_bar = L^-T sym(Phi(L^T L_bar)) L^-1, Murray 2016).  Returns (value, grad [q, D])."""
    Xb = np.asarray(Xb, dtype=np.float64)
    q, D = Xb.shape
    S = eps.shape[1]
    mean, cov = predict_joint(m, Xb[None])  # [1,q,1], [1,1,q,q]
    mu = mean[0, :, 0]
    Sigma = cov[0, 0] + jitter * np.eye(q)
    C = np.linalg.cholesky(Sigma)
    f = mu[:, None] + C @ eps  # [q, S]
    jstar = np.argmin(f, axis=0)
    imp = eta - f[jstar, np.arange(S)]
    active = imp > 0.0
    value = np.maximum(imp, 0.0).mean()
    G_mu = np.zeros(q)
    G_C = np.zeros((q, q))
    for s in np.nonzero(active)[0]:
        j = jstar[s]
        G_mu[j] -= 1.0 / S
        G_C[j, : j + 1] -= eps[: j + 1, s] / S
    P = np.tril(C.T @ G_C)
    P[np.diag_indices(q)] *= 0.5
    Msym = 0.5 * (P + P.T)
    Sbar = sla.solve_triangular(C.T, sla.solve_triangular(C.T, Msym, lower=False).T, lower=False).T  # C^-T M C^-1
    # d mu_j / d x_j and d Sigma[j,k] / d x_j
    ls = m.lengthscales
    Xt, Xq = m.X / ls, Xb / ls
    diff_n = Xq[:, None, :] - Xt[None, :, :]  # [q, N, D]
    dk_n = _kernel_dr2(m.kind, np.square(diff_n).sum(-1), m.variance)[:, :, None] * 2.0 * diff_n / ls  # [q, N, D]
    alpha = sla.cho_solve((m.L, True), m.err, check_finite=False)[:, 0]
    V = sla.cho_solve((m.L, True), kernel_matrix(m.kind, m.X, Xb, m.variance, ls), check_finite=False)  # [N, q]
    diff_q = Xq[:, None, :] - Xq[None, :, :]  # [q, q, D]
    dk_q = _kernel_dr2(m.kind, np.square(diff_q).sum(-1), m.variance)[:, :, None] * 2.0 * diff_q / ls
    dk_q[np.arange(q), np.arange(q)] = 0.0  # k(x, x) is constant
    grad = np.zeros((q, D))
    for j in range(q):
        w = G_mu[j] * alpha - 2.0 * (V @ Sbar[j])  # [N]
        grad[j] = dk_n[j].T @ w + 2.0 * (Sbar[j][:, None] * dk_q[j]).sum(0)
    return value, grad


# --------------------------------------------------------------------------------------------
# A5. Random Fourier features + theta posterior + trajectory
# (EXT gpflux 0.4.4 RandomFourierFeaturesCosine; trieste/models/gpflow/sampler.py:529-591,741-806,
#  901-936)
# --------------------------------------------------------------------------------------------
def rff_draw(kind: str, F: int, D: int, rng: np.random.Generator):
    """W [F, D]: N(0, I) for RBF; multivariate Student-t with nu = 2p+1 (1/3/5) for Matern
    (normal / sqrt(chi2_nu / nu) per feature).  b ~ U[0, 2 pi)."""
    W = rng.standard_normal((F, D))
    if kind != "rbf":
        nu = {"matern12": 1.0, "matern32": 3.0, "matern52": 5.0}[kind]
        W = W / np.sqrt(rng.chisquare(nu, size=(F, 1)) / nu)
    b = rng.uniform(0.0, 2.0 * math.pi, size=(F,))
    return W, b


def rff_features(X, W, b, variance, lengthscales):
    """phi(x) = sqrt(2 variance / F) cos((x / l) W^T + b): [M, D] -> [M, F]."""
    F = W.shape[0]
    return math.sqrt(2.0 * variance / F) * np.cos((X / lengthscales) @ W.T + b)


def rff_theta_posterior(m: GPRModel, W, b):
    """Returns (theta_mean [F], theta_chol_cov [F, F]); design space when F < n
    (sampler.py:529-557), gram space otherwise (:559-591; switch at :518-527)."""
    n = m.X.shape[0]
    F = W.shape[0]
    phi = rff_features(m.X, W, b, m.variance, m.lengthscales)  # [n, F]
    resid = m.y - m.mean_const
    if F < n:
        Dm = phi.T @ phi + m.noise * np.eye(F)
        Ld = np.linalg.cholesky(Dm)
        D_inv = sla.cho_solve((Ld, True), np.eye(F))
        mean = (D_inv @ (phi.T @ resid))[:, 0]
        chol_cov = np.linalg.cholesky(D_inv * m.noise)
    else:
        G = phi @ phi.T + m.noise * np.eye(n)
        Lg = np.linalg.cholesky(G)
        L_inv_phi = sla.solve_triangular(Lg, phi, lower=True)
        L_inv_y = sla.solve_triangular(Lg, resid, lower=True)
        mean = (L_inv_phi.T @ L_inv_y)[:, 0]
        cov = np.eye(F) - L_inv_phi.T @ L_inv_phi
        chol_cov = np.linalg.cholesky(cov)
    return mean, chol_cov


def rff_trajectory(Xq, W, b, theta, variance, lengthscales, mean_const, chunk: int = 65536):
    """``feature_decomposition_trajectory.__call__`` (sampler.py:901-936): Xq [M, B, D],
    theta [B, F] -> [M, B, 1]."""
    M, B, D = Xq.shape
    out = np.empty((M, B, 1))
    for s in range(0, M, chunk):
        x = Xq[s : s + chunk].reshape(-1, D)
        phi = rff_features(x, W, b, variance, lengthscales).reshape(-1, B, W.shape[0])
        out[s : s + chunk, :, 0] = (phi * theta[None]).sum(-1) + mean_const
    return out


def thompson_from_trajectory(Xc, W, b, thetas, variance, lengthscales, mean_const):
    """``ThompsonSamplerFromTrajectory.sample`` (trieste/acquisition/sampler.py:262-271): for each
    trajectory (row of thetas) argmin over the candidates; returns indices [q]."""
    idx = []
    for theta in thetas:
        f = rff_trajectory(Xc[:, None, :], W, b, theta[None], variance, lengthscales, mean_const)
        idx.append(int(np.argmin(f[:, 0, 0])))
    return np.array(idx)


# --------------------------------------------------------------------------------------------
# Optimiser-side reductions (trieste/acquisition/optimizer.py:124-150, 299-335)
# --------------------------------------------------------------------------------------------
def argmax_first(values: np.ndarray) -> int:
    """tf.math.argmax semantics: first maximal index."""
    return int(np.argmax(values))


def top_k(values: np.ndarray, k: int):
    """tf.math.top_k semantics: descending values, ties by lower index first."""
    k = min(k, values.shape[0])
    order = np.lexsort((np.arange(values.shape[0]), -values))[:k]
    return values[order], order


# --------------------------------------------------------------------------------------------
# Objectives used to synthesise the BASELINE configs (trieste/objectives/single_objectives.py)
# --------------------------------------------------------------------------------------------
def branin(x):  # :83-107
    x0 = x[..., :1] * 15.0 - 5.0
    x1 = x[..., 1:] * 15.0
    b = 5.1 / (4 * math.pi**2)
    c = 5 / math.pi
    t = 1 / (8 * math.pi)
    return (x1 - b * x0**2 + c * x0 - 6) ** 2 + 10 * (1 - t) * np.cos(x0) + 10


def scaled_branin(x):  # :110-124 (same internals, scale 1/51.95, translate -44.81)
    x0 = x[..., :1] * 15.0 - 5.0
    x1 = x[..., 1:] * 15.0
    b = 5.1 / (4 * math.pi**2)
    c = 5 / math.pi
    t = 1 / (8 * math.pi)
    return (1 / 51.95) * ((x1 - b * x0**2 + c * x0 - 6) ** 2 + 10 * (1 - t) * np.cos(x0) - 44.81)


def ackley(x):  # ackley_5 (:433-458) generalised to d dims (1/5.0 -> 1/d), SURVEY.md §8d
    d = x.shape[-1]
    x = (x - 0.5) * (32.768 * 2.0)
    e1 = -0.2 * np.sqrt((1.0 / d) * np.square(x).sum(-1))
    e2 = (1.0 / d) * np.cos(2.0 * math.pi * x).sum(-1)
    return (-20.0 * np.exp(e1) - np.exp(e2) + 20.0 + math.e)[..., None]


def hartmann_6(x):  # :476-501
    a = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array(
        [
            [10.0, 3.0, 17.0, 3.5, 1.7, 8.0],
            [0.05, 10.0, 17.0, 0.1, 8.0, 14.0],
            [3.0, 3.5, 1.7, 10.0, 17.0, 8.0],
            [17.0, 8.0, 0.05, 10.0, 0.1, 14.0],
        ]
    )
    P = np.array(
        [
            [0.1312, 0.1696, 0.5569, 0.0124, 0.8283, 0.5886],
            [0.2329, 0.4135, 0.8307, 0.3736, 0.1004, 0.9991],
            [0.2348, 0.1451, 0.3522, 0.2883, 0.3047, 0.6650],
            [0.4047, 0.8828, 0.8732, 0.5743, 0.1091, 0.0381],
        ]
    )
    inner = -(A * (x[..., None, :] - P) ** 2).sum(-1)
    return -(a * np.exp(inner)).sum(-1, keepdims=True)


def random_fourier_objective(x, terms: int = 64, seed: int = 2):
    """Fixed synthetic objective for C5: 64-term random Fourier function."""
    rng = np.random.default_rng(seed)
    d = x.shape[-1]
    w = rng.standard_normal((terms, d)) * 3.0
    ph = rng.uniform(0, 2 * math.pi, terms)
    a = rng.standard_normal(terms) / math.sqrt(terms)
    return (np.cos(x @ w.T + ph) * a).sum(-1, keepdims=True)


def synthetic_model(objective, N, D, kind="matern52", dtype=np.float64, seed=0, noise=None) -> GPRModel:
    """The synthetic configs of SURVEY.md §8d with ``build_gpr`` defaults
    (trieste/models/gpflow/builders.py:85-155,413-443): X ~ U[0,1]^D, lengthscale 0.2*sqrt(D),
    kernel variance Var(y), mean const mean(y), noise Var(y)/100."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, D))
    y = objective(X)
    var = float(np.var(y))
    if var <= 0:
        var = 1.0
    ls = np.full(D, 0.2 * math.sqrt(D))
    nz = var / 100.0 if noise is None else noise
    return build_model(kind, X.astype(dtype), y.astype(dtype), var, ls, nz, float(np.mean(y)))


# --------------------------------------------------------------------------------------------
# Decoupled (pathwise) trajectory sampler — trieste/models/gpflow/sampler.py:594-738, 809-855
# (exact-GP branch :668-677): f(x) = phi(x) w + sum_j v_j k(x, x_j) + m(x),
#   u = (y - m) + sqrt(noise) eps,   v = (K + noise I)^-1 (u - phi(X) w)
# --------------------------------------------------------------------------------------------
def decoupled_weights(m: GPRModel, W, b, prior_w, eps):
    """prior_w [B, F] ~ N(0, I), eps [B, N] ~ N(0, I)  ->  canonical weights v [B, N]."""
    phi_Z = rff_features(m.X, W, b, m.variance, m.lengthscales)  # [N, F]
    u = (m.y - m.mean_const)[:, 0][None, :] + math.sqrt(m.noise) * eps  # [B, N]
    diff = u - prior_w @ phi_Z.T  # [B, N]
    return sla.cho_solve((m.L, True), diff.T, check_finite=False).T


def decoupled_trajectory(m: GPRModel, Xq, W, b, prior_w, v, chunk: int = 32768):
    """Xq [M, B, D] -> [M, B, 1]."""
    M, B, D = Xq.shape
    out = np.empty((M, B, 1))
    for s in range(0, M, chunk):
        x = Xq[s : s + chunk]
        for bb in range(B):
            phi = rff_features(x[:, bb], W, b, m.variance, m.lengthscales)  # [m, F]
            kx = kernel_matrix(m.kind, x[:, bb], m.X, m.variance, m.lengthscales)  # [m, N]
            out[s : s + chunk, bb, 0] = phi @ prior_w[bb] + kx @ v[bb] + m.mean_const
    return out


def probability_below_threshold(mean, var, threshold):
    """trieste/acquisition/function/function.py:507-509: Normal(mean, sqrt(var)).cdf(threshold)."""
    return ndtr((threshold - mean) / np.sqrt(var))


# --------------------------------------------------------------------------------------------
# multiple_optimism_lower_confidence_bound — trieste/acquisition/function/function.py:1857-1911
# --------------------------------------------------------------------------------------------
def molcb_betas(batch_size: int, search_space_dim: int) -> np.ndarray:
    """:1898-1905: spread = 0.5 + 0.5 * (1..B) / (B + 1); betas = 5 * d * Normal(0,1).quantile(spread)."""
    spread = 0.5 + 0.5 * np.arange(1, batch_size + 1, dtype=np.float64) / (batch_size + 1.0)
    return 5.0 * search_space_dim * ssp.ndtri(spread)


def multiple_optimism_lower_confidence_bound(m: GPRModel, Xb: np.ndarray, search_space_dim: int) -> np.ndarray:
    """:1907-1911: x [..., B, D] -> -mean + sqrt(var) * betas, shape [..., B]."""
    Xb = np.asarray(Xb, dtype=np.float64)
    B, D = Xb.shape[-2], Xb.shape[-1]
    mean, var = predict(m, Xb.reshape(-1, D))
    mean, var = mean.reshape(Xb.shape[:-1]), var.reshape(Xb.shape[:-1])
    return -mean + np.sqrt(var) * molcb_betas(B, search_space_dim)


# --------------------------------------------------------------------------------------------
# conditional predictions (FastUpdateModel) — trieste/models/gpflow/models.py:355-425 (Chevalier et al. 2014,
# eqs. 8-10), the posterior the reference's Fantasizer evaluates through _fantasized_model
# (acquisition/function/greedy_batch.py:630-770)
# --------------------------------------------------------------------------------------------
def conditional_predict_f(m: GPRModel, Xq: np.ndarray, X_add: np.ndarray, y_add: np.ndarray):
    """models.py:383-425 written out: mean_add / cov_add at the additional points (:383-385), cross covariance
    (:386-390), L = chol(cov_add + noise I) (:392-397), A = L^-1 cov_cross, mean_new = mean_qp + A^T L^-1 (y_add - mean_add),
    var_new = var_qp - sum A^2 (:399-420).  Xq [M, D], X_add [N2, D], y_add [N2, 1]."""
    mean_add, cov_add = predict_f(m, X_add, full_cov=True)
    cov_cross = covariance_between_points(m, X_add, Xq)[0]  # [1, N2, M] -> [N2, M]
    L = sla.cholesky(cov_add + m.noise * np.eye(X_add.shape[0]), lower=True)
    A = sla.solve_triangular(L, cov_cross, lower=True)
    AM = sla.solve_triangular(L, np.asarray(y_add, dtype=np.float64) - mean_add, lower=True)
    mean_qp, var_qp = predict_f(m, Xq)
    return mean_qp + A.T @ AM, var_qp - np.sum(A * A, axis=0)[:, None]


# --------------------------------------------------------------------------------------------
# the reference's continuous optimiser engine — trieste/acquisition/optimizer.py:566-745: one
# scipy.optimize.minimize(method="l-bfgs-b", jac=True, bounds=...) per start on the NEGATED function (:721-738),
# default options (:719: none beyond SciPy's own), best run = argmax over the runs' values (:556-559)
# --------------------------------------------------------------------------------------------
def scipy_lbfgsb_multistart(value_and_gradient, starts: np.ndarray, lower, upper, options=None):
    """``value_and_gradient(x [n, D]) -> (f [n], g [n, D])`` of the function to MAXIMISE.  Returns
    (success [P] bool, fun [P] maximised values, x [P, D], nfev [P])."""
    from scipy import optimize as spo

    starts = np.asarray(starts, dtype=np.float64)
    P, D = starts.shape
    bounds = spo.Bounds(np.broadcast_to(np.asarray(lower, dtype=np.float64), (D,)), np.broadcast_to(np.asarray(upper, dtype=np.float64), (D,)))
    ok, fun, xs, nfev = np.zeros(P, bool), np.zeros(P), np.zeros((P, D)), np.zeros(P, np.int64)

    def neg(x):
        f, g = value_and_gradient(x[None, :])
        return -float(np.asarray(f).reshape(-1)[0]), -np.asarray(g, dtype=np.float64).reshape(-1)

    for p in range(P):
        res = spo.minimize(neg, starts[p], jac=True, bounds=bounds, method="l-bfgs-b", options=dict(options or {}))
        ok[p], fun[p], xs[p], nfev[p] = res.success, -res.fun, res.x, res.nfev
    return ok, fun, xs, nfev
