"""Analytic and Monte-Carlo single-objective acquisition functions of the hot path — mirrors
trieste/acquisition/function/function.py (EI :96-223, LCB :328-418, batch MC-EI :1074-1186).

Each callable keeps the reference's shape contract (``x: [..., 1, D] -> [..., 1]``; batch
functions ``[..., B, D] -> [..., 1]``) but evaluates predict + tail in ONE pass of fused GPU
kernels behind the C-ABI instead of ``model.predict`` followed by separate elementwise ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from .. import _lib
from ..data import Dataset
from ..models import GaussianProcessRegression, _flatten_leading, _ptr
from .interface import AcquisitionFunctionClass, SingleModelAcquisitionBuilder, SingleModelVectorizedAcquisitionBuilder

JITTER = 1e-6  # trieste/utils/misc.py:183


def _check_populated(dataset: Optional[Dataset]) -> Dataset:
    if dataset is None:
        raise ValueError("Dataset must be populated.")
    if len(dataset) == 0:
        raise ValueError("Dataset must be populated.")
    return dataset


def _to_host(x) -> np.ndarray:
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def _require_native(model) -> GaussianProcessRegression:
    if not isinstance(model, GaussianProcessRegression):
        raise ValueError(
            f"trieste_b200 acquisition functions need a trieste_b200.GaussianProcessRegression model; received {model!r}"
        )
    return model


class _FusedSingleQuery(AcquisitionFunctionClass):
    """Common machinery: squeeze the B=1 axis, run the fused kernel chain, restore shapes."""

    _acq: int = -1

    def __init__(self, model: GaussianProcessRegression, param: float):
        self._model = _require_native(model)
        self._param = float(param)

    def _before_call(self) -> None:
        """state that lives in the native handle and must be current before a launch (none by default)"""

    def _squeeze(self, x):
        self._before_call()
        x, _ = _lib.as_contiguous(x, self._model.dtype)
        if x.ndim < 2 or x.shape[-2] != 1:
            raise ValueError(
                f"This acquisition function only supports batch sizes of one; got input of shape {tuple(x.shape)}"
            )
        self._model._check_dim(x)
        return _flatten_leading(x.reshape(tuple(x.shape[:-2]) + (x.shape[-1],)), 1)

    def __call__(self, x):
        flat, lead = self._squeeze(x)
        M = flat.shape[0]
        out, po = _lib.empty_like_kind(flat, (M, 1), self._model.dtype)
        _lib.check(_lib.lib().tb_acq_eval(self._model.handle, self._acq, self._param, _ptr(flat), M, po, None))
        return out.reshape(lead + (1,))

    def value_and_gradient(self, x):
        """``tfp.math.value_and_gradient(fn, x)`` as used at acquisition/optimizer.py:621-629:
        returns (values [..., 1], d values / d x [..., 1, D])."""
        flat, lead = self._squeeze(x)
        M, D = flat.shape
        out, po = _lib.empty_like_kind(flat, (M, 1), self._model.dtype)
        grad, pg = _lib.empty_like_kind(flat, (M, D), self._model.dtype)
        _lib.check(_lib.lib().tb_acq_eval(self._model.handle, self._acq, self._param, _ptr(flat), M, po, pg))
        return out.reshape(lead + (1,)), grad.reshape(lead + (1, D))

    def maximize_from(self, starts, lower, upper, *, maxcor: int = 10, maxiter: int = 15000, maxls: int = 20,
                      gtol: float = 1e-5, ftol: float = 2.220446049250313e-09):
        """Device-side multi-start projected L-BFGS (``tb_acq_maximize``): every row of ``starts`` [P, D] is an
        independent local maximisation inside the box — the work of ``_perform_parallel_continuous_optimization`` and its
        SciPy greenlets (optimizer.py:566-745) without leaving the GPU between iterations.
        Returns (success [P] bool, values [P], x [P, D], nfev [P]); fp64 like the reference's SciPy side."""
        self._before_call()
        x0 = np.ascontiguousarray(_to_host(starts), dtype=np.float64)
        if x0.ndim != 2:
            raise ValueError(f"starts must be [P, D], got {x0.shape}")
        self._model._check_dim(x0)
        P, D = x0.shape
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lower, dtype=np.float64), (D,)))
        up = np.ascontiguousarray(np.broadcast_to(np.asarray(upper, dtype=np.float64), (D,)))
        x = np.empty((P, D))
        f = np.empty(P)
        ok = np.zeros(P, dtype=np.int32)
        nfev = np.zeros(P, dtype=np.int64)
        _lib.check(
            _lib.lib().tb_acq_maximize(
                self._model.handle, self._acq, self._param, lo.ctypes.data, up.ctypes.data, x0.ctypes.data, P,
                int(maxcor), int(maxiter), int(maxls), float(gtol), float(ftol),
                x.ctypes.data, f.ctypes.data, ok.ctypes.data, nfev.ctypes.data,
            )
        )
        return ok.astype(bool), f, x, nfev

    def fused_argmax(self, points):
        """points [M, D] -> (first-max index, value) without writing the M values to HBM
        (generate_random_search_optimizer / _get_max_discrete_points, optimizer.py:124-150)."""
        self._before_call()
        pts, _ = _lib.as_contiguous(points, self._model.dtype)
        if pts.ndim != 2:
            raise ValueError(f"points must be [M, D], got {tuple(pts.shape)}")
        self._model._check_dim(pts)
        best = C.c_double() if self._model.dtype == np.float64 else C.c_float()
        idx = C.c_int64()
        _lib.check(
            _lib.lib().tb_acq_argmax(
                self._model.handle, self._acq, self._param, _ptr(pts), pts.shape[0], None, C.byref(best), C.byref(idx)
            )
        )
        return int(idx.value), float(best.value)


class expected_improvement(_FusedSingleQuery):
    """function.py:190-223: ``(eta - mean) * cdf(eta) + variance * pdf(eta)``."""

    _acq = _lib.ACQ_EI

    def __init__(self, model, eta):
        super().__init__(model, float(np.asarray(eta).reshape(-1)[0]))

    def update(self, eta) -> None:
        self._param = float(np.asarray(eta).reshape(-1)[0])

    @property
    def eta(self) -> float:
        return self._param


class log_expected_improvement(expected_improvement):
    """log of :class:`expected_improvement`.  ABSENT in the reference at this commit;
    defined here (numerically stable for z << 0); parity unpinned."""

    _acq = _lib.ACQ_LOG_EI


class augmented_expected_improvement(expected_improvement):
    """function.py:283-325: EI times ``1 - sqrt(noise) / sqrt(noise + variance)`` (Huang et al. 2006); the noise
    variance is the model's likelihood variance, read from the native handle on every call (the reference re-assigns it
    in ``update``, :306-309)."""

    _acq = _lib.ACQ_AEI


class min_value_entropy_search(_FusedSingleQuery):
    """entropy.py:166-213: information gain about the objective minimum y* from evaluating at x, averaged over samples
    of y* (Wang & Jegelka 2017, adapted for minimisation).  The samples are pushed to the native handle before every
    launch (two functions may share one model)."""

    _acq = _lib.ACQ_MES

    def __init__(self, model, samples):
        super().__init__(model, 0.0)
        self.update(samples)

    def update(self, samples) -> None:
        s = np.asarray(samples, dtype=np.float64)
        if s.ndim != 2:
            raise ValueError(f"samples must have rank two, got shape {s.shape}")
        if s.shape[0] == 0:
            raise ValueError("samples must not be empty")
        self._samples = np.ascontiguousarray(s.reshape(-1))

    @property
    def samples(self) -> np.ndarray:
        return self._samples[:, None]

    def _before_call(self) -> None:
        _lib.check(
            _lib.lib().tb_acq_set_min_value_samples(
                self._model.handle, self._samples.ctypes.data_as(C.POINTER(C.c_double)), int(self._samples.size)
            )
        )


class _lcb(_FusedSingleQuery):
    def __init__(self, model, beta: float, negate: bool):
        if beta < 0:
            raise ValueError("Standard deviation scaling parameter beta must not be negative")
        super().__init__(model, beta)
        self._acq = _lib.ACQ_NEG_LCB if negate else _lib.ACQ_LCB


def lower_confidence_bound(model, beta: float):
    """function.py:389-418: ``mean - beta * sqrt(variance)``."""
    return _lcb(model, beta, negate=False)


def _eta_from_model(model, dataset: Dataset, search_space=None) -> float:
    """function.py:133-149: eta = min over the FEASIBLE training inputs of the posterior mean; with a constrained search
    space only the query points that satisfy the constraints count, and if none does eta = max of the mean."""
    _require_native(model)
    mean, _ = model.predict(np.asarray(dataset.query_points))
    mean = np.asarray(_to_host(mean))
    if search_space is not None and getattr(search_space, "has_constraints", False):
        feasible = np.asarray(search_space.is_feasible(np.asarray(dataset.query_points)), dtype=bool).reshape(-1)
        if not feasible.any():
            return float(np.max(mean, axis=0)[0])
        mean = mean[feasible]
    return float(np.min(mean, axis=0)[0])


class probability_below_threshold(_FusedSingleQuery):
    """function.py:481-513: ``Normal(mean, sqrt(var)).cdf(threshold)``."""

    _acq = _lib.ACQ_PBT

    def __init__(self, model, threshold):
        if np.ndim(threshold) != 0 and np.size(threshold) != 1:
            raise ValueError("threshold must be a scalar")
        super().__init__(model, float(np.asarray(threshold).reshape(-1)[0]))

    def update(self, threshold) -> None:
        self._param = float(np.asarray(threshold).reshape(-1)[0])


class ProbabilityOfImprovement(SingleModelAcquisitionBuilder):
    """function.py:47-93: probability of improving on eta = min posterior mean at the observed points."""

    def __repr__(self) -> str:
        return "ProbabilityOfImprovement()"

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        return probability_below_threshold(model, _eta_from_model(model, dataset))

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        if not isinstance(function, probability_below_threshold):
            raise ValueError(f"expected a probability_below_threshold function, got {function!r}")
        function.update(_eta_from_model(model, dataset))
        return function


class ProbabilityOfFeasibility(SingleModelAcquisitionBuilder):
    """function.py:421-478: probability that the constraint model is below ``threshold``."""

    def __init__(self, threshold: float):
        if np.ndim(threshold) != 0:
            raise ValueError("threshold must be a scalar")
        self._threshold = float(threshold)

    def __repr__(self) -> str:
        return f"ProbabilityOfFeasibility({self._threshold!r})"

    @property
    def threshold(self) -> float:
        return self._threshold

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        return probability_below_threshold(model, self._threshold)

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        return function  # no need to update anything (function.py:470-478)


class ExpectedImprovement(SingleModelAcquisitionBuilder):
    """function.py:96-187."""

    _fn_class = expected_improvement

    def __init__(self, search_space=None):
        self._search_space = search_space

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self._search_space!r})"

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        return self._fn_class(model, _eta_from_model(model, dataset, self._search_space))

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        if not isinstance(function, self._fn_class):
            raise ValueError(f"expected a {self._fn_class.__name__} function, got {function!r}")
        function.update(_eta_from_model(model, dataset, self._search_space))  # same object: no re-build
        return function


class LogExpectedImprovement(ExpectedImprovement):
    _fn_class = log_expected_improvement


class MinValueEntropySearch(SingleModelAcquisitionBuilder):
    """entropy.py:52-164.  The min-value samples come from ``min_value_sampler`` evaluated on the data plus
    ``grid_size`` random points of the search space (:134-137).  Default as in the reference (:111):
    ``ExactThompsonSampler(sample_min_value=True)`` — joint samples over all N + grid_size points, drawn on the device
    (``tb_gp_sample_joint``); above its 16384-point limit the :class:`GumbelSampler` takes over.  Any sampler with
    ``sample_min_value=True`` can be passed (``GumbelSampler``, ``ThompsonSamplerFromTrajectory``)."""

    def __init__(self, search_space, num_samples: int = 5, grid_size: int = 1000, min_value_sampler=None, seed=None):
        if num_samples <= 0:
            raise ValueError(f"num_samples must be positive, got {num_samples}")
        if grid_size <= 0:
            raise ValueError(f"grid_size must be positive, got {grid_size}")
        if min_value_sampler is not None:
            if not min_value_sampler.sample_min_value:
                raise ValueError(
                    "Minvalue Entropy Search requires a min_value_sampler that samples minimum values, "
                    "however the passed sampler has sample_min_value=False."
                )
        self._seed = seed
        self._draws = 0
        self._min_value_sampler = min_value_sampler  # None: chosen per draw (see _draw)
        self._search_space = search_space
        self._num_samples = num_samples
        self._grid_size = grid_size

    def __repr__(self) -> str:
        return (f"MinValueEntropySearch({self._search_space!r}, {self._num_samples!r}, {self._grid_size!r}, "
                f"{self._min_value_sampler!r})")

    MAX_EXACT_POINTS = 16384  # tb_gp_sample_joint's limit on the jointly sampled point set

    def _draw(self, model, dataset: Dataset) -> np.ndarray:
        grid = np.asarray(self._search_space.sample(self._grid_size))
        query_points = np.concatenate([np.asarray(dataset.query_points, dtype=grid.dtype), grid], axis=0)
        sampler = self._min_value_sampler
        # a seeded builder is reproducible: draw k of the builder uses seed + k for the samplers that take a per-call seed
        draw_seed = None if self._seed is None else int(self._seed) + self._draws
        self._draws += 1
        if sampler is None:
            # entropy.py:111: the reference default is ExactThompsonSampler(sample_min_value=True) — joint samples over the
            # data and the grid; beyond the device path's point limit the Gumbel sampler (marginals only) takes over
            from .sampler import ExactThompsonSampler, GumbelSampler

            if query_points.shape[0] <= self.MAX_EXACT_POINTS:
                return ExactThompsonSampler(sample_min_value=True).sample(model, self._num_samples, query_points, seed=draw_seed)
            sampler = GumbelSampler(sample_min_value=True, seed=draw_seed)
        return sampler.sample(model, self._num_samples, query_points)

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        return min_value_entropy_search(model, self._draw(model, dataset))

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        if not isinstance(function, min_value_entropy_search):
            raise ValueError(f"expected a min_value_entropy_search function, got {function!r}")
        function.update(self._draw(model, dataset))
        return function


class AugmentedExpectedImprovement(ExpectedImprovement):
    """function.py:225-280: eta = min posterior mean at the data, as for EI."""

    _fn_class = augmented_expected_improvement

    def __init__(self):
        super().__init__(None)

    def __repr__(self) -> str:
        return "AugmentedExpectedImprovement()"


class NegativeLowerConfidenceBound(SingleModelAcquisitionBuilder):
    """function.py:328-372: negated LCB so that maximisation minimises the bound."""

    def __init__(self, beta: float = 1.96):
        if beta < 0:
            raise ValueError(f"Confidence parameter's standard deviation scaling must not be negative, got {beta}")
        self._beta = beta

    def __repr__(self) -> str:
        return f"NegativeLowerConfidenceBound({self._beta!r})"

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        return _lcb(model, self._beta, negate=True)

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        return function  # no dependence on data (function.py:361-372)


class multiple_optimism_lower_confidence_bound(AcquisitionFunctionClass):
    """function.py:1857-1911 (MOLCB, Torossian et al. 2020): a VECTORISED function ``[..., B, D] -> [..., B]``; column b is
    the negated lower confidence bound ``-mean + beta_b sqrt(var)`` with ``beta_b = 5 d Phi^-1(0.5 + 0.5 b / (B + 1))``, b = 1..B,
    fixed at the first call (a later call with another batch size is an error, as in the reference).  Each column runs the
    fused predict + NegLCB kernels with its own beta (value and gradient), so ``batchify_vectorize`` optimises the B
    columns independently."""

    def __init__(self, model, search_space_dim: int):
        if search_space_dim <= 0:
            raise ValueError(f"search_space_dim must be positive, got {search_space_dim}")
        self._model = _require_native(model)
        self._search_space_dim = int(search_space_dim)
        self._betas: Optional[np.ndarray] = None  # [B], lazily initialised
        self._columns = []

    @property
    def betas(self) -> Optional[np.ndarray]:
        return self._betas

    def _prepare(self, x):
        if len(x.shape) < 2:
            raise ValueError(f"expected [..., B, D] query batches, got shape {tuple(x.shape)}")
        B = int(x.shape[-2])
        if B <= 0:
            raise ValueError("batch size must be positive")
        if self._betas is None:
            from statistics import NormalDist

            spread = 0.5 + 0.5 * np.arange(1, B + 1, dtype=np.float64) / (B + 1.0)
            self._betas = 5.0 * self._search_space_dim * np.array([NormalDist().inv_cdf(p) for p in spread])
            self._columns = [_lcb(self._model, float(b), negate=True) for b in self._betas]
        elif B != self._betas.shape[0]:
            raise ValueError(
                f"{type(self).__name__} requires a fixed batch size. Got batch size {B} but previous batch size was "
                f"{self._betas.shape[0]}."
            )
        return B

    def __call__(self, x):
        x = x if hasattr(x, "shape") else np.asarray(x)
        B = self._prepare(x)
        cols = [_to_host(self._columns[b](x[..., b : b + 1, :])) for b in range(B)]  # each [..., 1]
        return np.concatenate(cols, axis=-1)

    def value_and_gradient(self, x):
        """[..., B, D] -> (values [..., B], gradients [..., B, D]); column b only depends on x[..., b, :]."""
        x = x if hasattr(x, "shape") else np.asarray(x)
        B = self._prepare(x)
        vals, grads = [], []
        for b in range(B):
            v, g = self._columns[b].value_and_gradient(x[..., b : b + 1, :])
            vals.append(_to_host(v))
            grads.append(_to_host(g))
        return np.concatenate(vals, axis=-1), np.concatenate(grads, axis=-2)


class MultipleOptimismNegativeLowerConfidenceBound(SingleModelVectorizedAcquisitionBuilder):
    """function.py:1808-1854: builder of :class:`multiple_optimism_lower_confidence_bound`; nothing to update between
    steps."""

    def __init__(self, search_space):
        self._search_space = search_space

    def __repr__(self) -> str:
        return f"MultipleOptimismNegativeLowerConfidenceBound({self._search_space!r})"

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        return multiple_optimism_lower_confidence_bound(model, self._search_space.dimension)

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        if not isinstance(function, multiple_optimism_lower_confidence_bound):
            raise ValueError(f"expected a multiple_optimism_lower_confidence_bound function, got {function!r}")
        return function  # nothing to update


class batch_monte_carlo_expected_improvement(AcquisitionFunctionClass):
    """function.py:1150-1186: mean over S reparametrised joint samples of
    ``max(eta - min_q sample, 0)``."""

    def __init__(self, sample_size: int, model, eta, jitter: float):
        if not hasattr(model, "reparam_sampler"):
            raise ValueError(
                "The batch Monte-Carlo expected improvement acquisition function only supports models that "
                f"implement a reparam_sampler method; received {model!r}"
            )
        self._sample_size = sample_size
        self._model = _require_native(model)
        self._sampler = model.reparam_sampler(sample_size)
        self._eta = float(np.asarray(eta).reshape(-1)[0])
        self._jitter = jitter

    def update(self, eta) -> None:
        self._eta = float(np.asarray(eta).reshape(-1)[0])
        self._sampler.reset_sampler()

    def __call__(self, x):
        x, _ = _lib.as_contiguous(x, self._model.dtype)
        if x.ndim < 2:
            raise ValueError(f"expected [..., B, D] query batches, got shape {tuple(x.shape)}")
        self._model._check_dim(x)
        flat, lead = _flatten_leading(x, 2)
        nb, q = flat.shape[0], flat.shape[1]
        eps = np.ascontiguousarray(self._sampler._get_eps(q), dtype=self._model.dtype)  # [q, S] host, fixed until reset
        out, po = _lib.empty_like_kind(flat, (nb, 1), self._model.dtype)
        _lib.check(
            _lib.lib().tb_acq_batch_mc_ei(
                self._model.handle, _ptr(flat), nb, q, eps.ctypes.data, eps.shape[1], self._eta, self._jitter, po
            )
        )
        return out.reshape(lead + (1,))

    def value_and_gradient(self, x):
        """[..., B, D] -> (values [..., 1], d values / d x [..., B, D]): the reverse pass TensorFlow's autodiff performs
        when the reference maximises this function over ``space ** B`` (batchify_joint, optimizer.py:897-936)."""
        x, _ = _lib.as_contiguous(x, self._model.dtype)
        if x.ndim < 2:
            raise ValueError(f"expected [..., B, D] query batches, got shape {tuple(x.shape)}")
        self._model._check_dim(x)
        flat, lead = _flatten_leading(x, 2)
        nb, q, D = flat.shape
        eps = np.ascontiguousarray(self._sampler._get_eps(q), dtype=self._model.dtype)
        out, po = _lib.empty_like_kind(flat, (nb, 1), self._model.dtype)
        grad, pg = _lib.empty_like_kind(flat, (nb, q, D), self._model.dtype)
        _lib.check(
            _lib.lib().tb_acq_batch_mc_ei_grad(
                self._model.handle, _ptr(flat), nb, q, eps.ctypes.data, eps.shape[1], self._eta, self._jitter, po, pg
            )
        )
        return out.reshape(lead + (1,)), grad.reshape(lead + (q, D))


class monte_carlo_expected_improvement(batch_monte_carlo_expected_improvement):
    """function.py:883-920: ``mean_S max(eta - f_s(x), 0)`` over reparametrised samples, batch size one.  For a GPR the
    reference's ``model.reparam_sampler`` is the batch sampler (models.py:325-331), so this is the q = 1 case of the
    batch kernel chain."""

    def __call__(self, x):
        x_arr = x if hasattr(x, "shape") else np.asarray(x)
        if len(x_arr.shape) < 2 or x_arr.shape[-2] != 1:
            raise ValueError(
                f"This acquisition function only supports batch sizes of one; got input of shape {tuple(x_arr.shape)}"
            )
        return super().__call__(x)


class MonteCarloExpectedImprovement(SingleModelAcquisitionBuilder):
    """function.py:782-880: eta = min over the data of the sample mean at each training input."""

    def __init__(self, sample_size: int, *, jitter: float = JITTER):
        if sample_size <= 0:
            raise ValueError(f"sample_size must be positive, got {sample_size}")
        if jitter < 0:
            raise ValueError(f"jitter must be non-negative, got {jitter}")
        self._sample_size = sample_size
        self._jitter = jitter

    def __repr__(self) -> str:
        return f"MonteCarloExpectedImprovement({self._sample_size!r}, jitter={self._jitter!r})"

    def _eta(self, sampler, dataset: Dataset):
        x = np.asarray(dataset.query_points)
        samples = sampler.sample(x[..., None, :], jitter=self._jitter)  # [N, S, 1, 1]
        return np.min(np.mean(samples, axis=-3), axis=0)

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        if not hasattr(model, "reparam_sampler"):
            raise ValueError(
                f"MonteCarloExpectedImprovement only supports models with a reparam_sampler method; received {model!r}"
            )
        dataset = _check_populated(dataset)
        fn = monte_carlo_expected_improvement(self._sample_size, model, 0.0, self._jitter)
        fn._eta = float(np.asarray(self._eta(fn._sampler, dataset)).reshape(-1)[0])
        return fn

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        if not isinstance(function, monte_carlo_expected_improvement):
            raise ValueError(f"expected a monte_carlo_expected_improvement, got {function!r}")
        function._sampler.reset_sampler()
        function._eta = float(np.asarray(self._eta(function._sampler, dataset)).reshape(-1)[0])
        return function


class BatchMonteCarloExpectedImprovement(SingleModelAcquisitionBuilder):
    """function.py:1074-1147."""

    def __init__(self, sample_size: int, *, jitter: float = JITTER):
        if sample_size <= 0:
            raise ValueError(f"sample_size must be positive, got {sample_size}")
        if jitter < 0:
            raise ValueError(f"jitter must be non-negative, got {jitter}")
        self._sample_size = sample_size
        self._jitter = jitter

    def __repr__(self) -> str:
        return f"BatchMonteCarloExpectedImprovement({self._sample_size!r}, jitter={self._jitter!r})"

    def prepare_acquisition_function(self, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        mean, _ = model.predict(np.asarray(dataset.query_points))
        if mean.shape[-1] != 1:
            raise ValueError("Expected model with event shape [1].")
        eta = np.min(mean, axis=0)
        return batch_monte_carlo_expected_improvement(self._sample_size, model, eta, self._jitter)

    def update_acquisition_function(self, function, model, dataset: Optional[Dataset] = None):
        dataset = _check_populated(dataset)
        if not isinstance(function, batch_monte_carlo_expected_improvement):
            raise ValueError(f"expected a batch_monte_carlo_expected_improvement, got {function!r}")
        mean, _ = model.predict(np.asarray(dataset.query_points))
        function.update(np.min(mean, axis=0))
        return function
