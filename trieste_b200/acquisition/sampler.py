"""Thompson samplers over a discrete candidate set — mirrors trieste/acquisition/sampler.py
(``ThompsonSampler`` :34-82, ``ExactThompsonSampler`` :85-123, ``GumbelSampler`` :126-211,
``ThompsonSamplerFromTrajectory`` :214-273)."""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


class ThompsonSampler:
    """sampler.py:34-82: samples either the minimum values (``sample_min_value=True``) or the minimisers of
    the modelled function across a discrete set of points."""

    def __init__(self, sample_min_value: bool = False):
        self._sample_min_value = bool(sample_min_value)

    @property
    def sample_min_value(self) -> bool:
        return self._sample_min_value

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self._sample_min_value!r})"

    @staticmethod
    def _check(sample_size: int, at) -> np.ndarray:
        if sample_size <= 0:
            raise ValueError(f"sample_size must be positive, got {sample_size}")
        at_np = np.asarray(at, dtype=np.float64)
        if at_np.ndim != 2:
            raise ValueError(f"at must have shape [N, D], got {at_np.shape}")
        return at_np


class ExactThompsonSampler(ThompsonSampler):
    """sampler.py:85-123: joint samples of the model at all ``at`` points (``model.sample``: full covariance + Cholesky,
    O(N^3) in the number of points — the reference's default for DiscreteThompsonSampling and MinValueEntropySearch,
    practical up to a few thousand points; at most 16384 here), reduced per sample to the minimum value ([S, 1]) or
    the minimiser ([S, D])."""

    def sample(self, model, sample_size: int, at, select_output=None, seed: Optional[int] = None) -> np.ndarray:
        at_np = self._check(sample_size, at)
        samples = np.asarray(model.sample(at_np, sample_size, seed=seed), dtype=np.float64)[..., 0]  # [S, N]
        if self._sample_min_value:
            return samples.min(axis=1, keepdims=True)  # [S, 1]
        return at_np[np.argmin(samples, axis=1)]  # [S, D]


class ThompsonSamplerFromTrajectory(ThompsonSampler):
    """For each of ``sample_size`` draws: a fresh trajectory, evaluated on all M candidates, reduced to its minimum
    value ([S, 1]) or its minimiser ([S, D]) (sampler.py:252-273).  Evaluation + argmin are one fused GPU pass per
    trajectory; only the winning (value, index) returns to the host."""

    def sample(self, model, sample_size: int, at, select_output=None) -> np.ndarray:
        at_np = self._check(sample_size, at)
        if not hasattr(model, "trajectory_sampler"):
            raise ValueError(
                f"Thompson sampling from trajectory only supports models with a trajectory_sampler method; received {model!r}"
            )
        # The reference calls ``trajectory_sampler.get_trajectory()`` once per sample (sampler.py:262-263); every such call
        # shares the SAMPLER's feature functions (W, b drawn once in its __init__, models/gpflow/sampler.py:375, 398-400) and
        # draws fresh weights.  ``resample_trajectory`` does exactly that on one trajectory object without re-uploading W, b.
        trajectory_sampler = model.trajectory_sampler()
        trajectory = trajectory_sampler.get_trajectory()
        picked = []
        for i in range(sample_size):
            if i > 0:
                trajectory = trajectory_sampler.resample_trajectory(trajectory)
            vals, idx = trajectory.argmin_over(at_np)
            picked.append(np.array([vals[0]]) if self._sample_min_value else at_np[int(idx[0])])
        return np.stack(picked, axis=0)


def _log_ndtr(x: np.ndarray) -> np.ndarray:
    from scipy.special import log_ndtr

    return log_ndtr(x)


class GumbelSampler(ThompsonSampler):
    """sampler.py:126-211 (Wang & Jegelka 2017): approximate samples of the minimum value y* from a Gumbel
    distribution whose quartiles match the empirical cdf Pr(y* < y) = 1 - prod_i Phi(-(y - mu_i) / sd_i) built from the
    model's marginal predictions (``predict_y`` — Gaussian likelihood, :179-182) at ``at``.  The two quartiles are found
    by bisection; the per-point predictions are one batched GPU ``predict``."""

    def __init__(self, sample_min_value: bool = False, seed: Optional[int] = None):
        if not sample_min_value:
            raise ValueError(
                f"Gumbel samplers can only sample a function's minimal value, however received sample_min_value={sample_min_value}"
            )
        super().__init__(sample_min_value)
        self._rng = np.random.default_rng(seed)

    @staticmethod
    def fit(fmean: np.ndarray, fsd: np.ndarray):
        """Gumbel location / scale (a, b) from the quartiles of the empirical cdf (sampler.py:186-204)."""
        fmean = np.asarray(fmean, dtype=np.float64).reshape(-1)
        fsd = np.asarray(fsd, dtype=np.float64).reshape(-1)

        def probf(y: float) -> float:
            return 1.0 - math.exp(float(np.sum(_log_ndtr(-(y - fmean) / fsd))))

        left = float(np.min(fmean - 5.0 * fsd))
        right = float(np.max(fmean + 5.0 * fsd))

        def quantile(val: float) -> float:
            from scipy.optimize import bisect

            return bisect(lambda y: probf(y) - val, left, right, maxiter=10000)

        q1, q2 = quantile(0.25), quantile(0.75)
        l1 = math.log(math.log(4.0 / 3.0))
        l2 = math.log(math.log(4.0))
        b = (q1 - q2) / (l1 - l2)
        a = (q2 * l1 - q1 * l2) / (l1 - l2)
        return a, b

    def sample(self, model, sample_size: int, at, select_output=None) -> np.ndarray:
        """at [N, D] -> [sample_size, 1] samples of the minimum value."""
        at_np = self._check(sample_size, at)
        fmean, fvar = model.predict_y(at_np) if hasattr(model, "predict_y") else model.predict(at_np)
        a, b = self.fit(np.asarray(fmean, dtype=np.float64), np.sqrt(np.asarray(fvar, dtype=np.float64)))
        u = self._rng.uniform(size=sample_size)
        return (np.log(-np.log1p(-u)) * b + a)[:, None]
