"""Model-side samplers — mirrors trieste/models/gpflow/sampler.py
(BatchReparametrizationSampler :167-287, RandomFourierFeatureTrajectorySampler :452-591,
ResampleableRandomFourierFeatureFunctions :741-806, feature_decomposition_trajectory :858-953)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _lib
from .models import GaussianProcessRegression, _flatten_leading, _ptr

JITTER = 1e-6


def _reparam_sample(model: GaussianProcessRegression, at, eps: np.ndarray, jitter: float):
    """at [..., q, D], eps [q, S] -> samples [..., S, q, 1]."""
    x, _ = _lib.as_f64_contiguous(at)
    flat, lead = _flatten_leading(x, 2)
    nb, q = flat.shape[0], flat.shape[1]
    S = eps.shape[1]
    eps = np.ascontiguousarray(eps, dtype=np.float64)
    out, po = _lib.empty_like_kind(flat, (nb, S, q))
    _lib.check(_lib.lib().tb_gp_reparam_sample(model.handle, _ptr(flat), nb, q, eps.ctypes.data, S, jitter, po))
    return out.reshape(lead + (S, q, 1))


class BatchReparametrizationSampler:
    """sampler.py:167-287.  The base samples ``eps`` [L=1, q, S] are drawn once (NumPy generator —
    the reference uses tf.random.normal; RNG streams are never bit-compatible, so ``eps`` can also be
    injected with :meth:`set_eps`) and stay fixed until :meth:`reset_sampler`."""

    def __init__(self, sample_size: int, model: GaussianProcessRegression, seed: Optional[int] = None):
        if sample_size <= 0:
            raise ValueError(f"sample_size must be positive, got {sample_size}")
        if not hasattr(model, "predict_joint"):
            raise ValueError(f"BatchReparametrizationSampler only works with models that support predict_joint; received {model!r}")
        self._sample_size = sample_size
        self._model = model
        self._rng = np.random.default_rng(seed)
        self._eps: Optional[np.ndarray] = None  # [q, S]
        self._initialized = False

    def set_eps(self, eps: np.ndarray) -> None:
        eps = np.ascontiguousarray(np.asarray(eps, dtype=np.float64))
        if eps.ndim == 3:
            eps = eps[0]
        if eps.ndim != 2 or eps.shape[1] != self._sample_size:
            raise ValueError(f"eps must be [q, {self._sample_size}], got {eps.shape}")
        self._eps = eps
        self._initialized = True

    def _get_eps(self, batch_size: int) -> np.ndarray:
        if batch_size <= 0:
            raise ValueError("batch size must be positive")
        if not self._initialized or self._eps is None:
            self._eps = self._rng.standard_normal((batch_size, self._sample_size))
            self._initialized = True
        if self._eps.shape[0] != batch_size:
            raise ValueError(
                f"{type(self).__name__} requires a fixed batch size. Got batch size {batch_size} but previous "
                f"batch size was {self._eps.shape[0]}."
            )
        return self._eps

    def sample(self, at, *, jitter: float = JITTER):
        """at [..., B, D] -> [..., S, B, 1]."""
        if np.ndim(at) < 2:
            raise ValueError("at must have rank >= 2")
        if jitter < 0:
            raise ValueError(f"jitter must be non-negative, got {jitter}")
        eps = self._get_eps(int(np.shape(at)[-2]))
        return _reparam_sample(self._model, at, eps, jitter)

    def reset_sampler(self) -> None:
        self._initialized = False
