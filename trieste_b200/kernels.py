"""Hyper-parameter carriers standing in for the gpflow objects the reference model wraps
(``gpflow.kernels.{SquaredExponential,Matern12,Matern32,Matern52}``, ``gpflow.mean_functions.Constant``).
They hold numbers only; the arithmetic runs in the CUDA kernels."""
from __future__ import annotations

import numpy as np


class Stationary:
    kind = ""

    def __init__(self, variance: float = 1.0, lengthscales=1.0):
        self.variance = float(variance)
        self.lengthscales = np.atleast_1d(np.asarray(lengthscales, dtype=np.float64)).copy()
        if self.variance <= 0 or np.any(self.lengthscales <= 0):
            raise ValueError("kernel variance and lengthscales must be positive")

    def __repr__(self) -> str:
        return f"{type(self).__name__}(variance={self.variance!r}, lengthscales={self.lengthscales!r})"


class SquaredExponential(Stationary):
    kind = "rbf"


RBF = SquaredExponential


class Matern12(Stationary):
    kind = "matern12"


class Matern32(Stationary):
    kind = "matern32"


class Matern52(Stationary):
    kind = "matern52"


class Constant:
    """gpflow.mean_functions.Constant — built by ``build_gpr`` (builders.py:426-429)."""

    def __init__(self, c: float = 0.0):
        self.c = float(c)

    def __call__(self, X):
        X = np.asarray(X)
        return np.full(X.shape[:-1] + (1,), self.c, dtype=np.float64)
