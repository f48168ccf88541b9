"""Minimal search spaces: ``Box`` with i.i.d. uniform sampling (trieste/space.py:843-867) and
``DiscreteSearchSpace``.  Only what the acquisition optimisers of the hot path need."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


class SearchSpace:
    dimension: int

    def sample(self, num_samples: int, seed: Optional[int] = None) -> np.ndarray:  # pragma: no cover
        raise NotImplementedError

    @property
    def has_constraints(self) -> bool:
        return False


class Box(SearchSpace):
    def __init__(self, lower: Sequence[float], upper: Sequence[float]):
        self.lower = np.atleast_1d(np.asarray(lower, dtype=np.float64))
        self.upper = np.atleast_1d(np.asarray(upper, dtype=np.float64))
        if self.lower.shape != self.upper.shape or self.lower.ndim != 1 or self.lower.size == 0:
            raise ValueError(f"lower and upper must be non-empty 1-D of equal shape, got {self.lower.shape}, {self.upper.shape}")
        if np.any(self.lower > self.upper):
            raise ValueError("lower bound must not exceed upper bound")
        self._rng = np.random.default_rng()

    def __repr__(self) -> str:
        return f"Box({self.lower!r}, {self.upper!r})"

    @property
    def dimension(self) -> int:
        return int(self.lower.shape[0])

    def __pow__(self, n: int) -> "Box":
        """Cartesian power, as used by ``batchify_joint`` (acquisition/optimizer.py:924)."""
        return Box(np.tile(self.lower, n), np.tile(self.upper, n))

    def __mul__(self, other: "Box") -> "Box":
        return Box(np.concatenate([self.lower, other.lower]), np.concatenate([self.upper, other.upper]))

    def sample(self, num_samples: int, seed: Optional[int] = None) -> np.ndarray:
        if num_samples < 0:
            raise ValueError(f"num_samples must be non-negative, got {num_samples}")
        rng = self._rng if seed is None else np.random.default_rng(seed)
        u = rng.uniform(size=(num_samples, self.dimension))
        return self.lower + u * (self.upper - self.lower)

    def contains(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x)
        return np.all((x >= self.lower) & (x <= self.upper), axis=-1)


class DiscreteSearchSpace(SearchSpace):
    def __init__(self, points: np.ndarray):
        self.points = np.asarray(points, dtype=np.float64)
        if self.points.ndim != 2:
            raise ValueError(f"points must be rank 2, got {self.points.shape}")
        self._rng = np.random.default_rng()

    @property
    def dimension(self) -> int:
        return int(self.points.shape[1])

    def sample(self, num_samples: int, seed: Optional[int] = None) -> np.ndarray:
        rng = self._rng if seed is None else np.random.default_rng(seed)
        if num_samples >= len(self.points):
            return self.points
        return self.points[rng.choice(len(self.points), size=num_samples, replace=False)]
