"""Thompson samplers over a discrete candidate set — mirrors trieste/acquisition/sampler.py
(``ThompsonSamplerFromTrajectory`` :215-273)."""
from __future__ import annotations

import numpy as np


class ThompsonSamplerFromTrajectory:
    """For each of ``sample_size`` query points: draw a fresh trajectory, evaluate it on all M
    candidates, take the argmin (sample_min=True) and gather (sampler.py:262-271).  Evaluation +
    argmin are one fused GPU pass per trajectory; only the winning index returns to the host."""

    def __init__(self, sample_min: bool = True):
        self._sample_min = sample_min

    def __repr__(self) -> str:
        return f"ThompsonSamplerFromTrajectory({self._sample_min!r})"

    def sample(self, model, sample_size: int, at, select_output=None) -> np.ndarray:
        """at [M, D] -> [sample_size, D]."""
        if sample_size <= 0:
            raise ValueError(f"sample_size must be positive, got {sample_size}")
        at_np = np.asarray(at, dtype=np.float64)
        if at_np.ndim < 2:
            raise ValueError(f"at must have rank >= 2, got shape {at_np.shape}")
        if not hasattr(model, "trajectory_sampler"):
            raise ValueError(
                f"Thompson sampling from trajectory only supports models with a trajectory_sampler method; received {model!r}"
            )
        if not self._sample_min:
            raise NotImplementedError("sample_min=False (argmax) is not part of the hot path")
        trajectory_sampler = model.trajectory_sampler()
        trajectory = trajectory_sampler.get_trajectory()
        picked = []
        for i in range(sample_size):
            if i > 0:
                trajectory = trajectory_sampler.resample_trajectory(trajectory)
            _, idx = trajectory.argmin_over(at_np)
            picked.append(at_np[int(idx[0])])
        return np.stack(picked, axis=0)
