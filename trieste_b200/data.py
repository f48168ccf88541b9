"""``Dataset`` container — mirrors trieste/data.py:26-62 (query_points [N, D], observations [N, E])."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Dataset:
    query_points: np.ndarray
    observations: np.ndarray

    def __post_init__(self) -> None:
        qp = np.asarray(self.query_points)
        ob = np.asarray(self.observations)
        if qp.ndim < 2 or ob.ndim < 2:
            raise ValueError(
                f"query_points and observations must have rank >= 2, got shapes {qp.shape} and {ob.shape}"
            )
        if qp.shape[:-1] != ob.shape[:-1]:
            raise ValueError(
                f"Leading shapes of query_points and observations must match. Got shapes {qp.shape}, {ob.shape}."
            )
        object.__setattr__(self, "query_points", qp)
        object.__setattr__(self, "observations", ob)

    def __add__(self, rhs: "Dataset") -> "Dataset":
        return Dataset(
            np.concatenate([self.query_points, rhs.query_points], axis=0),
            np.concatenate([self.observations, rhs.observations], axis=0),
        )

    def __len__(self) -> int:
        return int(self.query_points.shape[0])

    def astuple(self):
        return self.query_points, self.observations
