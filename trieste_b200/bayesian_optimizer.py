"""Minimal Bayesian-optimisation driver — the shape of ``BayesianOptimizer.optimize``
(trieste/bayesian_optimizer.py:570-883) for the single-model, single-objective case of the README example
(README.md:33-66): per step ``rule.acquire`` -> observer -> ``model.update`` / ``model.optimize``.  The reference's
history records, checkpointing and TensorBoard logging are orchestration and out of scope."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from .acquisition.interface import OBJECTIVE
from .data import Dataset
from .rule import EfficientGlobalOptimization
from .space import SearchSpace


@dataclass
class OptimizationResult:
    dataset: Dataset
    model: object
    history: List[np.ndarray] = field(default_factory=list)  # query points of every step
    error: Optional[BaseException] = None

    def try_get_final_dataset(self) -> Dataset:
        if self.error is not None:
            raise self.error
        return self.dataset

    def try_get_optimal_point(self):
        """(query point, observation, index) of the best observation (bayesian_optimizer.py:260-280)."""
        ds = self.try_get_final_dataset()
        i = int(np.argmin(ds.observations[:, 0]))
        return ds.query_points[i], ds.observations[i], i


class BayesianOptimizer:
    def __init__(self, observer: Callable[[np.ndarray], np.ndarray], search_space: SearchSpace):
        self._observer = observer
        self._search_space = search_space

    def optimize(self, num_steps: int, dataset: Dataset, model, acquisition_rule=None) -> OptimizationResult:
        if num_steps < 0:
            raise ValueError(f"num_steps must be at least 0, got {num_steps}")
        rule = acquisition_rule if acquisition_rule is not None else EfficientGlobalOptimization()
        history: List[np.ndarray] = []
        try:  # the reference records the exception and returns the history so far (bayesian_optimizer.py:855-875)
            for _ in range(num_steps):
                query_points = rule.acquire(self._search_space, {OBJECTIVE: model}, {OBJECTIVE: dataset})
                observations = np.asarray(self._observer(query_points), dtype=np.float64).reshape(len(query_points), -1)
                dataset = dataset + Dataset(np.asarray(query_points, dtype=np.float64), observations)
                model.update(dataset)
                model.optimize(dataset)
                history.append(np.asarray(query_points))
        except Exception as e:  # noqa: BLE001
            return OptimizationResult(dataset, model, history, e)
        return OptimizationResult(dataset, model, history)
