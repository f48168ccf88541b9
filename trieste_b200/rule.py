"""The two acquisition rules that call the hot path — mirrors trieste/acquisition/rule.py
(``EfficientGlobalOptimization`` :209-399, ``DiscreteThompsonSampling`` :879-994).  Rules are
per-BO-step orchestration; everything else in rule.py is out of scope."""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np

from .acquisition.function import ExpectedImprovement
from .acquisition.interface import (
    OBJECTIVE,
    AcquisitionFunctionBuilder,
    GreedyAcquisitionFunctionBuilder,
    SingleModelAcquisitionBuilder,
    SingleModelGreedyAcquisitionBuilder,
    VectorizedAcquisitionFunctionBuilder,
)
from .acquisition.optimizer import automatic_optimizer_selector, batchify_joint, batchify_vectorize
from .acquisition.sampler import ExactThompsonSampler, ThompsonSamplerFromTrajectory  # noqa: F401
from .data import Dataset
from .space import SearchSpace


class EfficientGlobalOptimization:
    """rule.py:209-399: build (or update in place) the acquisition function, then maximise it."""

    def __init__(self, builder=None, optimizer=None, num_query_points: int = 1):
        if num_query_points <= 0:
            raise ValueError(f"Number of query points must be greater than 0, got {num_query_points}")
        if builder is None:
            if num_query_points != 1:
                raise ValueError("a batch acquisition builder must be given for num_query_points > 1")
            builder = ExpectedImprovement()
        if optimizer is None:
            optimizer = automatic_optimizer_selector
        if isinstance(builder, (SingleModelAcquisitionBuilder, SingleModelGreedyAcquisitionBuilder)):
            builder = builder.using(OBJECTIVE)
        if num_query_points > 1:  # rule.py:291-301
            if isinstance(builder, VectorizedAcquisitionFunctionBuilder):
                optimizer = batchify_vectorize(optimizer, num_query_points)  # batch elements optimised independently
            elif isinstance(builder, AcquisitionFunctionBuilder):
                optimizer = batchify_joint(optimizer, num_query_points)  # ... jointly over space ** q
            # a GreedyAcquisitionFunctionBuilder collects the batch sequentially in acquire()
        self._builder = builder
        self._optimizer = optimizer
        self._num_query_points = num_query_points
        self._acquisition_function = None

    def __repr__(self) -> str:
        return f"EfficientGlobalOptimization({self._builder!r}, {self._optimizer!r}, {self._num_query_points!r})"

    @property
    def acquisition_function(self):
        return self._acquisition_function

    def acquire(self, search_space: SearchSpace, models: Mapping[str, object],
                datasets: Optional[Mapping[str, Dataset]] = None) -> np.ndarray:
        if self._acquisition_function is None:
            self._acquisition_function = self._builder.prepare_acquisition_function(models, datasets=datasets)
        else:
            self._acquisition_function = self._builder.update_acquisition_function(
                self._acquisition_function, models, datasets=datasets
            )
        points = self._optimizer(search_space, self._acquisition_function)
        if isinstance(self._builder, GreedyAcquisitionFunctionBuilder):
            for _ in range(self._num_query_points - 1):  # rule.py:371-385: greedily allocate the remaining batch elements
                self._acquisition_function = self._builder.update_acquisition_function(
                    self._acquisition_function, models, datasets=datasets, pending_points=points, new_optimization_step=False
                )
                chosen_point = self._optimizer(search_space, self._acquisition_function)
                points = np.concatenate([points, chosen_point], axis=0)
        return points

    def acquire_single(self, search_space, model, dataset=None):
        return self.acquire(search_space, {OBJECTIVE: model}, None if dataset is None else {OBJECTIVE: dataset})


class DiscreteThompsonSampling:
    """rule.py:879-994: sample ``num_search_space_samples`` candidates, pick ``num_query_points`` by
    Thompson sampling from trajectories."""

    def __init__(self, num_search_space_samples: int, num_query_points: int, thompson_sampler=None):
        if not num_search_space_samples > 0:
            raise ValueError(f"Search space must be greater than 0, got {num_search_space_samples}")
        if not num_query_points > 0:
            raise ValueError(f"Number of query points must be greater than 0, got {num_query_points}")
        if thompson_sampler is not None:
            if thompson_sampler.sample_min_value:
                raise ValueError(
                    "Thompson sampling requires a thompson_sampler that samples minimizers, not just minimum values. "
                    "However the passed sampler has sample_min_value=True."
                )
        else:
            # rule.py:942-943: the reference default — exact joint samples, O(M^3) in the candidate count; pass
            # ThompsonSamplerFromTrajectory() for large candidate sets (BASELINE config 4)
            thompson_sampler = ExactThompsonSampler(sample_min_value=False)
        self._thompson_sampler = thompson_sampler
        self._num_search_space_samples = num_search_space_samples
        self._num_query_points = num_query_points

    def __repr__(self) -> str:
        return f"DiscreteThompsonSampling({self._num_search_space_samples!r}, {self._num_query_points!r}, {self._thompson_sampler!r})"

    def acquire(self, search_space: SearchSpace, models, datasets=None) -> np.ndarray:
        if OBJECTIVE not in models:
            raise ValueError(f"dict of models must contain the single key {OBJECTIVE}, got keys {list(models.keys())}")
        query_points = search_space.sample(self._num_search_space_samples)
        return self._thompson_sampler.sample(models[OBJECTIVE], self._num_query_points, query_points)

    def acquire_single(self, search_space, model, dataset=None):
        return self.acquire(search_space, {OBJECTIVE: model}, None if dataset is None else {OBJECTIVE: dataset})
