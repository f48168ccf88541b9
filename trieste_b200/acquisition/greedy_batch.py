"""``Fantasizer`` — greedy batches with any single-point acquisition function, mirrors
trieste/acquisition/function/greedy_batch.py:415-607 (builder) and :630-770 (``_fantasized_model``).

Every time a point of the batch has been chosen, its observation is "fantasised" (kriging believer: the model mean;
"sample": a posterior sample) and the model is conditioned on it.  For an exact GPR that conditional posterior IS the
posterior of the same GPR with the pending points appended to its data (Chevalier et al. 2014, eqs. 8-10; the reference
evaluates it through ``conditional_predict_*``, models/gpflow/models.py:355-525).  The B200-native form therefore keeps a
second device-resident model whose posterior cache is the base model's cache EXTENDED by the pending rows
(``tb_gp_append_data``, O(m N^2) per greedy step instead of conditioning every prediction on the host), so the fantasised
model runs the same fused predict + acquisition kernels — including the device-side multi-start optimiser — as the base
model."""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np

from ..data import Dataset
from ..models import GaussianProcessRegression, GPRSpec
from .function import ExpectedImprovement
from .interface import (
    OBJECTIVE,
    AcquisitionFunctionBuilder,
    GreedyAcquisitionFunctionBuilder,
    SingleModelAcquisitionBuilder,
    Tag,
)


def _generate_fantasized_data(fantasize_method: str, model, pending_points) -> Dataset:
    """greedy_batch.py:572-594."""
    pending_points = np.asarray(pending_points)
    if fantasize_method == "KB":
        fantasized_obs, _ = model.predict(pending_points)
    elif fantasize_method == "sample":
        fantasized_obs = model.sample(pending_points, num_samples=1)[0]
    else:
        raise NotImplementedError(f"fantasize_method must be KB or sample, received {fantasize_method!r}")
    return Dataset(pending_points, np.asarray(fantasized_obs))


class _fantasized_model(GaussianProcessRegression):
    """greedy_batch.py:630-770: the base model conditioned on additional (fantasised) data.  A native model of its own:
    data = base data + fantasised rows, same kernel / mean function / noise."""

    def __init__(self, model: GaussianProcessRegression, fantasized_data: Dataset):
        if not isinstance(model, GaussianProcessRegression):
            raise NotImplementedError(
                "Fantasizer only works with FastUpdateModel models that also support predict_joint, get_kernel and "
                f"get_observation_noise; received {model!r}"
            )
        self._base = model
        self._fantasized = self._check(fantasized_data)
        super().__init__(self._spec_from_base(), device=model.device, num_rff_features=model._num_rff_features,
                         use_decoupled_sampler=model._use_decoupled_sampler)
        if hasattr(model, "_engine"):
            self.set_engine(model._engine)

    def _check(self, data: Dataset) -> Dataset:
        X, Y = np.asarray(data.query_points), np.asarray(data.observations)
        if X.ndim != 2 or Y.ndim != 2 or Y.shape != (X.shape[0], 1):
            raise ValueError(
                f"fantasized data must have query_points [M, D] and observations [M, 1], got {X.shape} and {Y.shape}")
        return Dataset(X, Y)

    def _joined(self):
        base = self._base.get_internal_data()
        dt = self._base.dtype
        X = np.concatenate([np.asarray(base.query_points, dtype=dt), np.asarray(self._fantasized.query_points, dtype=dt)], axis=0)
        Y = np.concatenate([np.asarray(base.observations, dtype=dt), np.asarray(self._fantasized.observations, dtype=dt)], axis=0)
        return X, Y

    def _hyper_key(self):
        k = self._base.get_kernel()
        return (k.kind, float(k.variance), tuple(np.asarray(k.lengthscales, dtype=np.float64).reshape(-1)),
                float(self._base.get_observation_noise()), float(self._base.get_mean_function().c))

    def _spec_from_base(self) -> GPRSpec:
        self._key = self._hyper_key()
        return GPRSpec(self._joined(), self._base.get_kernel(), self._base.get_mean_function(), self._base.get_observation_noise())

    def update_fantasized_data(self, fantasized_data: Dataset) -> None:
        """greedy_batch.py:656-661.  The data become base + new fantasised rows: when that extends what this model already
        holds (kriging believer within one BO step: earlier pending points keep their values) the cached factors grow by a
        rank-m append; otherwise (new BO step, "sample") the cache is rebuilt."""
        self._fantasized = self._check(fantasized_data)
        if self._hyper_key() != self._key:  # the base model was re-trained: take its hyper-parameters
            self._key = self._hyper_key()
            self._spec.kernel = self._base.get_kernel()
            self._spec.mean_function = self._base.get_mean_function()
            self._spec.noise_variance = self._base.get_observation_noise()
            self._push_hyper()
        X, Y = self._joined()
        self.update(Dataset(X, Y))
        self.optimize(Dataset(X, Y))  # refreshes the cache only when update() could not append


class Fantasizer(GreedyAcquisitionFunctionBuilder):
    """greedy_batch.py:415-569."""

    def __init__(self, base_acquisition_function_builder=None, fantasize_method: str = "KB"):
        if fantasize_method not in ("KB", "sample"):
            raise ValueError(f"fantasize_method must be 'KB' or 'sample', got {fantasize_method!r}")
        if base_acquisition_function_builder is None:
            base_acquisition_function_builder = ExpectedImprovement()
        if isinstance(base_acquisition_function_builder, SingleModelAcquisitionBuilder):
            base_acquisition_function_builder = base_acquisition_function_builder.using(OBJECTIVE)
        self._builder: AcquisitionFunctionBuilder = base_acquisition_function_builder
        self._fantasize_method = fantasize_method
        self._base_acquisition_function = None
        self._fantasized_acquisition = None
        self._fantasized_models: Mapping[Tag, _fantasized_model] = {}

    def __repr__(self) -> str:
        return f"Fantasizer({self._builder!r}, {self._fantasize_method!r})"

    def _update_base_acquisition_function(self, models, datasets):
        if self._base_acquisition_function is not None:
            self._base_acquisition_function = self._builder.update_acquisition_function(
                self._base_acquisition_function, models, datasets)
        else:
            self._base_acquisition_function = self._builder.prepare_acquisition_function(models, datasets)
        return self._base_acquisition_function

    def _update_fantasized_acquisition_function(self, models, datasets, pending_points):
        pending_points = np.asarray(pending_points)
        if pending_points.ndim != 2:
            raise ValueError(f"pending_points must have rank 2, got shape {pending_points.shape}")
        fantasized_data = {
            tag: _generate_fantasized_data(self._fantasize_method, model, pending_points) for tag, model in models.items()
        }
        if datasets is None:
            datasets = fantasized_data
        else:
            datasets = {tag: data + fantasized_data[tag] for tag, data in datasets.items()}
        if self._fantasized_acquisition is None:
            self._fantasized_models = {tag: _fantasized_model(model, fantasized_data[tag]) for tag, model in models.items()}
            self._fantasized_acquisition = self._builder.prepare_acquisition_function(self._fantasized_models, datasets)
        else:
            for tag, model in self._fantasized_models.items():
                if model._base is not models[tag]:
                    raise ValueError("Fantasizer was prepared with a different model object for tag " + repr(tag))
                model.update_fantasized_data(fantasized_data[tag])
            self._fantasized_acquisition = self._builder.update_acquisition_function(
                self._fantasized_acquisition, self._fantasized_models, datasets)
        return self._fantasized_acquisition

    def prepare_acquisition_function(self, models, datasets=None, pending_points=None):
        for model in models.values():
            if not isinstance(model, GaussianProcessRegression):
                raise NotImplementedError(
                    "Fantasizer only works with FastUpdateModel models that also support predict_joint, get_kernel and "
                    f"get_observation_noise; received {model!r}"
                )
        if pending_points is None:
            return self._update_base_acquisition_function(models, datasets)
        return self._update_fantasized_acquisition_function(models, datasets, pending_points)

    def update_acquisition_function(self, function, models, datasets=None, pending_points=None,
                                    new_optimization_step: bool = True):
        if pending_points is None:
            return self._update_base_acquisition_function(models, datasets)
        return self._update_fantasized_acquisition_function(models, datasets, pending_points)
