"""The reference-side binding: a trieste model class backed by the B200-native engine.

This is the file a trieste maintainer would add (as `trieste/models/b200.py`): it converts `tf.Tensor` <-> NumPy at the protocol
edge and delegates everything else to `trieste_b200.GaussianProcessRegression`.  It needs `tensorflow` and `trieste` at import
time; neither can be installed in the build container, so `tests/test_gpu_adapter.py` executes it against minimal stand-ins
(a `tensorflow` stub with `constant` / `.numpy()`, and `trieste.models.interfaces` Protocol classes generated from the
committed `ast` fixture of the real file) — the adapter code itself is exactly what would run against the real packages.
"""
import numpy as np
import tensorflow as tf
from trieste.models.interfaces import (HasReparamSampler, HasTrajectorySampler, SupportsGetInternalData, SupportsGetKernel,
                                       SupportsGetObservationNoise, SupportsPredictJoint, SupportsPredictY,
                                       TrainableProbabilisticModel)

import trieste_b200 as tb

_KERNELS = {"SquaredExponential": tb.RBF, "Matern12": tb.Matern12, "Matern32": tb.Matern32, "Matern52": tb.Matern52}


def _np(x):
    return x.numpy() if hasattr(x, "numpy") else np.asarray(x)


class B200GaussianProcessRegression(TrainableProbabilisticModel, SupportsPredictJoint, SupportsPredictY, SupportsGetKernel,
                                    SupportsGetObservationNoise, SupportsGetInternalData, HasReparamSampler, HasTrajectorySampler):
    """Takes data and (trained) hyper-parameters from a `gpflow.models.GPR`; hyper-parameter training stays with gpflow."""

    def __init__(self, gpflow_gpr, device: int = 0):
        X, Y = (_np(t) for t in gpflow_gpr.data)
        k = gpflow_gpr.kernel
        kern = _KERNELS[type(k).__name__](float(_np(k.variance)), np.atleast_1d(_np(k.lengthscales)).astype(np.float64))
        c = getattr(gpflow_gpr.mean_function, "c", None)
        mean = tb.Constant(float(_np(c)) if c is not None else 0.0)
        self._gpflow = gpflow_gpr
        self._m = tb.GaussianProcessRegression(tb.GPRSpec((X, Y), kern, mean, float(_np(gpflow_gpr.likelihood.variance))), device=device)

    @property
    def native(self):
        """the trieste_b200 model: hand it to trieste_b200.acquisition builders for the fused predict + acquisition kernels"""
        return self._m

    def predict(self, query_points):
        mean, var = self._m.predict(_np(query_points))
        return tf.constant(mean), tf.constant(var)

    def predict_joint(self, query_points):
        mean, cov = self._m.predict_joint(_np(query_points))
        return tf.constant(mean), tf.constant(cov)

    def predict_y(self, query_points):
        mean, var = self._m.predict_y(_np(query_points))
        return tf.constant(mean), tf.constant(var)

    def sample(self, query_points, num_samples):
        return tf.constant(self._m.sample(_np(query_points), num_samples))

    def update(self, dataset):
        self._m.update(tb.Dataset(_np(dataset.query_points), _np(dataset.observations)))

    def optimize(self, dataset):
        self._m.optimize(None)

    def log(self, dataset=None):
        pass

    def get_kernel(self):
        return self._gpflow.kernel

    def get_observation_noise(self):
        return self._gpflow.likelihood.variance

    def get_internal_data(self):
        return self._m.get_internal_data()

    def reparam_sampler(self, num_samples):
        return self._m.reparam_sampler(num_samples)

    def trajectory_sampler(self):
        return self._m.trajectory_sampler()
