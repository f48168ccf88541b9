"""Shared test helpers: build matching (oracle model, native model) pairs."""
from __future__ import annotations

import numpy as np

from oracle import gp_oracle as o

KERNEL_CLASSES = {"rbf": "SquaredExponential", "matern12": "Matern12", "matern32": "Matern32", "matern52": "Matern52"}


def native_from_oracle(om, **kw):
    import trieste_b200 as tb

    kcls = getattr(tb, KERNEL_CLASSES[om.kind])
    spec = tb.GPRSpec((om.X, om.y), kcls(om.variance, om.lengthscales), tb.Constant(om.mean_const), om.noise)
    return tb.GaussianProcessRegression(spec, **kw)


def model_pair(objective, N, D, kind="matern52", seed=0, noise=None, engine=None):
    om = o.synthetic_model(objective, N, D, kind=kind, seed=seed, noise=noise)
    nm = native_from_oracle(om)
    if engine is not None:
        nm.set_engine(engine)
    return om, nm


def candidates(M, D, seed=1):
    return np.random.default_rng(seed).uniform(size=(M, D))
