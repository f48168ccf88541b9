"""Generates tests/golden/gpr_sklearn_*.npz — posterior mean/variance of an exact GPR computed by an
INDEPENDENT third implementation (scikit-learn GaussianProcessRegressor, fixed hyper-parameters,
optimizer=None).  The reference itself (trieste on TensorFlow/GPflow/TFP) cannot be imported in this
container, so these fixtures pin the oracle's GPR algebra, not GPflow's outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import gp_oracle as o  # noqa: E402  (only for the objective functions / config defaults)

CASES = [
    ("branin_n20_matern52", o.branin, 20, 2, "matern52", 1e-7),
    ("hartmann6_n200_matern52", o.hartmann_6, 200, 6, "matern52", None),
    ("hartmann6_n200_rbf", o.hartmann_6, 200, 6, "rbf", None),
    ("ackley10_n300_matern32", o.ackley, 300, 10, "matern32", None),
    # round 2: the non-differentiable Matern12, and the benchmark sizes (training data regenerated from the seed by the test:
    # the fixture stores the generator's name instead of X, y)
    ("hartmann6_n300_matern12", o.hartmann_6, 300, 6, "matern12", None),
    ("hartmann6_n1024_matern52", o.hartmann_6, 1024, 6, "matern52", None),
    ("ackley10_n4096_matern52", o.ackley, 4096, 10, "matern52", None),
]


def main():
    for name, obj, N, D, kind, noise in CASES:
        om = o.synthetic_model(obj, N, D, kind=kind, noise=noise)
        if kind == "rbf":
            base = RBF(length_scale=om.lengthscales, length_scale_bounds="fixed")
        else:
            nu = {"matern12": 0.5, "matern32": 1.5, "matern52": 2.5}[kind]
            base = Matern(length_scale=om.lengthscales, length_scale_bounds="fixed", nu=nu)
        k = ConstantKernel(om.variance, "fixed") * base
        g = GaussianProcessRegressor(k, alpha=om.noise, optimizer=None, normalize_y=False).fit(om.X, om.y - om.mean_const)
        Xq = np.random.default_rng(11).uniform(size=(64, D))
        mu, cov = g.predict(Xq, return_cov=True)
        data = dict(X=om.X, y=om.y) if N <= 300 else dict(generator=obj.__name__, N=N, D=D, seed=0)
        np.savez_compressed(
            os.path.join(HERE, f"gpr_sklearn_{name}.npz"),
            **data, kind=kind, variance=om.variance, lengthscales=om.lengthscales, noise=om.noise,
            mean_const=om.mean_const, Xq=Xq, mean=mu.reshape(-1) + om.mean_const, var=np.diag(cov).copy(), cov=cov,
        )
        print(name, "ok")


if __name__ == "__main__":
    main()
