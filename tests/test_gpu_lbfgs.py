"""Device-side multi-start projected L-BFGS (tb_acq_maximize; acquisition/optimizer.py:566-745) against the vectorised
host implementation of the same algorithm, first-order optimality and the reference's optimiser known answers
(tests/unit/acquisition/test_optimizer.py:86-168: maximisers of quadratics inside / on the boundary of a Box)."""
import numpy as np
import pytest

from oracle import gp_oracle as o
from tests.util import candidates, model_pair

pytestmark = pytest.mark.gpu


def _proj_grad(fn, x, lower, upper):
    _, grad = fn.value_and_gradient(x[:, None, :])
    return x - np.clip(x + grad[:, 0, :], lower, upper)


@pytest.mark.parametrize("which", ["neg_lcb", "ei", "log_ei", "mes"])
def test_device_lbfgs_reaches_first_order_points_and_matches_host(which, monkeypatch):
    from trieste_b200.acquisition import (expected_improvement, log_expected_improvement, lower_confidence_bound,
                                          min_value_entropy_search)
    from trieste_b200.acquisition.function import _lcb
    from trieste_b200.acquisition.optimizer import _perform_parallel_continuous_optimization

    om, nm = model_pair(o.hartmann_6, 300, 6)
    eta = o.ei_eta(om)
    fn = {
        "neg_lcb": lambda: _lcb(nm, 1.96, negate=True),
        "ei": lambda: expected_improvement(nm, eta),
        "log_ei": lambda: log_expected_improvement(nm, eta),
        "mes": lambda: min_value_entropy_search(nm, np.array([[eta - 0.1], [eta - 0.4]])),
    }[which]()
    lower, upper = np.zeros(6), np.ones(6)
    x0 = candidates(200, 6, seed=5)
    ok, val, x, nfev = fn.maximize_from(x0, lower, upper)
    assert ok.shape == (200,) and x.shape == (200, 6) and (x >= 0).all() and (x <= 1).all()
    assert ok.mean() > 0.9
    f0 = fn(x0[:, None, :])[:, 0]
    # never worse than the start (the start values come from the value-only kernels, the optimiser's from the
    # value+gradient kernels: tail quantities such as MES agree to ~1e-8 relative between the two digit engines)
    assert np.all(val >= f0 - 1e-6 * np.abs(f0) - 1e-9 * np.abs(f0).max())
    np.testing.assert_allclose(val, fn(x[:, None, :])[:, 0], rtol=1e-6, atol=1e-9 * np.abs(val).max())
    if which != "ei":  # plain EI is flat (~0, gradient ~0) far from the data: gtol is met immediately there
        pg = _proj_grad(fn, x[ok], lower, upper)
        assert np.abs(pg).max() < 1e-3 * max(1.0, np.abs(val).max())
    assert nfev.min() >= 1 and nfev.max() < 5000
    # the host implementation of the same algorithm from the same starts
    monkeypatch.setenv("TB_LBFGS", "host")
    s2, f2, x2, n2 = _perform_parallel_continuous_optimization(fn, lower, upper, x0[:, None, :], {})
    monkeypatch.delenv("TB_LBFGS")
    both = ok & s2[:, 0]
    scale = np.abs(f2).max()
    close = np.abs(val - f2[:, 0]) <= 1e-4 * scale
    assert close[both].mean() > 0.85  # a few starts may settle in different local optima (different history handling)
    assert abs(val.max() - f2.max()) <= 1e-4 * scale


def test_device_lbfgs_active_bounds_and_dimension_extremes():
    # maximiser on the boundary: data from f(x) = sum(x) make the LCB smallest at the origin corner
    import trieste_b200 as tb
    from trieste_b200.acquisition import NegativeLowerConfidenceBound

    for D in (1, 3, 32):
        rng = np.random.default_rng(D)
        X = rng.uniform(size=(80, D))
        y = X.sum(axis=1, keepdims=True)
        space = tb.Box([0.0] * D, [1.0] * D)
        ds = tb.Dataset(X, y)
        nm = tb.GaussianProcessRegression(tb.build_gpr(ds, space, likelihood_variance=1e-3))
        fn = NegativeLowerConfidenceBound(0.5).prepare_acquisition_function(nm, ds)
        x0 = rng.uniform(size=(40, D))
        ok, val, x, nfev = fn.maximize_from(x0, space.lower, space.upper)
        assert ok.all()
        assert (x >= 0).all() and (x <= 1).all()
        best = x[np.argmax(val)]
        if D <= 3:  # enough data to resolve the trend: no run may end below the value at the origin corner
            corner = fn(np.zeros((1, 1, D)))[0, 0]
            assert val.max() >= corner - 1e-6 * max(1.0, abs(corner)), (D, best, val.max(), corner)
        pg = _proj_grad(fn, x, space.lower, space.upper)
        assert np.abs(pg).max() < 1e-3


def test_device_lbfgs_through_the_continuous_optimizer_and_argument_checks():
    import trieste_b200 as tb
    from trieste_b200.acquisition import AugmentedExpectedImprovement, MinValueEntropySearch
    from trieste_b200.acquisition.optimizer import generate_continuous_optimizer

    om, nm = model_pair(o.hartmann_6, 200, 6)
    space = tb.Box([0.0] * 6, [1.0] * 6)
    ds = tb.Dataset(om.X, om.y)
    opt = generate_continuous_optimizer(num_initial_samples=1000, num_optimization_runs=8)
    for builder in (AugmentedExpectedImprovement(), MinValueEntropySearch(space, 3, 300, seed=0)):
        fn = builder.prepare_acquisition_function(nm, ds)
        x = opt(space, fn)
        assert x.shape == (1, 6) and space.contains(x).all()
        rnd = space.sample(1000, seed=1)
        assert fn(x[:, None, :])[0, 0] >= fn(rnd[:, None, :]).max() * (1 - 1e-6) - 1e-12
        assert opt.last_stats["spo_af_evaluations"] >= 1
    fn = AugmentedExpectedImprovement().prepare_acquisition_function(nm, ds)
    ok, val, x, nfev = fn.maximize_from(np.zeros((0, 6)), space.lower, space.upper)
    assert ok.shape == (0,) and x.shape == (0, 6)
    with pytest.raises(ValueError):
        fn.maximize_from(candidates(4, 6), space.lower, space.upper, maxcor=17)
    with pytest.raises(ValueError):
        fn.maximize_from(candidates(4, 6), space.lower, space.upper, maxls=0)
    with pytest.raises(ValueError):
        fn.maximize_from(candidates(4, 5), space.lower[:5], space.upper[:5])  # wrong input dimension
    # one iteration only: maxiter is honoured and reported as not converged (LCB is never flat, unlike EI far from data)
    from trieste_b200.acquisition import NegativeLowerConfidenceBound

    lcb = NegativeLowerConfidenceBound(1.96).prepare_acquisition_function(nm, ds)
    ok1, _, _, n1 = lcb.maximize_from(candidates(16, 6), space.lower, space.upper, maxiter=1, gtol=0.0, ftol=0.0)
    assert not ok1.any() and n1.max() <= 1 + 20 + 1


def test_device_lbfgs_single_precision_model():
    import trieste_b200 as tb
    from trieste_b200.acquisition import LogExpectedImprovement

    rng = np.random.default_rng(0)
    X = rng.uniform(size=(500, 8)).astype(np.float32)
    y = o.ackley(X.astype(np.float64)).astype(np.float32)
    space = tb.Box([0.0] * 8, [1.0] * 8)
    ds = tb.Dataset(X, y)
    nm = tb.GaussianProcessRegression(tb.build_gpr(ds, space))
    assert nm.dtype == np.float32
    fn = LogExpectedImprovement().prepare_acquisition_function(nm, ds)
    x0 = rng.uniform(size=(64, 8))
    ok, val, x, nfev = fn.maximize_from(x0, space.lower, space.upper)
    assert x.dtype == np.float64 and (x >= 0).all() and (x <= 1).all()
    f0 = fn(x0.astype(np.float32)[:, None, :])[:, 0]
    assert np.all(val >= f0 - 1e-3 * np.abs(f0).max())
    assert val.max() > f0.max()
