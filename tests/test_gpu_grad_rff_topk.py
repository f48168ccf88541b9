"""GPU parity: acquisition gradients, RFF trajectories / Thompson sampling, top-k, optimisers."""
import numpy as np
import pytest

from oracle import gp_oracle as o
from tests.util import candidates, model_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["matern52", "rbf", "matern32"])
@pytest.mark.parametrize("N,D", [(20, 2), (300, 6), (1024, 10)])
def test_ei_gradient_matches_oracle(kind, N, D):
    from trieste_b200.acquisition import expected_improvement

    obj = o.branin if D == 2 else (o.hartmann_6 if D == 6 else o.ackley)
    om, nm = model_pair(obj, N, D, kind=kind)
    Xq = candidates(300, D)
    eta = o.ei_eta(om)
    fn = expected_improvement(nm, eta)
    val, grad = fn.value_and_gradient(Xq[:, None, :])
    oval, ograd = o.ei_gradient(om, Xq, eta)
    assert val.shape == (300, 1) and grad.shape == (300, 1, D)
    np.testing.assert_allclose(val, oval, rtol=1e-6, atol=1e-15)
    scale = np.abs(ograd).max()
    np.testing.assert_allclose(grad[:, 0, :], ograd, rtol=1e-6, atol=1e-9 * scale)
    # the value-only path may run on the int8 engine, the gradient path runs on the native fp64 engine
    np.testing.assert_allclose(val, fn(Xq[:, None, :]), rtol=1e-6, atol=1e-15)


def test_lcb_and_logei_gradients_by_finite_differences():
    from trieste_b200.acquisition import log_expected_improvement, lower_confidence_bound

    om, nm = model_pair(o.hartmann_6, 200, 6)
    Xq = candidates(50, 6)
    for fn in [lower_confidence_bound(nm, 1.96), log_expected_improvement(nm, o.ei_eta(om))]:
        val, grad = fn.value_and_gradient(Xq[:, None, :])
        h = 1e-6
        for d in range(6):
            e = np.zeros(6)
            e[d] = h
            fd = (fn((Xq + e)[:, None, :]) - fn((Xq - e)[:, None, :])) / (2 * h)
            np.testing.assert_allclose(grad[:, 0, d], fd[:, 0], rtol=2e-4, atol=1e-6 * np.abs(grad).max())


@pytest.mark.parametrize("M,k", [(1, 1), (17, 5), (2048, 2048), (5000, 10), (100_000, 1000), (300_000, 7)])
def test_top_k_matches_tf_semantics(M, k):
    from trieste_b200.sampler import top_k

    rng = np.random.default_rng(M)
    v = rng.standard_normal(M)
    v[rng.integers(0, M, size=max(1, M // 10))] = v[0]  # ties
    tv, ti = top_k(v, k)
    ov, oi = o.top_k(v, k)
    np.testing.assert_array_equal(tv, ov)
    np.testing.assert_array_equal(ti, oi)


@pytest.mark.parametrize("kind", ["rbf", "matern52"])
@pytest.mark.parametrize("n,F", [(50, 200), (300, 128)])  # gram space (n <= F) and design space (F < n)
def test_rff_trajectory_matches_oracle(kind, n, F):
    om, nm = model_pair(o.hartmann_6, n, 6, kind=kind)
    from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler

    sampler = RandomFourierFeatureTrajectorySampler(nm, F, seed=0)
    W, b = o.rff_draw(kind, F, 6, np.random.default_rng(4))
    sampler._feature_functions.set_weights(W, b)
    mean, chol = sampler.theta_posterior()
    omean, ochol = o.rff_theta_posterior(om, W, b)
    np.testing.assert_allclose(mean.cpu().numpy(), omean, rtol=1e-6, atol=1e-8 * np.abs(omean).max())
    np.testing.assert_allclose((chol @ chol.T).cpu().numpy(), ochol @ ochol.T, rtol=1e-6, atol=1e-9)
    traj = sampler.get_trajectory()
    Xq = candidates(5000, 6)
    out = traj(Xq[:, None, :])
    theta = traj._weights_sample  # [1, F]
    ref = o.rff_trajectory(Xq[:, None, :], W, b, theta, om.variance, om.lengthscales, om.mean_const)
    assert out.shape == (5000, 1, 1)
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-9 * np.sqrt(om.variance))
    mv, mi = traj.argmin_over(Xq)
    assert int(mi[0]) == int(np.argmin(ref[:, 0, 0])) and mv[0] == out[mi[0], 0, 0]
    with pytest.raises(ValueError):
        traj(np.zeros((10, 2, 6)))  # batch size is fixed by the first call


def test_rff_design_equals_gram_space():
    # reference test restated (tests/unit/models/gpflow/test_sampler.py:530-542): both posteriors agree
    om, nm = model_pair(o.hartmann_6, 100, 6, kind="rbf")
    W, b = o.rff_draw("rbf", 100, 6, np.random.default_rng(4))
    md, cd = o.rff_theta_posterior(om, W[:99], b[:99])  # F = 99 < n: design
    # same features through the gram route by calling the native sampler with F = n (gram branch)
    from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler

    s = RandomFourierFeatureTrajectorySampler(nm, 100, seed=0)
    s._feature_functions.set_weights(W, b)
    mg, cg = s.theta_posterior()
    om_g, oc_g = o.rff_theta_posterior(om, W, b)
    np.testing.assert_allclose(mg.cpu().numpy(), om_g, rtol=1e-6, atol=1e-8)


def test_rff_batched_trajectories_and_thompson_sampler():
    from trieste_b200.acquisition.sampler import ThompsonSamplerFromTrajectory
    from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler

    om, nm = model_pair(o.hartmann_6, 100, 6)
    sampler = RandomFourierFeatureTrajectorySampler(nm, 256, seed=1)
    traj = sampler.get_trajectory()
    Xq = candidates(4000, 6)
    X3 = np.stack([Xq, Xq[::-1], Xq], axis=1)  # [N, B=3, D]
    out = traj(X3)
    W, b = sampler._feature_functions.W, sampler._feature_functions.b
    ref = o.rff_trajectory(X3, W, b, traj._weights_sample, om.variance, om.lengthscales, om.mean_const)
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-9 * np.sqrt(om.variance))
    mv, mi = traj.argmin_over(Xq)
    full = o.rff_trajectory(np.repeat(Xq[:, None, :], 3, 1), W, b, traj._weights_sample, om.variance, om.lengthscales, om.mean_const)
    np.testing.assert_array_equal(mi, np.argmin(full[:, :, 0], axis=0))
    pts = ThompsonSamplerFromTrajectory().sample(nm, 4, Xq)
    assert pts.shape == (4, 6)
    assert all(np.any(np.all(Xq == p, axis=1)) for p in pts)


def test_random_search_optimizer_and_initial_points():
    from trieste_b200 import Box, Dataset
    from trieste_b200.acquisition import ExpectedImprovement
    from trieste_b200.acquisition.optimizer import (
        generate_initial_points,
        generate_random_search_optimizer,
        sample_from_space,
    )

    om, nm = model_pair(o.hartmann_6, 200, 6)
    fn = ExpectedImprovement().prepare_acquisition_function(nm, Dataset(om.X, om.y))
    space = Box([0.0] * 6, [1.0] * 6)
    pt = generate_random_search_optimizer(20000)(space, fn)
    assert pt.shape == (1, 6) and space.contains(pt).all()
    # the fused argmax equals argmax of the evaluated values on the same points
    pts = space.sample(5000, seed=3)
    from trieste_b200.acquisition.optimizer import _get_max_discrete_points

    best = _get_max_discrete_points(pts[:, None, :], fn)
    vals = fn(pts[:, None, :])
    np.testing.assert_array_equal(best[0], pts[int(np.argmax(vals[:, 0]))])
    # streaming top-k over chunks == top-k over everything (optimizer.py:299-335)
    space._rng = np.random.default_rng(5)
    init = generate_initial_points(7, sample_from_space(3000, batch_size=1000), space, fn)
    assert init.shape == (7, 1, 6)
    space._rng = np.random.default_rng(5)
    allpts = np.concatenate([space.sample(1000) for _ in range(3)])
    v = fn(allpts[:, None, :])[:, 0]
    _, oi = o.top_k(v, 7)
    np.testing.assert_allclose(init[:, 0, :], allpts[oi], rtol=0, atol=0)
    with pytest.raises(ValueError):
        generate_random_search_optimizer(0)


def test_continuous_optimizer_finds_local_maximum():
    from trieste_b200 import Box, Dataset
    from trieste_b200.acquisition import NegativeLowerConfidenceBound
    from trieste_b200.acquisition.optimizer import generate_continuous_optimizer

    om, nm = model_pair(o.hartmann_6, 300, 6)
    fn = NegativeLowerConfidenceBound(1.96).prepare_acquisition_function(nm, Dataset(om.X, om.y))
    space = Box([0.0] * 6, [1.0] * 6)
    opt = generate_continuous_optimizer(num_initial_samples=2000, num_optimization_runs=16)
    x = opt(space, fn)
    assert x.shape == (1, 6) and space.contains(x).all()
    val, grad = fn.value_and_gradient(x[:, None, :])
    # first-order optimality of the projected gradient, and no worse than the best random sample
    pg = x - np.clip(x + grad[:, 0, :], space.lower, space.upper)
    assert np.abs(pg).max() < 1e-3
    rnd = space.sample(2000, seed=1)
    assert val[0, 0] >= fn(rnd[:, None, :]).max() - 1e-9


def test_efficient_global_optimization_rule_branin():
    # config 1 (README example) in miniature: EGO picks a point where EI is (near-)maximal
    from trieste_b200 import Box, Dataset, GaussianProcessRegression, build_gpr
    from trieste_b200.rule import EfficientGlobalOptimization

    space = Box([0.0, 0.0], [1.0, 1.0])
    X = space.sample(5, seed=0)
    y = o.branin(X)
    ds = Dataset(X, y)
    model = GaussianProcessRegression(build_gpr(ds, space, likelihood_variance=1e-7))
    rule = EfficientGlobalOptimization()
    q = rule.acquire_single(space, model, ds)
    assert q.shape == (1, 2) and space.contains(q).all()
    fn = rule.acquisition_function
    grid = space.sample(20000, seed=2)
    assert fn(q[:, None, :])[0, 0] >= fn(grid[:, None, :]).max() * (1 - 1e-6) - 1e-12
    # second step re-uses (updates) the same function object
    X2 = np.concatenate([X, q])
    ds2 = Dataset(X2, o.branin(X2))
    model.update(ds2)
    q2 = rule.acquire_single(space, model, ds2)
    assert rule.acquisition_function is fn and q2.shape == (1, 2)


@pytest.mark.parametrize("kind", ["rbf", "matern52"])
def test_decoupled_trajectory_matches_oracle(kind):
    # DecoupledTrajectorySampler (sampler.py:594-738) with injected W, b, prior weights and noise draws
    from trieste_b200.sampler import DecoupledTrajectorySampler

    om, nm = model_pair(o.hartmann_6, 150, 6, kind=kind)
    F, B = 256, 3
    s = DecoupledTrajectorySampler(nm, F, seed=0)
    W, b = o.rff_draw(kind, F, 6, np.random.default_rng(4))
    s._feature_functions.set_weights(W, b)
    rng = np.random.default_rng(7)
    pw, eps = rng.standard_normal((B, F)), rng.standard_normal((B, 150))
    v = s.canonical_weights(pw, eps)
    ov = o.decoupled_weights(om, W, b, pw, eps)
    np.testing.assert_allclose(v, ov, rtol=1e-7, atol=1e-9 * np.abs(ov).max())
    traj = s.get_trajectory()
    traj._weight_sampler = lambda nb: (pw, v)
    Xq = candidates(3000, 6)
    X3 = np.stack([Xq, Xq[::-1], Xq], axis=1)
    out = traj(X3)
    ref = o.decoupled_trajectory(om, X3, W, b, pw, ov)
    assert out.shape == (3000, 3, 1)
    np.testing.assert_allclose(out, ref, rtol=1e-8, atol=1e-8 * np.sqrt(om.variance))
    mv, mi = traj.argmin_over(Xq)
    full = o.decoupled_trajectory(om, np.repeat(Xq[:, None, :], 3, 1), W, b, pw, ov)
    np.testing.assert_array_equal(mi, np.argmin(full[:, :, 0], axis=0))
    # default trajectory sampler of the model is the decoupled one (models.py:342-345)
    assert isinstance(nm.trajectory_sampler(), DecoupledTrajectorySampler)


def test_decoupled_trajectory_moments():
    # trajectory moments vs predict (tests/unit/models/gpflow/test_models.py:638-681 restated, use_decoupled_sampler=True)
    om, nm = model_pair(o.hartmann_6, 60, 6, kind="rbf")
    from trieste_b200.sampler import DecoupledTrajectorySampler

    s = DecoupledTrajectorySampler(nm, 2000, seed=3)
    traj = s.get_trajectory()
    Xq = candidates(8, 6)
    S = 300
    f = traj(np.repeat(Xq[:, None, :], S, 1))[:, :, 0]
    mean, var = nm.predict(Xq)
    np.testing.assert_allclose(f.mean(1), mean[:, 0], atol=0.2 * np.sqrt(om.variance))
    np.testing.assert_allclose(f.var(1), var[:, 0], rtol=0.5, atol=0.02 * om.variance)
