"""GPU: BASELINE config 1 (README example, trieste README.md:33-66) end to end — Branin 2-D, 5 initial points,
ExpectedImprovement via EfficientGlobalOptimization, 15 BO steps — and the SAME loop on the oracle: with a shared
(seeded) candidate set per step, native and oracle must pick the same query point at every step."""
import numpy as np
import pytest

from oracle import gp_oracle as o

pytestmark = pytest.mark.gpu


def _loop_setup():
    import trieste_b200 as tb

    space = tb.Box([0.0, 0.0], [1.0, 1.0])
    X0 = space.sample(5, seed=0)
    ds = tb.Dataset(X0, o.branin(X0))
    spec = tb.build_gpr(ds, space, likelihood_variance=1e-7)  # docs/notebooks/expected_improvement.pct.py:94
    return tb, space, ds, spec


@pytest.mark.parametrize("engine", ["int8", "fp64"])
def test_config1_branin_ego_matches_oracle_step_by_step(engine):
    tb, space, ds, spec = _loop_setup()
    from trieste_b200.acquisition import ExpectedImprovement
    from trieste_b200.acquisition.optimizer import _get_max_discrete_points

    model = tb.GaussianProcessRegression(spec)
    model.set_engine(engine)
    builder = ExpectedImprovement()
    fn = None
    X, y = ds.query_points.copy(), ds.observations.copy()
    k = spec.kernel
    for step in range(15):
        cand = space.sample(5000, seed=100 + step)
        # native
        data = tb.Dataset(X, y)
        fn = builder.prepare_acquisition_function(model, data) if fn is None else builder.update_acquisition_function(fn, model, data)
        q_native = _get_max_discrete_points(cand[:, None, :], fn)
        # oracle on the same data / hyper-parameters / candidates
        om = o.build_model("matern52", X, y, k.variance, k.lengthscales, spec.noise_variance, spec.mean_function.c)
        ei = o.expected_improvement_at(om, cand, o.ei_eta(om))
        q_oracle = cand[int(np.argmax(ei[:, 0]))][None, :]
        np.testing.assert_array_equal(q_native, q_oracle, err_msg=f"different query point at BO step {step}")
        X = np.concatenate([X, q_native])
        y = np.concatenate([y, o.branin(q_native)])
        model.update(tb.Dataset(X, y))
    assert y.min() < 1.0  # Branin minimum 0.398; the 5 initial points give ~10 or worse
    assert y.min() < o.branin(space.sample(5, seed=0)).min()


def test_bayesian_optimizer_driver_runs_the_readme_example():
    tb, space, ds, spec = _loop_setup()
    from trieste_b200.bayesian_optimizer import BayesianOptimizer

    model = tb.GaussianProcessRegression(spec)
    result = BayesianOptimizer(o.branin, space).optimize(15, ds, model)
    final = result.try_get_final_dataset()
    assert len(final) == 20 and len(result.history) == 15
    x_best, y_best, _ = result.try_get_optimal_point()
    assert space.contains(x_best) and y_best[0] < 1.5

    def broken(x):
        raise RuntimeError("observer failed")

    res = BayesianOptimizer(broken, space).optimize(3, ds, tb.GaussianProcessRegression(spec))
    assert res.error is not None and len(res.history) == 0
    with pytest.raises(RuntimeError):
        res.try_get_final_dataset()


def test_batch_ego_with_monte_carlo_qei_and_joint_gradient_optimizer():
    # rule.py:291-297: a batch builder + num_query_points > 1 goes through batchify_joint over space ** q, driven by the
    # device-side value+gradient of the MC-qEI; each step appends q rows to the cached factors
    tb, space, ds, spec = _loop_setup()
    from trieste_b200.acquisition import BatchMonteCarloExpectedImprovement
    from trieste_b200.acquisition.optimizer import generate_continuous_optimizer
    from trieste_b200.bayesian_optimizer import BayesianOptimizer
    from trieste_b200.rule import EfficientGlobalOptimization

    model = tb.GaussianProcessRegression(spec)
    rule = EfficientGlobalOptimization(
        BatchMonteCarloExpectedImprovement(128),
        generate_continuous_optimizer(num_initial_samples=500, num_optimization_runs=8, optimizer_args={"maxiter": 40}),
        num_query_points=3,
    )
    result = BayesianOptimizer(o.branin, space).optimize(4, ds, model, rule)
    assert result.error is None, result.error
    final = result.try_get_final_dataset()
    assert len(final) == 5 + 4 * 3 and all(h.shape == (3, 2) for h in result.history)
    assert model.last_update_appended  # the BO steps extended the cache instead of refactorising
    assert final.observations.min() < ds.observations.min()
    # the appended cache equals a from-scratch one on the final data
    k = spec.kernel
    om = o.build_model("matern52", final.query_points, final.observations, k.variance, k.lengthscales, spec.noise_variance,
                       spec.mean_function.c)
    np.testing.assert_allclose(model.get_cholesky(), om.L, rtol=0, atol=1e-5 * np.sqrt(k.variance))  # noise 1e-7: cond ~1e9
