"""GPU parity of the INT8-tensor-core engine (Ozaki splitting): SAME stated fp64 tolerances as the native
DMMA engine (mean rtol 1e-9, variance atol 1e-9 sigma_f^2, EI rtol 1e-6)."""
import numpy as np
import pytest

from oracle import gp_oracle as o
from tests.util import candidates, model_pair

pytestmark = pytest.mark.gpu


def _check(om, nm, Xq):
    """Both product counts of the int8 engine ("int8" = picked from the a-priori estimate: 15 or 21; "int8x21" = pinned)
    against the oracle at the stated bar, and against the native engine: tighter still."""
    nm.set_engine("fp64")
    mean64, var64 = nm.predict(Xq)
    omean, ovar = o.predict_batched(om, Xq)
    sf = np.sqrt(om.variance)
    worst = 0.0
    for engine in ("int8", "int8x21"):
        nm.set_engine(engine)
        products, est = nm.engine_info()
        assert products == 21 if engine == "int8x21" else products in (15, 21)
        mean, var = nm.predict(Xq)
        np.testing.assert_allclose(mean, omean, rtol=1e-9, atol=1e-9 * sf)
        np.testing.assert_allclose(var, ovar, rtol=0, atol=1e-9 * om.variance)
        np.testing.assert_allclose(var, var64, rtol=0, atol=(5e-10 if products == 15 else 2e-10) * om.variance)
        if products == 15:  # the estimate that admitted the reduced mode must cover what is measured
            assert est <= 3e-10 and np.abs(var - var64).max() <= 3.0 * max(est, 1e-11) * om.variance
        assert var.min() >= 1e-12
        worst = max(worst, np.abs(var - ovar).max() / om.variance)
    nm.set_engine("int8")
    return worst


@pytest.mark.parametrize("kind", ["matern52", "rbf"])
@pytest.mark.parametrize("N,D", [(5, 2), (20, 2), (63, 3), (64, 3), (65, 3), (127, 6), (128, 6), (129, 6), (300, 6), (1024, 6)])
def test_int8_engine_predict_matches_oracle(kind, N, D):
    obj = o.branin if D == 2 else (o.hartmann_6 if D == 6 else o.ackley)
    om, nm = model_pair(obj, N, D, kind=kind)
    _check(om, nm, candidates(777, D))


def test_int8_engine_headline_n4096_and_near_training_points():
    om, nm = model_pair(o.ackley, 4096, 10)
    Xq = np.concatenate([candidates(2500, 10), om.X[:300] + 1e-7, om.X[300:500]])
    err = _check(om, nm, Xq)
    print(f"max |dvar| / sigma_f^2 at N=4096: {err:.3e}")


def test_int8_engine_picks_the_product_count_from_the_error_estimate():
    # the benchmark configurations run the 15-product single-pass kernel ...
    for obj, N, D in [(o.ackley, 4096, 10), (o.hartmann_6, 1024, 6)]:
        om, nm = model_pair(obj, N, D)
        products, est = nm.engine_info()
        assert products == 15 and 0 < est <= 3e-10, (N, products, est)
        nm.set_engine("int8x21")
        assert nm.engine_info()[0] == 21
        nm.set_engine("fp64")
        assert nm.engine_info()[0] == 0
    # ... a badly scaled factor (little noise on a smooth kernel: huge rows of Linv; estimate 3e-9) keeps the full 21 products
    var = o.synthetic_model(o.branin, 300, 2, kind="rbf").variance
    om, nm = model_pair(o.branin, 300, 2, kind="rbf", noise=1e-5 * var)
    products, est = nm.engine_info()
    assert products == 21, (products, est)
    _check(om, nm, candidates(500, 2))


def test_int8_engine_low_noise_clip_and_multi_chunk():
    om, nm = model_pair(o.branin, 20, 2, noise=1e-7)
    _check(om, nm, np.concatenate([om.X, candidates(500, 2)]))
    om, nm = model_pair(o.hartmann_6, 1024, 6)
    Xq = candidates(300_000, 6)
    nm.set_engine("int8")
    mean, var = nm.predict(Xq)
    idx = np.random.default_rng(3).choice(Xq.shape[0], 4096, replace=False)
    omean, ovar = o.predict(om, Xq[idx])
    np.testing.assert_allclose(mean[idx], omean, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(var[idx], ovar, rtol=0, atol=1e-9 * om.variance)


def test_int8_engine_ei_and_argmax():
    from trieste_b200 import Dataset
    from trieste_b200.acquisition import ExpectedImprovement

    om, nm = model_pair(o.hartmann_6, 1024, 6)
    nm.set_engine("int8")
    Xq = candidates(20000, 6)
    fn = ExpectedImprovement().prepare_acquisition_function(nm, Dataset(om.X, om.y))
    ei = fn(Xq[:, None, :])
    omean, ovar = o.predict(om, Xq)
    oei = o.expected_improvement(omean, ovar, o.ei_eta(om))
    big = oei > 1e-12
    np.testing.assert_allclose(ei[big], oei[big], rtol=1e-6)
    idx, best = fn.fused_argmax(Xq)
    assert idx == int(np.argmax(ei[:, 0]))
    # gradients fall back to the fp64 engine transparently
    val, grad = fn.value_and_gradient(Xq[:100, None, :])
    np.testing.assert_allclose(val, ei[:100], rtol=1e-6, atol=1e-15)
