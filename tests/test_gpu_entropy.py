"""Min-value entropy search (entropy.py:52-213) and the min-value samplers that feed it (acquisition/sampler.py:126-273)
against the oracle restatement."""
import numpy as np
import pytest

from oracle import gp_oracle as o
from tests.util import candidates, model_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("engine", ["int8", "fp64"])
@pytest.mark.parametrize("S", [1, 5, 64])
def test_min_value_entropy_search_matches_oracle(S, engine):
    from trieste_b200.acquisition import min_value_entropy_search

    om, nm = model_pair(o.hartmann_6, 300, 6, engine=engine)
    Xq = np.concatenate([candidates(3000, 6), om.X[:20]])  # training inputs: small variance, large |gamma|
    rng = np.random.default_rng(S)
    samples = (om.y.min() - np.abs(rng.normal(size=(S, 1))) * np.sqrt(om.variance))
    if S > 1:
        samples[0, 0] = om.y.min() - 30.0 * np.sqrt(om.variance)  # far tail: the erfcx branch
        samples[1, 0] = om.y.max()  # a "minimum" above the mean: gamma > 0 branch
    fn = min_value_entropy_search(nm, samples)
    omean, ovar = o.predict(om, Xq)
    ref = o.min_value_entropy_search(omean, ovar, samples)
    got = fn(Xq[:, None, :])
    assert got.shape == (Xq.shape[0], 1)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-12)
    idx, best = fn.fused_argmax(Xq)
    assert abs(best - ref.max()) <= 1e-6 * abs(ref.max()) and abs(ref[idx, 0] - ref.max()) <= 1e-6 * abs(ref.max())


def test_min_value_entropy_search_gradient_and_errors():
    from trieste_b200.acquisition import expected_improvement, min_value_entropy_search

    om, nm = model_pair(o.hartmann_6, 200, 6)
    samples = np.array([[om.y.min() - 0.1], [om.y.min() - 0.5], [om.y.min() + 0.05]])
    fn = min_value_entropy_search(nm, samples)
    Xq = candidates(60, 6)
    val, grad = fn.value_and_gradient(Xq[:, None, :])
    np.testing.assert_allclose(val, fn(Xq[:, None, :]), rtol=1e-6, atol=1e-12)
    h = 1e-6
    for d in range(6):
        e = np.zeros(6)
        e[d] = h
        fd = []
        for sgn in (1, -1):
            mu, v = o.predict(om, Xq + sgn * e)
            fd.append(o.min_value_entropy_search(mu, v, samples))
        np.testing.assert_allclose(grad[:, 0, d], ((fd[0] - fd[1]) / (2 * h))[:, 0], rtol=2e-4, atol=1e-6 * np.abs(grad).max())
    with pytest.raises(ValueError):
        min_value_entropy_search(nm, np.zeros(3))  # rank < 2 (entropy.py:181)
    with pytest.raises(ValueError):
        min_value_entropy_search(nm, np.zeros((0, 1)))  # empty (entropy.py:182)
    with pytest.raises(ValueError):
        fn(candidates(6, 6).reshape(3, 2, 6))  # batch size one only (entropy.py:194-197)
    # the samples live in the handle: another function on the same model must not disturb this one
    other = min_value_entropy_search(nm, samples - 3.0)
    a = fn(Xq[:, None, :])
    other(Xq[:, None, :])
    expected_improvement(nm, 0.0)(Xq[:, None, :])
    np.testing.assert_array_equal(fn(Xq[:, None, :]), a)
    fn.update(samples - 3.0)
    np.testing.assert_array_equal(fn(Xq[:, None, :]), other(Xq[:, None, :]))


def test_gumbel_sampler_and_builder():
    import trieste_b200 as tb
    from trieste_b200.acquisition import MinValueEntropySearch, min_value_entropy_search
    from trieste_b200.acquisition.sampler import GumbelSampler, ThompsonSamplerFromTrajectory

    om, nm = model_pair(o.hartmann_6, 150, 6)
    at = candidates(1000, 6)
    omean, ovar = o.predict(om, at)
    a_ref, b_ref = o.gumbel_fit(omean, np.sqrt(ovar + om.noise))  # predict_y (sampler.py:179-182)
    s = GumbelSampler(sample_min_value=True, seed=3).sample(nm, 7, at)
    u = np.random.default_rng(3).uniform(size=7)
    assert s.shape == (7, 1)
    np.testing.assert_allclose(s, o.gumbel_samples(a_ref, b_ref, u), rtol=1e-7)
    with pytest.raises(ValueError):
        GumbelSampler(sample_min_value=False)
    with pytest.raises(ValueError):
        GumbelSampler(True).sample(nm, 0, at)
    with pytest.raises(ValueError):
        MinValueEntropySearch(tb.Box([0.0] * 6, [1.0] * 6), min_value_sampler=ThompsonSamplerFromTrajectory(sample_min_value=False))
    space = tb.Box([0.0] * 6, [1.0] * 6)
    ds = tb.Dataset(om.X, om.y)
    builder = MinValueEntropySearch(space, num_samples=5, grid_size=500, seed=0)
    fn = builder.prepare_acquisition_function(nm, ds)
    assert isinstance(fn, min_value_entropy_search) and fn.samples.shape == (5, 1)
    omean, ovar = o.predict(om, at)
    np.testing.assert_allclose(fn(at[:, None, :]), o.min_value_entropy_search(omean, ovar, fn.samples), rtol=1e-6, atol=1e-12)
    before = fn.samples.copy()
    assert builder.update_acquisition_function(fn, nm, ds) is fn
    assert not np.array_equal(before, fn.samples)  # fresh draws
    # trajectory-based min-value samples (sampler.py:262-266): each is the minimum of one trajectory over the candidates
    ts = ThompsonSamplerFromTrajectory(sample_min_value=True).sample(nm, 3, at)
    assert ts.shape == (3, 1) and np.all(ts < omean.max())
    fn2 = MinValueEntropySearch(space, 3, 200, min_value_sampler=ThompsonSamplerFromTrajectory(True)).prepare_acquisition_function(nm, ds)
    assert np.isfinite(fn2(at[:, None, :])).all()
