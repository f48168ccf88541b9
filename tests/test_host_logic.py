"""CPU: host-side logic that needs no GPU — containers, search spaces, the vectorised L-BFGS, the
multi-rank argmax exchange (gloo, world_size 2)."""
import os

import numpy as np
import pytest

from trieste_b200.acquisition.optimizer import (
    FailedOptimizationError,
    _get_max_discrete_points,
    _perform_parallel_continuous_optimization,
    batchify_joint,
    generate_random_search_optimizer,
    sample_from_space,
)
from trieste_b200.data import Dataset
from trieste_b200.parallel import merge_best, shard_bounds
from trieste_b200.space import Box, DiscreteSearchSpace


class Quadratic:
    """f(x) = -sum a (x - c)^2 with analytic gradient: stands in for an acquisition function."""

    def __init__(self, c, a=None):
        self.c = np.asarray(c, dtype=float)
        self.a = np.ones_like(self.c) if a is None else np.asarray(a, dtype=float)
        self.calls = 0

    def __call__(self, x):
        x = np.asarray(x)
        return -(self.a * (x[..., 0, :] - self.c) ** 2).sum(-1, keepdims=True)

    def value_and_gradient(self, x):
        self.calls += 1
        d = x[:, 0, :] - self.c
        return -(self.a * d * d).sum(-1, keepdims=True), (-2 * self.a * d)[:, None, :]


def test_dataset_shapes_and_concat():
    d = Dataset(np.zeros((3, 2)), np.ones((3, 1)))
    assert len(d + d) == 6
    with pytest.raises(ValueError):
        Dataset(np.zeros((3, 2)), np.ones((4, 1)))
    with pytest.raises(ValueError):
        Dataset(np.zeros(3), np.ones(3))


def test_box_sampling_and_power():
    b = Box([0.0, -1.0], [1.0, 3.0])
    s = b.sample(1000, seed=0)
    assert s.shape == (1000, 2) and b.contains(s).all()
    assert (b**3).dimension == 6
    with pytest.raises(ValueError):
        Box([1.0], [0.0])
    with pytest.raises(ValueError):
        b.sample(-1)
    np.testing.assert_array_equal(b.sample(5, seed=3), b.sample(5, seed=3))


def test_random_search_generic_path_and_vectorised():
    space = Box([0.0, 0.0], [1.0, 1.0])
    f = Quadratic([0.25, 0.75])
    pt = generate_random_search_optimizer(4000)(space, f)
    assert pt.shape == (1, 2) and np.abs(pt[0] - [0.25, 0.75]).max() < 0.05

    def vec(x):  # [N, 2, D] -> [N, 2]
        return np.stack([-((x[:, 0] - 0.2) ** 2).sum(-1), -((x[:, 1] - 0.8) ** 2).sum(-1)], axis=1)

    pts = _get_max_discrete_points(space.sample(4000, seed=1)[:, None, :], (vec, 2))
    assert pts.shape == (2, 2) and np.abs(pts[0] - 0.2).max() < 0.05 and np.abs(pts[1] - 0.8).max() < 0.05
    with pytest.raises(ValueError):
        _get_max_discrete_points(space.sample(10)[:, None, :], (lambda x: np.zeros((10, 3)), 2))
    d = DiscreteSearchSpace(np.array([[0.0, 0.0], [0.3, 0.7], [1.0, 1.0]]))
    from trieste_b200.acquisition.optimizer import optimize_discrete

    np.testing.assert_array_equal(optimize_discrete(d, f), [[0.3, 0.7]])


def test_first_max_tie_rule():
    pts = np.array([[0.0], [1.0], [2.0], [3.0]])[:, None, :]
    best = _get_max_discrete_points(pts, lambda x: np.array([[1.0], [5.0], [5.0], [0.0]]))
    np.testing.assert_array_equal(best, [[1.0]])


def test_sample_from_space_chunks():
    space = Box([0.0], [1.0])
    chunks = list(sample_from_space(10, batch_size=4)(space))
    assert [c.shape[0] for c in chunks] == [4, 4, 2]
    with pytest.raises(ValueError):
        sample_from_space(0)


@pytest.mark.parametrize("D", [2, 6, 20])
def test_vectorised_lbfgs_interior_and_bound_optima(D):
    rng = np.random.default_rng(D)
    c = rng.uniform(-0.3, 1.3, size=D)  # some optima outside the unit box -> active bounds
    a = rng.uniform(0.5, 50.0, size=D)
    f = Quadratic(c, a)
    x0 = rng.uniform(size=(64, 1, D))
    success, fun, xs, nfev = _perform_parallel_continuous_optimization(f, np.zeros(D), np.ones(D), x0, {})
    assert success.all()
    np.testing.assert_allclose(xs[:, 0, :], np.broadcast_to(np.clip(c, 0, 1), (64, D)), atol=1e-4)
    assert nfev.max() < 60
    # every iteration is ONE batched evaluation for all active starts, not one per start
    assert f.calls < 80


def test_lbfgs_rosenbrock_from_many_starts():
    class Rosen:
        def __call__(self, x):
            return self.value_and_gradient(np.asarray(x))[0]

        def value_and_gradient(self, x):
            u, v = x[:, 0, 0], x[:, 0, 1]
            val = -((1 - u) ** 2 + 100 * (v - u * u) ** 2)
            g = -np.stack([-2 * (1 - u) - 400 * u * (v - u * u), 200 * (v - u * u)], axis=1)
            return val[:, None], g[:, None, :]

    x0 = np.random.default_rng(0).uniform(-1.5, 1.5, size=(32, 1, 2))
    success, fun, xs, _ = _perform_parallel_continuous_optimization(Rosen(), np.full(2, -2.0), np.full(2, 2.0), x0, {})
    assert success.mean() > 0.9
    np.testing.assert_allclose(xs[success[:, 0], 0, :], 1.0, atol=1e-3)


def test_failed_optimisation_raises():
    from trieste_b200.acquisition.optimizer import generate_continuous_optimizer

    class Bad:
        def __call__(self, x):
            return np.full(np.shape(x)[:-2] + (1,), np.nan)

        def value_and_gradient(self, x):
            return np.full((x.shape[0], 1), np.nan), np.full(x.shape, np.nan)

    import trieste_b200.acquisition.optimizer as opt

    orig = opt.generate_initial_points
    opt.generate_initial_points = lambda k, s, space, fn, vectorization=1: space.sample(k)[:, None, :]
    try:
        with pytest.raises(FailedOptimizationError):
            generate_continuous_optimizer(10, 2, num_recovery_runs=1)(Box([0.0], [1.0]), Bad())
    finally:
        opt.generate_initial_points = orig
    with pytest.raises(ValueError):
        generate_continuous_optimizer(5, 10)


def test_batchify_joint_reshapes():
    space = Box([0.0, 0.0], [1.0, 1.0])

    def joint(x):  # [..., q=3, D=2] -> [..., 1]: wants the 3 points at 0.1, 0.5, 0.9
        t = np.array([0.1, 0.5, 0.9])[:, None]
        return -((x - t) ** 2).sum((-1, -2))[..., None]

    opt = batchify_joint(generate_random_search_optimizer(20000), 3)
    pts = opt(space, joint)
    assert pts.shape == (3, 2)
    assert np.abs(pts - np.array([0.1, 0.5, 0.9])[:, None]).max() < 0.2


def test_shard_bounds_partition():
    for total in [0, 1, 7, 8, 1000003]:
        for world in [1, 2, 3, 8]:
            cuts = [shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_merge_best_tie_and_nan_rules():
    assert merge_best([(1.0, 5), (3.0, 9), (3.0, 2), (float("nan"), 0)]) == (3.0, 2)
    assert merge_best([(-np.inf, -1), (0.5, 7)]) == (0.5, 7)
    assert merge_best([]) == (-np.inf, -1)


def test_merge_topk_follows_tf_top_k():
    from trieste_b200.parallel import merge_topk

    v = np.array([0.5, 2.0, 2.0, np.nan, 1.0, 2.0, -np.inf])
    i = np.array([40, 30, 7, 1, 5, -1, 9])
    sel = merge_topk(v, i, 4)
    assert list(i[sel]) == [7, 30, 5, 40]  # ties on the value: lower global index first; NaN and padding (-1) never selected
    assert list(merge_topk(v, i, 100)) == list(merge_topk(v, i, 5))  # k larger than the valid count
    assert len(merge_topk(v, i, 0)) == 0


def _gloo_topk_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trieste_b200.parallel import allgather_topk, shard_bounds

        pts = np.random.default_rng(4).uniform(size=(301, 2))
        vals = np.round(np.sin(7 * pts[:, 0]) + pts[:, 1], 1)  # many ties
        lo, hi = shard_bounds(len(pts), rank, world)
        k = 9
        order = np.lexsort((np.arange(lo, hi), -vals[lo:hi]))[:k]  # the rank's own top-k (stands in for tb_topk)
        bv, bi, bp = allgather_topk(vals[lo:hi][order], lo + order, pts[lo:hi][order], k)
        # a rank with fewer than k candidates pads its payload
        sv, si, sp = allgather_topk(vals[lo : lo + 2], np.arange(lo, lo + 2), pts[lo : lo + 2], 3)
        q.put((rank, bv.tolist(), bi.tolist(), bp.tolist(), si.tolist()))
    finally:
        dist.destroy_process_group()


def test_allgather_topk_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_topk_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts = np.random.default_rng(4).uniform(size=(301, 2))
    vals = np.round(np.sin(7 * pts[:, 0]) + pts[:, 1], 1)
    want = np.lexsort((np.arange(301), -vals))[:9]
    small = sorted([0, 1, 151, 152], key=lambda j: (-vals[j], j))[:3]
    for rank, bv, bi, bp, si in res:
        assert bi == list(want)
        np.testing.assert_allclose(bv, vals[want])
        np.testing.assert_allclose(bp, pts[want])
        assert si == small
    assert res[0][1:] == res[1][1:]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from trieste_b200.parallel import sharded_argmax

        pts = np.random.default_rng(0).uniform(size=(1001, 3))
        vals = -((pts - 0.5) ** 2).sum(-1)
        vals[[10, 900]] = 1.0  # a tie straddling the two shards: global index 10 must win everywhere

        class Fn:  # stands in for the fused GPU argmax of one rank's shard
            def fused_argmax(self, p):
                lo = int(np.where((pts == p[0]).all(1))[0][0])
                v = vals[lo : lo + len(p)]
                i = int(np.argmax(v))
                return i, float(v[i])

        pt, bv, bi = sharded_argmax(Fn(), pts)

        # Thompson argmin over sharded candidates with a fake 2-trajectory evaluator
        from trieste_b200.parallel import sharded_multistart, sharded_thompson_argmin

        tv = np.stack([(pts - 0.25).sum(-1) ** 2, -pts[:, 0]], axis=1)  # [M, 2]

        class Traj:
            _batch_size = 2

            def argmin_over(self, p):
                lo = int(np.where((pts == p[0]).all(1))[0][0])
                v = tv[lo : lo + len(p)]
                idx = np.argmin(v, axis=0)
                return v[idx, np.arange(2)], idx

        tp, tvals, tidx = sharded_thompson_argmin(Traj(), pts)

        # multi-start sharding: each rank "optimises" its slice (identity) and reports its best
        ms_pt, ms_v, ms_i = sharded_multistart(lambda st: (st, -((st - 0.3) ** 2).sum(-1)), pts)
        q.put((rank, bi, bv, pt.tolist(), tidx.tolist(), tvals.tolist(), int(ms_i), float(ms_v)))
    finally:
        dist.destroy_process_group()


def test_sharded_argmax_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pts = np.random.default_rng(0).uniform(size=(1001, 3))
    tv = np.stack([(pts - 0.25).sum(-1) ** 2, -pts[:, 0]], axis=1)
    ms = -((pts - 0.3) ** 2).sum(-1)
    for rank, bi, bv, pt, tidx, tvals, ms_i, ms_v in res:
        assert bi == 10 and bv == 1.0
        np.testing.assert_allclose(pt[0], pts[10])
        assert tidx == list(np.argmin(tv, axis=0))
        np.testing.assert_allclose(tvals, tv.min(0))
        assert ms_i == int(np.argmax(ms)) and ms_v == ms.max()


def test_gumbel_fit_and_sampler_argument_checks():
    # acquisition/sampler.py:140-152,186-204 (host logic only: no model call)
    from oracle import gp_oracle as o
    from trieste_b200.acquisition.sampler import GumbelSampler, ThompsonSampler, ThompsonSamplerFromTrajectory

    rng = np.random.default_rng(0)
    mu, sd = rng.normal(size=40), rng.uniform(0.1, 2.0, size=40)
    np.testing.assert_allclose(GumbelSampler.fit(mu, sd), o.gumbel_fit(mu, sd), rtol=1e-12)
    with pytest.raises(ValueError):
        GumbelSampler()  # can only sample the minimum value
    assert ThompsonSamplerFromTrajectory().sample_min_value is False
    assert ThompsonSamplerFromTrajectory(True).sample_min_value is True
    assert "True" in repr(ThompsonSampler(True))
    with pytest.raises(ValueError):
        ThompsonSamplerFromTrajectory().sample(object(), 1, np.zeros((3, 2)))  # no trajectory_sampler
    with pytest.raises(ValueError):
        ThompsonSamplerFromTrajectory().sample(object(), 1, np.zeros(3))  # at must be [N, D]


def test_batchify_joint_passes_gradients_through_the_reshape():
    # optimizer.py:897-936: a batch function maximised over space ** q by the continuous optimiser
    from trieste_b200.acquisition.optimizer import batchify_joint, generate_continuous_optimizer

    centres = np.array([[0.2, 0.7], [0.9, 0.1], [0.5, 0.5]])

    class BatchQuadratic:
        def __call__(self, x):
            return self.value_and_gradient(x)[0]

        def value_and_gradient(self, x):
            x = np.asarray(x)  # [..., 3, 2]
            diff = x - centres
            return -np.sum(diff * diff, axis=(-1, -2))[..., None], -2.0 * diff

    space = Box([0.0, 0.0], [1.0, 1.0])

    def inner(expanded, f):  # the multi-start loop of generate_continuous_optimizer without its GPU top-k of the starts
        assert expanded.dimension == 6
        x0 = np.random.default_rng(0).uniform(size=(4, 1, 6))
        val, grad = f.value_and_gradient(x0)
        assert val.shape == (4, 1) and grad.shape == x0.shape
        ok, fun, xs, _ = _perform_parallel_continuous_optimization(f, expanded.lower, expanded.upper, x0, {})
        assert ok.all()
        return xs[np.argmax(fun[:, 0]), 0][None, :]

    pts = batchify_joint(inner, 3)(space, BatchQuadratic())
    assert pts.shape == (3, 2)
    np.testing.assert_allclose(pts, centres, atol=1e-4)
    with pytest.raises(ValueError):
        batchify_joint(generate_continuous_optimizer(), 0)


class _FakeModel:
    """predict / sample of a fixed independent Gaussian: mean = sum(x), var = 0.25 (host logic only)."""

    def predict(self, x):
        x = np.asarray(x, dtype=np.float64)
        return x.sum(-1, keepdims=True), np.full(x.shape[:-1] + (1,), 0.25)

    def sample(self, at, num_samples, seed=None):
        at = np.asarray(at, dtype=np.float64)
        z = np.random.default_rng(seed).standard_normal((num_samples, at.shape[0]))
        return (at.sum(-1)[None, :] + 0.5 * z)[..., None]


def test_independent_sampler_and_exact_thompson_sampler_host_logic():
    # models/gpflow/sampler.py:82-164 and acquisition/sampler.py:85-123 on a stand-in model (no GPU)
    from trieste_b200.acquisition.sampler import ExactThompsonSampler
    from trieste_b200.sampler import IndependentReparametrizationSampler

    m = _FakeModel()
    s = IndependentReparametrizationSampler(8, m, seed=0)
    x = np.random.default_rng(1).uniform(size=(5, 1, 3))
    out = s.sample(x, jitter=0.0)
    assert out.shape == (5, 8, 1, 1)
    eps = np.random.default_rng(0).standard_normal((8, 1))
    np.testing.assert_allclose(out[:, :, 0, 0], x.sum(-1) + 0.5 * eps[:, 0][None, :])
    np.testing.assert_array_equal(out, s.sample(x, jitter=0.0))
    s.set_eps(np.zeros(8))
    np.testing.assert_allclose(s.sample(x)[:, :, 0, 0], np.broadcast_to(x.sum(-1), (5, 8)))
    with pytest.raises(ValueError):
        s.set_eps(np.zeros(7))
    with pytest.raises(ValueError):
        s.sample(x, jitter=-1.0)
    at = np.random.default_rng(2).uniform(size=(50, 3))
    pts = ExactThompsonSampler().sample(m, 6, at, seed=3)
    mins = ExactThompsonSampler(True).sample(m, 6, at, seed=3)
    draws = m.sample(at, 6, seed=3)[..., 0]
    np.testing.assert_array_equal(pts, at[draws.argmin(1)])
    np.testing.assert_array_equal(mins[:, 0], draws.min(1))
    with pytest.raises(ValueError):
        ExactThompsonSampler().sample(m, 0, at)


def test_conditional_predict_host_algebra_reproduces_an_updated_model():
    # models.py:355-525: the N2 x N2 update algebra of conditional_predict_* (host side of the native model), with the
    # device calls replaced by the oracle: must equal the posterior of a model that has seen the additional data
    from oracle import gp_oracle as o
    from trieste_b200.data import Dataset
    from trieste_b200.models import GaussianProcessRegression as G

    class OracleBacked:
        def __init__(self, om):
            self.om, self._dtype = om, np.float64
            self._spec = type("S", (), {"noise_variance": om.noise})()

        def _check_dim(self, x):
            pass

        def predict(self, x):
            return o.predict(self.om, np.asarray(x))

        def predict_joint(self, x):
            return o.predict_joint(self.om, np.asarray(x))

        def covariance_between_points(self, a, b):
            return o.covariance_between_points(self.om, np.asarray(a), np.asarray(b))

        _conditional_parts = G._conditional_parts
        conditional_predict_f = G.conditional_predict_f
        conditional_predict_joint = G.conditional_predict_joint
        conditional_predict_y = G.conditional_predict_y

    for n2 in (1, 3, 40):
        full = o.synthetic_model(o.hartmann_6, 80 + n2, 6)
        head = o.build_model("matern52", full.X[:80], full.y[:80], full.variance, full.lengthscales, full.noise, full.mean_const)
        f = OracleBacked(head)
        Xq = np.random.default_rng(1).uniform(size=(33, 6))
        add = Dataset(full.X[80:], full.y[80:])
        mean, var = f.conditional_predict_f(Xq, add)
        omean, ovar = o.predict_f(full, Xq)
        np.testing.assert_allclose(mean, omean, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var, ovar, rtol=0, atol=1e-12 * full.variance)
        _, vy = f.conditional_predict_y(Xq, add)
        np.testing.assert_allclose(vy, ovar + full.noise, rtol=0, atol=1e-12 * full.variance)
        mj, cj = f.conditional_predict_joint(Xq, add)
        np.testing.assert_allclose(cj[0], o.predict_f(full, Xq, full_cov=True)[1], rtol=0, atol=1e-12 * full.variance)
        m2, v2 = f.conditional_predict_f(Xq, Dataset(np.stack([full.X[80:]] * 2), np.stack([full.y[80:], full.y[80:] + 1.0])))
        assert m2.shape == (2, 33, 1)
        np.testing.assert_allclose(m2[0], mean, rtol=1e-12)
        np.testing.assert_allclose(v2[1], var, rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError):
        f.conditional_predict_f(Xq[None], add)


def test_qmc_base_samples_and_skip_counter():
    # models/gpflow/sampler.py:53-79, 90-117, 239-254: Sobol base samples through the normal quantile; successive samplers with
    # qmc_skip=True take disjoint blocks of the sequence; no point is the origin (its quantile would be -inf)
    from trieste_b200.sampler import (BatchReparametrizationSampler, IndependentReparametrizationSampler,
                                      qmc_normal_samples)

    x = qmc_normal_samples(1024, 3)
    assert x.shape == (1024, 3) and np.all(np.isfinite(x))
    assert np.abs(x.mean(0)).max() < 0.01 and np.abs(x.std(0) - 1).max() < 0.01  # far tighter than 1024 random draws
    np.testing.assert_array_equal(qmc_normal_samples(8, 2)[4:], qmc_normal_samples(4, 2, skip=4))
    assert qmc_normal_samples(0, 3).shape == (0, 3) and qmc_normal_samples(5, 0).shape == (5, 0)

    class Model:  # predict / predict_joint of a unit Gaussian: the samplers' host arithmetic is what is under test
        def predict(self, x):
            return np.zeros(x.shape[:-1] + (1,)), np.ones(x.shape[:-1] + (1,))

        def predict_joint(self, x):
            q = x.shape[-2]
            return np.zeros(x.shape[:-1] + (1,)), np.broadcast_to(np.eye(q), x.shape[:-2] + (1, q, q)).copy()

    IndependentReparametrizationSampler.skip = 0
    s1 = IndependentReparametrizationSampler(16, Model(), qmc=True)
    a = s1.sample(np.zeros((1, 1, 2)), jitter=0.0)[0, :, 0, 0]
    s2 = IndependentReparametrizationSampler(16, Model(), qmc=True)
    b = s2.sample(np.zeros((1, 1, 2)), jitter=0.0)[0, :, 0, 0]
    assert IndependentReparametrizationSampler.skip == 32
    np.testing.assert_allclose(np.concatenate([a, b]), qmc_normal_samples(32, 1)[:, 0])
    s3 = IndependentReparametrizationSampler(16, Model(), qmc=True, qmc_skip=False)
    np.testing.assert_allclose(s3.sample(np.zeros((1, 1, 2)), jitter=0.0)[0, :, 0, 0], a)
    np.testing.assert_allclose(s1.sample(np.zeros((1, 1, 2)), jitter=0.0)[0, :, 0, 0], a)  # fixed until reset
    sb = BatchReparametrizationSampler(8, Model(), qmc=True)
    eps = sb._get_eps(3)
    assert eps.shape == (3, 8) and IndependentReparametrizationSampler.skip == 40
    np.testing.assert_allclose(eps.T, qmc_normal_samples(8, 3, skip=32))


def test_rank_m_append_algebra_restated():
    """The O(m N²) cache extension of ``tb_gp_append_data`` (csrc/factor.cuh ``append_*`` / ``kinv_grow_kernel``; the reference
    refactorises instead, models/gpflow/models.py:171-186 -> interface.py:108-112), restated in NumPy against a from-scratch
    factorisation: new rows of L (Y = Linv0 B), Schur-complement Cholesky of the m x m corner, new rows of Linv
    (-L22^-1 Y^T Linv0), alpha by two triangular products, and the bordered-inverse growth of the dense K^-1."""
    import scipy.linalg as sl
    from oracle import gp_oracle as o

    rng = np.random.default_rng(3)
    N0, m, D = 200, 5, 4
    X = rng.uniform(size=(N0 + m, D))
    y = np.sin(3 * X.sum(-1))
    var, ls, noise = 0.7, np.full(D, 0.5), 0.01
    K = o.kernel_matrix("matern52", X, X, var, ls) + noise * np.eye(N0 + m)
    L_full = np.linalg.cholesky(K)
    Linv_full = sl.solve_triangular(L_full, np.eye(N0 + m), lower=True)
    L0 = np.linalg.cholesky(K[:N0, :N0])
    Linv0 = sl.solve_triangular(L0, np.eye(N0), lower=True)
    B, C = K[:N0, N0:], K[N0:, N0:]
    Y = Linv0 @ B                                   # new rows of L, transposed
    L22 = np.linalg.cholesky(C - Y.T @ Y)           # Schur complement
    L = np.block([[L0, np.zeros((N0, m))], [Y.T, L22]])
    L22inv = sl.solve_triangular(L22, np.eye(m), lower=True)
    Linv = np.block([[Linv0, np.zeros((N0, m))], [-L22inv @ Y.T @ Linv0, L22inv]])
    np.testing.assert_allclose(L, L_full, atol=1e-12)
    np.testing.assert_allclose(Linv, Linv_full, atol=1e-10)
    err = y - y.mean()
    np.testing.assert_allclose(Linv.T @ (Linv @ err), np.linalg.solve(K, err), rtol=1e-9, atol=1e-9)  # alpha
    # dense K^-1: bordered inverse with W = K0^-1 B and the inverse Schur complement
    Kinv0 = Linv0.T @ Linv0
    W, Sinv = Kinv0 @ B, L22inv.T @ L22inv
    Kinv = np.block([[Kinv0 + W @ Sinv @ W.T, -W @ Sinv], [-Sinv @ W.T, Sinv]])
    np.testing.assert_allclose(Kinv, np.linalg.inv(K), rtol=1e-8, atol=1e-8)


def test_acquisition_builders_reject_what_the_reference_rejects():
    """Argument and data checks of the builders, as the reference's own tests state them (tests/unit/acquisition/function/
    test_function.py: *_raises_for_empty_data, *_raises_for_invalid_*, test_entropy.py: *_raises_for_invalid_init_params);
    every check fires before the native library is touched, so they run without a GPU."""
    import trieste_b200 as tb
    from trieste_b200.acquisition import (AugmentedExpectedImprovement, BatchMonteCarloExpectedImprovement, ExpectedImprovement,
                                          LogExpectedImprovement, MinValueEntropySearch, MonteCarloExpectedImprovement,
                                          NegativeLowerConfidenceBound, ProbabilityOfFeasibility, ProbabilityOfImprovement)
    from trieste_b200.acquisition.function import multiple_optimism_lower_confidence_bound
    from trieste_b200.acquisition.sampler import ThompsonSamplerFromTrajectory

    class NotAModel:  # never reached by the checks below
        pass

    empty = tb.Dataset(np.zeros((0, 2)), np.zeros((0, 1)))
    space = tb.Box([0.0, 0.0], [1.0, 1.0])
    for builder in (ExpectedImprovement(), LogExpectedImprovement(), AugmentedExpectedImprovement(), ProbabilityOfImprovement(),
                    MinValueEntropySearch(space), BatchMonteCarloExpectedImprovement(10), MonteCarloExpectedImprovement(10)):
        mc = isinstance(builder, MonteCarloExpectedImprovement)
        for data in (empty, None):  # function.py:136-137, 1113-1115 "Dataset must be populated."; MonteCarloExpectedImprovement
            # checks the model's reparam_sampler first, as the reference does (function.py:825-829), also a ValueError
            with pytest.raises(ValueError, match="reparam_sampler" if mc else "populated"):
                builder.prepare_acquisition_function(NotAModel(), dataset=data)
            with pytest.raises(ValueError):
                builder.update_acquisition_function(object(), NotAModel(), dataset=data)
    with pytest.raises(ValueError):
        NegativeLowerConfidenceBound(-0.1)  # function.py:346-349
    assert "1.96" in repr(NegativeLowerConfidenceBound())
    for bad in (0, -5):
        with pytest.raises(ValueError):
            BatchMonteCarloExpectedImprovement(bad)  # function.py:1088-1090
        with pytest.raises(ValueError):
            MonteCarloExpectedImprovement(bad)
        with pytest.raises(ValueError):
            MinValueEntropySearch(space, num_samples=bad)  # entropy.py:96-99
        with pytest.raises(ValueError):
            MinValueEntropySearch(space, grid_size=bad)
    with pytest.raises(ValueError):
        BatchMonteCarloExpectedImprovement(10, jitter=-1e-6)  # function.py:1092-1094
    with pytest.raises(ValueError):
        MonteCarloExpectedImprovement(10, jitter=-1e-6)
    with pytest.raises(ValueError):
        MinValueEntropySearch(space, min_value_sampler=ThompsonSamplerFromTrajectory(sample_min_value=False))  # entropy.py:113-118
    with pytest.raises(ValueError):
        ProbabilityOfFeasibility(np.array([0.5, 0.6]))  # threshold must be a scalar, function.py:447
    with pytest.raises(ValueError):
        multiple_optimism_lower_confidence_bound(NotAModel(), 0)  # search_space_dim must be positive, function.py:1872
    # a native function is required: the fused kernels have no generic-model path and say so
    with pytest.raises(ValueError, match="GaussianProcessRegression"):
        ExpectedImprovement().prepare_acquisition_function(NotAModel(), dataset=tb.Dataset(np.zeros((3, 2)), np.zeros((3, 1))))


def test_optimizer_factories_reject_what_the_reference_rejects():
    """tests/unit/acquisition/test_optimizer.py: test_generate_continuous_optimizer_raises_with_invalid_init_params,
    test_generate_random_search_optimizer_raises_with_invalid_sample_size, test_batchify_*_raises_with_invalid_batch_size,
    test_sample_from_space_raises (optimizer.py:215-221, 386-404, 914-916, 956-958, 984-986)."""
    from trieste_b200.acquisition.optimizer import (batchify_joint, batchify_vectorize, generate_continuous_optimizer,
                                                    generate_random_search_optimizer, sample_from_space)

    for kwargs in (dict(num_initial_samples=0), dict(num_initial_samples=-5), dict(num_optimization_runs=0),
                   dict(num_optimization_runs=-1), dict(num_initial_samples=5, num_optimization_runs=6), dict(num_recovery_runs=-1)):
        with pytest.raises(ValueError):
            generate_continuous_optimizer(**kwargs)
    generate_continuous_optimizer(num_recovery_runs=0)  # zero recovery runs is allowed
    for bad in (0, -3):
        with pytest.raises(ValueError):
            generate_random_search_optimizer(bad)
        with pytest.raises(ValueError):
            sample_from_space(bad)
        with pytest.raises(ValueError):
            sample_from_space(10, batch_size=bad)
        with pytest.raises(ValueError):
            batchify_joint(generate_random_search_optimizer(10), bad)
        with pytest.raises(ValueError):
            batchify_vectorize(generate_random_search_optimizer(10), bad)
    space = Box([0.0], [1.0])
    quad = lambda x: -((np.asarray(x) - 0.3) ** 2).sum(-1)  # noqa: E731  [..., 1, D] -> [..., 1]
    with pytest.raises(ValueError):  # an already vectorised function cannot be vectorised again (optimizer.py:962-966)
        batchify_vectorize(generate_random_search_optimizer(10), 2)(space, (quad, 2))
    with pytest.raises(ValueError):  # joint batches of a vectorised function are not defined (optimizer.py:921-925)
        batchify_joint(generate_random_search_optimizer(10), 2)(space, (quad, 2))


def test_rules_and_samplers_reject_what_the_reference_rejects():
    """tests/unit/acquisition/test_rule.py (test_discrete_thompson_sampling_raises_for_invalid_init_params,
    test_efficient_global_optimization_raises_for_no_query_points, ..._no_batch_fn_with_many_query_points) and
    tests/unit/models/gpflow/test_sampler.py (*_sampler_raises_for_invalid_sample_size, ..._for_negative_jitter,
    batch sampler needs predict_joint)."""
    from trieste_b200.acquisition.sampler import ThompsonSamplerFromTrajectory
    from trieste_b200.rule import DiscreteThompsonSampling, EfficientGlobalOptimization
    from trieste_b200.sampler import BatchReparametrizationSampler, IndependentReparametrizationSampler

    for bad in (0, -2):
        with pytest.raises(ValueError):
            EfficientGlobalOptimization(num_query_points=bad)  # rule.py:262-265
        with pytest.raises(ValueError):
            DiscreteThompsonSampling(bad, 1)  # rule.py:926-931
        with pytest.raises(ValueError):
            DiscreteThompsonSampling(100, bad)  # rule.py:933-938
        with pytest.raises(ValueError):
            IndependentReparametrizationSampler(bad, object())  # sampler.py:100-101
        with pytest.raises(ValueError):
            BatchReparametrizationSampler(bad, object())  # sampler.py:181-182
    with pytest.raises(ValueError):
        EfficientGlobalOptimization(num_query_points=3)  # no batch builder given (rule.py:267-275)
    with pytest.raises(ValueError):  # a sampler of minimum VALUES cannot pick query points (rule.py:940-946)
        DiscreteThompsonSampling(100, 2, thompson_sampler=ThompsonSamplerFromTrajectory(sample_min_value=True))
    with pytest.raises(ValueError):  # sampler.py:184-188: the batch sampler needs predict_joint
        BatchReparametrizationSampler(10, object())
    assert "EfficientGlobalOptimization(" in repr(EfficientGlobalOptimization())
