"""GPU, world size 2 over NCCL with the REAL kernels (round-1 verdict item 3): the package's sharding helpers
(`trieste_b200.parallel`) produce on every rank exactly what one GPU produces over the whole candidate set.
Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import trieste_b200 as tb
        from oracle import gp_oracle as o  # only for the objective / data generator (checker side)
        from trieste_b200.acquisition import ExpectedImprovement, LogExpectedImprovement
        from trieste_b200.parallel import sharded_argmax, sharded_multistart, sharded_thompson_argmin, sharded_topk
        from trieste_b200.sampler import top_k
        from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler

        om = o.synthetic_model(o.hartmann_6, 512, 6)
        spec = tb.GPRSpec((om.X, om.y), tb.Matern52(om.variance, om.lengthscales), tb.Constant(om.mean_const), om.noise)
        model = tb.GaussianProcessRegression(spec, device=rank)
        ds = tb.Dataset(om.X, om.y)
        pts = np.random.default_rng(5).uniform(size=(200_001, 6))  # identical on every rank, odd count: uneven shards
        fn = ExpectedImprovement().prepare_acquisition_function(model, ds)
        pt, bv, bi = sharded_argmax(fn, pts)
        idx1, val1 = fn.fused_argmax(pts)  # one GPU over everything
        # batch Thompson sampling: the same 4 trajectories on every rank (same seed)
        traj = RandomFourierFeatureTrajectorySampler(model, 512, seed=3).get_trajectory()
        traj._batch_size = 4
        traj.resample()
        traj._initialized = True
        tp, tvals, tidx = sharded_thompson_argmin(traj, pts)
        mv1, mi1 = traj.argmin_over(pts)
        # multi-start optimisation sharded over the ranks
        lfn = LogExpectedImprovement().prepare_acquisition_function(model, ds)
        starts = pts[:64]

        def optimise(s):
            ok, f, xs, nfev = lfn.maximize_from(s, np.zeros(6), np.ones(6), maxiter=20)
            return xs, f

        ms_pt, ms_v, ms_i = sharded_multistart(optimise, starts)
        xs_all, f_all = optimise(starts)
        # running top-k of generate_initial_points, sharded: local tb_topk + one all-gather of k tuples per rank
        tk_pts, tk_v, tk_i = sharded_topk(fn, pts, 16)
        tv1, ti1 = top_k(np.ascontiguousarray(np.asarray(fn(pts[:, None, :])).reshape(-1)), 16, device=rank)
        q.put((rank, int(bi), float(bv), pt.tolist(), int(idx1), float(val1), tidx.tolist(), tvals.tolist(), mi1.tolist(), mv1.tolist(),
               int(ms_i), float(ms_v), int(np.argmax(f_all)), float(f_all.max()),
               tk_i.tolist(), tk_v.tolist(), tk_pts.tolist(), np.asarray(ti1).tolist(), np.asarray(tv1).tolist()))
    except BaseException as exc:  # report instead of leaving the peer rank waiting in a collective until the timeout
        q.put(("error", rank, repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_helpers_over_nccl_world2():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = []
    for _ in procs:
        item = q.get(timeout=600)
        if item[0] == "error":  # a rank failed: its peer may be blocked in a collective — stop both, fail now
            for p in procs:
                p.terminate()
            pytest.fail(f"rank {item[1]} failed: {item[2]}")
        res.append(item)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pts = np.random.default_rng(5).uniform(size=(200_001, 6))
    for (rank, bi, bv, pt, idx1, val1, tidx, tvals, mi1, mv1, ms_i, ms_v, best_i, best_v, tk_i, tk_v, tk_pts, ti1, tv1) in res:
        assert tk_i == ti1 and tk_v == tv1  # the sharded top-k IS the single-GPU tb_topk over everything
        np.testing.assert_allclose(tk_pts, pts[tk_i])
        assert bi == idx1 and bv == val1  # the sharded winner IS the single-GPU first-max winner
        np.testing.assert_allclose(pt[0], pts[bi])
        assert tidx == mi1
        np.testing.assert_allclose(tvals, mv1, rtol=1e-12)
        assert ms_i == best_i and abs(ms_v - best_v) <= 1e-9 * max(1.0, abs(best_v))
    assert res[0][1:] == res[1][1:]  # identical on both ranks
