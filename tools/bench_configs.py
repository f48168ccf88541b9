"""Throughput of the BASELINE.json configs C2..C5 on one B200 (development / documentation aid; bench.py is the
contract benchmark of the headline metric).  Prints one JSON line per config.

    python tools/bench_configs.py [c2] [c3] [c4] [c5]
"""
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import trieste_b200 as tb
from trieste_b200.acquisition import BatchMonteCarloExpectedImprovement, ExpectedImprovement, LogExpectedImprovement
from trieste_b200.objectives import ackley, hartmann_6
from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def model(obj, N, D, dtype=np.float64, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(N, D)).astype(dtype)
    y = obj(X).astype(dtype)
    space = tb.Box([0.0] * D, [1.0] * D)
    ds = tb.Dataset(X, y)
    return tb.GaussianProcessRegression(tb.build_gpr(ds, space)), ds, space


def c2():
    m, ds, space = model(hartmann_6, 1024, 6)
    fn = ExpectedImprovement().prepare_acquisition_function(m, ds)
    M = 1_000_000
    xd = torch.rand(M, 6, dtype=torch.float64, device="cuda")
    xh = xd.cpu().numpy()
    td = timed(lambda: fn.fused_argmax(xd))
    th = timed(lambda: fn.fused_argmax(xh))
    return {"config": "C2 Hartmann6 GPR N=1024 fp64 EI random search over 1e6 candidates", "engine": m.engine,
            "device_resident_cand_per_s": M / td, "host_buffers_cand_per_s": M / th, "ms_device": td * 1e3}


def c3():
    m, ds, space = model(ackley, 4096, 10)
    q, S, B = 8, 512, 65536
    fn = BatchMonteCarloExpectedImprovement(S).prepare_acquisition_function(m, ds)
    fn._sampler.set_eps(np.random.default_rng(3).standard_normal((q, S)))
    xd = torch.rand(B, q, 10, dtype=torch.float64, device="cuda")
    td = timed(lambda: fn(xd), reps=2)
    Bg = 8192
    xg = xd[:Bg]
    tg = timed(lambda: fn.value_and_gradient(xg), reps=2)
    return {"config": "C3 Ackley-10 GPR N=4096 fp64 BatchMonteCarloExpectedImprovement q=8 S=512, 65536 q-batches",
            "batches_per_s": B / td, "points_per_s": B * q / td, "ms": td * 1e3,
            "value_and_gradient_batches_per_s": Bg / tg, "value_and_gradient_points_per_s": Bg * q / tg, "ms_grad": tg * 1e3}


def c4():
    m, ds, space = model(hartmann_6, 1024, 6)
    s = RandomFourierFeatureTrajectorySampler(m, 2048, seed=0)
    traj = s.get_trajectory()
    M = 1_250_000  # one GPU's shard of the 1e7 candidates of config 4
    xd = torch.rand(M, 6, dtype=torch.float64, device="cuda")
    td = timed(lambda: traj.argmin_over(xd))
    return {"config": "C4 Hartmann6 RFF F=2048 Thompson argmin over a 1.25e6-candidate shard (1e7 / 8 GPUs)",
            "cand_per_s": M / td, "ms": td * 1e3}


def c5():
    def rff_obj(x, terms=64, seed=2):
        rng = np.random.default_rng(seed)
        w = rng.standard_normal((terms, x.shape[-1])) * 3.0
        ph = rng.uniform(0, 2 * math.pi, terms)
        a = rng.standard_normal(terms) / math.sqrt(terms)
        return (np.cos(x @ w.T + ph) * a).sum(-1, keepdims=True)

    m, ds, space = model(rff_obj, 8192, 20, dtype=np.float32)
    fn = LogExpectedImprovement().prepare_acquisition_function(m, ds)
    M = 200_000
    xd = torch.rand(M, 1, 20, dtype=torch.float32, device="cuda")
    tf = timed(lambda: fn(xd), reps=2)
    R = 12_500  # one GPU's shard of the 1e5 multi-starts
    xs = torch.rand(R, 1, 20, dtype=torch.float32, device="cuda")
    tg = timed(lambda: fn.value_and_gradient(xs), reps=2)
    # the whole multi-start optimisation (30 L-BFGS iterations per start): device-side bookkeeping vs the host (NumPy) loop
    from trieste_b200.acquisition.optimizer import _perform_parallel_continuous_optimization
    x0 = xs.cpu().numpy().astype(np.float64)
    lo, up = np.zeros(20), np.ones(20)
    args = {"maxiter": 30}
    t0 = time.perf_counter(); okd, fd, _, nd = fn.maximize_from(x0[:, 0, :], lo, up, maxiter=30); td = time.perf_counter() - t0
    os.environ["TB_LBFGS"] = "host"
    t0 = time.perf_counter(); okh, fh, _, nh = _perform_parallel_continuous_optimization(fn, lo, up, x0, args); th = time.perf_counter() - t0
    del os.environ["TB_LBFGS"]
    return {"config": "C5 Synthetic-20D GPR N=8192 fp32 I/O log-EI (int8 engine, 10-product fp32 mode; gradient V = K^-1 k* as a dense digit GEMM)", "engine": m.engine,
            "forward_cand_per_s": M / tf, "value_and_gradient_starts_per_s": R / tg, "ms_forward": tf * 1e3, "ms_grad": tg * 1e3,
            "multistart_30_iterations": {"starts": R, "device_lbfgs_s": td, "host_lbfgs_s": th, "device_evals_max": int(nd.max()),
                                         "host_evals_max": int(nh.max()), "device_best": float(fd.max()), "host_best": float(fh.max()),
                                         "device_starts_per_s": R / td, "host_starts_per_s": R / th}}


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c2", "c3", "c4", "c5")] or ["c2", "c3", "c4", "c5"]
    for name in which:
        print(json.dumps(globals()[name]()), flush=True)
