"""Development aid: time the posterior-cache precompute (hand-written vs cuSOLVER/cuBLAS path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import trieste_b200 as tb
from trieste_b200.objectives import ackley

for N in (1024, 4096, 8192):
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(N, 10)); y = ackley(X)
    for mode in ("own", "cusolver"):
        os.environ["TB_FACTOR"] = mode
        m = tb.GaussianProcessRegression(tb.build_gpr(tb.Dataset(X, y), tb.Box([0.0] * 10, [1.0] * 10)))
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); m.update_posterior_cache(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print(f"N={N} {mode}: update_posterior_cache {min(ts)*1e3:.1f} ms")
        del m

# rank-m append (tb_gp_append_data) against the full refresh above
os.environ["TB_FACTOR"] = "own"
for N in (1024, 4096, 8192):
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(N + 50, 10)); y = ackley(X)
    spec = tb.build_gpr(tb.Dataset(X[:N], y[:N]), tb.Box([0.0] * 10, [1.0] * 10))
    m = tb.GaussianProcessRegression(spec)
    for step, add in ((1, 1), (2, 1), (3, 1), (4, 1), (5, 8), (6, 30)):
        n1 = m.get_internal_data().query_points.shape[0] + add
        torch.cuda.synchronize()
        t0 = time.perf_counter(); m.update(tb.Dataset(X[:n1], y[:n1])); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"N={n1 - add} append {add}: {dt*1e3:.2f} ms (appended={m.last_update_appended})")
    del m
