"""Development aid: per-dimension behaviour of the device-side L-BFGS against the host loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
import trieste_b200 as tb
from trieste_b200.acquisition import NegativeLowerConfidenceBound
from trieste_b200.acquisition.optimizer import _perform_parallel_continuous_optimization

for D in (1, 3, 8, 20, 32):
    rng = np.random.default_rng(D)
    X = rng.uniform(size=(80, D)); y = X.sum(axis=1, keepdims=True)
    space = tb.Box([0.0] * D, [1.0] * D)
    ds = tb.Dataset(X, y)
    nm = tb.GaussianProcessRegression(tb.build_gpr(ds, space, likelihood_variance=1e-3))
    fn = NegativeLowerConfidenceBound(0.5).prepare_acquisition_function(nm, ds)
    x0 = rng.uniform(size=(40, D))
    for cap in (50, 400):
        t0 = time.perf_counter()
        ok, val, x, nfev = fn.maximize_from(x0, space.lower, space.upper, maxiter=cap)
        dt = time.perf_counter() - t0
        _, grad = fn.value_and_gradient(x[:, None, :])
        pg = np.abs(x - np.clip(x + grad[:, 0, :], 0, 1)).max()
        print(f"D={D} device maxiter={cap}: {dt:.2f}s ok={ok.mean():.2f} nfev max={nfev.max()} med={np.median(nfev)} best={val.max():.6f} pg={pg:.2e} finite_grad={np.isfinite(grad).all()}", flush=True)
    os.environ["TB_LBFGS"] = "host"
    t0 = time.perf_counter()
    ok, val, x, nfev = _perform_parallel_continuous_optimization(fn, space.lower, space.upper, x0[:, None, :], {"maxiter": 400})
    print(f"D={D} host   maxiter=400: {time.perf_counter()-t0:.2f}s ok={ok.mean():.2f} nfev max={nfev.max()} med={np.median(nfev)} best={val.max():.6f}", flush=True)
    del os.environ["TB_LBFGS"]
