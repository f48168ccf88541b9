"""Per-round trace of the device multi-start L-BFGS (TB_LBFGS_TRACE=1) on the C5 shape: python tools/lbfgs_trace.py [starts]"""
import os, sys, math, time
os.environ["TB_LBFGS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import trieste_b200 as tb
from trieste_b200.acquisition import LogExpectedImprovement

def rff_obj(x, terms=64, seed=2):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((terms, x.shape[-1])) * 3.0
    ph = rng.uniform(0, 2 * math.pi, terms)
    a = rng.standard_normal(terms) / math.sqrt(terms)
    return (np.cos(x @ w.T + ph) * a).sum(-1, keepdims=True)

R = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500
rng = np.random.default_rng(0)
X = rng.uniform(size=(8192, 20)).astype(np.float32)
y = rff_obj(X.astype(np.float64)).astype(np.float32)
ds = tb.Dataset(X, y)
m = tb.GaussianProcessRegression(tb.build_gpr(ds, tb.Box([0.0] * 20, [1.0] * 20)))
fn = LogExpectedImprovement().prepare_acquisition_function(m, ds)
x0 = np.random.default_rng(10).uniform(size=(R, 20))
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ok, f, xs, nfev = fn.maximize_from(x0, np.zeros(20), np.ones(20), maxiter=30)
    torch.cuda.synchronize(); print("wall ms", 1e3 * (time.perf_counter() - t0), "mean nfev", float(np.mean(nfev)), file=sys.stderr)
