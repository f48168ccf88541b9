"""ncu launch-list driver: one value+gradient call of the C5 shape (N=8192, D=20, fp32 I/O, 12 500 starts) and one forward call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math
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

rng = np.random.default_rng(0)
X = rng.uniform(size=(8192, 20)).astype(np.float32)
y = rff_obj(X.astype(np.float64)).astype(np.float32)
ds = tb.Dataset(X, y)
m = tb.GaussianProcessRegression(tb.build_gpr(ds, tb.Box([0.0] * 20, [1.0] * 20)))
print("engine_info", m.engine_info())
fn = LogExpectedImprovement().prepare_acquisition_function(m, ds)
xs = torch.rand(12_500, 1, 20, dtype=torch.float32, device="cuda")
for _ in range(2):
    fn.value_and_gradient(xs)
xf = torch.rand(200_000, 1, 20, dtype=torch.float32, device="cuda")
fn(xf)
torch.cuda.synchronize()
