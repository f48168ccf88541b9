"""ncu driver: a few EI passes on the headline shape with the int8 engine."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
import trieste_b200 as tb
from trieste_b200.acquisition import expected_improvement
from trieste_b200.objectives import ackley
N, D, M = 4096, 10, 56832  # one chunk of the single-pass engine: 592 tiles x 96 candidates
rng = np.random.default_rng(0)
X = rng.uniform(size=(N, D)); y = ackley(X)
model = tb.GaussianProcessRegression(tb.build_gpr(tb.Dataset(X, y), tb.Box([0.0] * D, [1.0] * D)))
model.set_engine(sys.argv[1] if len(sys.argv) > 1 else "int8")
fn = expected_improvement(model, float(y.min()))
x = torch.rand(M, 1, D, dtype=torch.float64, device="cuda")
for _ in range(4):
    fn(x)
torch.cuda.synchronize()
