"""ncu driver for the non-headline kernels (round-1 verdict item 4): RFF trajectory evaluation, the joint / qEI kernels, the EI
gradient assembly, the device L-BFGS step and the cache-build (factorisation) kernels, each at its BASELINE config size.

    ncu --set full --clock-control none -k regex:"rff_eval|joint_kernel|qei_backward|qei_mix|grad_kernel|lbfgs_step|chol_|trinv_|kinv_kernel|kdot" \
        -c 40 -o gpurun_out/prof_misc python tools/prof_misc.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import __graft_entry__ as g

g.build()
import trieste_b200 as tb
from trieste_b200.acquisition import BatchMonteCarloExpectedImprovement, ExpectedImprovement, LogExpectedImprovement
from trieste_b200.objectives import ackley, hartmann_6
from trieste_b200.sampler import RandomFourierFeatureTrajectorySampler


def model(obj, N, D, dtype=np.float64):
    rng = np.random.default_rng(0)
    X = rng.uniform(size=(N, D)).astype(dtype)
    y = obj(X.astype(np.float64)).astype(dtype)
    ds = tb.Dataset(X, y)
    return tb.GaussianProcessRegression(tb.build_gpr(ds, tb.Box([0.0] * D, [1.0] * D))), ds


# cache build at N = 4096 (chol_*, trinv_*, and kinv_kernel through the first gradient request)
m, ds = model(ackley, 4096, 10)
fn = ExpectedImprovement().prepare_acquisition_function(m, ds)
x = torch.rand(8192, 1, 10, dtype=torch.float64, device="cuda")
fn.value_and_gradient(x)  # grad_kernel + the dense V GEMM
# device L-BFGS: a few rounds over 2048 starts
fn.maximize_from(np.random.default_rng(1).uniform(size=(2048, 10)), np.zeros(10), np.ones(10), maxiter=5)
# C3: joint kernels + qEI value and gradient
q, S = 8, 512
qfn = BatchMonteCarloExpectedImprovement(S).prepare_acquisition_function(m, ds)
qfn._sampler.set_eps(np.random.default_rng(3).standard_normal((q, S)))
xb = torch.rand(8192, q, 10, dtype=torch.float64, device="cuda")
qfn(xb)
qfn.value_and_gradient(xb[:2048])
# C4: RFF trajectory evaluation, F = 2048
m6, ds6 = model(hartmann_6, 1024, 6)
traj = RandomFourierFeatureTrajectorySampler(m6, 2048, seed=0).get_trajectory()
xc = torch.rand(1_250_000, 6, dtype=torch.float64, device="cuda")
traj.argmin_over(xc)
# decoupled trajectory (kdot_kernel)
dtraj = m6.trajectory_sampler().get_trajectory()
dtraj(xc[:200_000, None, :])
torch.cuda.synchronize()
