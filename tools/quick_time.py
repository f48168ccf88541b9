"""Quick device-resident timing of the EI path (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from oracle import gp_oracle as o
from tests.util import native_from_oracle
from trieste_b200 import _lib
from trieste_b200.acquisition import expected_improvement
import ctypes as C

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 10
M = int(sys.argv[3]) if len(sys.argv) > 3 else 37888 * 4
t0 = time.time()
om = o.synthetic_model(o.ackley if D != 6 else o.hartmann_6, N, D)
print("oracle model build %.2fs" % (time.time() - t0))
t0 = time.time()
nm = native_from_oracle(om)
print("native model build (incl. factorize) %.2fs" % (time.time() - t0))
fn = expected_improvement(nm, o.ei_eta(om))
x = torch.rand(M, 1, D, dtype=torch.float64, device="cuda")
_lib.lib().tb_gp_profile(nm.handle, 1)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    v = fn(x)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"N={N} D={D} M={M}: {dt*1e3:.1f} ms  {M/dt:.3e} cand/s  ({M/dt*N*N/1e12:.2f} TFLOP/s fp64 equiv)")
ms = C.c_double(); nl = C.c_int64(); fl = C.c_double()
_lib.lib().tb_gp_profile_read(nm.handle, C.byref(ms), C.byref(nl), C.byref(fl))
print(f"trigemm: {nl.value} launches, {ms.value:.1f} ms total, {fl.value/ms.value*1e-9:.2f} TFLOP/s")
