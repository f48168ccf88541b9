"""Development aid: time both engines on the headline shape."""
import sys, time, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.build()
from oracle import gp_oracle as o
from tests.util import native_from_oracle
from trieste_b200 import _lib
from trieste_b200.acquisition import expected_improvement

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
D = int(sys.argv[2]) if len(sys.argv) > 2 else 10
M = int(sys.argv[3]) if len(sys.argv) > 3 else 37888 * 8
om = o.synthetic_model(o.ackley if D != 6 else o.hartmann_6, N, D)
nm = native_from_oracle(om)
fn = expected_improvement(nm, o.ei_eta(om))
x = torch.rand(M, 1, D, dtype=torch.float64, device="cuda")
for eng in ["fp64", "int8", "int8"]:
    nm.set_engine(eng)
    _lib.lib().tb_gp_profile(nm.handle, 1)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        v = fn(x)
        torch.cuda.synchronize(); dt = time.time() - t0
    ms = C.c_double(); nl = C.c_int64(); fl = C.c_double()
    _lib.lib().tb_gp_profile_read(nm.handle, C.byref(ms), C.byref(nl), C.byref(fl))
    print(f"{eng}: N={N} M={M}: {dt*1e3:.1f} ms  {M/dt:.3e} cand/s; gemm {ms.value/3:.1f} ms/pass = {fl.value/ms.value*1e-9:.1f} TFLOP/s fp64-equivalent")
    if eng == "fp64": ref = v.clone()
print("max |EI_int8 - EI_fp64| =", float((v - ref).abs().max()), " max EI", float(ref.max()))
