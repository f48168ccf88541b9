"""The branch-free exp(-s) / sqrt / cos of csrc/fastmath.cuh are header-only and compile for the host: the same arithmetic
(magic-add range reduction, table + polynomial, coupled Newton step, Cody-Waite reduction + Taylor sine) is checked here against
long-double libm over the argument ranges the K* / RFF kernels use.  The device build differs only in the reciprocal-square-root
seed (MUFU.RSQ instead of 1/sqrtf) and in hardware FMA contraction, both inside the Newton step's convergence margin; the GPU
parity tests pin the device side."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_fastmath_error_bounds_on_the_host(tmp_path):
    exe = str(tmp_path / "fastmath_check")
    subprocess.run(["g++", "-O2", "-x", "c++", "-DFM_ITERS=2000000", "-o", exe, os.path.join(ROOT, "tools", "fastmath_check.cu")],
                   check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    # the harness exits non-zero above 4e-16 (exp, relative), 2.3e-16 (sqrt, relative), 1e-13 (cos, absolute)
    assert res.returncode == 0, res.stdout + res.stderr
    out = res.stdout
    assert "exp_neg(0) = 1," in out  # exact at the origin: k(x, x) = variance exactly
    worst_exp = float(out.split("exp_neg: max rel err ")[1].split()[0])
    worst_sqrt = float(out.split("sqrt_pos: max rel err ")[1].split()[0])
    worst_cos = float(out.split("cos_fast: max ABS err ")[1].split()[0])
    assert worst_exp < 4e-16 and worst_sqrt < 2.3e-16 and worst_cos < 6e-16
