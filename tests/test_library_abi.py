"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls without a GPU), and the product path fails loudly without a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "trieste_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_all_exported(native_lib):
    from trieste_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(native_lib, name), f"{name} declared in include/trieste_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_version_and_error_string(native_lib):
    assert b"sm_100a" in native_lib.tb_version()
    assert isinstance(native_lib.tb_last_error(), bytes)


def test_fails_loudly_without_gpu(native_lib):
    from trieste_b200 import _lib

    if _lib.device_count() > 0:
        pytest.skip("a GPU is visible")
    import numpy as np

    import trieste_b200 as tb

    spec = tb.GPRSpec((np.zeros((3, 2)), np.zeros((3, 1))), tb.Matern52(1.0, [1.0, 1.0]), tb.Constant(0.0), 0.1)
    with pytest.raises(_lib.NativeLibraryError):
        tb.GaussianProcessRegression(spec)
