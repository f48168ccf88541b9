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


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under trieste_b200/ or integration/ may import, open or execute anything under
    oracle/ (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs do, as the checker)."""
    import ast
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    offenders = []
    for top in ("trieste_b200", "integration"):
        for dirpath, _, files in os.walk(os.path.join(root, top)):
            for f in files:
                path = os.path.join(dirpath, f)
                if f.endswith(".py"):
                    tree = ast.parse(open(path).read())
                    for node in ast.walk(tree):
                        names = []
                        if isinstance(node, ast.Import):
                            names = [a.name for a in node.names]
                        elif isinstance(node, ast.ImportFrom):
                            names = [node.module or ""]
                        if any(n == "oracle" or n.startswith("oracle.") for n in names):
                            offenders.append(path)
                elif f.endswith((".cu", ".cuh", ".h", ".cpp")):
                    if "oracle/" in open(path, errors="ignore").read():
                        offenders.append(path)
    assert not offenders, offenders
    # dynamic imports: no importlib / __import__ of the oracle either
    for dirpath, _, files in os.walk(os.path.join(root, "trieste_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src, f


def test_every_header_symbol_is_mapped_in_integration_md():
    """INTEGRATION.md's table names, for each C-ABI entry point, the reference interface it stands behind."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "trieste_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    symbols = sorted(set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", header)))
    assert len(symbols) >= 30
    missing = [s for s in symbols if s not in doc]
    assert not missing, missing
