import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session", autouse=True)
def _built_extension():
    """(Re)build the in-tree C-ABI library when it is missing or older than its sources (no-op otherwise)."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def native_lib(_built_extension):
    from trieste_b200 import _lib

    return _lib.lib()
