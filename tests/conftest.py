import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries; built on demand so that `pytest` works from a clean checkout."""
    import __graft_entry__ as g
    from cupss_b200 import capi
    need = [capi.ENGINE_LIB, capi.PRODUCT_LIB]
    import cases
    if os.path.isdir("/root/reference/src"):
        need += [cases.ORACLE_F, cases.ORACLE_U, cases.SHIM]
    if not all(os.path.exists(p) for p in need):
        g.build()
    return True
