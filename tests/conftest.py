import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dragon():
    from tetsim_b200 import mesh
    return mesh.load_dragon()


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test session (built in-tree by __graft_entry__.build)."""
    from tetsim_b200 import build
    build.build()
    import oracle
    oracle.build()
