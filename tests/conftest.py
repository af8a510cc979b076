import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_libs():
    """Build librthost.so / librtgpu.so / liboracle.so once per session if they are missing."""
    import __graft_entry__ as g
    g.ensure_built()
    return True
