import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vt_ctx():
    """One CUDA context for the whole GPU session. Fails loudly (no skip) when the library or the device is missing."""
    import voxeltoy_b200 as vt
    ctx = vt.Context(0)
    yield ctx
    ctx.close()
