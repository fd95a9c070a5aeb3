import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference or the prebuilt oracle/_ref library")


@pytest.fixture(scope="session")
def built():
    """Builds (if needed) the C oracle, the host library and the CUDA library once per session."""
    import __graft_entry__ as ge
    ge.build(quiet=True)
    return True
