import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (oracle, and oracle/_ref where /root/reference exists) and the product library."""
    from jackalope_b200 import build as jbuild
    jbuild.build()
    from oracle import harness
    harness.build()           # after the product: oracle/_ref/libjlp_glue.so links against it
    yield


@pytest.fixture(scope="session")
def ctx():
    import jackalope_b200 as J
    c = J.Context(0)
    yield c
    c.close()
