import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_ctx():
    """One lg_ctx for the whole GPU session.  No fallback: a missing library or GPU is an error."""
    from ligero_b200 import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
