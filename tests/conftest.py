import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def engine():
    """One engine context on cuda:0 for the whole GPU test session (fails loudly if the
    CUDA library is missing: there is no CPU fallback)."""
    import pem_spgemm_b200 as pem
    ctx = pem.Context(0)
    yield ctx
    ctx.close()
