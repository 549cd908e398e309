import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def solver():
    """One sdv handle for the whole GPU session (fails loudly when the CUDA library or the GPU is missing)."""
    from sadvio_b200 import api

    s = api.Solver()
    yield s
    s.close()
