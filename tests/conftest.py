import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cvb():
    """The product library bound to cuda:0.  GPU tests fail (not skip) if it cannot initialise."""
    import compv_b200
    compv_b200.init(0)
    return compv_b200
