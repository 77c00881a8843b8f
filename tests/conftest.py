import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name), allow_pickle=False)


@pytest.fixture(scope="session")
def lib():
    """The product C-ABI library (loads without a GPU; compute calls need one)."""
    from contrastors_b200 import _lib
    return _lib.load()
