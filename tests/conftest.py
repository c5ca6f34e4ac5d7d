import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "psi_cases.npz"))


@pytest.fixture(scope="session")
def cuda_solver_lib():
    """Build (if needed) and load the CUDA library; GPU tests fail loudly without it."""
    from dyobav_mpcnwta_warehouse_b200.csrc import build
    from dyobav_mpcnwta_warehouse_b200 import _lib
    build.build()
    return _lib.load()
