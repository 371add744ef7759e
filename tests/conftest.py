import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_cases.npz")))


@pytest.fixture(scope="session")
def designs():
    from ccgp_b200 import workloads
    return dict(np.load(workloads.DESIGNS_PATH))


@pytest.fixture(scope="session")
def engine():
    import ccgp_b200
    eng = ccgp_b200.Engine(0)   # raises loudly if libccgp.so or the GPU is missing
    yield eng
    eng.close()


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1.0)
