import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def relerr(a, b, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


@pytest.fixture(scope="session")
def golden_dense():
    return load_golden("dense_nodes")


@pytest.fixture(scope="session")
def golden_vecchia():
    return load_golden("vecchia_nodes")


@pytest.fixture(scope="session")
def golden_jd():
    return load_golden("jd")


@pytest.fixture(scope="session")
def golden_ess():
    return load_golden("ess_replay")


@pytest.fixture(scope="session")
def golden_e2e():
    return load_golden("e2e")


@pytest.fixture(scope="session")
def golden_loo():
    return load_golden("loo")


@pytest.fixture(scope="session")
def golden_metric():
    return load_golden("metric")


@pytest.fixture(scope="session")
def golden_update():
    return load_golden("update")


@pytest.fixture(scope="session")
def golden_lik():
    return load_golden("likelihood")
