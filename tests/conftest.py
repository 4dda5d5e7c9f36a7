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


# ---- parity against the reference's own noise floor (tests/golden/floor.npz, `make_golden.py floor`) ---------------
_FLOOR = {}
PARITY_ROWS = []   # (key, max error, max |reference|, floor, bound used, error / bound)


def floor_ok(key, got, ref, K=64.0, rel=1e-9):
    """True when max|got - ref| <= max(rel * max|ref|, K * floor[key]).
    `rel` is BASELINE.json's tolerance (1e-9 relative); where a quantity is so ill-conditioned that the UNMODIFIED
    reference moves by more than that when its BLAS thread count changes, when its inputs move by one ulp, or
    when it is simply run in another process (floor.npz records the largest of the three), the bound is K times
    that movement instead.  Every comparison is logged; DGPB_PARITY_REPORT=<path> writes the table."""
    if not _FLOOR:
        _FLOOR.update({k: v for k, v in load_golden("floor").items()})
    got, ref = np.asarray(got, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    err = float(np.max(np.abs(got - ref))) if got.size else 0.0
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    floor = float(np.max(_FLOOR[key]))
    bound = max(rel * scale, K * floor)
    PARITY_ROWS.append((key, err, scale, floor, bound, err / bound if bound > 0 else (0.0 if err == 0 else np.inf)))
    return err <= bound


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("DGPB_PARITY_REPORT")
    if not path or not PARITY_ROWS:
        return
    with open(path, "w") as f:
        f.write("| quantity | max abs error (GPU vs reference fixture) | max abs reference | reference-vs-reference floor | "
                "error / max abs reference | error / floor |\n|---|---|---|---|---|---|\n")
        for key, err, scale, floor, bound, ratio in PARITY_ROWS:
            f.write(f"| {key} | {err:.3e} | {scale:.3e} | {floor:.3e} | {err / scale if scale else 0:.2e} | "
                    f"{err / floor if floor else float('inf') if err else 0:.2f} |\n")


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
