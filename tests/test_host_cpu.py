"""CPU-only checks of the boundary and the host logic (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dgp_b200 import _lib

    header = open(os.path.join(ROOT, "include", "dgpb.h")).read()
    declared = set(re.findall(r"\b(dgpb_[a-z_A-Z0-9]+)\s*\(", header))
    declared -= {"dgpb_ws", "dgpb_node"}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dgpb.h but not exported by libdgpb.so"
    assert set(_lib.exported_symbols()) == declared
    assert lib.dgpb_sizeof_node() == ctypes.sizeof(_lib.DgpbNode)
    assert lib.dgpb_version() >= 100


def test_no_cpu_fallback_without_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import dgp_b200 as D

    k = D.kernel(length=np.array([1.0]))
    k.input, k.output = np.random.rand(5, 1), np.random.rand(5, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        k.k_matrix()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dgp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_priors_and_bookkeeping_match_oracle():
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    for prior in ("ga", "inv_ga"):
        k = D.kernel(length=np.array([0.7, 1.9]), nugget=3e-3, nugget_est=True, prior_name=prior)
        lp = k.log_prior()
        lpf = k.log_prior_fod()
        assert np.isclose(lp, O.log_prior(k.length, k.nugget, prior, k.prior_coef, True))
        assert np.allclose(lpf, O.log_prior_fod(k.length, k.nugget, prior, k.prior_coef, True))
    k = D.kernel(length=np.array([1.0, 2.0]), nugget=1e-3, nugget_est=True)
    assert np.allclose(k.log_t(), np.log([1.0, 2.0, 1e-3]))
    k.update(np.log([3.0, 4.0, 1e-2]))
    assert np.allclose(k.length, [3, 4]) and np.allclose(k.nugget, [1e-2])
    assert D.combine([1], [2]) == [[1], [2]]


def test_pickle_roundtrip_keeps_numpy_state(tmp_path):
    import dgp_b200 as D

    k = D.kernel(length=np.array([1.0]))
    k.input, k.output = np.random.rand(4, 1), np.random.rand(4, 1)
    k.Rinv, k.Rinv_y = np.eye(4), np.ones(4)
    D.write(k, str(tmp_path / "node"))
    k2 = D.read(str(tmp_path / "node"))
    assert np.array_equal(k2.Rinv, np.eye(4)) and np.array_equal(k2.input, k.input)


def test_input_validation_matches_reference():
    import dgp_b200 as D

    with pytest.raises(Exception):
        D.dgp(np.zeros(5), np.zeros((5, 1)))
    with pytest.raises(ValueError):
        D.kernel(length=np.array([1.0]), name="rbf")
