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


def test_likelihood_host_formulas_match_oracle_and_summary():
    """Host-side pieces of the likelihood layers that need no GPU: closed-form observable moments against the oracle
    restatement, the Gauss-Hermite expectation against direct quadrature, `summary` over a hierarchy with a
    likelihood node, and the structure checks of `dgp`."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dgp_b200 as D
    import dgp_oracle as O
    from scipy.integrate import quad
    from scipy.special import gammaln

    rng = np.random.default_rng(0)
    m, v = rng.normal(size=(6, 2)), rng.uniform(0.01, 0.5, size=(6, 2))
    for name, cls, cols in (("Poisson", D.Poisson, 1), ("Hetero", D.Hetero, 2), ("NegBin", D.NegBin, 2)):
        got = cls.prediction(m[:, :cols], v[:, :cols])
        ref = O.LikNode(name, np.arange(cols), None).prediction(m[:, :cols], v[:, :cols])
        assert np.allclose(got[0], np.ravel(ref[0]), rtol=1e-13) and np.allclose(got[1], np.ravel(ref[1]), rtol=1e-13)
    mu, var, y = np.array([[0.3], [1.0]]), np.array([[0.2], [0.05]]), np.array([[1.0], [3.0]])
    q = D.emulator._expected_likelihood(D.Poisson.pllik, mu, var, y)
    for i in range(2):
        f = lambda t: (np.exp(y[i, 0] * t - np.exp(t) - gammaln(y[i, 0] + 1))
                       * np.exp(-(t - mu[i, 0]) ** 2 / (2 * var[i, 0])) / np.sqrt(2 * np.pi * var[i, 0]))
        assert abs(q[i, 0] - quad(f, -10, 10)[0]) <= 1e-7
    # ZIP / ZINB reduce to Poisson / NegBin moments when the zero-inflation probability vanishes
    big = np.full((6, 1), -40.0)
    zm, zv = np.hstack((m[:, :1], big)), np.hstack((v[:, :1], np.zeros((6, 1))))
    assert np.allclose(D.ZIP.prediction(zm, zv)[0], D.Poisson.prediction(m[:, :1], v[:, :1])[0], rtol=1e-12)
    zm3, zv3 = np.hstack((m, big)), np.hstack((v, np.zeros((6, 1))))
    assert np.allclose(D.ZINB.prediction(zm3, zv3)[0], D.NegBin.prediction(m, v)[0], rtol=1e-12)
    cat = D.Categorical(num_classes=2, link="probit")
    p_mean, p_var = cat.prediction(np.zeros((3, 1)), np.ones((3, 1)))
    assert np.allclose(p_mean, 0.5) and np.all(p_var > 0) and np.all(p_var < 0.25)
    layers = D.combine([D.kernel(length=np.array([1.0]))], [D.kernel(length=np.array([1.0]), scale_est=True)],
                       [D.Poisson(input_dim=np.array([0]))])
    text = D.summary(layers, tablefmt="plain")
    assert "Poisson" in text and "Likelihood3.1" in text and "NA" in text
    X, Y = rng.uniform(size=(8, 1)), rng.poisson(3.0, size=(8, 1)).astype(float)
    with pytest.raises(NotImplementedError):   # likelihood nodes only as a final layer
        D.dgp(X, Y, D.combine([D.kernel(length=np.array([1.0]))], [D.Poisson()], [D.kernel(length=np.array([1.0]))]))


def test_bench_leg_budget_keeps_stated_sizes_when_they_fit():
    """bench.py's predict legs run the number of test points BASELINE.json states unless a timed probe says the call
    would exceed --leg-seconds; then the largest multiple of the shard granule that fits (never fewer than the probe)."""
    import sys
    sys.path.insert(0, ROOT)
    from bench import fit_points

    assert fit_points(10000, 1024, 6.0, 60.0, 256) == 10000               # 58.6 s: fits
    assert fit_points(1000000, 1024, 0.32, 60.0, 256) == 192000           # 312 s stated -> 60 s worth, multiple of 256
    assert fit_points(1000000, 1024, 0.32, 0.0, 256) == 1000000           # no budget: stated size
    assert fit_points(500, 1024, 1.0, 60.0, 256) == 500                   # fewer points than the probe
    assert fit_points(100000, 8192, 100.0, 60.0, 2048) == 8192            # never fewer than the probe
    assert fit_points(100000, 8192, 0.0, 60.0, 2048) == 100000            # degenerate timer
    for stated, probe, t, budget, g in ((123457, 1024, 1.7, 60.0, 256), (10 ** 6, 4096, 0.9, 30.0, 1024)):
        m = fit_points(stated, probe, t, budget, g)
        assert probe <= m <= stated and (m == stated or m % g == 0) and m * t / probe <= budget * 1.0001


def test_m_step_rendezvous_chunks_and_sizes_once(monkeypatch):
    """The M-step's rendezvous (`_GradBatcher`) on the host, with the library replaced by a stand-in: 40 optimiser
    threads (more than the 32 matrices one batched launch takes) get their own results back round after round, no call
    exceeds the limit, the batch size is derived from free device memory ONCE per M-step (cudaMemGetInfo costs
    milliseconds), and a thread that retires no longer holds the others up."""
    import threading

    from dgp_b200 import _lib as L
    import importlib
    dgp_mod = importlib.import_module("dgp_b200.dgp")

    calls, mem_queries = [], []

    class FakeLib:
        def dgpb_ws_bytes(self, ws):
            return 0

        def dgpb_last_error(self):
            return b""

        def dgpb_nllik_grad_dense_batch(self, ws, arr, B, n, res_ptr, ldo, status_ptr, stream):
            calls.append(B)
            res = np.ctypeslib.as_array(ctypes.cast(res_ptr, ctypes.POINTER(ctypes.c_double)), shape=(B, ldo))
            for b in range(B):
                res[b, :] = arr[b].scale * 10 + np.arange(ldo)       # recognisable per node
            return L.DGPB_OK

    class FakeCuda:
        @staticmethod
        def set_device(dev):
            pass

        @staticmethod
        def mem_get_info(dev):
            mem_queries.append(dev)
            return (64 << 30, 180 << 30)

    class FakeTorch:
        cuda = FakeCuda

    monkeypatch.setattr(L, "load", lambda: FakeLib())
    monkeypatch.setattr(L, "torch_mod", lambda: FakeTorch)
    monkeypatch.setattr(L, "stream", lambda: None)

    nthreads, rounds, P = 40, 3, 4
    batcher = dgp_mod._GradBatcher(nthreads, None, 0)
    got = [[] for _ in range(nthreads)]

    def worker(i):
        node = L.DgpbNode()
        node.scale = float(i)
        for r in range(rounds if i % 2 else rounds - 1):      # even threads finish one round early
            out = batcher.evaluate(node, 100, P, rid=i)
            got[i].append(out.copy())
        batcher.retire()

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(nthreads)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not any(t.is_alive() for t in threads)
    for i in range(nthreads):
        assert len(got[i]) == (rounds if i % 2 else rounds - 1)
        for out in got[i]:
            assert np.array_equal(out, i * 10 + np.arange(P + 2))
    assert max(calls) <= dgp_mod._GradBatcher.MAX_BATCH and sum(calls) == 20 * rounds + 20 * (rounds - 1)
    assert calls[:2] == [32, 8]                                # 40 requests of the first round: 32 + 8
    assert len(mem_queries) == 1
    assert batcher.stats[0] == len(calls) and batcher.stats[1] == sum(calls)
