"""GPU parity tests: the CUDA path (through the C-ABI / the reference-shaped Python API) against
(a) the golden vectors produced by the unmodified reference and (b) the CPU oracle on seeded inputs.

Tolerances follow BASELINE.json: kernel matrices / log-likelihoods / gradients / predictions <= 1e-9
relative in FP64 where the problem is well conditioned; for the reference's default 1e-6 nugget
(cond(K) ~ 1e7-1e9) the comparison is made against the cancellation scale of the quantity (the sum of the
absolute values of its terms), which is the reference's own BLAS-thread noise floor.
Index arrays are compared bit-exactly; ESS accept/shrink decisions must be identical.
"""
import numpy as np
import pytest

from conftest import floor_ok, relerr

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _case(g, ci):
    p = f"c{ci}_"
    d = {k[len(p):]: g[k] for k in g.files if k.startswith(p)}
    d["name"] = str(d["name"])
    return d


def _node_from_case(c, with_stats=False):
    import dgp_b200 as D

    nugget_est, scale_est, d_loc, d_glob = [int(v) for v in c["flags"][:4]]
    k = D.kernel(length=c["length"].copy(), scale=c["scale"][0], nugget=c["nugget"][0], name=c["name"],
                 nugget_est=bool(nugget_est), scale_est=bool(scale_est),
                 connect=np.arange(d_glob) if d_glob else None)
    X = c["X"]
    k.input = X[:, :d_loc].copy()
    k.input_dim = np.arange(d_loc)
    if d_glob:
        k.global_input = X[:, d_loc:].copy()
    k.output = c["y"].copy()
    k.D = d_loc + d_glob
    k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
    if with_stats:
        k.scale = c["scale_after"].copy()
        k.Rinv, k.Rinv_y = c["Rinv"], c["Rinv_y"]
    return k


# ------------------------------------------------------------------------------------------------ 1
def test_kmatrix_and_derivatives_vs_reference(golden_dense):
    g = golden_dense
    for ci in range(int(g["ncases"])):
        c = _case(g, ci)
        k = _node_from_case(c)
        K, fod = k.k_matrix(fod_eval=True)
        assert relerr(K, c["K"], 1e-300) <= 1e-12, ci
        assert fod.shape == c["fod"].shape
        assert np.max(np.abs(fod - c["fod"])) <= 1e-12, ci
        assert relerr(k.k_matrix(), c["K_plain"], 1e-300) <= 1e-12


def test_kmatrix_symmetry_and_ragged_sizes():
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(3)
    for n, d, name in ((1, 1, "sexp"), (2, 3, "matern2.5"), (63, 2, "sexp"), (64, 5, "matern2.5"), (65, 4, "sexp"),
                       (200, 16, "sexp"), (333, 7, "matern2.5")):
        k = D.kernel(length=rng.uniform(0.5, 2, d), name=name, nugget=1e-4)
        k.input = rng.uniform(0, 1, (n, d))
        K = k.k_matrix()
        assert np.array_equal(K, K.T)
        assert relerr(K, O.k_matrix(k.input, k.length, 1e-4, name), 1e-300) <= 1e-12


# ------------------------------------------------------------------------------------------------ 2
def test_dense_loglik_gradient_vs_reference(golden_dense):
    g = golden_dense
    for ci in range(int(g["ncases"])):
        c = _case(g, ci)
        k = _node_from_case(c)
        # 1e-9 relative, or -- where cond(K) makes the reference itself move by more than that -- a multiple of the
        # reference-vs-reference floor recorded in floor.npz
        ll = k.log_likelihood_func()
        assert floor_ok(f"dense_c{ci}_loglik", ll, c["loglik"]), (ci, ll, c["loglik"])
        f, gr = k.llik(k.log_t().copy())
        assert floor_ok(f"dense_c{ci}_nllik", f, c["nllik"]), ci
        assert floor_ok(f"dense_c{ci}_nllik_grad", gr, c["nllik_grad"]), (ci, gr, c["nllik_grad"])
        well = c["nugget"][0] >= 1e-3
        assert abs(k.scale[0] - c["scale_after"][0]) <= (1e-9 if well else 1e-7) * abs(c["scale_after"][0])


def test_dense_pipeline_vs_oracle_midsize():
    """n not a multiple of the 64-column panel / 128-row tile, several kernels; oracle = LAPACK."""
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(11)
    for n, d, name, ard in ((130, 3, "sexp", False), (257, 4, "matern2.5", True), (700, 6, "sexp", True)):
        length = rng.uniform(0.6, 1.5, d if ard else 1)
        k = D.kernel(length=length.copy(), name=name, nugget=1e-3, scale=0.7, nugget_est=True, scale_est=True)
        k.input = rng.uniform(0, 1, (n, d))
        k.output = np.sin(k.input.sum(1, keepdims=True) * 2) + 0.1 * rng.standard_normal((n, 1))
        k.D = d
        ll = k.log_likelihood_func()
        ll0 = O.loglik_dense(k.input, k.output, length, 0.7, 1e-3, name)
        assert abs(ll - ll0) <= 1e-9 * abs(ll0), (n, ll, ll0)
        f0, g0, s0 = O.nllik_grad_dense(k.input, k.output, length, 0.7, 1e-3, name, True, True)
        k.prior_name = None
        f, gr = k.llik(k.log_t().copy())
        assert abs(f[0] - f0) <= 1e-9 * max(1.0, abs(f0))
        assert np.max(np.abs(gr - g0)) <= 1e-8 * max(1.0, np.max(np.abs(g0))), (n, gr, g0)
        assert abs(k.scale[0] - s0) <= 1e-9 * s0
        # compute_stats: K * Rinv = I and Rinv_y = Rinv y
        k.compute_stats()
        K = O.k_matrix(k.input, k.length, k.nugget, name)
        Ri = k.Rinv
        assert np.array_equal(Ri, Ri.T)
        assert np.max(np.abs(K @ Ri - np.eye(n))) <= 1e-8
        assert relerr(k.Rinv_y, np.linalg.solve(K, k.output[:, 0]), 1e-3 * np.max(np.abs(k.Rinv_y))) <= 1e-7
        # prior draw: chol(scale K) z
        from dgp_b200 import _lib as L
        import ctypes
        z = rng.standard_normal(n)
        bufs = k._upload()  # keep the device buffers alive while the node descriptor points at them
        node = k._node(bufs)
        zd, nud = L.to_dev(z), L.empty((n,))
        L.check(L.load().dgpb_mvn_draw(L.workspace(), ctypes.byref(node), n, L.ptr(zd), L.ptr(nud), L.stream()))
        ref = np.linalg.cholesky(k.scale[0] * K) @ z
        assert relerr(nud.cpu().numpy(), ref, 1e-6) <= 1e-9


def test_not_positive_definite_raises_linalgerror():
    import dgp_b200 as D

    k = D.kernel(length=np.array([50.0]), nugget=-0.5)
    x = np.linspace(0, 1, 40)[:, None]
    k.input = np.concatenate((x, x[:5]), 0)  # diagonal 0.5 under off-diagonals ~1 -> indefinite
    k.output = np.ones((45, 1))
    with pytest.raises(np.linalg.LinAlgError):
        k.log_likelihood_func()


def test_potrf_and_gemm_building_blocks():
    from dgp_b200 import _lib as L
    import ctypes

    lib = L.load()
    rng = np.random.default_rng(5)
    for n in (64, 100, 513, 1500):
        A = rng.standard_normal((n, n))
        A = A @ A.T / n + np.eye(n)
        Ad = L.to_dev(A)
        info = ctypes.c_int(-1)
        L.check(lib.dgpb_potrf(L.workspace(), L.ptr(Ad), n, ctypes.byref(info), L.stream()))
        assert info.value == 0
        Lc = np.tril(Ad.cpu().numpy())
        assert np.max(np.abs(Lc @ Lc.T - A)) <= 1e-12 * n
    M, N, K = 300, 258, 96
    A, B = rng.standard_normal((M, K)), rng.standard_normal((N, K))
    C, Ad, Bd = L.empty((M, N)), L.to_dev(A), L.to_dev(B)
    L.check(lib.dgpb_dgemm_nt(L.ptr(Ad), L.ptr(Bd), L.ptr(C), M, N, K, L.stream()))
    assert np.max(np.abs(C.cpu().numpy() - A @ B.T)) <= 1e-12 * K


def test_hyper_block_factorisation_variants():
    """The two-level blocked factorisation (hyper-blocks of 128 / 256 / 512 columns, forced on at small n through
    dgpb_tune) gives the same factor, log-likelihood, gradient and inverse as LAPACK for every setting -- and the
    SAME BITS for every setting: each entry is updated in ascending column order in groups of four whatever the
    blocking, so a batch of three matrices (256-column hyper-blocks) and a batch of eight (512) agree exactly,
    which is what keeps a chain shared by several GPUs identical to the chain on one."""
    import ctypes
    import dgp_b200 as D
    from dgp_b200 import _lib as L
    from oracle import dgp_oracle as O

    lib = L.load()
    rng = np.random.default_rng(17)
    n = 1100
    A = rng.standard_normal((n, n))
    A = A @ A.T / n + np.eye(n)
    d = 5
    length = rng.uniform(0.6, 1.5, d)
    X = rng.uniform(0, 1, (n, d))
    y = np.sin(X.sum(1, keepdims=True) * 2) + 0.1 * rng.standard_normal((n, 1))
    ll0 = O.loglik_dense(X, y, length, 0.7, 1e-3, "sexp")
    f0, g0, s0 = O.nllik_grad_dense(X, y, length, 0.7, 1e-3, "sexp", True, True)
    K = O.k_matrix(X, length, 1e-3, "sexp")
    bits = []
    try:
        L.check(lib.dgpb_tune(b"hb_small_b", 0))   # a single matrix would otherwise be capped at 256 columns
        for hb, min_w in ((128, 0), (256, 0), (512, 0), (512, 700), (1024, 0)):
            L.check(lib.dgpb_tune(b"hb", hb))
            L.check(lib.dgpb_tune(b"hb_min_w", min_w))
            Ad = L.to_dev(A)
            info = ctypes.c_int(-1)
            L.check(lib.dgpb_potrf(L.workspace(), L.ptr(Ad), n, ctypes.byref(info), L.stream()))
            assert info.value == 0
            Lc = np.tril(Ad.cpu().numpy())
            assert np.max(np.abs(Lc @ Lc.T - A)) <= 1e-12 * n, (hb, min_w)
            k = D.kernel(length=length.copy(), name="sexp", nugget=1e-3, scale=0.7, nugget_est=True, scale_est=True)
            k.input, k.output, k.D, k.prior_name = X, y.copy(), d, None
            ll = k.log_likelihood_func()
            assert abs(ll - ll0) <= 1e-9 * abs(ll0), (hb, min_w, ll, ll0)
            f, gr = k.llik(k.log_t().copy())
            assert abs(f[0] - f0) <= 1e-9 * max(1.0, abs(f0)), (hb, min_w)
            assert np.max(np.abs(gr - g0)) <= 1e-8 * max(1.0, np.max(np.abs(g0))), (hb, min_w, gr, g0)
            k.compute_stats()
            assert np.max(np.abs(K @ k.Rinv - np.eye(n))) <= 1e-8, (hb, min_w)
            bits.append((Lc, ll, f[0], gr.copy()))
        for Lc, ll, f, gr in bits[1:]:
            assert np.array_equal(Lc, bits[0][0]) and ll == bits[0][1] and f == bits[0][2]
            assert np.array_equal(gr, bits[0][3])
        with pytest.raises(ValueError):
            L.check(lib.dgpb_tune(b"no_such_knob", 1))
    finally:
        L.check(lib.dgpb_tune(b"hb", 512))
        L.check(lib.dgpb_tune(b"hb_min_w", 2560))
        L.check(lib.dgpb_tune(b"hb_small_b", 3))


# ------------------------------------------------------------------------------------------------ 5
def _gp_scales(c, xt):
    from oracle import dgp_oracle as O

    r = O.k_cross(c["X"], xt, c["length"], c["name"])
    Sm = np.abs(r) @ np.abs(c["Rinv_y"])
    Sv = np.einsum("ti,ij,tj->t", np.abs(r), np.abs(c["Rinv"]), np.abs(r)) * c["scale_after"][0]
    return Sm, Sv


def test_predictions_given_reference_stats(golden_dense):
    g = golden_dense
    for ci in range(int(g["ncases"])):
        c = _case(g, ci)
        k = _node_from_case(c, with_stats=True)
        d_glob = int(c["flags"][3])
        well = c["nugget"][0] >= 1e-3
        zt = c["zt"] if d_glob else None
        m, v = k.gp_prediction(c["xt"], zt)
        xt = c["xt"] if zt is None else np.concatenate((c["xt"], zt), 1)
        Sm, Sv = _gp_scales(c, xt)
        assert np.all(np.abs(m - c["gp_m"]) <= 1e-13 * Sm), ci
        assert np.all(np.abs(v - c["gp_v"]) <= 1e-13 * Sv), ci
        if well:
            assert relerr(m, c["gp_m"], 1e-3) <= 1e-9
            assert np.max(np.abs(v - c["gp_v"])) <= 1e-9 * c["scale_after"][0]
        m2, v2 = k.linkgp_prediction(c["lk_m_in"], c["lk_v_in"], zt)
        a = np.abs(c["Rinv_y"])
        Slm = a.sum()
        Slv = a.sum() ** 2 + c["scale_after"][0] * np.abs(c["Rinv"]).sum()
        assert np.all(np.abs(m2 - c["lk_m"]) <= 1e-13 * Slm), (ci, np.max(np.abs(m2 - c["lk_m"])))
        assert np.all(np.abs(v2 - c["lk_v"]) <= 1e-13 * Slv), (ci, np.max(np.abs(v2 - c["lk_v"])))
        if well:
            assert relerr(m2, c["lk_m"], 1e-3) <= 1e-9, ci
            assert np.max(np.abs(v2 - c["lk_v"])) <= 1e-9 * c["scale_after"][0], ci


def test_linkgp_tensor_core_kernel_equals_vector_kernel():
    """link_gp with the squared-exponential kernel evaluates the J exponents as a DMMA product of test-point
    coefficients and pair features; same moments as the vector-pipe kernel and the oracle, with and without
    connected global inputs, ragged n / M."""
    from dgp_b200 import _lib as L
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    lib = L.load()
    rng = np.random.default_rng(41)
    for n, M, Dw, Dz, ard in ((150, 37, 3, 0, False), (333, 70, 8, 8, False), (257, 16, 5, 2, True), (64, 100, 1, 0, False)):
        D_all = Dw + Dz
        length = rng.uniform(0.5, 1.5, D_all if ard else 1)
        k = D.kernel(length=length.copy(), name="sexp", nugget=1e-3, scale=1.4,
                     connect=np.arange(Dz) if Dz else None)
        k.input = rng.uniform(0, 1, (n, Dw))
        k.input_dim = np.arange(Dw)
        if Dz:
            k.global_input = rng.uniform(0, 1, (n, Dz))
        k.output = np.sin(k.input.sum(1, keepdims=True) * 2)
        k.D = D_all
        k.compute_stats()
        m_in, v_in = rng.uniform(0, 1, (M, Dw)), rng.uniform(1e-4, 0.05, (M, Dw))
        z = rng.uniform(0, 1, (M, Dz)) if Dz else None
        L.check(lib.dgpb_tune(b"linkgp_mma", 0))
        try:
            m0, v0 = k.linkgp_prediction(m_in, v_in, z)
        finally:
            L.check(lib.dgpb_tune(b"linkgp_mma", 1))
        m1, v1 = k.linkgp_prediction(m_in, v_in, z)
        assert relerr(m1, m0, 1e-9) <= 1e-10, (n, M, Dw, Dz)
        assert np.max(np.abs(v1 - v0)) <= 1e-9 * max(1.0, np.max(np.abs(v0))), (n, M, Dw, Dz, np.max(np.abs(v1 - v0)))


def test_linkgp_matern_tabulated_kernel_equals_direct_kernel():
    """Matern-2.5 link_gp: the J integrals assembled from per-point tables of their transcendental factors equal
    the direct per-pair closed form (and the oracle), including zero input variances, connected global inputs,
    coincident coordinates and ragged sizes."""
    from dgp_b200 import _lib as L
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    lib = L.load()
    rng = np.random.default_rng(43)
    for n, M, Dw, Dz, ard in ((90, 21, 2, 0, False), (200, 33, 5, 0, True), (130, 9, 3, 2, True), (33, 70, 1, 0, False)):
        D_all = Dw + Dz
        length = rng.uniform(0.5, 1.5, D_all if ard else 1)
        k = D.kernel(length=length.copy(), name="matern2.5", nugget=1e-3, scale=1.4,
                     connect=np.arange(Dz) if Dz else None)
        k.input = rng.uniform(0, 1, (n, Dw))
        k.input[1] = k.input[0]          # coincident training coordinates (x1 == x2 off the diagonal)
        k.input_dim = np.arange(Dw)
        if Dz:
            k.global_input = rng.uniform(0, 1, (n, Dz))
        k.output = np.sin(k.input.sum(1, keepdims=True) * 2)
        k.D = D_all
        k.compute_stats()
        m_in, v_in = rng.uniform(-0.2, 1.2, (M, Dw)), rng.uniform(1e-4, 0.05, (M, Dw))
        v_in[0, 0] = 0.0                  # deterministic input in one dimension
        z = rng.uniform(0, 1, (M, Dz)) if Dz else None
        L.check(lib.dgpb_tune(b"linkgp_matern_tab", 0))
        try:
            m0, v0 = k.linkgp_prediction(m_in, v_in, z)
        finally:
            L.check(lib.dgpb_tune(b"linkgp_matern_tab", 1))
        m1, v1 = k.linkgp_prediction(m_in, v_in, z)
        assert relerr(m1, m0, 1e-9) <= 1e-10, (n, M, Dw, Dz)
        assert np.max(np.abs(v1 - v0)) <= 1e-9 * max(1.0, np.max(np.abs(v0))), (n, M, Dw, Dz, np.max(np.abs(v1 - v0)))
        if Dz == 0 and n <= 100:
            lfull = np.full(Dw, length[0]) if len(length) == 1 else length
            m2, v2 = O.link_gp(m_in, v_in, None, k.input, None, k.Rinv, k.Rinv_y, None, None, 1.4, lfull, 1e-3, "matern2.5")
            # two coincident training rows make R^-1 ~ 1/nugget large: the variance is a difference of O(1e3) terms
            assert relerr(m1, m2, 1e-3) <= 1e-9 and np.max(np.abs(v1 - v2)) <= 1e-12 * np.max(np.abs(k.Rinv)) * n


def test_linkgp_wide_inputs_vs_oracle():
    """Dw beyond the register-tiled template sizes and n not a multiple of the pair tile."""
    import dgp_b200 as D
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(21)
    for name, Dw, Dz, n in (("sexp", 9, 0, 70), ("sexp", 13, 2, 45), ("sexp", 3, 1, 100), ("matern2.5", 2, 1, 50)):
        length = rng.uniform(0.8, 1.6, Dw + Dz)
        k = D.kernel(length=length.copy(), name=name, nugget=1e-3, scale=1.2,
                     connect=np.arange(Dz) if Dz else None)
        k.input = rng.uniform(0, 1, (n, Dw))
        if Dz:
            k.global_input = rng.uniform(0, 1, (n, Dz))
        X = k._X()
        k.output = np.cos(X.sum(1, keepdims=True))
        Rinv, Rinv_y = O.compute_stats(X, k.output, length, 1e-3, name)
        k.Rinv, k.Rinv_y = Rinv, Rinv_y
        M = 19
        m_in, v_in = rng.uniform(0, 1, (M, Dw)), rng.uniform(1e-4, 0.03, (M, Dw))
        z = rng.uniform(0, 1, (M, Dz)) if Dz else None
        m, v = k.linkgp_prediction(m_in, v_in, z)
        R2, P = O.sexp_stats(k.input, length[:Dw]) if name == "sexp" else (None, None)
        m0, v0 = O.link_gp(m_in, v_in, z, k.input, k.global_input, Rinv, Rinv_y, R2, P, 1.2, length, 1e-3, name)
        assert relerr(m, m0, 1e-3) <= 1e-9, (name, Dw)
        assert np.max(np.abs(v - v0)) <= 1e-9 * 1.2, (name, Dw, np.max(np.abs(v - v0)))


# ------------------------------------------------------------------------------------------------ 4
def test_nn_indices_bit_exact(golden_vecchia):
    from dgp_b200 import vecchia as V
    from oracle import dgp_oracle as O

    g = golden_vecchia
    for j in range(4):
        x, m = g[f"nn{j}_x"], int(g[f"nn{j}_m"])
        assert np.array_equal(V.nn(x, m), g[f"nn{j}_NN"]), j
        assert np.array_equal(V.get_pred_nn(g[f"nn{j}_q"], x, 50), g[f"nn{j}_pred"]), j
    rng = np.random.default_rng(2)
    x = rng.uniform(0, 1, (5000, 10))
    q = rng.uniform(0, 1, (777, 10))
    assert np.array_equal(V.nn(x, 25), O.nn_ordered(x, 25))
    assert np.array_equal(V.get_pred_nn(q, x, 50), O.knn(q, x, 50))
    # m >= n shortcut and tiny inputs
    assert np.array_equal(V.get_pred_nn(q[:7], x[:5], 50), O.knn(q[:7], x[:5], 50))
    assert np.array_equal(V.nn(x[:3], 25), O.nn_ordered(x[:3], 25))
    assert np.array_equal(V.nn(x[:1], 25), O.nn_ordered(x[:1], 25))


def test_knn_tensor_core_screen_equals_scalar_search():
    """The neighbour search screens candidates on the tensor path (FP64 DMMA, or split-TF32 HMMA) and ranks the
    survivors with the reference arithmetic; its index arrays must equal those of the scalar exact kernel bit for
    bit -- inputs with a large common offset (the TF32 screen shifts by the first candidate), random inputs, ragged
    sizes, both list capacities (m <= 29 / m <= 61), the ordered variant, and inputs full of exact distance ties
    (lattice points, duplicates) where the screen has to hand queries back to the scalar kernel."""
    from dgp_b200 import _lib as L
    from dgp_b200 import vecchia as V

    lib = L.load()
    rng = np.random.default_rng(23)

    default_mode = 5

    def run(fn, mode):
        L.check(lib.dgpb_tune(b"knn_mma", mode))
        try:
            return fn()
        finally:
            L.check(lib.dgpb_tune(b"knn_mma", default_mode))

    def both(fn):
        # scalar exact, FP64 DMMA screen, warp-level split-TF32 screen, tcgen05 split-TF32 screen (TMA + TMEM)
        ref, dmma, tf32, tc5 = run(fn, 0), run(fn, 1), run(fn, 3), run(fn, 5)
        assert np.array_equal(dmma, tf32) and np.array_equal(dmma, tc5)
        return ref, tc5

    assert np.array_equal(run(lambda: V.get_pred_nn(rng.uniform(0, 1, (64, 4)), rng.uniform(0, 1, (300, 4)), 5), 1).shape, (64, 5))
    for n, M, D, m in ((1000, 333, 3, 5), (5000, 1500, 10, 25), (3001, 700, 20, 50), (700, 129, 31, 29),
                       (300, 64, 10, 61), (40, 17, 2, 30), (129, 1, 1, 25)):
        off = 1000.0 if n == 5000 else 0.0
        x, q = rng.uniform(0, 1, (n, D)) + off, rng.uniform(0, 1, (M, D)) + off
        a, b = both(lambda: V.get_pred_nn(q, x, m))
        assert a.shape == (M, min(m, n)) and np.array_equal(a, b), (n, M, D, m)
        a, b = both(lambda: V.nn(x, m))
        assert np.array_equal(a, b), ("ordered", n, D, m)
        d2 = ((q[:, None, :] - x[None, :, :]) ** 2).sum(-1) if n * M <= 2_000_000 else None
        if d2 is not None:
            assert np.array_equal(np.sort(a_idx := np.argsort(d2, 1, kind="stable")[:, :min(m, n)], 1),
                                  np.sort(V.get_pred_nn(q, x, m), 1)), (n, M, D, m)
    # exact ties: integer lattice with duplicated points, queries on lattice sites
    g = np.stack(np.meshgrid(*[np.arange(6.0)] * 3, indexing="ij"), -1).reshape(-1, 3)
    x = np.concatenate((g, g[:50]), 0)
    q = g[::3] + 0.0
    a, b = both(lambda: V.get_pred_nn(q, x, 25))
    assert np.array_equal(a, b)
    a, b = both(lambda: V.nn(x, 10))
    assert np.array_equal(a, b)


def test_vecchia_register_kernel_equals_shared_memory_kernel():
    """Blocks of <= 32 points are factored in registers (lane = row); same numbers as the shared-memory kernel
    and as the oracle, for both kernels, ragged early blocks (-1 padded rows) and the prediction form."""
    from dgp_b200 import _lib as L
    from dgp_b200 import vecchia as V
    from oracle import dgp_oracle as O

    lib = L.load()
    rng = np.random.default_rng(31)

    def both(fn):
        L.check(lib.dgpb_tune(b"vecchia_small", 0))
        try:
            ref = fn()
        finally:
            L.check(lib.dgpb_tune(b"vecchia_small", 1))
        return ref, fn()

    for n, D, m, name in ((300, 3, 5, "sexp"), (500, 10, 25, "sexp"), (257, 7, 31, "matern2.5"), (40, 2, 1, "matern2.5")):
        X = rng.uniform(0, 1, (n, D))
        y = np.sin(3 * X.sum(1)) + 0.05 * rng.standard_normal(n)
        length = rng.uniform(0.5, 1.5, D)
        NN = V.nn(X / length, m)
        a, b = both(lambda: V.vecchia_llik(X, y, NN, 1.3, length, 1e-3, None, name))
        assert abs(a - b) <= 1e-11 * abs(a), (n, D, m, a, b)
        ref = O.vecchia_llik(X, y.reshape(-1, 1), NN, 1.3, length, 1e-3, np.ones(n), name)
        assert abs(b - ref) <= 1e-9 * abs(ref), (n, D, m, b, ref)
        # prediction form: gp_vecch through the node API
        import dgp_b200 as Dm
        k = Dm.kernel(length=length.copy(), name=name, nugget=1e-3, scale=1.3)
        k.input, k.output, k.D, k.vecch, k.pred_m = X, y.reshape(-1, 1), D, True, m
        xt = rng.uniform(0, 1, (77, D))
        (m0, v0), (m1, v1) = both(lambda: k.gp_prediction(xt, None))
        assert relerr(m1, m0, 1e-9) <= 1e-10 and relerr(v1, v0, 1e-12) <= 1e-9, (n, D, m)


def test_vecchia_kernels_vs_reference(golden_vecchia):
    import dgp_b200 as D
    from dgp_b200 import vecchia as V

    g = golden_vecchia
    for ci in range(int(g["ncases"])):
        c = _case(g, ci)
        nugget_est, scale_est, d_loc, d_glob, m = [int(v) for v in c["flags"]]
        X, y, o, NN = c["X"], c["y"], c["ord"], c["NNarray"]
        assert np.array_equal(V.nn((X / c["length"])[o], m), NN), ci
        fk = f"vecchia_c{ci}_"
        ll = V.vecchia_llik(X[o], y[o], NN, c["scale"][0], c["length"], c["nugget"][0], None, c["name"])
        assert floor_ok(fk + "llik", ll, c["llik"]), ci
        Lm = V.L_matrix(X[o], NN, c["length"], c["nugget"][0], c["name"])
        assert floor_ok(fk + "Lmatrix", Lm, c["Lmatrix"]), ci
        draw = V.fmvn_sp(X[o], NN, c["scale"][0], c["length"], c["nugget"][0], c["name"], z=c["z"])
        # the draw divides by the entries of L: its floor entry only saw the solve, not a perturbed L
        assert floor_ok(fk + "draw", draw, c["draw"], rel=1e-9 if nugget_est else 1e-6), ci
        k = D.kernel(length=c["length"].copy(), scale=c["scale"][0], nugget=c["nugget"][0], name=c["name"],
                     nugget_est=bool(nugget_est), scale_est=bool(scale_est),
                     connect=np.arange(d_glob) if d_glob else None)
        k.input, k.output = X[:, :d_loc].copy(), y.copy()
        if d_glob:
            k.global_input = X[:, d_loc:].copy()
        k.vecch, k.m, k.ord, k.NNarray, k.rev_ord = True, m, o, NN, np.argsort(o)
        f, gr = k.llik_vecch(k.log_t().copy())
        assert floor_ok(fk + "nllik", f, c["nllik"]), ci
        assert floor_ok(fk + "nllik_grad", gr, c["nllik_grad"]), (ci, gr)
        assert abs(k.scale[0] - c["scale_after"][0]) <= 1e-9 * abs(c["scale_after"][0])
        # predictions (neighbour search + block kernels)
        k.pred_m = c["pred_NN"].shape[1]
        zt = c["zt"] if d_glob else None
        m1, v1 = k.gp_prediction(c["xt"], zt)
        assert floor_ok(fk + "gp_m", m1, c["gp_m"]), ci
        # the predictive variance sigma^2 (1 + eta - r'K^-1 r) cancels to ~eta at the training points' scale: its
        # absolute error is bounded at the scale of the cancelling terms (sigma^2), as for the dense `gp`
        assert floor_ok(fk + "gp_v", v1, c["gp_v"], rel=1e-9 * c["scale_after"][0] / max(1e-300, np.max(c["gp_v"]))), ci
        m2, v2 = k.linkgp_prediction(c["lk_m_in"], c["lk_v_in"], zt)
        assert floor_ok(fk + "lk_m", m2, c["lk_m"]), ci
        assert floor_ok(fk + "lk_v", v2, c["lk_v"]), (ci, float(np.max(np.abs(v2 - c["lk_v"]))))


# ------------------------------------------------------------------------------------------------ 3
def _load_layers(g, prefix, widths, name, vecch):
    import dgp_b200 as D

    layers = []
    for l, w in enumerate(widths):
        layer = []
        for k in range(w):
            p = f"{prefix}L{l}K{k}_"
            node = D.kernel(length=g[p + "length"].copy(), scale=g[p + "scale"][0], nugget=g[p + "nugget"][0], name=name,
                            scale_est=(l == len(widths) - 1))
            node.input = g[p + "input"].copy()
            node.output = g[p + "output"].copy()
            node.input_dim = np.arange(node.input.shape[1])
            if p + "global_input" in g.files:
                node.global_input = g[p + "global_input"].copy()
                node.connect = np.arange(node.global_input.shape[1])
            node.D = node.input.shape[1] + (0 if node.global_input is None else node.global_input.shape[1])
            node.para_path = np.atleast_2d(np.concatenate((node.scale, node.length, node.nugget)))
            node.vecch = vecch
            if vecch:
                node.ord, node.NNarray = g[p + "ord"], g[p + "NNarray"]
                node.rev_ord = np.argsort(node.ord)
                node.m = node.NNarray.shape[1] - 1
            layer.append(node)
        layers.append(layer)
    return layers


@pytest.mark.parametrize("ess_batch,ess_trsv,ess_prefetch", [(8, 1, 1), (1, 1, 1), (3, 0, 1), (32, 1, 0), (8, 0, 0)])
def test_ess_replay_identical_decisions(golden_ess, ess_batch, ess_trsv, ess_prefetch):
    """Replays the reference's ESS sweeps with its own normal / uniform draws.  `ess_batch` is the size of the
    speculative proposal wave (1 = one proposal at a time), `ess_trsv` switches the threshold-by-triangular-solve
    shortcut (cached factors of upper nodes whose inputs did not move), `ess_prefetch` the speculative assembly of
    the next wave during the current factorisation: every setting must consume exactly the
    reference's uniforms and try exactly its angles."""
    from dgp_b200 import _lib as L
    from dgp_b200.imputation import _DeviceLayers

    L.check(L.load().dgpb_tune(b"ess_batch", ess_batch))
    L.check(L.load().dgpb_tune(b"ess_trsv", ess_trsv))
    L.check(L.load().dgpb_tune(b"ess_prefetch", ess_prefetch))
    try:
        _ess_replay(golden_ess, _DeviceLayers)
    finally:
        L.check(L.load().dgpb_tune(b"ess_batch", 8))
        L.check(L.load().dgpb_tune(b"ess_trsv", 1))
        L.check(L.load().dgpb_tune(b"ess_prefetch", 1))


def _ess_replay(golden_ess, _DeviceLayers):
    g = golden_ess
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        widths, name, vecch = [int(w) for w in g[p + "widths"]], str(g[p + "name"]), bool(g[p + "vecch"])
        layers = _load_layers(g, p + "pre_", widths, name, vecch)
        dev = _DeviceLayers(layers)
        Z, U, sweeps = g[p + "Z"], g[p + "U"], int(g[p + "sweeps"])
        zi = ui = 0
        values = []
        for _ in range(sweeps):
            for l in range(len(widths) - 1):
                M = widths[l]
                nprop, thetas = dev.ess_call(l, list(range(M)), list(range(widths[l + 1])), Z[zi:zi + M],
                                             np.concatenate((U[ui:], np.full(4, 0.5))))
                zi += M
                values.append(U[ui])
                values.extend(thetas)
                ui += 1 + nprop
        assert zi == len(Z) and ui == len(U), (ci, ui, len(U))  # same draws consumed == identical decisions
        assert np.allclose(values, g[p + "draw_values"], rtol=1e-12, atol=0), ci
        dev.write_back()
        post = _load_layers(g, p + "post_", widths, name, vecch)
        for l in range(len(widths)):
            for k in range(widths[l]):
                assert floor_ok(f"ess_c{ci}_post_L{l}K{k}", layers[l][k].output, post[l][k].output), (ci, l, k)
                assert relerr(layers[l][k].input, post[l][k].input, 1e-4) <= 1e-6, (ci, l, k)
        # M-step from the reference's imputed state reproduces its optimiser path
        for l in range(len(widths)):
            for k in range(widths[l]):
                node = post[l][k]
                node.maximise()
                got = np.concatenate((node.scale, node.length, node.nugget))
                assert np.allclose(got, g[p + f"mstep_L{l}K{k}"], rtol=5e-4), (ci, l, k, got)


# ------------------------------------------------------------------------------------------------ e2e
def _snapshot_layers(g, prefix, name_of=lambda l, k: "sexp", scale_est_last=True):
    import dgp_b200 as D

    layers, l = [], 0
    while f"{prefix}L{l}K0_input" in g.files:
        layer, k = [], 0
        while f"{prefix}L{l}K{k}_input" in g.files:
            p = f"{prefix}L{l}K{k}_"
            node = D.kernel(length=g[p + "length"].copy(), scale=g[p + "scale"][0], nugget=g[p + "nugget"][0],
                            name=name_of(l, k))
            node.input, node.output = g[p + "input"].copy(), g[p + "output"].copy()
            node.input_dim = np.arange(node.input.shape[1])
            if p + "global_input" in g.files:
                node.global_input = g[p + "global_input"].copy()
                node.connect = np.arange(node.global_input.shape[1])
            node.vecch = False
            layer.append(node)
            k += 1
        layers.append(layer)
        l += 1
    return layers


def _frozen_emulator(g, tag, name):
    import dgp_b200 as D

    emu = D.emulator.__new__(D.emulator)
    emu.all_layer_set = []
    for s in range(int(g[f"{tag}_nimp"])):
        layers = _snapshot_layers(g, f"{tag}_S{s}_", lambda l, k: name)
        for layer in layers:
            for node in layer:
                node.compute_stats()
        emu.all_layer_set.append(layers)
    emu.all_layer = emu.all_layer_set[0]
    emu.n_layer = len(emu.all_layer)
    emu.vecch = False
    return emu


def test_end_to_end_predict_with_frozen_imputations(golden_e2e):
    """emulator.predict on the reference's own imputed states (step function = BASELINE config 1, and the
    2-layer Matern model = config-2 shape): K^-1 is recomputed on the GPU, so this checks kernel build +
    factorisation + gp + link_gp + aggregation end to end."""
    g = golden_e2e
    emu = _frozen_emulator(g, "step", "sexp")
    mu, var = emu.predict(g["step_xt"])
    assert mu.shape == g["step_mu"].shape
    assert floor_ok("e2e_step_mu", mu, g["step_mu"])
    assert floor_ok("e2e_step_var", var, g["step_var"])
    emu2 = _frozen_emulator(g, "mat", "matern2.5")
    mu2, var2 = emu2.predict(g["mat_xt"])
    assert floor_ok("e2e_mat_mu", mu2, g["mat_mu"])
    assert floor_ok("e2e_mat_var", var2, g["mat_var"])


def _frozen_linked_system(g):
    import dgp_b200 as D

    kinds = {0: "matern2.5", 1: "matern2.5", 2: "sexp"}
    sets = []
    for s in range(int(g["lgp_nimp"])):
        one = []
        for e in range(3):
            layers = _snapshot_layers(g, f"lgp_S{s}_E{e}_", lambda l, k, e=e: kinds[e])
            for layer in layers:
                for node in layer:
                    node.compute_stats()
            cont = D.container.__new__(D.container)
            if len(layers) == 1:
                cont.type, cont.structure = 'gp', layers[0][0]
            else:
                cont.type, cont.structure = 'dgp', layers
            cont.vecch = False
            cont.local_input_idx = np.array([0, 1]) if e == 0 else np.array([0])
            one.append([cont])
        sets.append(one)
    system = D.lgp.__new__(D.lgp)
    system.L, system.all_layer, system.all_layer_set, system.num_model = 3, sets[0], sets, [1, 1]
    return system


def test_linked_system_with_frozen_imputations(golden_e2e):
    g = golden_e2e
    system = _frozen_linked_system(g)
    mu, var = system.predict(g["lgp_xt"])
    assert np.max(np.abs(mu[0] - g["lgp_mu"])) <= 5e-6 * max(1.0, np.max(np.abs(g["lgp_mu"])))
    assert np.max(np.abs(var[0] - g["lgp_var"])) <= 5e-6 * max(1.0, np.max(g["lgp_var"]))


def test_linked_system_with_likelihood_emulator(golden_lik):
    """lgp whose second emulator is a DGP + Poisson likelihood (linkgp.py:569-571) on the reference's imputed states."""
    import dgp_b200 as D

    g = golden_lik
    sets = []
    for s in range(int(g["lgp_nimp"])):
        first = _snapshot_layers(g, f"lgp_S{s}_E0_", lambda l, k: "matern2.5")
        second = _snapshot_layers(g, f"lgp_S{s}_E1_", lambda l, k: "sexp")
        for layer in first + second:
            for node in layer:
                node.compute_stats()
        lik = D.Poisson(input_dim=np.arange(1))
        lik.output, lik.input = g["lgp_Y2"].copy(), second[-1][0].output.copy()
        c1, c2 = D.container.__new__(D.container), D.container.__new__(D.container)
        c1.type, c1.structure, c1.vecch, c1.local_input_idx = 'gp', first[0][0], False, np.array([0, 1])
        c2.type, c2.structure, c2.vecch, c2.local_input_idx = 'dgp', second + [[lik]], False, np.array([0])
        sets.append([[c1], [c2]])
    system = D.lgp.__new__(D.lgp)
    system.L, system.all_layer, system.all_layer_set, system.num_model = 2, sets[0], sets, [1]
    mu, var = system.predict(g["lgp_xt"])
    assert mu[0].shape == g["lgp_mu"].shape
    # the CPU restatement of the same chain differs from the reference by 3.4e-5 here (three emulators with
    # cond(K) ~ 2e7-3e7 feeding an exponential)
    assert np.max(np.abs(mu[0] - g["lgp_mu"]) / np.abs(g["lgp_mu"])) <= 2e-4
    assert np.max(np.abs(var[0] - g["lgp_var"]) / np.abs(g["lgp_var"])) <= 1e-3


def test_batched_m_step_equals_one_node_at_a_time(monkeypatch):
    """The M-step serves every round of L-BFGS-B requests of all dense nodes with one batched sliding-window
    factorisation; each node must follow exactly the parameter path it follows alone."""
    import copy
    import dgp_b200 as D

    rng = np.random.default_rng(5)
    n, d = 180, 3
    X = rng.uniform(0, 1, (n, d))
    Y = np.stack([np.sin(4 * X.sum(1)), X[:, 0] * X[:, 1]], 1)
    np.random.seed(3)
    D.nb_seed(3)
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(3)]
    l2 = [D.kernel(length=np.array([1.0, 1.2, 0.8, 1.1, 0.9, 1.0]), name="matern2.5", scale_est=True, nugget_est=True,
                   nugget=1e-3, connect=np.arange(d)) for _ in range(2)]
    model = D.dgp(X, Y, D.combine(l1, l2))
    model.imp.sample(burnin=2)
    twin = copy.deepcopy(model)
    monkeypatch.setenv("DGPB_MSTEP_BATCH", "0")
    twin._m_step()
    monkeypatch.setenv("DGPB_MSTEP_BATCH", "1")
    model._m_step()
    for la, lb in zip(model.all_layer, twin.all_layer):
        for ka, kb in zip(la, lb):
            assert np.array_equal(ka.length, kb.length) and np.array_equal(ka.scale, kb.scale)
            assert np.array_equal(ka.nugget, kb.nugget)


def test_leave_one_out_vs_reference(golden_loo):
    """gp.loo (dense closed form and Vecchia) and emulator.loo (dense emulator conditioning on all other points,
    Vecchia emulator on the m nearest others) against the reference's values on its own imputed states."""
    import dgp_b200 as D

    g = golden_loo
    X, Y = g["gp_X"], g["gp_Y"]
    for tag, name in (("se", "sexp"), ("ma", "matern2.5")):
        length = np.array([0.7, 0.9])
        gd = D.gp(X, Y, D.kernel(length=length.copy(), scale=1.3, nugget=1e-4, name=name))
        mu, s2 = gd.loo()
        assert mu.shape == (len(X), 1) and s2.shape == (len(X), 1)
        assert relerr(mu, g[f"gp_{tag}_dense_mu"], 1e-2) <= 1e-7, tag        # nugget 1e-4: cond(K) ~ 1e6
        assert relerr(s2, g[f"gp_{tag}_dense_var"], 1e-300) <= 1e-7, tag
        gv = D.gp(X, Y, D.kernel(length=length.copy(), scale=1.3, nugget=1e-4, name=name), vecchia=True, m=8)
        mu, s2 = gv.loo(m=6)
        assert relerr(mu, g[f"gp_{tag}_vecch_mu"], 1e-3) <= 1e-8, tag
        assert relerr(s2, g[f"gp_{tag}_vecch_var"], 1e-300) <= 1e-8, tag
        samples = gv.loo(method="sampling", sample_size=7, m=6)
        assert samples.shape == (len(X), 7) and np.all(np.isfinite(samples))
    Xe = g["emu_X"]
    for tag, vec in (("dense", False), ("vecch", True)):
        emu = D.emulator.__new__(D.emulator)
        emu.all_layer_set = []
        for s in range(int(g[f"emu_{tag}_nimp"])):
            layers = _snapshot_layers(g, f"emu_{tag}_S{s}_", lambda l, k: "matern2.5")
            for layer in layers:
                for node in layer:
                    node.vecch = vec
                    if not vec:
                        node.compute_stats()
            emu.all_layer_set.append(layers)
        emu.all_layer, emu.n_layer, emu.vecch = emu.all_layer_set[0], 2, vec
        mu, s2 = emu.loo(Xe, m=5)
        ref_mu, ref_var = g[f"emu_{tag}_mu"], g[f"emu_{tag}_var"]
        assert mu.shape == ref_mu.shape
        assert np.max(np.abs(mu - ref_mu)) <= 1e-7 * max(1.0, np.max(np.abs(ref_mu))), tag
        assert np.max(np.abs(s2 - ref_var)) <= 1e-7 * max(1.0, np.max(np.abs(ref_var))), tag
        for layers in emu.all_layer_set:   # the context restores the nodes
            assert all(node.vecch == vec and not node.loo_state for layer in layers for node in layer)
    big = D.emulator.__new__(D.emulator)
    big.vecch, big.all_layer = False, [[type("K", (), {"input": np.zeros((100, 1))})()]]
    with pytest.raises(NotImplementedError):
        big.loo(np.zeros((100, 1)))


def test_design_criteria_vs_reference(golden_metric):
    """emulator.metric (ALM / MICE / VIGF, emulation.py:323-420) on the reference's own imputed states: a 2-layer
    Matern and a 3-layer squared-exponential DGP with two outputs.  K^-1 is recomputed on the GPU (nugget 1e-6,
    cond ~1e10), hence the same absolute tolerance on the moments as the end-to-end predict test; the MICE score is
    a log ratio, compared where the reference's own variance is well above that floor."""
    import dgp_b200 as D

    g = golden_metric
    xc = g["x_cand"]
    for tag, name in (("ma2", "matern2.5"), ("se3", "sexp")):
        emu = _frozen_emulator(g, tag, name)
        alm = emu.metric(xc, method="ALM", score_only=True)
        assert alm.shape == g[f"{tag}_alm"].shape
        assert np.max(np.abs(alm - g[f"{tag}_alm"])) <= 2e-6, tag
        idx, val = emu.metric(xc, method="ALM")
        assert np.array_equal(idx, np.argmax(g[f"{tag}_alm"], axis=0)) and np.allclose(val, alm[idx, [0, 1]])
        solid = np.min(g[f"{tag}_mice_var"], axis=0) > 1e-4
        assert solid.sum() > 20
        for key, ns in (("mice", 1.0), ("mice_small", 1e-3)):
            mice = emu.metric(xc, method="MICE", nugget_s=ns, score_only=True)
            assert mice.shape == g[f"{tag}_{key}"].shape and np.all(np.isfinite(mice))
            assert np.max(np.abs(mice - g[f"{tag}_{key}"])[solid]) <= 1e-3, (tag, key)
            assert np.array_equal(np.argmax(mice, axis=0), np.argmax(g[f"{tag}_{key}"], axis=0)), (tag, key)
        model = type("M", (), {"X": g["X"]})()
        vigf = emu.metric(xc, method="VIGF", obj=model, score_only=True)
        ref = g[f"{tag}_vigf"]
        # squared biases reach O(1) here and inherit the 2e-6 absolute floor of the moments (see the docstring)
        floor = 2e-6 * max(1.0, float(np.max(g[f"{tag}_vigf_bias"])))
        assert np.max(np.abs(vigf - ref)) <= floor, tag
        idx, val = emu.pmetric(xc, method="VIGF", obj=model)
        assert np.array_equal(idx, g[f"{tag}_vigf_idx"])
        assert np.max(np.abs(val - g[f"{tag}_vigf_val"])) <= floor, tag
        with pytest.raises(Exception):
            emu.metric(xc, method="VIGF")
        with pytest.raises(Exception):
            emu.metric(xc[:, 0], method="ALM")


def test_warm_start_update_xy(golden_update):
    """dgp.update_xy (dgp.py:824-1095): the deterministic carry-over of the latent layers against the reference
    (conditional means at added rows through the gp / gp_vecch prediction kernels; row selection when the design
    shrinks), then the public call on all four branches."""
    import dgp_b200 as D

    g = golden_update
    for tag, vec, name in (("dense_ma", False, "matern2.5"), ("dense_se", False, "sexp"), ("vecch_se", True, "sexp")):
        model = D.dgp.__new__(D.dgp)
        model.all_layer = _snapshot_layers(g, f"{tag}_before_", lambda l, k: name)
        for layer in model.all_layer:
            for node in layer:
                node.vecch, node.m = vec, 6
        model.n_layer, model.vecch, model.m, model.block = 3, vec, 6, True
        model.check_rep, model.indices, model.nn_method, model.ord_fun = True, None, 'exact', None
        model.X, model.Y, model.n_data = g["X_new"], g["Y_new"], len(g["X_new"])
        model.update_all_layer_larger(g[f"{tag}_sub_idx"])
        tol = 1e-6 if not vec else 1e-8          # dense: R^-1 y at nugget 1e-6 (cond ~1e10)
        for stage in ("larger", "smaller"):
            if stage == "smaller":
                keep = g[f"{tag}_keep"]
                model.X, model.Y, model.n_data = model.X[keep], model.Y[keep], len(keep)
                model.update_all_layer_smaller(keep)
            for l, layer in enumerate(model.all_layer):
                for k, node in enumerate(layer):
                    p = f"{tag}_{stage}_L{l}K{k}_"
                    assert node.input.shape == g[p + "input"].shape, (tag, stage, l, k)
                    assert np.max(np.abs(node.input - g[p + "input"])) <= tol, (tag, stage, l, k)
                    assert np.max(np.abs(node.output - g[p + "output"])) <= tol, (tag, stage, l, k)
                    if node.connect is not None:
                        assert np.array_equal(node.global_input, g[p + "global_input"])
                    if vec:
                        assert node.NNarray.shape == (len(model.X), 7) and sorted(node.ord) == list(range(len(model.X)))
    # public call: grown design, shrunk design, unrelated design, reset
    np.random.seed(5)
    D.nb_seed(5)
    Xo, Yo, Xn, Yn = g["X_old"], g["Y_old"], g["X_new"], g["Y_new"]
    model = D.dgp(Xo, Yo)
    model.train(N=2, disable=True)
    latent_before = model.all_layer[0][0].output.copy()
    theta = model.all_layer[1][0].length.copy()
    model.update_xy(Xn, Yn)
    assert model.n_data == len(Xn) and model.all_layer[0][0].input.shape == Xn.shape
    assert model.all_layer[-1][0].output.shape == Yn.shape and np.array_equal(model.all_layer[-1][0].output, Yn)
    assert np.array_equal(model.all_layer[1][0].length, theta)          # hyper-parameters are carried over
    assert model.all_layer[0][0].output.shape == (len(Xn), 1) and latent_before.shape == (len(Xo), 1)
    model.train(N=1, disable=True)
    model.update_xy(Xn[:15], Yn[:15])
    assert model.all_layer[1][0].input.shape == (15, 2) and model.n_data == 15
    model.update_xy(Xo[:20] + 0.01, Yo[:20])
    assert model.all_layer[0][1].output.shape == (20, 1)
    model.update_xy(Xo, Yo, reset=True)
    assert np.array_equal(model.all_layer[1][0].length, model.all_layer[1][0].para_path[0, 1:-1])
    model.train(N=1, disable=True)
    assert model.N == 4 and np.all(np.isfinite(model.all_layer[1][0].para_path))
    with pytest.raises(Exception):
        model.update_xy(Xo[:, 0], Yo)


LIK_CASES = (("poi", "Poisson", 1, None), ("nb", "NegBin", 2, None), ("het", "Hetero", 2, None),
             ("catl", "Categorical", 1, "logit"), ("catp", "Categorical", 1, "probit"),
             ("cats", "Categorical", 3, "softmax"), ("catr", "Categorical", 3, "robustmax"),
             ("zip", "ZIP", 2, None), ("zinb", "ZINB", 3, None))


def _lik_model(g, prefix, likname, width, Y, stats=False, link=None):
    import dgp_b200 as D

    layers = _snapshot_layers(g, prefix, lambda l, k: "sexp" if l == 0 else "matern2.5")
    assert len(layers) == 2 and len(layers[1]) == width
    if stats:
        for layer in layers:
            for node in layer:
                node.compute_stats()
    if likname == "Categorical":
        lik = D.Categorical(num_classes=2 if width == 1 else width, input_dim=np.arange(width), link=link)
    else:
        lik = getattr(D, likname)(input_dim=np.arange(width))
    lik.output = Y.copy()
    lik.input = np.hstack([node.output for node in layers[1]])
    return layers + [[lik]], lik


def test_likelihood_layers_vs_reference(golden_lik):
    """Poisson / NegBin / Hetero final layers: device log-likelihood, ESS sweeps replayed with the
    reference's own draws through dgpb_ess_block_lik (Hetero: node-wise, the mean drawn from its exact Gaussian
    conditional by the shifted factorisation), and emulator predictions on the reference's imputed states."""
    from dgp_b200.imputation import _DeviceLayers
    import dgp_b200 as D

    g = golden_lik
    for tag, likname, width, link in LIK_CASES:
        p = f"{tag}_"
        layers, lik = _lik_model(g, p + "pre_", likname, width, g[p + "Y"], link=link)
        assert np.array_equal(lik.input, g[p + "lik_input_pre"])
        ref = float(g[p + "llik_pre"])
        assert abs(lik.llik() - ref) <= 1e-10 * abs(ref), tag
        dev = _DeviceLayers(layers)
        Z, U, SD = g[p + "Z"], g[p + "U"], g[p + "SD"]
        pad = np.full(4, 0.5)
        zi = ui = si = 0
        values = []
        for _ in range(int(g[p + "sweeps"])):
            nprop, th = dev.ess_call(0, [0, 1], list(range(width)), Z[zi:zi + 2], np.concatenate((U[ui:], pad)))
            zi += 2
            values.append(U[ui]); values.extend(th); ui += 1 + nprop
            if likname == "Hetero":
                dev.hetero_update(1, 0, 0, sd=SD[si])
                si += 1
                nprop, th = dev.lik_call(1, [1], [0], Z[zi:zi + 1], np.concatenate((U[ui:], pad)))
                zi += 1
            else:
                nprop, th = dev.lik_call(1, list(range(width)), [0], Z[zi:zi + width], np.concatenate((U[ui:], pad)))
                zi += width
            values.append(U[ui]); values.extend(th); ui += 1 + nprop
        assert zi == len(Z) and ui == len(U) and si == len(SD), (tag, ui, len(U))
        assert np.allclose(values, g[p + "draw_values"], rtol=1e-12, atol=0), tag
        dev.write_back()
        for l in range(2):
            for k, node in enumerate(layers[l]):
                assert relerr(node.output, g[f"{p}post_L{l}K{k}_output"], 1e-4) <= 1e-6, (tag, l, k)
        ref = float(g[p + "llik_post"])
        assert abs(lik.llik() - ref) <= 1e-6 * abs(ref), tag
        # predictions on the reference's imputed states
        emu = D.emulator.__new__(D.emulator)
        emu.all_layer_set = [_lik_model(g, f"{p}S{s}_", likname, width, g[p + "Y"], stats=True, link=link)[0]
                             for s in range(int(g[p + "nimp"]))]
        emu.all_layer, emu.n_layer, emu.vecch = emu.all_layer_set[0], 3, False
        # K-class probabilities are Monte-Carlo estimates from numpy's RNG (same seed and draw order as the fixture);
        # robustmax counts arg-max wins out of 1000 draws, so a latent difference of 1e-6 may move a count
        rtol = 5e-3 if link == "robustmax" else 2e-5
        np.random.seed(20261017 + 77)
        mu, var = emu.predict(g[p + "xt"])
        assert mu.shape == g[p + "mu"].shape
        assert np.max(np.abs(mu - g[p + "mu"]) / (1e-3 + np.abs(g[p + "mu"]))) <= rtol, tag
        assert np.max(np.abs(var - g[p + "var"]) / (1e-3 + np.abs(g[p + "var"]))) <= 2 * rtol, tag
        np.random.seed(20261017 + 77)
        mus, vars_ = emu.predict(g[p + "xt"], full_layer=True)
        assert len(mus) == 3
        assert np.max(np.abs(mus[-2] - g[p + "mu_full_gp"])) <= 5e-6 * max(1.0, np.max(np.abs(g[p + "mu_full_gp"]))), tag
        assert np.max(np.abs(mus[-1] - g[p + "mu_full_last"]) / (1e-3 + np.abs(g[p + "mu_full_last"]))) <= rtol, tag
        avg, per = emu.nllik(g[p + "xt_sorted"], g[p + "yt"])
        assert per.shape == g[p + "nllik"].shape
        assert np.max(np.abs(per - g[p + "nllik"])) <= 2e-5 * max(1.0, np.max(np.abs(g[p + "nllik"]))), tag
        assert abs(avg - float(g[p + "nllik_avg"])) <= 2e-5 * max(1.0, abs(float(g[p + "nllik_avg"]))), tag
        shuffled = np.arange(len(per))[::-1]     # scored point by point: the order of the test set does not matter
        _, per_rev = emu.nllik(g[p + "xt_sorted"][shuffled], g[p + "yt"][shuffled])
        assert np.allclose(per_rev, per[shuffled], rtol=1e-9), tag
        # design criteria act on the last GP layer (emulation.py:344-349, 373-413)
        model = type("M", (), {"X": g[p + "X"]})()
        alm = emu.metric(g[p + "xt"], method="ALM", score_only=True)
        assert alm.shape == g[p + "alm"].shape and np.max(np.abs(alm - g[p + "alm"])) <= 5e-6, tag
        mice = emu.metric(g[p + "xt"], method="MICE", score_only=True)
        solid = g[p + "alm"] > 1e-3
        assert solid.sum() >= 5 and np.max(np.abs(mice - g[p + "mice"])[solid]) <= 5e-3, tag
        vigf = emu.metric(g[p + "xt"], method="VIGF", obj=model, score_only=True)
        assert np.max(np.abs(vigf - g[p + "vigf"])) <= 1e-5 * max(1.0, np.max(g[p + "vigf"])), tag
        samples = emu.predict(g[p + "xt"], method="sampling", sample_size=3)
        n_out = g[p + "mu"].shape[1]
        assert len(samples) == n_out and samples[0].shape == (len(g[p + "xt"]), 3 * len(emu.all_layer_set))
        full = emu.predict(g[p + "xt"], method="sampling", sample_size=2, full_layer=True)
        assert len(full) == 3 and full[1][0].shape == (len(g[p + "xt"]), 2 * len(emu.all_layer_set))
        assert len(full[2]) == n_out
        if likname == "Categorical":
            per_imp_mu, _ = emu.predict(g[p + "xt"], aggregation=False)
            assert len(per_imp_mu) == len(emu.all_layer_set) and per_imp_mu[0].shape == g[p + "mu"].shape
            assert np.all(np.stack(samples, 0) >= 0) and np.all(np.stack(samples, 0) <= 1)


def test_vecchia_dgp_with_likelihood_layer(golden_lik):
    """Vecchia GP layers under a Poisson node through the public API: sparse prior draws feed dgpb_ess_block_lik,
    block likelihoods drive the middle pair, predictions run on the Vecchia kernels.  The Vecchia path of the
    reference draws every random number from numpy's global generator, and this implementation consumes that stream
    in the same order -- so the WHOLE run (construction, 3 SEM iterations, emulator with 2 imputations, prediction)
    reproduces the reference's, started from the same seed."""
    import dgp_b200 as D

    g = golden_lik
    seed = 21
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    D.nb_seed(seed)
    n, d = 300, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = rng.poisson(np.exp(1.0 + np.sin(3 * X[:, 0]) + X[:, 1])).astype(float).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([0.5]), name="sexp") for _ in range(d)]
    l2 = [D.kernel(length=np.array([0.5]), name="sexp", scale_est=True, connect=np.arange(d))]
    model = D.dgp(X, Y, D.combine(l1, l2, [D.Poisson()]), vecchia=True, m=12)
    model.train(N=3, disable=True)
    theta = np.concatenate([np.concatenate((k.scale, k.length, k.nugget)) for layer in model.all_layer[:-1]
                            for k in layer])
    assert np.allclose(theta, g["vpoi_theta"], rtol=1e-6), theta
    emu = D.emulator(model.estimate(), N=2)
    xt = rng.uniform(0, 1, size=(40, d))
    mu, var = emu.predict(xt, m=20)
    assert mu.shape == (40, 1) and np.all(np.isfinite(mu)) and np.all(var > 0)
    assert np.max(np.abs(mu - g["vpoi_mu"]) / g["vpoi_mu"]) <= 1e-6
    assert np.max(np.abs(var - g["vpoi_var"]) / g["vpoi_var"]) <= 1e-5


def test_single_gp_layer_under_likelihood(golden_lik):
    """One GP layer feeding a Poisson node: predictions and the 2-layer branches of the design criteria
    (emulation.py:362-372, 402-403) on the reference's imputed states."""
    import dgp_b200 as D

    g = golden_lik
    p = "poi2_"
    sets = []
    for s in range(int(g[p + "nimp"])):
        layers = _snapshot_layers(g, f"{p}S{s}_", lambda l, k: "sexp")
        assert len(layers) == 1 and len(layers[0]) == 1
        layers[0][0].compute_stats()
        lik = D.Poisson(input_dim=np.arange(1))
        lik.output, lik.input = g[p + "Y"].copy(), layers[0][0].output.copy()
        sets.append(layers + [[lik]])
    emu = D.emulator.__new__(D.emulator)
    emu.all_layer_set, emu.all_layer, emu.n_layer, emu.vecch = sets, sets[-1], 2, False
    xt = g[p + "xt"]
    # the fitted scale is ~3.6e4 (near-linear latent surface), so the latent variance scale (1 + nugget - r'R^-1 r)
    # ~ 0.04 carries the cancellation error of the bracket times the scale: two algebraically equal CPU evaluations
    # of the reference's formula differ by 4e-5 here.  Tolerances below are that floor, 2e-9 x scale.
    floor = 2e-9 * float(sets[0][0][0].scale[0])
    assert 1e-5 < floor < 1e-3
    mu, var = emu.predict(xt)
    assert np.max(np.abs(mu - g[p + "mu"]) / np.abs(g[p + "mu"])) <= 2e-5 + floor
    # Poisson variance = mean + (e^v - 1) mean^2: the latent-variance floor enters times mean^2
    assert np.all(np.abs(var - g[p + "var"]) <= 1e-4 * g[p + "var"] + 2 * floor * g[p + "mu"] ** 2)
    alm = emu.metric(xt, method="ALM", score_only=True)
    assert np.max(np.abs(alm - g[p + "alm"])) <= 5e-6 + floor
    vigf = emu.metric(xt, method="VIGF", obj=type("M", (), {"X": g[p + "X"]})(), score_only=True)
    assert np.max(np.abs(vigf - g[p + "vigf"])) <= 1e-3 * max(1.0, np.max(g[p + "vigf"]))
    assert np.array_equal(np.argmax(vigf, axis=0), np.argmax(g[p + "vigf"], axis=0))
    # 2-layer MICE: latent variance of ONE state over its smoothed variance
    mice = emu.metric(xt, method="MICE", score_only=True)
    assert mice.shape == g[p + "mice"].shape and np.all(np.isfinite(mice))
    assert np.max(np.abs(mice - g[p + "mice"]) / g[p + "mice"]) <= 1e-2


@pytest.mark.parametrize("likname", ["Poisson", "NegBin", "Hetero", "Categorical2", "Categorical3", "ZIP", "ZINB"])
def test_likelihood_public_api(likname):
    """dgp(X, Y, combine(..., [likelihood])) -> train -> estimate -> emulator -> predict through the public API."""
    import dgp_b200 as D

    rng = np.random.default_rng(11)
    np.random.seed(11)
    D.nb_seed(11)
    n, d = 40, 2
    X = rng.uniform(0, 1, size=(n, d))
    gx = np.sin(3 * X[:, 0]) + X[:, 1]
    width = {"Poisson": 1, "Categorical2": 1, "Categorical3": 3, "ZINB": 3}.get(likname, 2)
    if likname == "Hetero":
        Y = (gx + np.exp(-1.5 + X[:, 0]) * rng.standard_normal(n)).reshape(-1, 1)
    elif likname == "Categorical2":
        Y = np.where(gx > 0.9, "yes", "no").reshape(-1, 1)             # labels go through the class encoder
    elif likname == "Categorical3":
        Y = np.digitize(gx, [0.6, 1.2]).reshape(-1, 1) + 5
    else:
        Y = rng.poisson(np.exp(1.0 + gx)).astype(float).reshape(-1, 1)
        if likname in ("ZIP", "ZINB"):
            Y[rng.uniform(size=n) < 0.3] = 0.0
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(d)]
    l2 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, connect=np.arange(d)) for _ in range(width)]
    make = (lambda: D.Categorical()) if likname.startswith("Categorical") else getattr(D, likname)
    model = D.dgp(X, Y, D.combine(l1, l2, [make()]))
    assert model.all_layer[-1][0].input.shape == (n, width)
    if likname.startswith("Categorical"):
        top = model.all_layer[-1][0]
        assert top.num_classes == (2 if width == 1 else 3) and top.link == ("logit" if width == 1 else "softmax")
        assert set(np.unique(top.output)) == set(range(top.num_classes))
        assert all(k.scale[0] == 1.0 for k in model.all_layer[-2])     # the scale of the first burn-in is restored
    model.train(N=3, disable=True)
    assert model.all_layer[1][0].para_path.shape[0] == 4
    emu = D.emulator(model.estimate(), N=2)
    xt = rng.uniform(0, 1, size=(12, d))
    mu, var = emu.predict(xt)
    n_out = 3 if likname == "Categorical3" else 1
    assert mu.shape == (12, n_out) and np.all(np.isfinite(mu)) and np.all(var >= 0)
    if likname != "Hetero":
        assert np.all(mu > 0)          # a count mean / class probabilities
    if likname == "Categorical3":
        assert np.allclose(mu.sum(1), 1.0, atol=1e-9)
    idx, score = emu.metric(xt)
    assert idx.shape == (width,) and np.all(score > 0)
    yt = np.abs(np.round(mu[:, [0]])) if not likname.startswith("Categorical") else (np.arange(12) % 2).reshape(-1, 1)
    avg, per = emu.nllik(xt, yt)
    assert per.shape == (12,) and np.isfinite(avg)
    with pytest.raises(Exception):
        D.dgp(X, Y, D.combine(l1, [D.kernel(length=np.array([1.0]))] * (1 if width == 2 else 2), [make()]))


def test_gp_design_criteria_and_update(golden_metric):
    """gp.metric (ALM / MICE / VIGF, gp.py:271-324) and gp.update_xy (gp.py:144-181), dense and Vecchia."""
    import dgp_b200 as D

    g = golden_metric
    xc = g["x_cand"]
    for tag, name in (("se", "sexp"), ("ma", "matern2.5")):
        for vtag, vec in (("dense", False), ("vecch", True)):
            q = f"gp_{tag}_{vtag}_"
            em = D.gp(g["gp_X"], g["gp_Y"], D.kernel(length=np.array([0.6, 0.8]), scale=1.2, nugget=1e-4, name=name),
                      vecchia=vec, m=10)
            tol = 1e-7 if not vec else 1e-9        # dense: R^-1 at nugget 1e-4
            for method, key in (("ALM", "alm"), ("MICE", "mice"), ("VIGF", "vigf")):
                score = em.metric(xc, method=method, score_only=True, m=12)
                ref = g[q + key]
                assert score.shape == ref.shape, (q, key)
                assert np.max(np.abs(score - ref)) <= tol * max(1.0, np.max(np.abs(ref))), (q, key)
                idx, val = em.pmetric(xc, method=method, m=12)
                assert idx[0] == np.argmax(ref[:, 0]) and abs(val[0] - ref.max()) <= tol * max(1.0, ref.max())
            em.update_xy(g["gp_X2"], g["gp_Y2"])
            assert em.n_data == len(g["gp_X2"]) and em.kernel.input.shape == g["gp_X2"].shape
            mu, var = em.ppredict(xc, m=12)
            assert np.max(np.abs(mu - g[q + "upd_mu"])) <= 10 * tol and np.max(np.abs(var - g[q + "upd_var"])) <= 10 * tol
            em.update_xy(g["gp_X"], g["gp_Y"], reset=True)
            assert np.array_equal(em.kernel.length, np.array([0.6, 0.8]))
    with pytest.raises(Exception):
        em.metric(xc[:, 0])


def test_public_api_train_and_predict_smoke():
    """The user-facing path runs: dgp(X,Y).train -> estimate -> emulator -> predict; the fit is sane."""
    import dgp_b200 as D

    np.random.seed(7)
    D.nb_seed(7)
    X = np.linspace(0, 1, 12)[:, None]
    Y = np.where(X > 0.5, 1.0, -1.0)
    model = D.dgp(X, Y)
    model.train(N=8, disable=True)
    assert model.N == 8 and model.all_layer[0][0].para_path.shape[0] == 9
    emu = D.emulator(model.estimate(), N=2)
    mu, var = emu.predict(np.linspace(0, 1, 33)[:, None])
    assert mu.shape == (33, 1) and var.shape == (33, 1)
    assert np.all(np.isfinite(mu)) and np.all(var >= 0)
    assert np.mean(np.sign(mu[:, 0]) == np.sign(np.linspace(0, 1, 33) - 0.5 + 1e-9)) > 0.8
    with pytest.raises(Exception):
        emu.predict(np.linspace(0, 1, 5))


def test_vecchia_dgp_public_api_roundtrip(tmp_path):
    """Vecchia DGP through the public API (config-4 shape, small): train (ordered kNN, sparse prior draws, block
    likelihoods, Vecchia M-step), emulator, predict, sampling, LOO, write/read; predictions survive pickling."""
    import dgp_b200 as D

    rng = np.random.default_rng(12)
    np.random.seed(12)
    D.nb_seed(12)
    n, d = 400, 3
    X = rng.uniform(0, 1, (n, d))
    f = lambda x: np.sin(2 * np.pi * x[:, 0] * x[:, 1]) + x[:, 2] ** 2
    Y = (f(X) + 0.02 * rng.standard_normal(n)).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(d)]
    l2 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, nugget_est=True, nugget=1e-2,
                   connect=np.arange(d))]
    model = D.dgp(X, Y, D.combine(l1, l2), vecchia=True, m=12)
    model.train(N=4, disable=True)
    assert model.N == 4 and model.timing["i_step"] > 0 and model.timing["m_step"] > 0
    emu = D.emulator(model.estimate(burnin=0), N=2)
    xt = rng.uniform(0, 1, (257, d))
    mu, var = emu.predict(xt, m=12)
    assert mu.shape == (257, 1) and np.all(np.isfinite(mu)) and np.all(var > 0)
    assert np.sqrt(np.mean((mu[:, 0] - f(xt)) ** 2)) < 0.5
    emu.group_first_layer = False      # one launch per node instead of the multi-node first-layer kernel
    mu1, var1 = emu.predict(xt, m=12)
    emu.group_first_layer = True
    assert relerr(mu, mu1, 1e-6) <= 1e-9 and relerr(var, var1, 1e-300) <= 1e-9
    draws = emu.predict(xt[:9], method="sampling", sample_size=5, m=12)
    assert len(draws) == 1 and draws[0].shape == (9, 10)
    lmu, lvar = emu.loo(X, m=10)
    assert lmu.shape == (n, 1) and np.all(np.isfinite(lmu)) and np.all(lvar > 0)
    D.write(emu, str(tmp_path / "emu"))
    emu2 = D.read(str(tmp_path / "emu"))
    mu2, var2 = emu2.predict(xt, m=12)
    assert np.array_equal(mu, mu2) and np.array_equal(var, var2)


def test_knn_properties_at_config4_scale():
    """BASELINE config-4 sized neighbour search (n = 100k training points, d = 10, m = 25): every row is sorted by
    distance, sampled rows equal a numpy brute-force search bit for bit, the ordered search only returns earlier
    points in descending index order, and no query needed a different answer from the scalar exact kernel."""
    from dgp_b200 import vecchia as V

    rng = np.random.default_rng(77)
    n, M, d, m = 100000, 20000, 10, 25
    x, q = rng.uniform(0, 1, (n, d)), rng.uniform(0, 1, (M, d))
    NN = V.get_pred_nn(q, x, m)
    assert NN.shape == (M, m) and NN.min() >= 0 and NN.max() < n
    rows = rng.choice(M, 48, replace=False)
    for r in rows:
        d2 = np.zeros(n)
        for k in range(d):          # the reference arithmetic: ascending dimension, no FMA
            d2 = d2 + (q[r, k] - x[:, k]) * (q[r, k] - x[:, k])
        order = np.lexsort((np.arange(n), d2))[:m]
        assert np.array_equal(NN[r], order), r
        assert np.all(np.diff(d2[NN[r]]) >= 0)
    sub = rng.choice(M, 2000, replace=False)
    dsub = ((q[sub, None, :] - x[NN[sub]]) ** 2).sum(-1)
    assert np.all(np.diff(dsub, axis=1) >= -1e-15)
    NNo = V.nn(x[:30000], m)
    assert NNo.shape == (30000, m + 1) and np.array_equal(NNo[:, 0], np.arange(30000))
    body = NNo[:, 1:]
    assert np.all((body < np.arange(30000)[:, None]) | (body == -1))
    valid = body >= 0
    assert np.array_equal(valid.sum(1), np.minimum(np.arange(30000), m))
    assert np.all((np.diff(body, axis=1) < 0) | ~valid[:, 1:])


def test_property_checks_at_baseline_scale():
    """Size-independent properties at a BASELINE-sized node (n=2000, D=10): the factor reproduces K, the
    inverse is an inverse, log-likelihood agrees with an independent FP64 computation (torch/cuSOLVER used
    as a CHECKER only)."""
    import ctypes
    import dgp_b200 as D
    from dgp_b200 import _lib as L

    rng = np.random.default_rng(99)
    n, d = 2000, 10
    k = D.kernel(length=np.full(d, 1.3), name="matern2.5", nugget=1e-4, scale=1.7)
    k.input = rng.uniform(0, 1, (n, d))
    k.output = np.sin(k.input.sum(1, keepdims=True))
    K = torch.from_numpy(k.k_matrix()).cuda()
    Lt = torch.linalg.cholesky(1.7 * K)
    y = torch.from_numpy(k.output).cuda()
    w = torch.linalg.solve_triangular(Lt, y, upper=False)
    ll_ref = float(-0.5 * (2 * torch.log(torch.diagonal(Lt)).sum() + (w * w).sum()))
    ll = k.log_likelihood_func()
    assert abs(ll - ll_ref) <= 1e-9 * abs(ll_ref), (ll, ll_ref)
    k.compute_stats()
    Ri = k._Rinv
    resid = (K @ Ri - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item()
    assert resid <= 1e-8, resid
    # linearity of the predictor in y: predicting with 2y doubles the mean, leaves the variance
    xt = rng.uniform(0, 1, (257, d))
    m1, v1 = k.gp_prediction(xt, None)
    k.output = 2 * k.output
    k.compute_stats()
    m2, v2 = k.gp_prediction(xt, None)
    assert relerr(m2, 2 * m1, 1e-6) <= 1e-8 and np.max(np.abs(v2 - v1)) <= 1e-9 * 1.7


# ------------------------------------------------------------------------------------------------ headline shape
def _headline_node(rng, n=5000, scale_est=False):
    """One GP node of BASELINE config 3's upper layers: 8 latent inputs + 8 connected global inputs, n = 5000,
    squared exponential, one length-scale; nugget 1e-3 (the conditioning at which 1e-9 is the target)."""
    import dgp_b200 as D

    k = D.kernel(length=np.array([1.1]), name="sexp", nugget=1e-3, scale=1.3, scale_est=scale_est,
                 connect=np.arange(8))
    k.input = rng.uniform(0, 1, (n, 8))
    k.global_input = rng.uniform(0, 1, (n, 8))
    k.input_dim = np.arange(8)
    X = np.concatenate((k.input, k.global_input), 1)
    k.output = (np.sin(X.sum(1)) + X[:, 0] * X[:, 9] + 0.05 * rng.standard_normal(n)).reshape(-1, 1)
    k.D = 16
    k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
    return k, X


def test_headline_shape_likelihood_gradient_statistics():
    """n = 5000, D = 16 with the PRODUCTION tunables (K = 512 hyper-blocks, graded ramp, critical-path stream): plain
    layout (ESS log-likelihood) and augmented layout (M-step objective + gradient, K^-1) against LAPACK (the oracle)
    to BASELINE.json's 1e-9."""
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(5000)
    k, X = _headline_node(rng, scale_est=True)
    ll = k.log_likelihood_func()
    ll0 = O.loglik_dense(X, k.output, k.length, 1.3, 1e-3, "sexp")
    assert abs(ll - ll0) <= 1e-9 * abs(ll0), (ll, ll0)
    k.prior_name = None
    f, gr = k.llik(k.log_t().copy())
    f0, g0, s0 = O.nllik_grad_dense(X, k.output, k.length, 1.3, 1e-3, "sexp", True, False)
    assert abs(f[0] - f0) <= 1e-9 * abs(f0), (f, f0)
    assert np.max(np.abs(gr - g0)) <= 1e-9 * max(1.0, np.max(np.abs(g0))), (gr, g0)
    assert abs(k.scale[0] - s0) <= 1e-9 * s0
    k.compute_stats()
    Kd = torch.from_numpy(O.k_matrix(X, k.length, 1e-3, "sexp")).cuda()
    resid = (Kd @ k._Rinv - torch.eye(len(X), dtype=torch.float64, device="cuda")).abs().max().item()
    assert resid <= 1e-9, resid
    from scipy.linalg import cho_factor, cho_solve
    a0 = cho_solve(cho_factor(Kd.cpu().numpy(), lower=True), k.output[:, 0])
    assert relerr(k.Rinv_y, a0, 1e-3 * np.max(np.abs(a0))) <= 1e-9


def test_headline_shape_ess_wave_of_eight():
    """One blocked ESS update of the config-3 layer pair (8 latent nodes feeding 8 upper nodes, n = 5000: a wave is 8
    matrices of one candidate angle) checked against LAPACK: the threshold, the likelihood sums at the accepted and at
    the last rejected angle decide exactly as the device did, and the accepted latent columns are
    f cos(theta) + chol(K) z sin(theta) to 1e-9."""
    import dgp_b200 as D
    from dgp_b200.imputation import _DeviceLayers
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(5001)
    n, d, w = 5000, 8, 8
    Xg = rng.uniform(0, 1, (n, d))
    F = np.stack([np.sin(3 * Xg[:, j] + 0.3 * j) + 0.3 * Xg[:, (j + 1) % d] for j in range(w)], 1)
    Yup = np.stack([np.sin(F.sum(1) + j) + 0.1 * Xg[:, j] for j in range(w)], 1)

    def build(make):
        l1 = [make(length=np.array([0.9]), name="sexp", nugget=1e-3) for _ in range(w)]
        l2 = [make(length=np.array([1.2 + 0.05 * j]), name="sexp", nugget=1e-3, connect=np.arange(d)) for j in range(w)]
        for j, nd in enumerate(l1):
            nd.input, nd.input_dim, nd.output = Xg.copy(), np.arange(d), F[:, [j]].copy()
        for j, nd in enumerate(l2):
            nd.input, nd.input_dim, nd.global_input, nd.output = F.copy(), np.arange(w), Xg.copy(), Yup[:, [j]].copy()
        return [l1, l2]

    layers = build(lambda **kw: D.kernel(**kw))
    for layer in layers:
        for nd in layer:
            nd.D = nd.input.shape[1] + (0 if nd.global_input is None else nd.global_input.shape[1])
            nd.vecch = False
    z, u = rng.standard_normal((w, n)), rng.uniform(size=64)
    dev = _DeviceLayers(layers)
    nprop, thetas = dev.ess_call(0, list(range(w)), list(range(w)), z, u)
    dev.write_back()
    got = np.concatenate([nd.output for nd in layers[0]], 1)
    # ---- LAPACK side
    ref = build(lambda **kw: O.Node(**kw))
    Lc = np.linalg.cholesky(O.k_matrix(Xg, np.array([0.9]), 1e-3, "sexp"))   # the 8 targets share inputs and theta
    nu = Lc @ z.T

    def upper_sum(fp):
        s = 0.0
        for nd in ref[1]:
            nd.input = fp
            s += nd.loglik()
        return s

    log_y = upper_sum(F) + np.log(u[0])
    # replay the bracket rule with the device's angles: uniforms consumed = 1 + proposals, angles as the rule draws them
    th, lo, hi = 2 * np.pi * u[1], 2 * np.pi * u[1] - 2 * np.pi, 2 * np.pi * u[1]
    for i in range(nprop):
        assert abs(thetas[i] - th) <= 1e-12 * max(1.0, abs(th)), i
        if th < 0:
            lo = th
        else:
            hi = th
        th = lo + (hi - lo) * u[2 + i]
    acc = F * np.cos(thetas[-1]) + nu * np.sin(thetas[-1])
    assert upper_sum(acc) > log_y                      # the accepted angle is accepted by LAPACK too
    if nprop > 1:                                      # and the one before it rejected
        rej = F * np.cos(thetas[-2]) + nu * np.sin(thetas[-2])
        assert not (upper_sum(rej) > log_y)
    assert relerr(got, acc, 1e-3) <= 1e-9


def test_headline_shape_linked_prediction():
    """link_gp (squared exponential, DMMA exponent kernel) for 64 Gaussian test inputs at n = 5000, Dw = 8 latent +
    Dz = 8 global dimensions against the oracle's closed form."""
    from oracle import dgp_oracle as O

    rng = np.random.default_rng(5002)
    k, X = _headline_node(rng)
    k.compute_stats()
    M = 64
    mt, vt, zt = rng.uniform(0, 1, (M, 8)), rng.uniform(1e-4, 0.05, (M, 8)), rng.uniform(0, 1, (M, 8))
    m2, v2 = k.linkgp_prediction(mt, vt, zt)
    Rinv, Rinv_y = O.compute_stats(X, k.output, k.length, 1e-3, "sexp")
    R2, P = O.sexp_stats(k.input, k.length)
    m0, v0 = O.link_gp(mt, vt, zt, k.input, k.global_input, Rinv, Rinv_y, R2, P, 1.3, k.length, 1e-3, "sexp")
    assert relerr(m2, m0, 1e-3) <= 1e-9
    assert np.max(np.abs(v2 - v0)) <= 1e-9 * 1.3


# ------------------------------------------------------------------------------------------------ sampling
def test_sampling_reproduces_the_reference_draws():
    """method='sampling'  with numpy's global generator seeded as the fixture script seeded it: the
    samples are the reference's, draw for draw (same order of normal variates, moments from the GPU)."""
    import dgp_b200 as D
    from conftest import load_golden

    g = load_golden("sampling")
    emu = _frozen_emulator(g, "emu", "sexp")
    np.random.seed(4242)
    last = np.asarray(emu.predict(g["emu_xt"], method="sampling", sample_size=4))
    assert last.shape == g["emu_last"].shape
    # a sample is mu + sqrt(var) z: near-zero predictive variances (e2e floor 8e-8 absolute) pass through a square root
    assert np.max(np.abs(last - g["emu_last"])) <= 5e-5 * max(1.0, np.max(np.abs(g["emu_last"])))
    np.random.seed(4243)
    full = emu.predict(g["emu_xt"], method="sampling", sample_size=2, full_layer=True)
    for l, per in enumerate(full):
        ref = g[f"emu_full_L{l}"]
        assert np.asarray(per).shape == ref.shape, l
        assert np.max(np.abs(np.asarray(per) - ref)) <= 5e-5 * max(1.0, np.max(np.abs(ref))), l
    gp1 = D.gp(g["gp_X"], g["gp_Y"], D.kernel(length=np.array([0.7, 0.9]), scale=1.3, nugget=1e-4, name="matern2.5"))
    np.random.seed(4244)
    s = gp1.predict(g["emu_xt"], method="sampling", sample_size=5)
    assert s.shape == g["gp_samples"].shape and np.max(np.abs(s - g["gp_samples"])) <= 1e-7
    # linked system on the frozen imputations of e2e.npz
    system = _frozen_linked_system(load_golden("e2e"))
    np.random.seed(4245)
    ls = np.asarray(system.predict(load_golden("e2e")["lgp_xt"], method="sampling", sample_size=3)[0])
    assert ls.shape == g["lgp_last"].shape
    assert np.max(np.abs(ls - g["lgp_last"])) <= 5e-5 * max(1.0, np.max(np.abs(g["lgp_last"])))
    np.random.seed(4246)
    fl = system.predict(load_golden("e2e")["lgp_xt"], method="sampling", sample_size=2, full_layer=True)
    for l, per in enumerate(fl):
        ref = g[f"lgp_full_L{l}"]
        assert np.asarray(per[0]).shape == ref.shape, l
        assert np.max(np.abs(np.asarray(per[0]) - ref)) <= 5e-5 * max(1.0, np.max(np.abs(ref))), l


# ------------------------------------------------------------------------------------------------ advisor findings
def test_m_step_with_more_nodes_than_one_batched_launch():
    """A DGP may have any number of dense GP nodes: 16 + 16 + 2 = 34 nodes exceed the library's 32 matrices per batched
    launch, so the M-step's rendezvous has to serve them in chunks (the reference has no such limit)."""
    import dgp_b200 as D

    rng = np.random.default_rng(34)
    np.random.seed(34)
    D.nb_seed(34)
    n, d = 48, 16
    X = rng.uniform(0, 1, (n, d))
    Y = np.stack([np.sin(X.sum(1)), np.cos(X[:, 0] * 3)], 1)
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(16)]
    l2 = [D.kernel(length=np.array([1.0]), name="sexp", connect=np.arange(2)) for _ in range(16)]
    l3 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, connect=np.arange(2)) for _ in range(2)]
    model = D.dgp(X, Y, [l1, l2, l3])
    model.train(2, disable=True)
    for layer in model.all_layer:
        for node in layer:
            assert node.para_path.shape[0] == 3 and np.all(np.isfinite(node.para_path))


def test_vecchia_mode_switches_skip_likelihood_nodes():
    """emulator.to_vecchia / remove_vecchia / loo and container.to_vecchia / remove_vecchia (lgp.set_vecchia) on
    structures that hold likelihood nodes (emulation.py:62-107, linkgp.py:64-90 guard with kernel.type == 'gp')."""
    import dgp_b200 as D

    rng = np.random.default_rng(7)
    np.random.seed(7)
    D.nb_seed(7)
    n = 60
    X = rng.uniform(0, 1, (n, 2))
    Y = rng.poisson(np.exp(1.0 + np.sin(3 * X[:, 0]) + X[:, 1])).astype(float).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([0.5]), name="sexp") for _ in range(2)]
    l2 = [D.kernel(length=np.array([0.5]), name="sexp", scale_est=True, connect=np.arange(2))]
    model = D.dgp(X, Y, D.combine(l1, l2, [D.Poisson()]))
    model.train(2, disable=True)
    emu = D.emulator(model.estimate(), N=2)
    xt = rng.uniform(0, 1, (9, 2))
    mu0, var0 = emu.predict(xt)
    emu.to_vecchia()
    assert emu.vecch and all(k.vecch for one in emu.all_layer_set for layer in one for k in layer if k.type == 'gp')
    mu1, _ = emu.predict(xt, m=n)          # conditioning on every point: the Vecchia form equals the dense one
    assert np.max(np.abs(mu1 - mu0)) <= 1e-5 * max(1.0, np.max(np.abs(mu0)))
    emu.remove_vecchia()
    mu2, var2 = emu.predict(xt)
    assert np.max(np.abs(mu2 - mu0)) <= 1e-9 * max(1.0, np.max(np.abs(mu0)))
    assert np.max(np.abs(var2 - var0)) <= 1e-9 * max(1.0, np.max(np.abs(var0)))
    cont = D.container(model.estimate(), np.array([0, 1]))
    cont.to_vecchia()
    assert cont.vecch
    cont.remove_vecchia()
    assert not cont.vecch and all(not k.vecch for layer in cont.structure for k in layer if k.type == 'gp')


# ------------------------------------------------------------------------------------------------ small models
def test_small_model_istep_in_one_launch(golden_ess):
    """`dgpb_ess_sweeps_small` (n <= 64: all sweeps of an I-step in ONE kernel launch, csrc/ess_small.cu) replays the
    reference's sweeps from the fixture with its own normals and uniforms: it consumes exactly the reference's
    uniforms (identical accept / shrink decisions) and ends in the reference's latent layers."""
    from dgp_b200.imputation import _DeviceLayers

    g = golden_ess
    done = 0
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        widths, name, vecch = [int(w) for w in g[p + "widths"]], str(g[p + "name"]), bool(g[p + "vecch"])
        if vecch:
            continue
        layers = _load_layers(g, p + "pre_", widths, name, vecch)
        dev = _DeviceLayers(layers)
        assert dev.small_ok(True)
        Z, U, sweeps = g[p + "Z"], g[p + "U"], int(g[p + "sweeps"])
        nprop = dev.sweeps_small(sweeps, z=Z, u=np.concatenate((U, np.full(8, 0.5))))
        assert dev.last_uniforms == len(U), (ci, dev.last_uniforms, len(U))
        assert nprop == len(U) - sweeps * (len(widths) - 1), ci      # one uniform per proposal + one per threshold
        dev.write_back()
        post = _load_layers(g, p + "post_", widths, name, vecch)
        for l in range(len(widths)):
            for k in range(widths[l]):
                assert floor_ok(f"ess_c{ci}_post_L{l}K{k}", layers[l][k].output, post[l][k].output), (ci, l, k)
        done += 1
    assert done == 2


def test_small_model_path_equals_general_path(monkeypatch):
    """The public API takes the one-launch path for small models: same seeds, same chain as the general (wave) path --
    same number of proposals, same uniforms consumed, same imputed layers to rounding."""
    import dgp_b200 as D

    def run(small):
        monkeypatch.setenv("DGPB_ESS_SMALL", "1" if small else "0")
        np.random.seed(77)
        D.nb_seed(77)
        X = np.linspace(0, 1, 10)[:, None]
        Y = np.array([[-1.0] if i < 0.5 else [1.0] for i in X[:, 0]])
        layers = [[D.kernel(length=np.array([1.0]), name="sexp")], [D.kernel(length=np.array([1.0]), name="sexp")],
                  [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True)]]
        model = D.dgp(X, Y, D.combine(*layers))
        first = [[k.output.copy() for k in layer] for layer in model.all_layer]   # after the 10 sweeps of the constructor
        model.train(5, disable=True)
        return model, first, np.random.uniform(size=2)

    a, fa, ra = run(True)
    b, fb, rb = run(False)
    assert a.imp.n_proposals == b.imp.n_proposals and np.array_equal(ra, rb)
    for la, lb in zip(fa, fb):          # the I-step alone: rounding-level agreement
        for ka, kb in zip(la, lb):
            assert np.allclose(ka, kb, rtol=1e-9, atol=1e-11)
    # five SEM iterations later (n = 10, nugget 1e-6: the M-step amplifies the last bits) the chains still coincide
    for la, lb in zip(a.all_layer, b.all_layer):
        for ka, kb in zip(la, lb):
            assert np.allclose(ka.output, kb.output, rtol=1e-4, atol=1e-6)
            assert np.allclose(ka.para_path, kb.para_path, rtol=1e-3, atol=1e-6)


# ------------------------------------------------------------------------------------------------ Hetero + Vecchia
def test_hetero_exact_draw_under_vecchia():
    """the mean process of a Hetero likelihood under the Vecchia approximation
    (`dgpb_hetero_vecchia_draw`): conditioning sets bit-exact (kernel.ord_nn(pointer=True)), the sparse U and the draw
    with the reference's recorded normals against the reference fixture."""
    import dgp_b200 as D
    from conftest import load_golden
    from dgp_b200 import _lib as L

    g = load_golden("hetvecch")
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        X, o, name = g[p + "X"], g[p + "ord"], str(g[p + "name"])
        n, m = len(X), int(g[p + "m"])
        k = D.kernel(length=g[p + "length"].copy(), scale=g[p + "scale"][0], nugget=1e-6, name=name)
        k.input, k.output = X.copy(), g[p + "y"].reshape(-1, 1).copy()
        k.vecch, k.m = True, m
        k.ord_nn(ord=o, NNarray=np.zeros((n, m + 1), dtype=np.int64), pointer=True)
        assert np.array_equal(k.imp_NNarray, g[p + "imp_NN"]), ci
        Xo, NN = L.to_dev(np.ascontiguousarray(X[o])), L.to_dev(g[p + "imp_NN"], np.int64)
        gam, yo, sd = L.to_dev(g[p + "gamma"][o]), L.to_dev(g[p + "y"][o]), L.to_dev(g[p + "sd"])
        f, U = L.empty((n,)), L.empty(tuple(g[p + "imp_NN"].shape))
        larr, lptr = L.length_host(k.length)
        L.check(L.load().dgpb_hetero_vecchia_draw(L.ptr(Xo), L.ptr(NN), n, X.shape[1], NN.shape[1], lptr, len(larr),
                                                  float(k.scale[0]), L.KIND[name], L.ptr(gam), L.ptr(yo), L.ptr(sd),
                                                  L.ptr(f), L.ptr(U), L.stream()))
        Uref = g[p + "U_rev"][:, ::-1]
        assert relerr(U.cpu().numpy(), Uref, 1e-3 * np.max(np.abs(Uref))) <= 1e-6, ci   # cond ~ 1e10 rows, see the oracle test
        fr = f.cpu().numpy()[np.argsort(o)]
        assert relerr(fr, g[p + "f"], 1e-3 * np.max(np.abs(g[p + "f"]))) <= 1e-6, ci
        # the reference-shaped call (device vectors in data order)
        lik = D.Hetero(input_dim=np.arange(2))
        f2 = lik.posterior_vecch_dev(k, n, L.to_dev(np.log(g[p + "gamma"])), L.to_dev(g[p + "y"]), g[p + "sd"])
        assert relerr(f2.cpu().numpy(), g[p + "f"], 1e-3 * np.max(np.abs(g[p + "f"]))) <= 1e-6, ci


def test_vecchia_dgp_with_hetero_layer():
    """A whole public-API run of a Vecchia DGP under a Hetero likelihood from one seed (under Vecchia numpy's
    generator drives every draw): hyper-parameters, latent layer and predictions follow the reference's run."""
    import dgp_b200 as D
    from conftest import load_golden

    g = load_golden("hetvecch")
    seed = 33
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    D.nb_seed(seed)
    n, d = 200, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = (np.sin(4 * X[:, 0]) + X[:, 1] + np.exp(-1.5 + X[:, 0]) * rng.standard_normal(n)).reshape(-1, 1)
    l1 = [D.kernel(length=np.array([0.5]), name="sexp") for _ in range(d)]
    l2 = [D.kernel(length=np.array([0.5]), name="sexp", scale_est=True, connect=np.arange(d)) for _ in range(2)]
    model = D.dgp(X, Y, D.combine(l1, l2, [D.Hetero()]), vecchia=True, m=10)
    assert model.all_layer[1][0].imp_NNarray is not None and model.all_layer[1][1].imp_NNarray is None
    model.train(N=3, disable=True)
    theta = np.concatenate([np.concatenate((k.scale, k.length, k.nugget)) for layer in model.all_layer[:-1]
                            for k in layer])
    assert np.allclose(theta, g["api_theta"], rtol=1e-5), (theta, g["api_theta"])
    latent = np.concatenate([k.output for k in model.all_layer[1]], 1)
    assert np.allclose(latent, g["api_latent"], rtol=1e-5, atol=1e-6)
    emu = D.emulator(model.estimate(), N=2)
    xt = rng.uniform(0, 1, size=(30, d))
    mu, var = emu.predict(xt, m=15)
    assert np.max(np.abs(mu - g["api_mu"])) <= 1e-5 * max(1.0, np.max(np.abs(g["api_mu"])))
    assert np.max(np.abs(var - g["api_var"]) / g["api_var"]) <= 1e-4
