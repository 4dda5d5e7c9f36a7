"""Generate golden input/output vectors from the UNMODIFIED reference (`/root/reference/dgpsi`).

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py [dense jd vecchia ess e2e loo metric update lik floor sampling hetvecch]

Writes `tests/golden/*.npz` (one file per group: dense_nodes, jd, vecchia_nodes, ess_replay, e2e, loo, metric,
update, likelihood; every group is seeded, re-running reproduces the committed files bit for bit).  Every fixture stores the exact inputs fed to the reference function
together with what it returned, so `tests/` can check (a) the oracle (`oracle/dgp_oracle.py`) on CPU
and (b) the CUDA path on a B200 against the reference's own numbers.  The reference ships no tests of
its own so these files ARE the pin.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import import_reference  # noqa: E402

dgpsi = import_reference()
from dgpsi import functions as F  # noqa: E402
from dgpsi import imputation as IMP  # noqa: E402
from dgpsi import vecchia as V  # noqa: E402
from dgpsi.kernel_class import kernel  # noqa: E402

SEED = 20261017


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def make_node(rng, n, d_loc, d_glob, name, ard, nugget=1e-6, scale=1.0, nugget_est=False, scale_est=False):
    D = d_loc + d_glob
    length = rng.uniform(0.4, 1.6, size=D if ard else 1)
    k = kernel(length=length, scale=scale, nugget=nugget, name=name, nugget_est=nugget_est, scale_est=scale_est,
               connect=np.arange(d_glob) if d_glob else None)
    k.input = rng.uniform(0, 1, size=(n, d_loc))
    k.input_dim = np.arange(d_loc)
    if d_glob:
        k.global_input = rng.uniform(0, 1, size=(n, d_glob))
    X = k.input if not d_glob else np.concatenate((k.input, k.global_input), 1)
    k.output = (np.sin(3 * X.sum(1)) + 0.3 * rng.standard_normal(n)).reshape(-1, 1)
    k.D = D
    k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
    return k, X


# ---------------------------------------------------------------- 1+2: dense kernels / likelihoods
def gen_dense():
    rng = np.random.default_rng(SEED)
    out = {}
    cases = []
    ci = 0
    for name in ("sexp", "matern2.5"):
        for ard in (False, True):
            for d_glob in (0, 2):
                for nugget_est, scale_est, nugget in ((False, False, 1e-6), (True, True, 1e-3)):
                    n, d_loc = 37 + 3 * ci, 2
                    k, X = make_node(rng, n, d_loc, d_glob, name, ard, nugget=nugget, scale=1.3,
                                     nugget_est=nugget_est, scale_est=scale_est)
                    p = f"c{ci}_"
                    out[p + "X"] = X
                    out[p + "y"] = k.output.copy()
                    out[p + "length"] = k.length.copy()
                    out[p + "scale"] = k.scale.copy()
                    out[p + "nugget"] = k.nugget.copy()
                    out[p + "flags"] = np.array([nugget_est, scale_est, d_loc, d_glob], dtype=np.int64)
                    out[p + "name"] = np.array(name)
                    K, fod = k.k_matrix(fod_eval=True)
                    out[p + "K"], out[p + "fod"] = K, fod
                    assert np.array_equal(k.k_matrix(), K) or np.allclose(k.k_matrix(), K, rtol=1e-15, atol=0)
                    out[p + "K_plain"] = k.k_matrix()
                    out[p + "loglik"] = np.atleast_1d(k.log_likelihood_func()).astype(np.float64)
                    x0 = k.log_t()
                    f, g = k.llik(x0.copy())
                    out[p + "nllik"], out[p + "nllik_grad"] = np.atleast_1d(f), np.atleast_1d(g)
                    out[p + "scale_after"] = np.atleast_1d(k.scale).copy()
                    k.compute_stats()
                    out[p + "Rinv"], out[p + "Rinv_y"] = k.Rinv, k.Rinv_y
                    # prediction: plain GP on deterministic inputs
                    M = 23
                    xt = rng.uniform(-0.1, 1.1, size=(M, d_loc))
                    zt = rng.uniform(-0.1, 1.1, size=(M, d_glob)) if d_glob else None
                    m1, v1 = k.gp_prediction(xt, zt)
                    out[p + "xt"] = xt
                    if d_glob:
                        out[p + "zt"] = zt
                    out[p + "gp_m"], out[p + "gp_v"] = m1, v1
                    # prediction: linked GP on Gaussian inputs (one zero-variance column entry)
                    mt = rng.uniform(0, 1, size=(M, d_loc))
                    vt = rng.uniform(1e-4, 0.05, size=(M, d_loc))
                    vt[0, 0] = 0.0
                    m2, v2 = k.linkgp_prediction(mt, vt, zt)
                    out[p + "lk_m_in"], out[p + "lk_v_in"] = mt, vt
                    out[p + "lk_m"], out[p + "lk_v"] = m2, v2
                    cases.append(ci)
                    ci += 1
    out["ncases"] = np.array(ci)
    save("dense_nodes", **out)


# ---------------------------------------------------------------- Jd / Jd0 scalar known answers
def gen_jd():
    rng = np.random.default_rng(SEED + 1)
    N = 400
    X1, X2 = rng.uniform(-0.5, 1.5, N), rng.uniform(-0.5, 1.5, N)
    zm = rng.uniform(-0.2, 1.2, N)
    zv = 10 ** rng.uniform(-6, -0.5, N)
    ell = 10 ** rng.uniform(-0.7, 0.7, N)
    jd = np.array([V.Jd(X1[i], X2[i], zm[i], zv[i], ell[i]) for i in range(N)])
    jd0 = np.array([V.Jd0(X1[i], zm[i], zv[i], ell[i]) for i in range(N)])
    save("jd", X1=X1, X2=X2, zm=zm, zv=zv, ell=ell, jd=jd, jd0=jd0)


# ---------------------------------------------------------------- 4: Vecchia
def gen_vecchia():
    rng = np.random.default_rng(SEED + 2)
    out = {}
    ci = 0
    for name in ("sexp", "matern2.5"):
        for ard in (False, True):
            for d_glob in (0, 1):
                nugget_est, scale_est = bool(ci % 2), bool((ci // 2) % 2)
                n, d_loc, m = 150 + 10 * ci, 2, 8
                k, X = make_node(rng, n, d_loc, d_glob, name, ard, nugget=1e-3 if nugget_est else 1e-6,
                                 scale=0.8, nugget_est=nugget_est, scale_est=scale_est)
                k.vecch, k.m = True, m
                np.random.seed(SEED + ci)
                k.ord_nn()
                p = f"c{ci}_"
                out[p + "X"], out[p + "y"] = X, k.output.copy()
                out[p + "length"], out[p + "scale"], out[p + "nugget"] = k.length.copy(), k.scale.copy(), k.nugget.copy()
                out[p + "flags"] = np.array([nugget_est, scale_est, d_loc, d_glob, m], dtype=np.int64)
                out[p + "name"] = np.array(name)
                out[p + "ord"], out[p + "NNarray"] = k.ord.copy(), k.NNarray.copy()
                out[p + "llik"] = np.atleast_1d(k.log_likelihood_func_vecch())
                Xo = X[k.ord]
                Lm = V.L_matrix(Xo, k.NNarray, k.length, k.nugget[0], name)
                out[p + "Lmatrix"] = Lm
                z = rng.standard_normal(n)
                out[p + "z"] = z
                out[p + "draw"] = V.forward_solve_sp(Lm / np.sqrt(k.scale[0]), k.NNarray, z)
                x0 = k.log_t()
                f, g = k.llik_vecch(x0.copy())
                out[p + "nllik"], out[p + "nllik_grad"] = np.atleast_1d(f), np.atleast_1d(g)
                out[p + "scale_after"] = np.atleast_1d(k.scale).copy()
                # predictions
                M, pm = 40, 12
                k.pred_m = pm
                xt = rng.uniform(0, 1, size=(M, d_loc))
                zt = rng.uniform(0, 1, size=(M, d_glob)) if d_glob else None
                xq = xt if zt is None else np.concatenate((xt, zt), 1)
                out[p + "pred_NN"] = V.get_pred_nn(xq / k.length, X / k.length, pm)
                m1, v1 = k.gp_prediction(xt, zt)
                out[p + "xt"] = xt
                if d_glob:
                    out[p + "zt"] = zt
                out[p + "gp_m"], out[p + "gp_v"] = m1, v1
                mt = rng.uniform(0, 1, size=(M, d_loc))
                vt = rng.uniform(1e-4, 0.05, size=(M, d_loc))
                xq2 = mt if zt is None else np.concatenate((mt, zt), 1)
                out[p + "lk_NN"] = V.get_pred_nn(xq2 / k.length, X / k.length, pm)
                m2, v2 = k.linkgp_prediction(mt, vt, zt)
                out[p + "lk_m_in"], out[p + "lk_v_in"] = mt, vt
                out[p + "lk_m"], out[p + "lk_v"] = m2, v2
                ci += 1
    out["ncases"] = np.array(ci)
    # a bigger ordered-NN / kNN index fixture (bit-exact integers)
    for j, (n, d, m) in enumerate(((1000, 3, 25), (700, 10, 25), (60, 2, 25), (20, 2, 25))):
        x = rng.uniform(0, 1, size=(n, d))
        out[f"nn{j}_x"] = x
        out[f"nn{j}_m"] = np.array(m)
        out[f"nn{j}_NN"] = V.nn(x, m)
        q = rng.uniform(0, 1, size=(333, d))
        out[f"nn{j}_q"] = q
        out[f"nn{j}_pred"] = V.get_pred_nn(q, x, 50)
    save("vecchia_nodes", **out)


# ---------------------------------------------------------------- 3: ESS with injected draws
class Injector:
    """Replays pre-generated draws through the names `dgpsi.imputation` imported
    (imputation.py:1-4): `fmvn`, `fmvn_sp`, `uniform`."""

    def __init__(self, Z, U):
        self.Z, self.U = Z, U
        self.zi = self.ui = 0
        self.thetas = []
        self.block_log = []

    def fmvn(self, cov):
        z = self.Z[self.zi]
        self.zi += 1
        return np.linalg.cholesky(cov) @ z

    def fmvn_sp(self, X, NNarray, scale, length, nugget, name):
        z = self.Z[self.zi]
        self.zi += 1
        L = V.L_matrix(X, NNarray, length, nugget, name) / np.sqrt(scale)
        return V.forward_solve_sp(L, NNarray, z)

    def uniform(self, low=0.0, high=1.0):
        u = self.U[self.ui]
        self.ui += 1
        val = low + (high - low) * u
        self.thetas.append(val)
        return val


def build_ref_dgp(rng, n, d, widths, name, vecchia=False, m=8):
    X = rng.uniform(0, 1, size=(n, d))
    Y = np.stack([np.sin(4 * X.sum(1)) + X[:, 0] * X[:, -1], np.cos(3 * X[:, 0]) - X[:, -1] ** 2], 1)[:, : widths[-1]]
    layers = []
    for li, w in enumerate(widths):
        last = li == len(widths) - 1
        layers.append([kernel(length=np.array([1.0 + 0.1 * kk]), name=name, scale_est=last,
                              connect=np.arange(d) if li > 0 else None) for kk in range(w)])
    model = dgpsi.dgp(X, Y, dgpsi.combine(*layers), vecchia=vecchia, m=m)
    return X, Y, model


def snapshot(all_layer, prefix, out):
    for l, layer in enumerate(all_layer):
        for k, node in enumerate(layer):
            p = f"{prefix}L{l}K{k}_"
            out[p + "input"] = node.input.copy()
            out[p + "output"] = node.output.copy()
            out[p + "length"], out[p + "scale"], out[p + "nugget"] = node.length.copy(), node.scale.copy(), node.nugget.copy()
            if node.global_input is not None:
                out[p + "global_input"] = node.global_input.copy()
            if node.vecch:
                out[p + "ord"], out[p + "NNarray"] = node.ord.copy(), node.NNarray.copy()


def gen_ess():
    out = {}
    for ci, (name, vecch, widths) in enumerate((("sexp", False, (3, 2, 2)), ("matern2.5", False, (3, 1)),
                                                ("sexp", True, (3, 2)))):
        rng = np.random.default_rng(SEED + 10 + ci)
        np.random.seed(SEED + ci)
        dgpsi.nb_seed(SEED + ci)
        n, d = (45, 3) if not vecch else (120, 3)
        X, Y, model = build_ref_dgp(rng, n, d, widths, name, vecchia=vecch)
        p = f"c{ci}_"
        out[p + "X"], out[p + "Y"] = X, Y
        out[p + "name"] = np.array(name)
        out[p + "widths"] = np.array(widths)
        out[p + "vecch"] = np.array(vecch)
        snapshot(model.all_layer, p + "pre_", out)
        sweeps = 3
        nblk = sweeps * (len(widths) - 1)
        Z = rng.standard_normal((nblk * max(widths), n))
        U = rng.uniform(size=4000)
        inj = Injector(Z, U)
        old = IMP.fmvn, IMP.fmvn_sp, IMP.uniform
        IMP.fmvn, IMP.fmvn_sp, IMP.uniform = inj.fmvn, inj.fmvn_sp, inj.uniform
        try:
            model.imp.sample(burnin=sweeps - 1)
        finally:
            IMP.fmvn, IMP.fmvn_sp, IMP.uniform = old
        out[p + "Z"], out[p + "U"] = Z[: inj.zi], U[: inj.ui]
        out[p + "draw_values"] = np.array(inj.thetas)
        out[p + "sweeps"] = np.array(sweeps)
        snapshot(model.all_layer, p + "post_", out)
        # one M-step on the imputed state (dgp.py:1391-1398)
        for l, layer in enumerate(model.all_layer):
            for k, node in enumerate(layer):
                node.maximise()
                out[p + f"mstep_L{l}K{k}"] = np.concatenate((node.scale, node.length, node.nugget))
    out["ncases"] = np.array(3)
    save("ess_replay", **out)


# ---------------------------------------------------------------- end-to-end: step function (config 1) + 2-layer
def gen_e2e():
    out = {}
    np.random.seed(SEED)
    dgpsi.nb_seed(SEED)
    X = np.linspace(0, 1, 10)[:, None]
    Y = np.array([[-1.0] if i < 0.5 else [1.0] for i in X[:, 0]])
    layer1 = [kernel(length=np.array([1.0]), name="sexp")]
    layer2 = [kernel(length=np.array([1.0]), name="sexp")]
    layer3 = [kernel(length=np.array([1.0]), name="sexp", scale_est=True)]
    model = dgpsi.dgp(X, Y, dgpsi.combine(layer1, layer2, layer3))
    model.train(N=60, disable=True)
    final = model.estimate()
    emu = dgpsi.emulator(final, N=4)
    xt = np.linspace(0, 1, 41)[:, None]
    mu, var = emu.predict(xt)
    out["step_X"], out["step_Y"], out["step_xt"] = X, Y, xt
    out["step_mu"], out["step_var"] = mu, var
    out["step_nimp"] = np.array(len(emu.all_layer_set))
    for s, al in enumerate(emu.all_layer_set):
        snapshot(al, f"step_S{s}_", out)
    # 2-layer Matern with global connection (config-2 shape, small n) -- frozen imputations
    rng = np.random.default_rng(SEED + 5)
    n, d = 60, 3
    Xm = rng.uniform(0, 1, size=(n, d))
    Ym = (np.sin(2 * np.pi * Xm[:, 0] * Xm[:, 1]) + (Xm[:, 2] - 0.5) ** 2).reshape(-1, 1)
    l1 = [kernel(length=np.array([1.0]), name="matern2.5") for _ in range(d)]
    l2 = [kernel(length=np.array([1.0]), name="matern2.5", scale_est=True, connect=np.arange(d))]
    model2 = dgpsi.dgp(Xm, Ym, dgpsi.combine(l1, l2))
    model2.train(N=15, disable=True)
    emu2 = dgpsi.emulator(model2.estimate(), N=3)
    xt2 = rng.uniform(0, 1, size=(57, d))
    mu2, var2 = emu2.predict(xt2)
    out["mat_X"], out["mat_Y"], out["mat_xt"] = Xm, Ym, xt2
    out["mat_mu"], out["mat_var"] = mu2, var2
    out["mat_nimp"] = np.array(len(emu2.all_layer_set))
    for s, al in enumerate(emu2.all_layer_set):
        snapshot(al, f"mat_S{s}_", out)
    # linked system GP -> DGP -> GP (config-5 shape, small n)
    n = 40
    X1 = rng.uniform(0, 1, size=(n, 2))
    Y1 = (np.sin(3 * X1[:, 0]) + X1[:, 1] ** 2).reshape(-1, 1)
    g1 = dgpsi.gp(X1, Y1, kernel(length=np.array([1.0, 1.0]), name="matern2.5", scale_est=True))
    g1.train()
    X2 = rng.uniform(-0.2, 2.0, size=(n, 1))
    Y2 = np.tanh(2 * (X2 - 0.9))
    d2 = dgpsi.dgp(X2, Y2, dgpsi.combine([kernel(length=np.array([1.0]), name="matern2.5")],
                                         [kernel(length=np.array([1.0]), name="matern2.5", scale_est=True,
                                                 connect=np.arange(1))]))
    d2.train(N=12, disable=True)
    X3 = rng.uniform(-1.1, 1.1, size=(n, 1))
    Y3 = X3**2 - 0.3 * X3
    g3 = dgpsi.gp(X3, Y3, kernel(length=np.array([1.0]), name="sexp", scale_est=True))
    g3.train()
    c1 = dgpsi.container(g1.export(), np.array([0, 1]))
    c2 = dgpsi.container(d2.estimate(), np.array([0]))
    c3 = dgpsi.container(g3.export(), np.array([0]))
    system = dgpsi.lgp(dgpsi.combine([c1], [c2], [c3]), N=3)
    xt3 = rng.uniform(0, 1, size=(31, 2))
    mu3, var3 = system.predict(xt3)
    out["lgp_xt"], out["lgp_mu"], out["lgp_var"] = xt3, mu3[0], var3[0]
    out["lgp_nimp"] = np.array(len(system.all_layer_set))
    for s, one in enumerate(system.all_layer_set):
        for l, layer in enumerate(one):
            cont = layer[0]
            if cont.type == "gp":
                snapshot([[cont.structure]], f"lgp_S{s}_E{l}_", out)
            else:
                snapshot(cont.structure, f"lgp_S{s}_E{l}_", out)
    save("e2e", **out)


# ---------------------------------------------------------------- leave-one-out
def gen_loo():
    out = {}
    rng = np.random.default_rng(SEED + 9)
    np.random.seed(SEED + 9)
    dgpsi.nb_seed(SEED + 9)
    n, d = 50, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = (np.sin(4 * X[:, 0]) + X[:, 1] ** 2 + 0.05 * rng.standard_normal(n)).reshape(-1, 1)
    out["gp_X"], out["gp_Y"] = X, Y
    for tag, name in (("se", "sexp"), ("ma", "matern2.5")):
        length = np.array([0.7, 0.9])
        g = dgpsi.gp(X, Y, kernel(length=length.copy(), scale=1.3, nugget=1e-4, name=name))
        mu, s2 = g.loo()
        out[f"gp_{tag}_dense_mu"], out[f"gp_{tag}_dense_var"] = mu, s2
        gv = dgpsi.gp(X, Y, kernel(length=length.copy(), scale=1.3, nugget=1e-4, name=name), vecchia=True, m=8)
        mu, s2 = gv.loo(m=6)
        out[f"gp_{tag}_vecch_mu"], out[f"gp_{tag}_vecch_var"] = mu, s2
    # DGP emulators with frozen imputations: dense (LOO conditions on all other points) and Vecchia
    n = 30
    Xd = rng.uniform(0, 1, size=(n, d))
    Yd = (np.sin(2 * np.pi * Xd[:, 0] * Xd[:, 1]) + (Xd[:, 1] - 0.5) ** 2).reshape(-1, 1)
    out["emu_X"], out["emu_Y"] = Xd, Yd
    for tag, vec in (("dense", False), ("vecch", True)):
        l1 = [kernel(length=np.array([1.0]), name="matern2.5") for _ in range(d)]
        l2 = [kernel(length=np.array([1.0]), name="matern2.5", scale_est=True, connect=np.arange(d))]
        model = dgpsi.dgp(Xd, Yd, dgpsi.combine(l1, l2), vecchia=vec, m=6)
        model.train(N=8, disable=True)
        emu = dgpsi.emulator(model.estimate(), N=2)
        mu, s2 = emu.loo(Xd, m=5)
        out[f"emu_{tag}_mu"], out[f"emu_{tag}_var"] = mu, s2
        out[f"emu_{tag}_nimp"] = np.array(len(emu.all_layer_set))
        for s, al in enumerate(emu.all_layer_set):
            snapshot(al, f"emu_{tag}_S{s}_", out)
    save("loo", **out)


def gen_metric():
    """Sequential-design criteria of emulation.py:323-420 (ALM, MICE, VIGF) on frozen imputations."""
    out = {}
    rng = np.random.default_rng(SEED + 10)
    np.random.seed(SEED + 10)
    dgpsi.nb_seed(SEED + 10)
    n, d = 30, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = np.stack([np.sin(2 * np.pi * X[:, 0] * X[:, 1]) + (X[:, 1] - 0.5) ** 2, np.cos(3 * X[:, 0]) * X[:, 1]], 1)
    xc = rng.uniform(0, 1, size=(40, d))
    out["X"], out["Y"], out["x_cand"] = X, Y, xc
    for tag, name, depth in (("ma2", "matern2.5", 2), ("se3", "sexp", 3)):
        layers = [[kernel(length=np.array([1.0]), name=name) for _ in range(d)]]
        if depth == 3:
            layers.append([kernel(length=np.array([1.0]), name=name, connect=np.arange(d)) for _ in range(2)])
        layers.append([kernel(length=np.array([1.0]), name=name, scale_est=True, connect=np.arange(d))
                       for _ in range(2)])
        model = dgpsi.dgp(X, Y, dgpsi.combine(*layers))
        model.train(N=8, disable=True)
        emu = dgpsi.emulator(model.estimate(), N=3)
        out[f"{tag}_alm"] = emu.metric(xc, method="ALM", score_only=True)
        out[f"{tag}_mice"] = emu.metric(xc, method="MICE", nugget_s=1.0, score_only=True)
        out[f"{tag}_mice_small"] = emu.metric(xc, method="MICE", nugget_s=1e-3, score_only=True)
        out[f"{tag}_vigf"] = emu.metric(xc, method="VIGF", obj=model, score_only=True)
        idx, val = emu.metric(xc, method="VIGF", obj=model)
        out[f"{tag}_vigf_idx"], out[f"{tag}_vigf_val"] = np.asarray(idx), np.asarray(val)
        # the per-imputation moments the criteria are built from (oracle check of the host arithmetic)
        pin, s2 = emu.predict_mice(xc, False, m=50)
        out[f"{tag}_mice_input"], out[f"{tag}_mice_var"] = np.asarray(pin), np.asarray(s2)
        index = np.argmin(((xc[:, None, :] - X[None, :, :]) ** 2).sum(-1), axis=1)
        bias, s2v = emu.predict_vigf(xc, index, False, m=50)
        out[f"{tag}_vigf_bias"], out[f"{tag}_vigf_var"] = np.asarray(bias), np.asarray(s2v)
        out[f"{tag}_nimp"] = np.array(len(emu.all_layer_set))
        for s, al in enumerate(emu.all_layer_set):
            snapshot(al, f"{tag}_S{s}_", out)
    # single GP emulator: gp.metric (gp.py:271-324) and gp.update_xy (gp.py:144-181)
    Xg = rng.uniform(0, 1, size=(45, d))
    Yg = (np.sin(4 * Xg[:, 0]) + Xg[:, 1] ** 2).reshape(-1, 1)
    Xg2 = rng.uniform(0, 1, size=(52, d))
    Yg2 = (np.sin(4 * Xg2[:, 0]) + Xg2[:, 1] ** 2).reshape(-1, 1)
    out["gp_X"], out["gp_Y"], out["gp_X2"], out["gp_Y2"] = Xg, Yg, Xg2, Yg2
    for tag, name in (("se", "sexp"), ("ma", "matern2.5")):
        for vtag, vec in (("dense", False), ("vecch", True)):
            g = dgpsi.gp(Xg, Yg, kernel(length=np.array([0.6, 0.8]), scale=1.2, nugget=1e-4, name=name),
                         vecchia=vec, m=10)
            q = f"gp_{tag}_{vtag}_"
            out[q + "alm"] = g.metric(xc, method="ALM", score_only=True, m=12)
            out[q + "mice"] = g.metric(xc, method="MICE", score_only=True, m=12)
            out[q + "vigf"] = g.metric(xc, method="VIGF", score_only=True, m=12)
            g.update_xy(Xg2, Yg2)
            mu2, var2 = g.predict(xc, m=12)
            out[q + "upd_mu"], out[q + "upd_var"] = mu2, var2
    save("metric", **out)


def gen_update():
    """Warm start with a grown / shrunk design: the deterministic part of dgp.update_xy (dgp.py:824-1095), i.e.
    update_all_layer_larger (conditional means at the added rows) and update_all_layer_smaller."""
    out = {}
    rng = np.random.default_rng(SEED + 11)
    np.random.seed(SEED + 11)
    dgpsi.nb_seed(SEED + 11)
    n, n_new, d = 24, 31, 2
    Xall = rng.uniform(0, 1, size=(n_new, d))
    f = lambda X: (np.sin(2 * np.pi * X[:, 0] * X[:, 1]) + (X[:, 1] - 0.5) ** 2).reshape(-1, 1)
    perm = rng.permutation(n_new)          # the grown design holds the old rows in shuffled positions
    Xnew = Xall[perm]
    Xold = Xall[:n]
    out["X_old"], out["Y_old"], out["X_new"], out["Y_new"] = Xold, f(Xold), Xnew, f(Xnew)
    for tag, vec, name in (("dense_ma", False, "matern2.5"), ("dense_se", False, "sexp"), ("vecch_se", True, "sexp")):
        l1 = [kernel(length=np.array([1.0]), name=name) for _ in range(d)]
        l2 = [kernel(length=np.array([1.0]), name=name, connect=np.arange(d)) for _ in range(2)]
        l3 = [kernel(length=np.array([1.0]), name=name, scale_est=True, connect=np.arange(d))]
        model = dgpsi.dgp(Xold, f(Xold), dgpsi.combine(l1, l2, l3), vecchia=vec, m=6)
        model.train(N=6, disable=True)
        snapshot(model.all_layer, f"{tag}_before_", out)
        # preamble of update_xy (dgp.py:835-858) without the random burn-in that follows
        model.Y, model.X = f(Xnew), Xnew
        model.indices = None
        model.n_data = n_new
        sub_idx = np.where((model.X == Xold[:, None]).all(-1))[1]
        out[f"{tag}_sub_idx"] = sub_idx
        model.update_all_layer_larger(sub_idx)
        snapshot(model.all_layer, f"{tag}_larger_", out)
        keep = np.sort(rng.choice(n_new, 17, replace=False))
        out[f"{tag}_keep"] = keep
        model.Y, model.X = model.Y[keep], model.X[keep]
        model.n_data = len(keep)
        model.update_all_layer_smaller(keep)
        snapshot(model.all_layer, f"{tag}_smaller_", out)
    save("update", **out)


def gen_lik():
    """Likelihood layers (likelihood_class.py): ESS sweeps with injected draws for Poisson / NegBin / Hetero final
    layers (Hetero: node-wise updates with the exact conditional draw of the mean, its normals recorded), the
    log-likelihoods, and predictions on frozen imputations."""
    out = {}
    cases = (("poi", dgpsi.Poisson, 1), ("nb", dgpsi.NegBin, 2), ("het", dgpsi.Hetero, 2),
             ("catl", lambda: dgpsi.Categorical(link="logit"), 1), ("catp", lambda: dgpsi.Categorical(link="probit"), 1),
             ("cats", lambda: dgpsi.Categorical(link="softmax"), 3),
             ("catr", lambda: dgpsi.Categorical(link="robustmax"), 3),
             ("zip", dgpsi.ZIP, 2), ("zinb", dgpsi.ZINB, 3))
    import os
    only = os.environ.get("GOLDEN_LIK_ONLY")
    for ci, (tag, Lik, width) in enumerate(cases):
        if only and tag not in only.split(","):
            continue
        print("case", tag, flush=True)
        rng = np.random.default_rng(SEED + 20 + ci)
        np.random.seed(SEED + 20 + ci)
        dgpsi.nb_seed(SEED + 20 + ci)
        n, d = 28, 2
        X = rng.uniform(0, 1, size=(n, d))
        g = np.sin(3 * X[:, 0]) + X[:, 1]
        if tag == "het":
            Y = (g + np.exp(-1.5 + X[:, 0]) * rng.standard_normal(n)).reshape(-1, 1)
        elif tag == "poi":
            Y = rng.poisson(np.exp(1.0 + g)).astype(float).reshape(-1, 1)
        elif tag in ("zip", "zinb"):
            counts = rng.poisson(np.exp(1.0 + g)) if tag == "zip" else rng.negative_binomial(3.0, 3.0 / (3.0 + np.exp(1.0 + g)))
            Y = np.where(rng.uniform(size=n) < 0.3, 0, counts).astype(float).reshape(-1, 1)
        elif tag in ("catl", "catp"):
            Y = (g + 0.3 * rng.standard_normal(n) > 0.9).astype(int).reshape(-1, 1)
        elif tag in ("cats", "catr"):
            Y = np.digitize(g + 0.3 * rng.standard_normal(n), [0.6, 1.2]).reshape(-1, 1)
        else:
            Y = rng.negative_binomial(3.0, 3.0 / (3.0 + np.exp(1.0 + g))).astype(float).reshape(-1, 1)
        l1 = [kernel(length=np.array([1.0]), name="sexp") for _ in range(d)]
        l2 = [kernel(length=np.array([1.0]), name="matern2.5", scale_est=True, connect=np.arange(d))
              for _ in range(width)]
        # dgp.py:527-532 leaves the NegBin dispersion column of np.empty() unset when there are no replicates; pin
        # the uninitialised memory to zeros while the reference object is built so the run is reproducible
        real_empty = np.empty
        np.empty = lambda *a, **k: np.zeros(*a, **k)
        try:
            model = dgpsi.dgp(X, Y, dgpsi.combine(l1, l2, [Lik()]))
        finally:
            np.empty = real_empty
        model.train(N=4, disable=True)
        p = f"{tag}_"
        out[p + "X"], out[p + "Y"] = X, Y
        snapshot(model.all_layer[:-1], p + "pre_", out)
        out[p + "llik_pre"] = np.array(model.all_layer[-1][0].llik())
        out[p + "lik_input_pre"] = model.all_layer[-1][0].input.copy()
        sweeps = 3
        Z = rng.standard_normal((sweeps * 2 * 3, n))
        U = rng.uniform(size=4000)
        inj = Injector(Z, U)
        SD = []
        real_randn = np.random.randn

        def randn(*shape):
            sd = real_randn(*shape)
            SD.append(sd.copy())
            return sd

        old = IMP.fmvn, IMP.fmvn_sp, IMP.uniform
        IMP.fmvn, IMP.fmvn_sp, IMP.uniform = inj.fmvn, inj.fmvn_sp, inj.uniform
        np.random.randn = randn
        try:
            model.imp.sample(burnin=sweeps - 1)
        finally:
            IMP.fmvn, IMP.fmvn_sp, IMP.uniform = old
            np.random.randn = real_randn
        out[p + "Z"], out[p + "U"] = Z[: inj.zi], U[: inj.ui]
        out[p + "SD"] = np.asarray(SD) if SD else np.zeros((0, n, 2))
        out[p + "draw_values"] = np.array(inj.thetas)
        out[p + "sweeps"] = np.array(sweeps)
        snapshot(model.all_layer[:-1], p + "post_", out)
        out[p + "llik_post"] = np.array(model.all_layer[-1][0].llik())
        # predictions on frozen imputations
        model.train(N=4, disable=True)
        emu = dgpsi.emulator(model.estimate(), N=2)
        xt = rng.uniform(0, 1, size=(15, d))
        np.random.seed(SEED + 77)          # Categorical.prediction draws from numpy's RNG (K > 2 classes)
        mu, var = emu.predict(xt)
        out[p + "xt"], out[p + "mu"], out[p + "var"] = xt, mu, var
        np.random.seed(SEED + 77)
        mus, vars_ = emu.predict(xt, full_layer=True)
        out[p + "mu_full_last"], out[p + "var_full_last"] = mus[-1], vars_[-1]
        out[p + "mu_full_gp"], out[p + "var_full_gp"] = mus[-2], vars_[-2]
        out[p + "nimp"] = np.array(len(emu.all_layer_set))
        for s_, al in enumerate(emu.all_layer_set):
            snapshot(al[:-1], f"{p}S{s_}_", out)
        # emulation.py:872-874,909 indexes the predictions by the inverse map of np.unique even when no row repeats
        # (then the predictions are still in the caller's order): only inputs already in np.unique order are scored
        # against their own outputs, so the fixture uses such inputs
        order = np.lexsort(xt.T[::-1])
        xs = xt[order]
        assert np.array_equal(xs, np.unique(xt, axis=0))
        if tag.startswith("cat"):
            yt = (np.arange(len(xt)) % (2 if width == 1 else 3)).reshape(-1, 1)
        else:
            yt = (np.round(np.abs(mu)) if tag != "het" else mu + 0.1)[order]
        avg, per = emu.nllik(xs, yt)
        out[p + "xt_sorted"], out[p + "yt"], out[p + "nllik_avg"], out[p + "nllik"] = xs, yt, np.array(avg), per
        # design criteria of an emulator with a likelihood layer (emulation.py:344-349, 373-392, 393-413)
        out[p + "alm"] = emu.metric(xt, method="ALM", score_only=True)
        out[p + "mice"] = emu.metric(xt, method="MICE", score_only=True)
        out[p + "vigf"] = emu.metric(xt, method="VIGF", obj=model, score_only=True)
    # one GP layer under a Poisson node: the 2-layer branches (emulation.py:362-372, 402-403)
    if not only:
        rng = np.random.default_rng(SEED + 29)
        np.random.seed(SEED + 29)
        dgpsi.nb_seed(SEED + 29)
        n, d = 26, 2
        X = rng.uniform(0, 1, size=(n, d))
        Y = rng.poisson(np.exp(1.0 + np.sin(3 * X[:, 0]) + X[:, 1])).astype(float).reshape(-1, 1)
        model = dgpsi.dgp(X, Y, dgpsi.combine([kernel(length=np.array([1.0]), name="sexp", scale_est=True)],
                                              [dgpsi.Poisson()]))
        model.train(N=4, disable=True)
        emu = dgpsi.emulator(model.estimate(), N=2)
        xt = rng.uniform(0, 1, size=(15, d))
        p = "poi2_"
        out[p + "X"], out[p + "Y"], out[p + "xt"] = X, Y, xt
        out[p + "mu"], out[p + "var"] = emu.predict(xt)
        out[p + "alm"] = emu.metric(xt, method="ALM", score_only=True)
        out[p + "mice"] = emu.metric(xt, method="MICE", score_only=True)
        out[p + "vigf"] = emu.metric(xt, method="VIGF", obj=model, score_only=True)
        out[p + "nimp"] = np.array(len(emu.all_layer_set))
        for s_, al in enumerate(emu.all_layer_set):
            snapshot(al[:-1], f"{p}S{s_}_", out)
    # a linked system whose second emulator is a DGP + Poisson likelihood (linkgp.py:569-571)
    if not only:
        rng = np.random.default_rng(SEED + 31)
        np.random.seed(SEED + 31)
        dgpsi.nb_seed(SEED + 31)
        n = 30
        X1 = rng.uniform(0, 1, size=(n, 2))
        Y1 = (np.sin(3 * X1[:, 0]) + X1[:, 1]).reshape(-1, 1)
        g1 = dgpsi.gp(X1, Y1, kernel(length=np.array([1.0]), name="matern2.5", scale_est=True))
        g1.train()
        X2 = rng.uniform(-0.2, 2.0, size=(n, 1))
        Y2 = rng.poisson(np.exp(0.5 + X2)).astype(float)
        d2 = dgpsi.dgp(X2, Y2, dgpsi.combine([kernel(length=np.array([1.0]), name="sexp")],
                                             [kernel(length=np.array([1.0]), name="sexp")],
                                             [dgpsi.Poisson()]))   # fixed unit scales: an estimated scale of 1e5
        d2.train(N=6, disable=True)                                # makes the variances pure cancellation noise
        system = dgpsi.lgp(dgpsi.combine([dgpsi.container(g1.export(), np.array([0, 1]))],
                                         [dgpsi.container(d2.estimate(), np.array([0]))]), N=2)
        xt = rng.uniform(0, 1, size=(17, 2))
        mu, var = system.predict(xt)
        p = "lgp_"
        out[p + "xt"], out[p + "mu"], out[p + "var"], out[p + "Y2"] = xt, mu[0], var[0], Y2
        out[p + "nimp"] = np.array(len(system.all_layer_set))
        for s_, one in enumerate(system.all_layer_set):
            snapshot([[one[0][0].structure]], f"{p}S{s_}_E0_", out)
            snapshot(one[1][0].structure[:-1], f"{p}S{s_}_E1_", out)
    # Vecchia GP layers under a Poisson node through the public API: the Vecchia path draws every random number from
    # numpy's global generator (vecchia.py:137, imputation.py:79-119), so a reimplementation that consumes the same
    # stream reproduces the whole run -- construction, 3 SEM iterations, emulator, prediction
    if not only:
        seed = 21
        rng = np.random.default_rng(seed)
        np.random.seed(seed)
        dgpsi.nb_seed(seed)
        n, d = 300, 2
        X = rng.uniform(0, 1, size=(n, d))
        Y = rng.poisson(np.exp(1.0 + np.sin(3 * X[:, 0]) + X[:, 1])).astype(float).reshape(-1, 1)
        l1 = [kernel(length=np.array([0.5]), name="sexp") for _ in range(d)]
        l2 = [kernel(length=np.array([0.5]), name="sexp", scale_est=True, connect=np.arange(d))]
        model = dgpsi.dgp(X, Y, dgpsi.combine(l1, l2, [dgpsi.Poisson()]), vecchia=True, m=12)
        model.train(N=3, disable=True)
        emu = dgpsi.emulator(model.estimate(), N=2)
        xt = rng.uniform(0, 1, size=(40, d))
        mu, var = emu.predict(xt, m=20)
        out["vpoi_mu"], out["vpoi_var"] = mu, var
        out["vpoi_theta"] = np.concatenate([np.concatenate((k.scale, k.length, k.nugget)) for layer in model.all_layer[:-1]
                                            for k in layer])
    save("likelihood", **out)


# ---------------------------------------------------------------- noise floor of the reference against itself
def _ulp_jitter(rng, a):
    """Every entry moved to a neighbouring double (relative change <= 2.3e-16), at random up or down."""
    a = np.asarray(a, dtype=np.float64)
    return np.where(rng.uniform(size=a.shape) < 0.5, np.nextafter(a, -np.inf), np.nextafter(a, np.inf))


def _spread(base, variants):
    """Largest absolute deviation of any variant from the base run, element-wise."""
    base = {k: np.atleast_1d(np.asarray(v, dtype=np.float64)) for k, v in base.items()}
    out = {k: np.zeros_like(v) for k, v in base.items()}
    for var in variants:
        for k, v in var.items():
            out[k] = np.maximum(out[k], np.abs(np.atleast_1d(np.asarray(v, dtype=np.float64)) - base[k]))
    return out


def gen_floor():
    """SURVEY.md section 7 hard part 1(b): how much does the UNMODIFIED reference move when nothing meaningful
    changes?  Three sources, element-wise maximum of all:
      (a) the BLAS thread count (1 thread against all cores: another summation order inside LAPACK / BLAS),
      (b) every input entry moved to a neighbouring double (8 random draws),
      (c) the committed fixture against a fresh run of the same code on the same inputs (another process, possibly
          another host CPU: numba compiles for the machine it runs on).
    Written to floor.npz with the keys of the fixtures they belong to; the GPU parity tests bound every error that
    exceeds the 1e-9 relative target by a multiple of this floor.  Reads the committed fixtures (their exact inputs)
    and never rewrites them."""
    from threadpoolctl import threadpool_limits

    rng = np.random.default_rng(SEED + 77)
    out = {}

    # ---- dense nodes
    g = np.load(os.path.join(HERE, "dense_nodes.npz"))
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        nugget_est, scale_est, d_loc, d_glob = [int(v) for v in g[p + "flags"][:4]]
        name = str(g[p + "name"])

        def run(X, y, xt, zt, mt, vt):
            k = kernel(length=g[p + "length"].copy(), scale=g[p + "scale"][0], nugget=g[p + "nugget"][0], name=name,
                       nugget_est=bool(nugget_est), scale_est=bool(scale_est),
                       connect=np.arange(d_glob) if d_glob else None)
            k.input, k.input_dim = X[:, :d_loc].copy(), np.arange(d_loc)
            if d_glob:
                k.global_input = X[:, d_loc:].copy()
            k.output, k.D = y.copy(), d_loc + d_glob
            k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
            r = {"loglik": k.log_likelihood_func()}
            f, gr = k.llik(k.log_t().copy())
            r["nllik"], r["nllik_grad"] = f, gr
            k.compute_stats()
            r["Rinv_y"] = k.Rinv_y
            r["gp_m"], r["gp_v"] = k.gp_prediction(xt, zt)
            r["lk_m"], r["lk_v"] = k.linkgp_prediction(mt, vt, zt)
            return r

        X, y, xt = g[p + "X"], g[p + "y"], g[p + "xt"]
        zt = g[p + "zt"] if d_glob else None
        mt, vt = g[p + "lk_m_in"], g[p + "lk_v_in"]
        base = run(X, y, xt, zt, mt, vt)
        variants = [{key: g[p + key] for key in base if p + key in g.files}]
        with threadpool_limits(limits=1):
            variants.append(run(X, y, xt, zt, mt, vt))
        for _ in range(8):
            variants.append(run(_ulp_jitter(rng, X), _ulp_jitter(rng, y), _ulp_jitter(rng, xt),
                                None if zt is None else _ulp_jitter(rng, zt), _ulp_jitter(rng, mt), vt))
        for key, v in _spread(base, variants).items():
            out[f"dense_{p}{key}"] = v

    # ---- Vecchia nodes (orderings and neighbour sets fixed: the integers of the fixture)
    g = np.load(os.path.join(HERE, "vecchia_nodes.npz"))
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        nugget_est, scale_est, d_loc, d_glob, m = [int(v) for v in g[p + "flags"]]
        name = str(g[p + "name"])

        def runv(X, y, xt, zt, mt, vt, z):
            k = kernel(length=g[p + "length"].copy(), scale=g[p + "scale"][0], nugget=g[p + "nugget"][0], name=name,
                       nugget_est=bool(nugget_est), scale_est=bool(scale_est),
                       connect=np.arange(d_glob) if d_glob else None)
            k.input, k.input_dim = X[:, :d_loc].copy(), np.arange(d_loc)
            if d_glob:
                k.global_input = X[:, d_loc:].copy()
            k.output, k.D = y.copy(), d_loc + d_glob
            k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
            k.vecch, k.m = True, m
            k.ord, k.NNarray = g[p + "ord"], g[p + "NNarray"]
            k.rev_ord = np.argsort(k.ord)
            r = {"llik": k.log_likelihood_func_vecch()}
            Lm = V.L_matrix(X[k.ord], k.NNarray, k.length, k.nugget[0], name)
            r["Lmatrix"] = Lm
            r["draw"] = V.forward_solve_sp(Lm / np.sqrt(k.scale[0]), k.NNarray, z)
            f, gr = k.llik_vecch(k.log_t().copy())
            r["nllik"], r["nllik_grad"] = f, gr
            k.pred_m = g[p + "pred_NN"].shape[1]
            r["gp_m"], r["gp_v"] = k.gp_prediction(xt, zt)
            r["lk_m"], r["lk_v"] = k.linkgp_prediction(mt, vt, zt)
            return r

        X, y, xt, z = g[p + "X"], g[p + "y"], g[p + "xt"], g[p + "z"]
        zt = g[p + "zt"] if d_glob else None
        mt, vt = g[p + "lk_m_in"], g[p + "lk_v_in"]
        base = runv(X, y, xt, zt, mt, vt, z)
        # (c) the committed fixture itself: the same reference code run in another process / on another host CPU
        # (numba compiles for the host it runs on), which moves the ill-conditioned Matern variances by up to 2e-5
        variants = [{key: g[p + key] for key in base if p + key in g.files}]
        with threadpool_limits(limits=1):
            variants.append(runv(X, y, xt, zt, mt, vt, z))
        for _ in range(8):
            variants.append(runv(_ulp_jitter(rng, X), _ulp_jitter(rng, y), _ulp_jitter(rng, xt),
                                 None if zt is None else _ulp_jitter(rng, zt), _ulp_jitter(rng, mt), vt, z))
        for key, v in _spread(base, variants).items():
            out[f"vecchia_{p}{key}"] = v

    # ---- end to end: emulator.predict on the frozen imputations of e2e.npz
    g = np.load(os.path.join(HERE, "e2e.npz"))

    def ref_layers(prefix, name, jitter):
        layers, l = [], 0
        while f"{prefix}L{l}K0_input" in g.files:
            layer, kk = [], 0
            while f"{prefix}L{l}K{kk}_input" in g.files:
                q = f"{prefix}L{l}K{kk}_"
                node = kernel(length=g[q + "length"].copy(), scale=g[q + "scale"][0], nugget=g[q + "nugget"][0], name=name)
                node.input, node.output = jitter(g[q + "input"]), jitter(g[q + "output"])
                node.input_dim = np.arange(node.input.shape[1])
                if q + "global_input" in g.files:
                    node.global_input = jitter(g[q + "global_input"])
                    node.connect = np.arange(node.global_input.shape[1])
                node.vecch = False
                node.compute_stats()
                layer.append(node)
                kk += 1
            layers.append(layer)
            l += 1
        return layers

    def run_emu(tag, name, jitter):
        emu = dgpsi.emulator.__new__(dgpsi.emulator)
        emu.all_layer_set = [ref_layers(f"{tag}_S{s}_", name, jitter) for s in range(int(g[f"{tag}_nimp"]))]
        emu.all_layer = emu.all_layer_set[0]
        emu.n_layer = len(emu.all_layer)
        emu.vecch = False
        mu, var = emu.predict(jitter(g[f"{tag}_xt"]))
        return {"mu": mu, "var": var}

    for tag, name in (("step", "sexp"), ("mat", "matern2.5")):
        base = run_emu(tag, name, lambda a: np.array(a, copy=True))
        assert np.allclose(base["mu"], g[f"{tag}_mu"], rtol=0, atol=1e-5), tag
        variants = [{"mu": g[f"{tag}_mu"], "var": g[f"{tag}_var"]}]
        with threadpool_limits(limits=1):
            variants.append(run_emu(tag, name, lambda a: np.array(a, copy=True)))
        for _ in range(8):
            variants.append(run_emu(tag, name, lambda a: _ulp_jitter(rng, a)))
        for key, v in _spread(base, variants).items():
            out[f"e2e_{tag}_{key}"] = v

    # ---- ESS sweeps with injected draws: latent columns after the sweeps (decisions must not change)
    g = np.load(os.path.join(HERE, "ess_replay.npz"))
    for ci in range(int(g["ncases"])):
        p = f"c{ci}_"
        widths, name, vecch = [int(w) for w in g[p + "widths"]], str(g[p + "name"]), bool(g[p + "vecch"])
        sweeps = int(g[p + "sweeps"])

        def run_ess(jitter):
            layers = []
            for l, w in enumerate(widths):
                layer = []
                for kk in range(w):
                    q = f"{p}pre_L{l}K{kk}_"
                    node = kernel(length=g[q + "length"].copy(), scale=g[q + "scale"][0], nugget=g[q + "nugget"][0],
                                  name=name, scale_est=(l == len(widths) - 1))
                    node.input, node.output = jitter(g[q + "input"]), jitter(g[q + "output"])
                    node.input_dim = np.arange(node.input.shape[1])
                    if q + "global_input" in g.files:
                        node.global_input = jitter(g[q + "global_input"])
                        node.connect = np.arange(node.global_input.shape[1])
                    node.D = node.input.shape[1] + (0 if node.global_input is None else node.global_input.shape[1])
                    node.vecch = vecch
                    if vecch:
                        node.ord, node.NNarray = g[q + "ord"], g[q + "NNarray"]
                        node.rev_ord = np.argsort(node.ord)
                        node.m = node.NNarray.shape[1] - 1
                    layer.append(node)
                layers.append(layer)
            # the layers feed each other: inputs of layer l+1 ARE the outputs of layer l
            for l in range(1, len(widths)):
                col = np.concatenate([nd.output for nd in layers[l - 1]], 1)
                for nd in layers[l]:
                    nd.input = col[:, nd.input_dim].copy()
            inj = Injector(g[p + "Z"], np.concatenate((g[p + "U"], np.full(64, 0.5))))
            old = IMP.fmvn, IMP.fmvn_sp, IMP.uniform
            IMP.fmvn, IMP.fmvn_sp, IMP.uniform = inj.fmvn, inj.fmvn_sp, inj.uniform
            try:
                IMP.imputer(layers, True).sample(burnin=sweeps - 1)
            finally:
                IMP.fmvn, IMP.fmvn_sp, IMP.uniform = old
            if inj.ui != len(g[p + "U"]):
                return None   # a decision flipped under the perturbation: not a rounding-level comparison
            return {f"post_L{l}K{kk}": layers[l][kk].output for l in range(len(widths)) for kk in range(widths[l])}

        base = run_ess(lambda a: np.array(a, copy=True))
        assert base is not None, ci
        variants = [{key: g[f"{p}{key}_output"] for key in base}]
        with threadpool_limits(limits=1):
            variants.append(run_ess(lambda a: np.array(a, copy=True)))
        for _ in range(8):
            variants.append(run_ess(lambda a: _ulp_jitter(rng, a)))
        variants = [v for v in variants if v is not None]
        for key, v in _spread(base, variants).items():
            out[f"ess_{p}{key}"] = v
        out[f"ess_{p}variants"] = np.array(len(variants))
    save("floor", **out)


# ---------------------------------------------------------------- method='sampling' with the reference's own draws
def gen_sampling():
    """`predict(method='sampling')` of a DGP emulator (two outputs; final layer only and every
    layer), of a linked system and of a single GP, with numpy's global generator seeded right before each call --
    the draws ARE the injected normals: the GPU path must return the same samples from the same seed."""
    out = {}
    rng = np.random.default_rng(SEED + 31)
    np.random.seed(SEED + 31)
    dgpsi.nb_seed(SEED + 31)
    n, d = 30, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = np.stack([np.sin(2 * np.pi * X[:, 0] * X[:, 1]) + (X[:, 1] - 0.5) ** 2, np.cos(3 * X[:, 0]) * X[:, 1]], 1)
    layers = [[kernel(length=np.array([1.0]), name="sexp") for _ in range(d)],
              [kernel(length=np.array([1.0]), name="sexp", connect=np.arange(d)) for _ in range(2)],
              [kernel(length=np.array([1.0]), name="sexp", scale_est=True, connect=np.arange(d)) for _ in range(2)]]
    model = dgpsi.dgp(X, Y, dgpsi.combine(*layers))
    model.train(N=6, disable=True)
    emu = dgpsi.emulator(model.estimate(), N=3)
    xt = rng.uniform(0, 1, size=(17, d))
    out["emu_xt"] = xt
    out["emu_nimp"] = np.array(len(emu.all_layer_set))
    for s, al in enumerate(emu.all_layer_set):
        snapshot(al, f"emu_S{s}_", out)
    np.random.seed(4242)
    out["emu_last"] = np.asarray(emu.predict(xt, method="sampling", sample_size=4))
    np.random.seed(4243)
    full = emu.predict(xt, method="sampling", sample_size=2, full_layer=True)
    for l, per in enumerate(full):
        out[f"emu_full_L{l}"] = np.asarray(per)
    # single GP
    g = dgpsi.gp(X, Y[:, [0]], kernel(length=np.array([0.7, 0.9]), scale=1.3, nugget=1e-4, name="matern2.5"))
    out["gp_X"], out["gp_Y"] = X, Y[:, [0]]
    np.random.seed(4244)
    out["gp_samples"] = g.predict(xt, method="sampling", sample_size=5)
    # linked system GP -> DGP -> GP on the frozen imputations of e2e.npz (config-5 shape)
    e = np.load(os.path.join(HERE, "e2e.npz"))
    kinds = {0: "matern2.5", 1: "matern2.5", 2: "sexp"}
    sets = []
    for s in range(int(e["lgp_nimp"])):
        one = []
        for em in range(3):
            lay, l = [], 0
            while f"lgp_S{s}_E{em}_L{l}K0_input" in e.files:
                q = f"lgp_S{s}_E{em}_L{l}K0_"
                node = kernel(length=e[q + "length"].copy(), scale=e[q + "scale"][0], nugget=e[q + "nugget"][0],
                              name=kinds[em])
                node.input, node.output = e[q + "input"].copy(), e[q + "output"].copy()
                node.input_dim = np.arange(node.input.shape[1])
                if q + "global_input" in e.files:
                    node.global_input = e[q + "global_input"].copy()
                    node.connect = np.arange(node.global_input.shape[1])
                node.vecch = False
                node.compute_stats()
                lay.append([node])
                l += 1
            cont = dgpsi.container.__new__(dgpsi.container)
            if len(lay) == 1:
                cont.type, cont.structure = "gp", lay[0][0]
            else:
                cont.type, cont.structure = "dgp", lay
            cont.vecch = False
            cont.local_input_idx = np.array([0, 1]) if em == 0 else np.array([0])
            one.append([cont])
        sets.append(one)
    system = dgpsi.lgp.__new__(dgpsi.lgp)
    system.L, system.all_layer, system.all_layer_set, system.num_model = 3, sets[0], sets, [1, 1]
    np.random.seed(4245)
    out["lgp_last"] = np.asarray(system.predict(e["lgp_xt"], method="sampling", sample_size=3)[0])
    np.random.seed(4246)
    fl = system.predict(e["lgp_xt"], method="sampling", sample_size=2, full_layer=True)
    for l, per in enumerate(fl):
        out[f"lgp_full_L{l}"] = np.asarray(per[0])
    save("sampling", **out)


# ---------------------------------------------------------------- Hetero exact draw under Vecchia
def gen_hetvecch():
    """Latent-Vecchia draw of the mean process under a heteroskedastic Gaussian likelihood: the conditioning sets
    (kernel.ord_nn(pointer=True), kernel_class.py:268-275), the sparse U (U_matrix, vecchia.py:426-445) and the draw
    (U_matrix_sp + Hetero.post_het_vecch, likelihood_class.py:165-183) with its normals recorded; then a whole public
    API run of a Vecchia DGP with a Hetero layer from one seed (numpy's generator drives everything under Vecchia)."""
    from dgpsi.likelihood_class import Hetero
    out = {}
    rng = np.random.default_rng(SEED + 41)
    ci = 0
    for name in ("sexp", "matern2.5"):
        for ard in (False, True):
            n, d, m = 90 + 20 * ci, 2, 8
            k, X = make_node(rng, n, d, 0, name, ard, nugget=1e-6, scale=0.7 + 0.4 * ci)
            k.vecch, k.m = True, m
            np.random.seed(SEED + 100 + ci)
            k.ord_nn(pointer=True)
            p = f"c{ci}_"
            gamma = np.exp(rng.uniform(-3, 0.5, size=n))          # data order
            y = k.output[:, 0] + np.sqrt(gamma) * rng.standard_normal(n)
            Xo = X[k.ord]
            out[p + "X"], out[p + "ord"], out[p + "imp_NN"] = X, k.ord.copy(), k.imp_NNarray.copy()
            out[p + "length"], out[p + "scale"], out[p + "name"] = k.length.copy(), k.scale.copy(), np.array(name)
            out[p + "m"] = np.array(m)
            out[p + "gamma"], out[p + "y"] = gamma, y
            g2 = np.concatenate((gamma[k.ord], gamma[k.ord]))
            NNarray = k.imp_NNarray
            Cond = NNarray > n - 1
            U = V.U_matrix(np.vstack((Xo, Xo)), NNarray[:, ::-1], Cond[:, ::-1], k.length, 0.0, k.scale[0], g2, name)
            out[p + "U_rev"] = U                                  # rows in revNNarray order
            U_l, U_ol = V.U_matrix_sp(Xo, NNarray, k.scale[0], k.length, 0.0, name, g2, k.imp_pointer_row,
                                      k.imp_pointer_col)
            np.random.seed(4300 + ci)
            sd = np.random.randn(n)
            np.random.seed(4300 + ci)
            f_ord = Hetero.post_het_vecch(U_l, U_ol, y[k.ord])
            out[p + "sd"], out[p + "f"] = sd, f_ord[k.rev_ord]
            ci += 1
    out["ncases"] = np.array(ci)
    # public API run: Vecchia DGP + Hetero
    seed = 33
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    dgpsi.nb_seed(seed)
    n, d = 200, 2
    X = rng.uniform(0, 1, size=(n, d))
    Y = (np.sin(4 * X[:, 0]) + X[:, 1] + np.exp(-1.5 + X[:, 0]) * rng.standard_normal(n)).reshape(-1, 1)
    l1 = [kernel(length=np.array([0.5]), name="sexp") for _ in range(d)]
    l2 = [kernel(length=np.array([0.5]), name="sexp", scale_est=True, connect=np.arange(d)) for _ in range(2)]
    model = dgpsi.dgp(X, Y, dgpsi.combine(l1, l2, [dgpsi.Hetero()]), vecchia=True, m=10)
    model.train(N=3, disable=True)
    out["api_theta"] = np.concatenate([np.concatenate((k.scale, k.length, k.nugget)) for layer in model.all_layer[:-1]
                                       for k in layer])
    out["api_latent"] = np.concatenate([k.output for k in model.all_layer[1]], 1)
    emu = dgpsi.emulator(model.estimate(), N=2)
    xt = rng.uniform(0, 1, size=(30, d))
    mu, var = emu.predict(xt, m=15)
    out["api_mu"], out["api_var"] = mu, var
    save("hetvecch", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["dense", "jd", "vecchia", "ess", "e2e", "loo", "metric", "update", "lik", "floor", "sampling", "hetvecch"]
    for w in which:
        {"dense": gen_dense, "jd": gen_jd, "vecchia": gen_vecchia, "ess": gen_ess, "e2e": gen_e2e, "loo": gen_loo,
         "metric": gen_metric, "update": gen_update, "lik": gen_lik, "floor": gen_floor, "sampling": gen_sampling, "hetvecch": gen_hetvecch}[w]()
