"""CPU oracle for the dgpsi stochastic-imputation (SI) hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain numpy/scipy restatement of the reference's
algorithm (mingdeyu/DGP, `dgpsi` 2.6.0).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it; the product package
`dgp_b200` never does (it fails loudly when the CUDA library is missing).

Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the *unmodified reference itself*, generated in the build
container by `tests/golden/make_golden.py` and committed as `tests/golden/*.npz`
(`tests/test_oracle_golden.py` checks every function below against them).

Every function cites the reference lines it restates (paths relative to the reference root).
All arithmetic is IEEE FP64, arrays are C-ordered, index arrays are int64.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.special import erf

SQRT5 = math.sqrt(5.0)

# --------------------------------------------------------------------------------------------
# 1. kernel matrices                                           dgpsi/kernel_class.py:304-359
# --------------------------------------------------------------------------------------------


def _scaled(X, length):
    """X / length with length broadcast (len 1 or D) -- kernel_class.py:324."""
    return np.asarray(X, dtype=np.float64) / np.asarray(length, dtype=np.float64)


def k_matrix(X, length, nugget, name, fod_eval=False, nugget_est=False, wdiag=None):
    """Correlation matrix K (n x n) and optionally dK/dlog(theta) (P x n x n).

    kernel_class.py:304-359 with functions.py:16-93 (Matern coefficient loops) and
    functions.py:36-45 (`fod_exp`).  `X` is [input | global_input] already concatenated.
    sexp:      K_ij = exp(-sum_d ((x_id-x_jd)/l_d)^2)
    matern2.5: K_ij = prod_d(1+sqrt5 r_d+5/3 r_d^2) * exp(-sqrt5 sum_d r_d),  r_d=|x_id-x_jd|/l_d
    diagonal = 1 + nugget (* wdiag_i when replicates are pooled).
    """
    length = np.atleast_1d(np.asarray(length, dtype=np.float64))
    Xl = _scaled(X, length)
    n, D = Xl.shape
    ard = len(length) != 1
    nfod = (D if ard else 1) if fod_eval else 0
    fod = np.zeros((nfod, n, n)) if fod_eval else None
    if name == "sexp":
        dist = np.zeros((n, n))
        for d in range(D):
            diff = Xl[:, d][:, None] - Xl[:, d][None, :]
            dist += diff * diff
        K = np.exp(-dist)
        if fod_eval:
            if ard:
                for d in range(D):
                    diff = Xl[:, d][:, None] - Xl[:, d][None, :]
                    fod[d] = 2.0 * diff**2 * K
            else:
                fod[0] = 2.0 * dist * K
    elif name == "matern2.5":
        coef = np.ones((n, n))
        s = np.zeros((n, n))
        cd = []
        for d in range(D):
            r = np.abs(Xl[:, d][:, None] - Xl[:, d][None, :])
            poly = 1.0 + SQRT5 * r + (5.0 / 3.0) * r**2
            coef *= poly
            s += r
            if fod_eval:
                cd.append((5.0 / 3.0) * (r**2) * (1.0 + SQRT5 * r) / poly)
        K = coef * np.exp(-SQRT5 * s)
        if fod_eval:
            if ard:
                for d in range(D):
                    fod[d] = cd[d] * K
            else:
                fod[0] = sum(cd) * K
    else:
        raise ValueError(name)
    if fod_eval:
        for p in range(nfod):
            np.fill_diagonal(fod[p], 0.0)
    nug = float(np.atleast_1d(nugget)[0])
    diag_add = nug * (np.ones(n) if wdiag is None else np.asarray(wdiag, dtype=np.float64))
    if fod_eval and nugget_est:
        fod = np.concatenate((fod, np.diag(diag_add)[None, :, :]), axis=0)
    np.fill_diagonal(K, 1.0 + diag_add)
    return (K, fod) if fod_eval else K


def k_vec(W, x, length, name):
    """Cross-correlation vector k(W, x) for one test point -- vecchia.py:244-265 (`K_vec_nb`)."""
    Wl, xl = _scaled(W, length), _scaled(x, length)
    if name == "sexp":
        return np.exp(-np.sum((Wl - xl) ** 2, axis=1))
    r = np.abs(Wl - xl)
    return np.prod(1.0 + SQRT5 * r + (5.0 / 3.0) * r**2, axis=1) * np.exp(-SQRT5 * np.sum(r, axis=1))


def k_cross(W, x, length, name):
    """k(W, x_t) for all test points at once: (M x n).  Same formula as `k_vec`."""
    Wl, xl = _scaled(W, length), _scaled(x, length)
    M, n = xl.shape[0], Wl.shape[0]
    if name == "sexp":
        dist = np.zeros((M, n))
        for d in range(Wl.shape[1]):
            diff = xl[:, d][:, None] - Wl[:, d][None, :]
            dist += diff * diff
        return np.exp(-dist)
    coef, s = np.ones((M, n)), np.zeros((M, n))
    for d in range(Wl.shape[1]):
        r = np.abs(xl[:, d][:, None] - Wl[:, d][None, :])
        coef *= 1.0 + SQRT5 * r + (5.0 / 3.0) * r**2
        s += r
    return coef * np.exp(-SQRT5 * s)


# --------------------------------------------------------------------------------------------
# 2. dense likelihood, gradient, statistics                    dgpsi/kernel_class.py:361-509,735-764
# --------------------------------------------------------------------------------------------


def log_prior(length, nugget, prior_name, prior_coef, nugget_est, cl=None):
    """kernel_class.py:367-381 with functions.py:95-100 (`g`).  `prior_coef` as STORED by the
    reference constructor (shape -1 for 'ga', +1 for 'inv_ga'; kernel_class.py:93-104)."""
    if prior_name is None:
        return 0.0
    length, nugget = np.atleast_1d(length), np.atleast_1d(nugget)
    a, b = prior_coef[0], prior_coef[1]
    if prior_name == "ref":
        t = np.sum(cl / length) + nugget
        return (a * np.log(t) - b * t)[0]

    def g(x):
        if prior_name == "ga":
            return np.sum(a * np.log(x) - b * x)
        return np.sum(-a * np.log(x) - b / x)

    lp = g(length)
    if nugget_est:
        lp += g(nugget)
    return lp


def log_prior_fod(length, nugget, prior_name, prior_coef, nugget_est, cl=None):
    """kernel_class.py:361-365,383-401: derivative of the log prior wrt log-parameters."""
    length, nugget = np.atleast_1d(length), np.atleast_1d(nugget)
    if prior_name is None:
        return np.zeros(len(length) + (1 if nugget_est else 0))
    a, b = prior_coef[0], prior_coef[1]
    if prior_name == "ref":
        t = np.sum(cl / length) + nugget
        fod = (b - a / t) * cl / length
        if nugget_est:
            fod = np.concatenate((fod, (a / t - b) * nugget))
        return fod

    def gf(x):
        return a - b * x if prior_name == "ga" else -a + b / x

    fod = gf(length)
    if nugget_est:
        fod = np.concatenate((fod, gf(nugget)))
    return fod


def loglik_dense(X, y, length, scale, nugget, name):
    """ESS log-likelihood of one dense GP node -- kernel_class.py:481-488.
    cov = scale*K; L = chol(cov); -0.5*(2 sum log|L_ii| + y' cov^-1 y)."""
    cov = float(np.atleast_1d(scale)[0]) * k_matrix(X, length, nugget, name)
    L = cholesky(cov, lower=True, check_finite=False)
    logdet = 2.0 * np.sum(np.log(np.abs(np.diag(L))))
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    quad = (y.T @ cho_solve((L, True), y, check_finite=False))[0, 0]
    return -0.5 * (logdet + quad)


def nllik_grad_dense(X, y, length, scale, nugget, name, scale_est, nugget_est):
    """M-step objective and gradient WITHOUT the prior terms -- kernel_class.py:403-445
    (no-replicate branch).  Returns (nllik, grad (P,), scale) where scale is the profiled
    sigma^2 = y'K^-1y/n when `scale_est` (kernel_class.py:430) and the input value otherwise."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    n = len(y)
    K, Kt = k_matrix(X, length, nugget, name, fod_eval=True, nugget_est=nugget_est)
    L = cholesky(K, lower=True, check_finite=False)
    KinvKt = np.array([cho_solve((L, True), Kt_i, check_finite=False) for Kt_i in Kt])
    tr = np.array([np.trace(M) for M in KinvKt])
    logdet = 2.0 * np.sum(np.log(np.abs(np.diag(L))))
    KinvY = cho_solve((L, True), y, check_finite=False)
    YKinvKtKinvY = (y.T @ KinvKt @ KinvY).flatten()
    YKinvY = (y.T @ KinvY)[0, 0]
    P1, P2 = -0.5 * tr, 0.5 * YKinvKtKinvY
    if scale_est:
        scale = YKinvY / n
        nllik = 0.5 * (logdet + n * np.log(scale))
    else:
        scale = float(np.atleast_1d(scale)[0])
        nllik = 0.5 * (logdet + YKinvY / scale)
    grad = -P1 - P2 / scale
    return float(nllik), grad, float(scale)


def compute_stats(X, y, length, nugget, name):
    """R^-1 and R^-1 y used by the predictors -- kernel_class.py:735-748."""
    R = k_matrix(X, length, nugget, name)
    L = np.linalg.cholesky(R)
    Rinv = cho_solve((L, True), np.eye(len(R)), check_finite=False)
    Rinv_y = cho_solve((L, True), np.asarray(y, dtype=np.float64).reshape(-1, 1), check_finite=False).flatten()
    return Rinv, Rinv_y


def sexp_stats(X_local, length_local):
    """R2sexp (n x n) and Psexp (d x n x n) on the local dims -- kernel_class.py:752-764,
    functions.py:259-272 (`Pmatrix`)."""
    Xl = _scaled(X_local, length_local)
    n, d = Xl.shape
    dist = np.zeros((n, n))
    P = np.empty((d, n, n))
    for k in range(d):
        diff = Xl[:, k][:, None] - Xl[:, k][None, :]
        dist += diff * diff
        P[k] = Xl[:, k][:, None] + Xl[:, k][None, :]
    R2 = np.exp(-dist / 2.0)
    np.fill_diagonal(R2, 1.0)
    return R2, P


# --------------------------------------------------------------------------------------------
# 5. closed-form prediction                                    dgpsi/functions.py:379-506
# --------------------------------------------------------------------------------------------


def gp_predict(x, W, Rinv, Rinv_y, scale, length, nugget, name):
    """functions.py:379-394 (`gp`): m = r'R^-1y, v = |scale (1 + nugget - r'R^-1 r)|.
    `x` (M x D) and `W` (n x D) already include the global columns."""
    r = k_cross(W, x, length, name)  # M x n
    m = r @ Rinv_y
    rRr = np.einsum("ti,ti->t", r, r @ Rinv.T)
    v = np.abs(float(np.atleast_1d(scale)[0]) * (1.0 + float(np.atleast_1d(nugget)[0]) - rRr))
    return m, v


def trace_sum(A, B):
    """functions.py:496-506: sum_kl A_kl B_kl over the lower triangle with off-diagonals doubled."""
    il = np.tril_indices(len(A), -1)
    return np.sum(np.diag(A) * np.diag(B)) + 2.0 * np.sum(A[il] * B[il])


def quad(A, b):
    """vecchia.py:990-1000: b'Ab via the lower triangle."""
    il = np.tril_indices(len(A), -1)
    return np.sum(np.diag(A) * b**2) + 2.0 * np.sum(A[il] * b[il[0]] * b[il[1]])


def IJ_sexp(X, z_m, z_v, length, R2sexp, Psexp):
    """functions.py:432-451."""
    n, d = X.shape
    Xz = X - z_m
    Ic, Jc = 1.0, 1.0
    Jexp = np.zeros((n, n))
    for k in range(d):
        div = 2.0 * z_v[k] / length[k] ** 2
        Ic *= 1.0 + div
        Jc *= 1.0 + 2.0 * div
        Jexp += (Psexp[k] - 2.0 * z_m[k] / length[k]) ** 2 / (2.0 + 4.0 * div)
    Ic, Jc = 1.0 / math.sqrt(Ic), 1.0 / math.sqrt(Jc)
    J = Jc * np.exp(-Jexp) * R2sexp
    I = Ic * np.exp(-np.sum(Xz**2 / (2.0 * z_v + length**2), axis=1))
    return I, J


def IJ_sexp_direct(X, z_m, z_v, length):
    """Squared-exponential branch of `IJ_nb` -- vecchia.py:845-869 (no R2sexp/Psexp tables)."""
    Xz = X - z_m
    Ic = 1.0 / math.sqrt(np.prod(1.0 + 2.0 * z_v / length**2))
    Jc = 1.0 / math.sqrt(np.prod(1.0 + 4.0 * z_v / length**2))
    I = Ic * np.exp(-np.sum(Xz**2 / (2.0 * z_v + length**2), axis=1))
    s = Xz[:, None, :] + Xz[None, :, :]
    dd = Xz[:, None, :] - Xz[None, :, :]
    e = np.sum(s**2 / (8.0 * z_v + 2.0 * length**2) + dd**2 / (2.0 * length**2), axis=2)
    ediag = np.sum(2.0 * Xz**2 / (4.0 * z_v + length**2), axis=1)
    e[np.diag_indices(len(X))] = ediag
    return I, Jc * np.exp(-e)


def _matern_plain(zX, ell):
    """(1+sqrt5|a|/l+5a^2/(3l^2)) exp(-sqrt5|a|/l) -- functions.py:471."""
    a = np.abs(zX)
    return (1.0 + SQRT5 * a / ell + 5.0 * zX**2 / (3.0 * ell**2)) * np.exp(-SQRT5 * a / ell)


def I_matern_dim(x, z_m, z_v, ell):
    """One-dimensional factor of the Matern-2.5 I integral for all training points `x` (n,) --
    functions.py:463-471."""
    zX = z_m - x
    if z_v == 0:
        return _matern_plain(zX, ell)
    muA, muB = zX - SQRT5 * z_v / ell, zX + SQRT5 * z_v / ell
    tA = np.exp((5.0 * z_v - 2.0 * SQRT5 * ell * zX) / (2.0 * ell**2)) * (
        (1.0 + SQRT5 * muA / ell + 5.0 * (muA**2 + z_v) / (3.0 * ell**2)) * 0.5 * (1.0 + erf(muA / math.sqrt(2.0 * z_v)))
        + (SQRT5 + (5.0 * muA) / (3.0 * ell)) * math.sqrt(0.5 * z_v / math.pi) / ell * np.exp(-0.5 * muA**2 / z_v)
    )
    tB = np.exp((5.0 * z_v + 2.0 * SQRT5 * ell * zX) / (2.0 * ell**2)) * (
        (1.0 - SQRT5 * muB / ell + 5.0 * (muB**2 + z_v) / (3.0 * ell**2)) * 0.5 * (1.0 + erf(-muB / math.sqrt(2.0 * z_v)))
        + (SQRT5 - (5.0 * muB) / (3.0 * ell)) * math.sqrt(0.5 * z_v / math.pi) / ell * np.exp(-0.5 * muB**2 / z_v)
    )
    return tA + tB


def Jd(X1, X2, z_m, z_v, ell):
    """One-dimensional Matern-2.5 J integral  int k(X1,z) k(X2,z) N(z; z_m, z_v) dz  for
    X1 != X2 (arrays broadcast) -- vecchia.py:915-959.  P1/P2/P3 are the three integration
    regions (z above both points, between them, below both)."""
    x1, x2 = np.minimum(X1, X2), np.maximum(X1, X2)
    l2, l3, l4 = ell**2, ell**3, ell**4
    sv = math.sqrt(2.0 * z_v)
    g = math.sqrt(0.5 * z_v / math.pi)

    def mom(mu):  # raw Gaussian moments used by every E*A* combination
        return mu, mu**2 + z_v, mu**3 + 3.0 * z_v * mu, mu**4 + 6.0 * z_v * mu**2 + 3.0 * z_v**2

    def tail(mu, x):  # coefficient polynomials multiplying E*2, E*3, E*4 in the boundary terms
        return (mu + x, mu**2 + 2.0 * z_v + x**2 + mu * x,
                mu**3 + x**3 + x * mu**2 + mu * x**2 + 3.0 * z_v * x + 5.0 * z_v * mu)

    E30 = 1.0 + (25.0 * x1**2 * x2**2 - 3.0 * SQRT5 * (3.0 * l3 + 5.0 * ell * x1 * x2) * (x1 + x2)
                 + 15.0 * l2 * (x1**2 + x2**2 + 3.0 * x1 * x2)) / (9.0 * l4)
    E31 = (18.0 * SQRT5 * l3 + 15.0 * SQRT5 * ell * (x1**2 + x2**2) - (75.0 * l2 + 50.0 * x1 * x2) * (x1 + x2)
           + 60.0 * SQRT5 * ell * x1 * x2) / (9.0 * l4)
    E32 = 5.0 * (5.0 * x1**2 + 5.0 * x2**2 + 15.0 * l2 - 9.0 * SQRT5 * ell * (x1 + x2) + 20.0 * x1 * x2) / (9.0 * l4)
    E33 = 10.0 * (3.0 * SQRT5 * ell - 5.0 * x1 - 5.0 * x2) / (9.0 * l4)
    E34 = 25.0 / (9.0 * l4)
    muC = z_m - 2.0 * SQRT5 * z_v / ell
    c1, c2, c3, c4 = mom(muC)
    E3A31 = E30 + c1 * E31 + c2 * E32 + c3 * E33 + c4 * E34
    t1, t2, t3 = tail(muC, x2)
    E3A32 = E31 + t1 * E32 + t2 * E33 + t3 * E34
    P1 = np.exp((10.0 * z_v + SQRT5 * ell * (x1 + x2 - 2.0 * z_m)) / l2) * (
        0.5 * E3A31 * (1.0 + erf((muC - x2) / sv)) + E3A32 * g * np.exp(-0.5 * (x2 - muC) ** 2 / z_v))

    E40 = 1.0 + (25.0 * x1**2 * x2**2 + 3.0 * SQRT5 * (3.0 * l3 - 5.0 * ell * x1 * x2) * (x2 - x1)
                 + 15.0 * l2 * (x1**2 + x2**2 - 3.0 * x1 * x2)) / (9.0 * l4)
    E41 = 5.0 * (3.0 * SQRT5 * ell * (x2**2 - x1**2) + 3.0 * l2 * (x1 + x2) - 10.0 * x1 * x2 * (x1 + x2)) / (9.0 * l4)
    E42 = 5.0 * (5.0 * x1**2 + 5.0 * x2**2 - 3.0 * l2 - 3.0 * SQRT5 * ell * (x2 - x1) + 20.0 * x1 * x2) / (9.0 * l4)
    E43 = -50.0 * (X1 + X2) / (9.0 * l4)
    E44 = 25.0 / (9.0 * l4)
    m1, m2, m3, m4 = mom(z_m)
    E4A41 = E40 + m1 * E41 + m2 * E42 + m3 * E43 + m4 * E44
    a1, a2, a3 = tail(z_m, x1)
    b1, b2, b3 = tail(z_m, x2)
    E4A42 = E41 + a1 * E42 + a2 * E43 + a3 * E44
    E4A43 = E41 + b1 * E42 + b2 * E43 + b3 * E44
    P2 = np.exp(-SQRT5 * (x2 - x1) / ell) * (
        0.5 * E4A41 * (erf((x2 - z_m) / sv) - erf((x1 - z_m) / sv))
        + E4A42 * g * np.exp(-0.5 * (x1 - z_m) ** 2 / z_v) - E4A43 * g * np.exp(-0.5 * (x2 - z_m) ** 2 / z_v))

    E50 = 1.0 + (25.0 * x1**2 * x2**2 + 3.0 * SQRT5 * (3.0 * l3 + 5.0 * ell * x1 * x2) * (x1 + x2)
                 + 15.0 * l2 * (x1**2 + x2**2 + 3.0 * x1 * x2)) / (9.0 * l4)
    E51 = (18.0 * SQRT5 * l3 + 15.0 * SQRT5 * ell * (x1**2 + x2**2) + (75.0 * l2 + 50.0 * x1 * x2) * (x1 + x2)
           + 60.0 * SQRT5 * ell * x1 * x2) / (9.0 * l4)
    E52 = 5.0 * (5.0 * x1**2 + 5.0 * x2**2 + 15.0 * l2 + 9.0 * SQRT5 * ell * (x1 + x2) + 20.0 * x1 * x2) / (9.0 * l4)
    E53 = 10.0 * (3.0 * SQRT5 * ell + 5.0 * x1 + 5.0 * x2) / (9.0 * l4)
    E54 = 25.0 / (9.0 * l4)
    muD = z_m + 2.0 * SQRT5 * z_v / ell
    d1, d2, d3, d4 = mom(muD)
    E5A51 = E50 - d1 * E51 + d2 * E52 - d3 * E53 + d4 * E54
    u1, u2, u3 = tail(muD, x1)
    E5A52 = E51 - u1 * E52 + u2 * E53 - u3 * E54
    P3 = np.exp((10.0 * z_v - SQRT5 * ell * (x1 + x2 - 2.0 * z_m)) / l2) * (
        0.5 * E5A51 * (1.0 + erf((x1 - muD) / sv)) + E5A52 * g * np.exp(-0.5 * (x1 - muD) ** 2 / z_v))
    return P1 + P2 + P3


def Jd0(x1, z_m, z_v, ell):
    """Diagonal (X1 == X2) case of `Jd` -- vecchia.py:961-988 (the middle region vanishes)."""
    l2, l3, l4 = ell**2, ell**3, ell**4
    sv = math.sqrt(2.0 * z_v)
    g = math.sqrt(0.5 * z_v / math.pi)
    E30 = 1.0 + (25.0 * x1**4 - 6.0 * SQRT5 * (3.0 * l3 + 5.0 * ell * x1**2) * x1 + 75.0 * l2 * (x1**2)) / (9.0 * l4)
    E31 = (18.0 * SQRT5 * l3 + 90.0 * SQRT5 * ell * x1**2 - (150.0 * l2 + 100.0 * x1**2) * x1) / (9.0 * l4)
    E32 = 5.0 * (30.0 * x1**2 + 15.0 * l2 - 18.0 * SQRT5 * ell * x1) / (9.0 * l4)
    E33 = 10.0 * (3.0 * SQRT5 * ell - 10.0 * x1) / (9.0 * l4)
    E34 = 25.0 / (9.0 * l4)
    muC = z_m - 2.0 * SQRT5 * z_v / ell
    E3A31 = E30 + muC * E31 + (muC**2 + z_v) * E32 + (muC**3 + 3.0 * z_v * muC) * E33 + (
        muC**4 + 6.0 * z_v * muC**2 + 3.0 * z_v**2) * E34
    E3A32 = E31 + (muC + x1) * E32 + (muC**2 + 2.0 * z_v + x1**2 + muC * x1) * E33 + (
        muC**3 + x1**3 + x1 * muC**2 + muC * x1**2 + 3.0 * z_v * x1 + 5.0 * z_v * muC) * E34
    P1 = np.exp((10.0 * z_v + SQRT5 * ell * (2.0 * x1 - 2.0 * z_m)) / l2) * (
        0.5 * E3A31 * (1.0 + erf((muC - x1) / sv)) + E3A32 * g * np.exp(-0.5 * (x1 - muC) ** 2 / z_v))
    E50 = 1.0 + (25.0 * x1**4 + 6.0 * SQRT5 * (3.0 * l3 + 5.0 * ell * x1**2) * x1 + 75.0 * l2 * (x1**2)) / (9.0 * l4)
    E51 = (18.0 * SQRT5 * l3 + 90.0 * SQRT5 * ell * x1**2 + (150.0 * l2 + 100.0 * x1**2) * x1) / (9.0 * l4)
    E52 = 5.0 * (30.0 * x1**2 + 15.0 * l2 + 18.0 * SQRT5 * ell * x1) / (9.0 * l4)
    E53 = 10.0 * (3.0 * SQRT5 * ell + 10.0 * x1) / (9.0 * l4)
    E54 = 25.0 / (9.0 * l4)
    muD = z_m + 2.0 * SQRT5 * z_v / ell
    E5A51 = E50 - muD * E51 + (muD**2 + z_v) * E52 - (muD**3 + 3.0 * z_v * muD) * E53 + (
        muD**4 + 6.0 * z_v * muD**2 + 3.0 * z_v**2) * E54
    E5A52 = E51 - (muD + x1) * E52 + (muD**2 + 2.0 * z_v + x1**2 + muD * x1) * E53 - (
        muD**3 + x1**3 + x1 * muD**2 + muD * x1**2 + 3.0 * z_v * x1 + 5.0 * z_v * muD) * E54
    P3 = np.exp((10.0 * z_v - SQRT5 * ell * (2.0 * x1 - 2.0 * z_m)) / l2) * (
        0.5 * E5A51 * (1.0 + erf((x1 - muD) / sv)) + E5A52 * g * np.exp(-0.5 * (x1 - muD) ** 2 / z_v))
    return P1 + P3


def IJ_matern(X, z_m, z_v, length):
    """functions.py:453-494 (identical maths in vecchia.py:870-906)."""
    n, d = X.shape
    I = np.ones(n)
    J = np.ones((n, n))
    for k in range(d):
        xk = X[:, k]
        I *= I_matern_dim(xk, z_m[k], z_v[k], length[k])
        if z_v[k] != 0:
            with np.errstate(all="ignore"):
                Jk = Jd(xk[None, :], xk[:, None], z_m[k], z_v[k], length[k])
            Jk[np.diag_indices(n)] = Jd0(xk, z_m[k], z_v[k], length[k])
        else:
            Ik = _matern_plain(z_m[k] - xk, length[k])
            Jk = Ik[:, None] * Ik[None, :]
        J *= Jk
    return I, J


def link_gp(m, v, z, w1, global_w1, Rinv, Rinv_y, R2sexp, Psexp, scale, length, nugget, name):
    """Linked-GP predictive moments for Gaussian inputs N(m_t, diag v_t) -- functions.py:396-430.
    mean_t = I'R^-1y;  var_t = |(R^-1y)'J(R^-1y) - mean_t^2 + scale (1 + nugget - tr(R^-1 J))|."""
    M = m.shape[0]
    Dw = w1.shape[1]
    length = np.atleast_1d(np.asarray(length, dtype=np.float64))
    Dz = 0 if z is None else z.shape[1]
    if len(length) == 1:
        length = np.full(Dw + Dz, length[0])
    lw = length[:Dw]
    mo, vo = np.zeros(M), np.zeros(M)
    for t in range(M):
        if name == "sexp" and R2sexp is not None and Psexp is not None:
            I, J = IJ_sexp(w1, m[t], v[t], lw, R2sexp, Psexp)
        elif name == "sexp":
            I, J = IJ_sexp_direct(w1, m[t], v[t], lw)
        else:
            I, J = IJ_matern(w1, m[t], v[t], lw)
        if z is not None:
            Iz = k_vec(global_w1, z[t], length[Dw:], name)
            I = I * Iz
            J = J * np.outer(Iz, Iz)
        tr = trace_sum(Rinv, J)
        mt = I @ Rinv_y
        mo[t] = mt
        vo[t] = np.abs(quad(J, Rinv_y) - mt**2 + scale * (1.0 + nugget - tr))
    return mo, vo


def aggregate(means, variances):
    """Mixture over imputations -- emulation.py:846-847, linkgp.py:493-494."""
    mu = np.mean(means, axis=0)
    sigma2 = np.mean(np.square(means) + variances, axis=0) - mu**2
    return mu, sigma2


# --------------------------------------------------------------------------------------------
# 4. Vecchia                                                   dgpsi/vecchia.py
# --------------------------------------------------------------------------------------------


def _sqdist_rows(a, b):
    """sum_k (a_ik - b_jk)^2 accumulated in ascending k with separate multiply and add
    (the order the CPU kNN libraries use; SURVEY.md section 7 hard part 5)."""
    d = np.zeros((a.shape[0], b.shape[0]))
    for k in range(a.shape[1]):
        diff = a[:, k][:, None] - b[:, k][None, :]
        d += diff * diff
    return d


def nn_ordered(x, m, chunk=2048):
    """Ordered nearest neighbours -- vecchia.py:42-109 (`nn`): row i holds {i} and the m nearest
    earlier points j<i, sorted by index DESCENDING and padded with -1."""
    n = x.shape[0]
    m = min(m, n - 1)
    NN = np.full((n, m + 1), -1, dtype=np.int64)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        d = _sqdist_rows(x[s:e], x[:e])
        for r in range(s, e):
            row = d[r - s, : r + 1]
            k = min(m + 1, r + 1)
            idx = np.argsort(row, kind="stable")[:k]
            NN[r, :k] = idx
    return np.fliplr(np.sort(NN, axis=1))


def knn(query, x, m, chunk=1024):
    """Plain k-nearest neighbours, distance ascending -- vecchia.py:20-40 (`get_pred_nn`)."""
    n = x.shape[0]
    m = min(m, n)
    if m == n:
        k = query.shape[0]
        return (np.arange(m) + np.arange(k)[:, None]) % m
    out = np.empty((query.shape[0], m), dtype=np.int64)
    for s in range(0, query.shape[0], chunk):
        d = _sqdist_rows(query[s:s + chunk], x)
        out[s:s + chunk] = np.argsort(d, axis=1, kind="stable")[:, :m]
    return out


def _block_K(xi, length, nugget_i, name):
    """`K_matrix_nb(xi, length, 0, name)` + `add_to_diag_square` -- vecchia.py:292-321,594-599."""
    K = k_matrix(xi, length, 0.0, name)
    K[np.diag_indices(len(K))] += nugget_i
    return K


def vecchia_llik(X, y, NNarray, scale, length, nugget, nugget_diag, name):
    """vecchia.py:164-180.  X, y already in Vecchia order.  Note: no n*log(scale) term."""
    quad_s, logdet = 0.0, 0.0
    for i in range(X.shape[0]):
        idx = NNarray[i]
        idx = idx[idx >= 0][::-1]
        Ki = _block_K(X[idx], length, nugget * nugget_diag[idx], name)
        Li = np.linalg.cholesky(Ki)
        w = solve_triangular(Li, y[idx, 0], lower=True)
        quad_s += w[-1] ** 2
        logdet += 2.0 * np.log(np.abs(Li[-1, -1]))
    return -0.5 * (logdet + quad_s / scale)


def vecchia_nllik(X, y, NNarray, scale, length, nugget, nugget_diag, name, scale_est, nugget_est):
    """vecchia.py:182-242 (no-replicate branch, origin_n == n).  Returns (nllik, grad, scale)
    WITHOUT prior terms."""
    n = X.shape[0]
    length = np.atleast_1d(length)
    p = len(length) + (1 if nugget_est else 0)
    dquad, dlogdet = np.zeros(p), np.zeros(p)
    quad_s, logdet = 0.0, 0.0
    for i in range(n):
        idx = NNarray[i]
        idx = idx[idx >= 0][::-1]
        xi, yi = X[idx], y[idx, 0]
        nug_i = nugget * nugget_diag[idx]
        Ki, dKi = k_matrix(xi, length, 0.0, name, fod_eval=True, nugget_est=False)
        Ki[np.diag_indices(len(Ki))] = 1.0 + nug_i
        if nugget_est:
            dKi = np.concatenate((dKi, np.diag(nug_i)[None]), axis=0)
        Li = np.linalg.cholesky(Ki)
        w = solve_triangular(Li, yi, lower=True)
        e = np.zeros(len(idx))
        e[-1] = 1.0
        b = solve_triangular(Li.T, e, lower=False)
        for k in range(p):
            LidK = solve_triangular(Li, dKi[k] @ b, lower=True)
            si = w @ LidK
            dquad[k] += 2.0 * si * w[-1] - LidK[-1] * w[-1] ** 2
            dlogdet[k] += LidK[-1]
        quad_s += w[-1] ** 2
        logdet += 2.0 * np.log(np.abs(Li[-1, -1]))
    if scale_est:
        scale = quad_s / n
        nllik = 0.5 * (logdet + n * np.log(scale))
    else:
        nllik = 0.5 * (logdet + quad_s / scale)
    return float(nllik), 0.5 * (dlogdet - dquad / scale), float(scale)


def L_matrix(X, NNarray, length, nugget, name):
    """Rows of the sparse inverse Cholesky factor -- vecchia.py:409-424."""
    n, mm = NNarray.shape
    out = np.zeros((n, mm))
    for i in range(n):
        idx = NNarray[i]
        idx = idx[idx >= 0][::-1]
        b = len(idx)
        Ki = k_matrix(X[idx], length, nugget, name)
        Li = np.linalg.cholesky(Ki)
        e = np.zeros(b)
        e[-1] = 1.0
        out[i, :b] = solve_triangular(Li.T, e, lower=False)[::-1]
    return out


def forward_solve_sp(L, NNarray, b):
    """Sequential sparse lower-triangular solve -- vecchia.py:111-120."""
    n, mm = L.shape
    x = np.zeros(n)
    for i in range(n):
        k = min(i + 1, mm)
        x[i] = (b[i] - np.dot(L[i, 1:k], x[NNarray[i, 1:k]])) / L[i, 0]
    return x


def mvn_draw_vecchia(X, NNarray, scale, length, nugget, name, z):
    """`fmvn_sp` with the standard-normal vector injected -- vecchia.py:133-140."""
    L = L_matrix(X, NNarray, length, nugget, name) / np.sqrt(scale)
    return forward_solve_sp(L, NNarray, z)


def gp_vecch(x, w, NNarray, y, scale, length, nugget, nugget_diag, name):
    """vecchia.py:635-654."""
    M = x.shape[0]
    mo, vo = np.zeros(M), np.zeros(M)
    for t in range(M):
        idx = NNarray[t]
        idx = idx[idx >= 0]
        Xi = np.vstack((w[idx], x[t:t + 1]))
        nug = np.append(nugget * nugget_diag[idx], nugget)
        Li = np.linalg.cholesky(_block_K(Xi, length, nug, name))
        mo[t] = Li[-1, :-1] @ solve_triangular(Li[:-1, :-1], y[idx, 0], lower=True)
        vo[t] = scale * Li[-1, -1] ** 2
    return mo, vo


def loo_gp_dense(Y, Rinv, Rinv_y, scale):
    """Closed-form leave-one-out moments of a dense GP -- gp.py:353-359."""
    s2 = (1.0 / np.diag(Rinv)).reshape(-1, 1)
    return Y - Rinv_y.reshape(-1, 1) * s2, scale * s2


def loo_gp_vecch(X, Y, m, scale, length, nugget, name):
    """gp.loo in Vecchia mode -- gp.py:343-352 + loo_gp_vecch vecchia.py:657-673: every training point predicted
    from its m nearest OTHER training points (the first neighbour returned by the search is the point itself)."""
    Xs = _scaled(X, length)
    NN = knn(Xs, Xs, m + 1)[:, 1:]
    mu, var = gp_vecch(X, X, NN, Y, scale, length, nugget, np.ones(len(X)), name)
    return mu.reshape(-1, 1), var.reshape(-1, 1)


def mice_var(x, x_extra, input_dim, connect, name, length, scale, nugget, nugget_s):
    """Smoothed predictive variances of a GP node on the candidate set -- functions.py:244-256."""
    from scipy.linalg import pinvh
    kin = x[:, input_dim]
    if connect is not None:
        kin = np.concatenate((kin, x_extra[:, connect]), 1)
    R = k_matrix(kin, length, max(nugget_s, nugget), name)
    return (scale / np.diag(pinvh(R, check_finite=False))).reshape(-1, 1)


def mice_score(pred_inputs, variances, x_cand, last_layers, nugget_s):
    """MICE criterion from the per-imputation moments -- emulation.py:378-392.  last_layers[s][k] =
    (input_dim, connect, name, length, scale, nugget) of output node k in imputation s."""
    S = len(pred_inputs)
    score = np.zeros_like(variances[0])
    for s in range(S):
        smooth = np.hstack([mice_var(pred_inputs[s], x_cand, *node, nugget_s) for node in last_layers[s]])
        with np.errstate(divide='ignore'):
            score += np.log(variances[s] / smooth)
    return score / S


def vigf_score(bias, sigma2):
    """VIGF criterion from the squared biases and variances of the imputations -- emulation.py:409-413."""
    bias, sigma2 = np.asarray(bias), np.asarray(sigma2)
    E1 = np.mean(np.square(bias) + 6 * bias * sigma2 + 3 * np.square(sigma2), axis=0)
    E2 = np.mean(bias + sigma2, axis=0)
    return E1 - E2 ** 2


def cond_mean(x, z, w1, global_w1, y, length, nugget, name):
    """GP conditional mean at new inputs -- functions.py:301-309 with R^-1 y from the Cholesky solve of
    dgp.py:913-915."""
    if z is not None:
        x, w1 = np.concatenate((x, z), 1), np.concatenate((w1, global_w1), 1)
    Lc = np.linalg.cholesky(k_matrix(w1, length, nugget, name))
    return k_cross(w1, x, length, name) @ cho_solve((Lc, True), y).ravel()


def cond_mean_vecch(x, z, w1, global_w1, y, scale, length, nugget, name, m=50):
    """vecchia.py:624-633."""
    if z is not None:
        x, w1 = np.concatenate((x, z), 1), np.concatenate((w1, global_w1), 1)
    NN = knn(_scaled(x, length), _scaled(w1, length), m)
    return gp_vecch(x, w1, NN, y, scale, length, nugget, np.ones(len(y)), name)[0]


def grow_layers(nodes, X, Y, sub_idx, vecch):
    """dgp.update_all_layer_larger -- dgp.py:886-1012 (GP nodes, no replicates).  nodes[l][k] is a dict with
    input, global_input (or None), output, length, scale, nugget, name, input_dim, connect; updated in place."""
    In = X.copy()
    mask = np.zeros(len(X), dtype=bool)
    mask[sub_idx] = True
    for l, layer in enumerate(nodes):
        last = l == len(nodes) - 1
        Out = np.empty((len(In), len(layer)))
        for k, nd in enumerate(layer):
            if not last:
                xin = In[~mask][:, nd["input_dim"]]
                zin = None if nd["connect"] is None else X[~mask][:, nd["connect"]]
                if vecch:
                    mu = cond_mean_vecch(xin, zin, nd["input"], nd["global_input"], nd["output"], nd["scale"][0],
                                         nd["length"], nd["nugget"][0], nd["name"])
                else:
                    mu = cond_mean(xin, zin, nd["input"], nd["global_input"], nd["output"], nd["length"],
                                   nd["nugget"][0], nd["name"])
                Out[sub_idx, k] = nd["output"].ravel()
                Out[~mask, k] = mu
                nd["output"] = Out[:, [k]].copy()
            else:
                nd["output"] = Y[:, [k]].copy()
            nd["input"] = In[:, nd["input_dim"]].copy()
            if nd["connect"] is not None:
                nd["global_input"] = X[:, nd["connect"]].copy()
        if not last:
            In = Out.copy()
    return nodes


def link_gp_vecch(m, v, z, w1, global_w1, NNarray, y, scale, length, nugget, nugget_diag, name):
    """vecchia.py:758-796 with `IJ_nb` (vecchia.py:838-907)."""
    M = m.shape[0]
    Dw = w1.shape[1]
    length = np.atleast_1d(np.asarray(length, dtype=np.float64))
    Dz = 0 if z is None else z.shape[1]
    if len(length) == 1:
        length = np.full(Dw + Dz, length[0])
    lw = length[:Dw]
    mo, vo = np.zeros(M), np.zeros(M)
    for t in range(M):
        idx = NNarray[t]
        idx = idx[idx >= 0]
        wi = w1[idx]
        if name == "sexp":
            I, J = IJ_sexp_direct(wi, m[t], v[t], lw)
        else:
            I, J = IJ_matern(wi, m[t], v[t], lw)
        if z is not None:
            gi = global_w1[idx]
            Iz = k_vec(gi, z[t], length[Dw:], name)
            I, J = I * Iz, J * np.outer(Iz, Iz)
            Ki = k_matrix(np.concatenate((wi, gi), 1), length, 0.0, name)
        else:
            Ki = k_matrix(wi, length, 0.0, name)
        Ki[np.diag_indices(len(Ki))] += nugget * nugget_diag[idx]
        tr = np.trace(np.linalg.solve(Ki, J))
        Li = np.linalg.cholesky(Ki)
        a = cho_solve((Li, True), y[idx, 0])
        mt = I @ a
        mo[t] = mt
        vo[t] = np.abs(quad(J, a) - mt**2 + scale * (1.0 + nugget - tr))
    return mo, vo


# --------------------------------------------------------------------------------------------
# 3. elliptical slice sampling                                 dgpsi/imputation.py:44-119
# --------------------------------------------------------------------------------------------


class Node:
    """Minimal GP-node record for the ESS / SEM restatement (mirrors the attributes the
    reference's `kernel` uses on this path, kernel_class.py:86-144)."""

    def __init__(self, length, scale=1.0, nugget=1e-6, name="sexp", scale_est=False, nugget_est=False,
                 input_dim=None, connect=None, prior_name="ga", prior_coef=None):
        self.length = np.atleast_1d(np.asarray(length, dtype=np.float64)).copy()
        self.scale = float(scale)
        self.nugget = float(nugget)
        self.name = name
        self.scale_est, self.nugget_est = scale_est, nugget_est
        self.input_dim, self.connect = input_dim, connect
        self.prior_name = prior_name
        if prior_coef is None:
            prior_coef = np.array([1.6, 0.3])
        pc = np.array(prior_coef, dtype=np.float64)
        if prior_name == "ga":
            pc[0] -= 1.0
        elif prior_name == "inv_ga":
            pc[0] += 1.0
        self.prior_coef = pc
        self.input = self.global_input = self.output = None
        self.vecch = False
        self.ord = self.rev_ord = self.NNarray = None
        self.m = 25

    def X(self):
        if self.global_input is not None:
            return np.concatenate((self.input, self.global_input), 1)
        return self.input

    def loglik(self):
        if self.vecch:
            X = self.X()
            return vecchia_llik(X[self.ord], self.output[self.ord], self.NNarray, self.scale, self.length,
                                self.nugget, np.ones(len(self.output)), self.name)
        return loglik_dense(self.X(), self.output, self.length, self.scale, self.nugget, self.name)

    def prior_draw(self, z):
        """fmvn / fmvn_sp with injected standard normals -- functions.py:113-121, vecchia.py:133-140."""
        if self.vecch:
            X = self.X()
            return mvn_draw_vecchia(X[self.ord], self.NNarray, self.scale, self.length, self.nugget, self.name,
                                    z)[self.rev_ord]
        return np.linalg.cholesky(self.scale * k_matrix(self.X(), self.length, self.nugget, self.name)) @ z

    def set_ord_nn(self, ord):
        self.ord = np.asarray(ord)
        self.rev_ord = np.argsort(self.ord)
        self.NNarray = nn_ordered((self.X() / self.length)[self.ord], self.m)

    def objective(self, x):
        """`llik` / `llik_vecch` including prior terms -- kernel_class.py:403-479."""
        theta = np.exp(x)
        if self.nugget_est:
            self.length, self.nugget = theta[:-1], float(theta[-1])
        else:
            self.length = theta
        if self.vecch:
            X = self.X()
            f, g, s = vecchia_nllik(X[self.ord], self.output[self.ord], self.NNarray, self.scale, self.length,
                                    self.nugget, np.ones(len(self.output)), self.name, self.scale_est,
                                    self.nugget_est)
        else:
            f, g, s = nllik_grad_dense(self.X(), self.output, self.length, self.scale, self.nugget, self.name,
                                       self.scale_est, self.nugget_est)
        self.scale = s
        f -= log_prior(self.length, self.nugget, self.prior_name, self.prior_coef, self.nugget_est)
        g = g - log_prior_fod(self.length, self.nugget, self.prior_name, self.prior_coef, self.nugget_est)
        return f, g

    def maximise(self):
        """L-BFGS-B M-step with the reference's options -- kernel_class.py:516-578 ('ga' prior branch)."""
        from scipy.optimize import Bounds, minimize

        x0 = np.log(np.concatenate((self.length, [self.nugget]))) if self.nugget_est else np.log(self.length)
        D = self.input.shape[1] + (0 if self.global_input is None else self.global_input.shape[1])
        opts = {"maxiter": 100, "maxfun": int(np.max((30, 20 + 5 * D)))}
        if self.nugget_est:
            lb = np.concatenate((-np.inf * np.ones(len(x0) - 1), np.log([1e-8])))
            minimize(self.objective, x0, method="L-BFGS-B", jac=True, bounds=Bounds(lb, np.inf * np.ones(len(x0))),
                     options=opts)
        else:
            minimize(self.objective, x0, method="L-BFGS-B", jac=True, options=opts)


class LikNode:
    """Likelihood node of a final layer -- likelihood_class.py: Poisson :39-48, Hetero :110-116, NegBin :264-272."""

    def __init__(self, name, input_dim, output, link=None, eps=1e-3):
        self.name, self.input_dim, self.output = name, np.asarray(input_dim), output
        self.link, self.eps = link, eps
        self.input = None

    def loglik(self):
        from scipy.special import gammaln, log_ndtr
        y, f = self.output.flatten(), self.input
        if self.name == "Categorical":   # likelihood_class.py:333-359
            K = f.shape[1]
            if self.link == "logit":
                return np.sum(y * f[:, 0] - np.logaddexp(0, f[:, 0]))
            if self.link == "probit":
                return np.sum(y * log_ndtr(f[:, 0]) + (1 - y) * log_ndtr(-f[:, 0]))
            if self.link == "robustmax":
                hit = np.argmax(f, axis=1) == y.astype(int)
                return np.sum(np.where(hit, np.log(1.0 - self.eps), np.log(self.eps / (K - 1))))
            top = np.max(f, axis=1)
            lse = np.log(np.sum(np.exp(f - top[:, None]), axis=1)) + top
            return np.sum(f[np.arange(len(y)), y.astype(int)] - lse)
        if self.name in ("ZIP", "ZINB"):   # likelihood_class.py:497-525, 653-693
            from scipy.special import expit
            if self.name == "ZIP":
                count, fpi = -np.exp(f[:, 0]) + y * f[:, 0] - gammaln(y + 1.0), f[:, 1]
            else:
                n, a, fpi = np.exp(-f[:, 1]), f[:, 0] + f[:, 1], f[:, 2]
                count = gammaln(y + n) - gammaln(n) - gammaln(y + 1.0) + y * a - (y + n) * np.logaddexp(0.0, a)
            pi = expit(fpi)
            ll = np.where(y == 0, np.logaddexp(np.log(pi), np.log1p(-pi) + count), np.log1p(-pi) + count)
            return np.sum(ll)
        if self.name == "Poisson":
            return np.sum(y * f[:, 0] - np.exp(f[:, 0]) - gammaln(y + 1))
        if self.name == "Hetero":
            with np.errstate(divide="ignore"):
                return np.sum(-0.5 * (np.log(2 * np.pi) + f[:, 1] + np.exp(np.log((y - f[:, 0]) ** 2) - f[:, 1])))
        n = np.exp(-f[:, 1])
        a = f[:, 0] + f[:, 1]
        return np.sum(gammaln(y + n) - gammaln(n) - gammaln(y + 1.0) + y * a - (y + n) * np.logaddexp(0.0, a))

    def prediction(self, m, v):
        """Moments of the observable from Gaussian moments of the latent inputs -- likelihood_class.py:65-78,
        124-128, 283-287."""
        if self.name == "Poisson":
            return (np.exp(m + v / 2)).flatten(), (np.exp(m + v / 2) + (np.exp(v) - 1) * np.exp(2 * m + v)).flatten()
        if self.name == "Hetero":
            return m[:, 0], np.exp(m[:, 1] + v[:, 1] / 2) + v[:, 0]
        return (np.exp(m[:, 0] + v[:, 0] / 2),
                np.exp(2 * m[:, 0] + v[:, 0]) * (np.exp(v[:, 0]) - 1) + np.exp(m[:, 0] + v[:, 0] / 2)
                + np.exp(m[:, 1] + v[:, 1] / 2) * np.exp(2 * m[:, 0] + 2 * v[:, 0]))


def post_het1(v, Gamma, y, sd):
    """Exact conditional draw of the Hetero mean process with injected normals sd (n x 2) --
    likelihood_class.py:185-210."""
    Lc = cholesky(v + np.diag(Gamma), lower=True)
    L1 = cholesky(v, lower=True)
    mu = v @ cho_solve((Lc, True), y.flatten())
    u = L1 @ sd[:, 0]
    w = np.sqrt(Gamma) * sd[:, 1]
    return -v @ cho_solve((Lc, True), u + w) + (mu + u)


def ess_block(targets, uppers, z, u):
    """One blocked ESS update of a latent layer with INJECTED randomness -- imputation.py:44-119.

    targets: list of Node producing the layer;  uppers: list of Node fed by it.
    z: (M x n) standard normals (row k -> target k);  u: iterator/array of U(0,1) draws consumed
    in the reference's order: threshold, initial angle, then one per rejection.
    Returns (thetas tried, number of uniforms consumed)."""
    M, n = len(targets), len(targets[0].output)
    f, nu = np.zeros((n, M)), np.zeros((n, M))
    for k, t in enumerate(targets):
        f[:, k] = t.output.flatten()
        nu[:, k] = t.prior_draw(z[k])
    u = list(u)
    ui = 0
    log_y = sum(up.loglik() for up in uppers) + np.log(u[ui]); ui += 1
    theta = u[ui] * 2.0 * np.pi; ui += 1
    tmin, tmax = theta - 2.0 * np.pi, theta
    thetas = []
    while True:
        thetas.append(theta)
        fp = f * np.cos(theta) + nu * np.sin(theta)  # update_f, functions.py:203-208
        for up in uppers:
            up.input = fp[:, up.input_dim]
        if sum(up.loglik() for up in uppers) > log_y:
            for k, t in enumerate(targets):
                t.output[:, 0] = fp[:, k]
            return thetas, ui
        if theta < 0.0:
            tmin = theta
        else:
            tmax = theta
        theta = tmin + (tmax - tmin) * u[ui]; ui += 1


def ess_one(target, k, uppers, z, u):
    """Node-wise ESS update of target node k of its layer with INJECTED randomness -- imputation.py:166-221: only
    column `input_dim == k` of every linked upper node moves.  Returns (thetas tried, uniforms consumed)."""
    f = target.output.flatten()
    nu = target.prior_draw(z)
    u = list(u)
    ui = 0
    log_y = sum(up.loglik() for up in uppers) + np.log(u[ui]); ui += 1
    theta = u[ui] * 2.0 * np.pi; ui += 1
    tmin, tmax = theta - 2.0 * np.pi, theta
    thetas = []
    while True:
        thetas.append(theta)
        fp = f * np.cos(theta) + nu * np.sin(theta)
        for up in uppers:
            up.input[:, np.asarray(up.input_dim) == k] = fp.reshape(-1, 1)
        if sum(up.loglik() for up in uppers) > log_y:
            target.output[:, 0] = fp
            return thetas, ui
        if theta < 0.0:
            tmin = theta
        else:
            tmax = theta
        theta = tmin + (tmax - tmin) * u[ui]; ui += 1


def ess_sweeps(all_layer, burnin, rng, max_u=64):
    """`imputer.sample(burnin)` (imputation.py:22-42, block=True) drawing from `rng`
    (numpy Generator): per block update M*n normals then `max_u` uniforms are drawn up front so
    the stream position does not depend on the number of rejections."""
    n_prop = 0
    for _ in range(burnin + 1):
        for l in range(len(all_layer) - 1):
            tg, up = all_layer[l], all_layer[l + 1]
            z = rng.standard_normal((len(tg), len(tg[0].output)))
            u = rng.random(max_u)
            th, _ = ess_block(tg, up, z, u)
            n_prop += len(th)
    return n_prop


def build_dgp(X, Y, all_layer):
    """Generic branch of `dgp.initialize` -- dgp.py:565-691: wire inputs/outputs of every node.
    Latent layers start as copies of their input (same width), or column-resampled when wider."""
    In = X
    L = len(all_layer)
    for l, layer in enumerate(all_layer):
        nk = len(layer)
        if l != L - 1:
            if In.shape[1] == nk:
                Out = In.copy()
            elif In.shape[1] < nk:
                Out = np.concatenate((In, In[:, np.random.choice(In.shape[1], nk - In.shape[1])]), 1)
            else:
                raise NotImplementedError("oracle: narrowing layers use KPCA in the reference (dgp.py:568-574)")
        for k, node in enumerate(layer):
            if node.input_dim is None:
                node.input_dim = np.arange(In.shape[1])
            node.input = In[:, node.input_dim].copy()
            if node.connect is not None:
                node.global_input = X[:, node.connect]
            node.output = Y[:, [k]].copy() if l == L - 1 else Out[:, [k]].copy()
        if l != L - 1:
            In = Out.copy()
    return all_layer


def sem_iteration(all_layer, rng, ess_burn=10):
    """One stochastic-EM iteration -- dgp.py:1381-1398: I-step (ess_burn+1 sweeps) then the
    M-step over every GP node.  Returns the number of ESS proposals evaluated."""
    n_prop = ess_sweeps(all_layer, ess_burn, rng)
    for layer in all_layer:
        for node in layer:
            node.maximise()
    return n_prop


# ------------------------------------------------------------------------------------------------
# Hetero likelihood: exact draw of the mean process under the Vecchia approximation
# ------------------------------------------------------------------------------------------------
def imp_nn_array(Xo_scaled, m):
    """Conditioning sets of the latent-Vecchia draw -- kernel_class.py:268-273.  Xo_scaled: inputs / length in
    Vecchia order.  Row i = [i + n (latent f_i), i (observed y_i), the m - 1 nearest other points: index + n when
    they come earlier in the ordering (latent), plain index when later (observed)]."""
    n = Xo_scaled.shape[0]
    NNs = knn(Xo_scaled, Xo_scaled, m)[:, 1:]
    prev = NNs < np.arange(n)[:, None]
    NNs = np.where(prev, NNs + n, NNs)
    return np.hstack((np.arange(n).reshape(-1, 1) + n, np.arange(n).reshape(-1, 1), NNs))


def hetero_u_matrix(Xo, imp_NN, scale, length, name, gamma_o):
    """Entries of the sparse U in imp_NN's own order (n x m1) -- U_matrix, vecchia.py:426-445, whose rows are these
    reversed: per row K = scale corr(x) + diag(gamma on the observed entries + 1e-10), u = chol(K)^-T e_last."""
    n, m1 = imp_NN.shape
    out = np.zeros((n, m1))
    for i in range(n):
        idx = imp_NN[i]
        idx = idx[idx >= 0][::-1]
        b = len(idx)
        Ki = scale * k_matrix(Xo[idx % n], length, 0.0, name)
        Ki[np.diag_indices(b)] += np.where(idx >= n, 0.0, gamma_o[idx % n]) + 1e-10
        Li = np.linalg.cholesky(Ki)
        e = np.zeros(b)
        e[-1] = 1.0
        out[i, :b] = solve_triangular(Li.T, e, lower=False)[::-1]
    return out


def hetero_vecchia_draw(Xo, imp_NN, scale, length, name, gamma_o, y_o, sd):
    """Hetero.post_het_vecch (likelihood_class.py:165-183) on the U of `hetero_u_matrix`: L = U_latent^T,
    mu = -L^-1 U_obs^T y, sample = L^-1 sd; returns mu + sample in Vecchia order."""
    n, m1 = imp_NN.shape
    U = hetero_u_matrix(Xo, imp_NN, scale, length, name, gamma_o)
    f = np.zeros(n)
    for i in range(n):   # forward substitution, row i of L = the latent entries of row i of imp_NN (diagonal first)
        acc = sd[i]
        for c in range(1, m1):
            e = imp_NN[i, c]
            if e < 0:
                continue
            if e >= n:
                acc -= U[i, c] * f[e - n]
            else:
                acc -= U[i, c] * y_o[e]
        f[i] = acc / U[i, 0]
    return f
