"""Vecchia (nearest-neighbour conditioning) entry points with the reference's function names and argument
meaning (dgpsi/vecchia.py), each a thin wrapper over libdgpb.so.  numpy in / numpy out unless the name ends
in `_dev`."""
from __future__ import annotations

import numpy as np

from . import _lib as L


def get_pred_nn_dev(query, x, m=50):
    """Plain kNN on device tensors (distance ascending, ties -> smaller index) -- vecchia.py:20-40."""
    query, x = query.contiguous(), x.contiguous()
    M, D = query.shape
    n = x.shape[0]
    m = min(int(m), n)
    NN = L.empty((M, m), "i8")
    L.check(L.load().dgpb_knn(L.ptr(query), M, L.ptr(x), n, D, m, L.ptr(NN), L.stream()))
    return NN


def get_pred_nn(query, x, m=50, method='exact', **_):
    """vecchia.py:20-40.  Only the exact search exists here (the reference's default)."""
    if method != 'exact':
        raise NotImplementedError("dgp_b200 implements the exact neighbour search only")
    return L.to_host(get_pred_nn_dev(L.to_dev(query, np.float64), L.to_dev(x, np.float64), m))


def nn(x, m, method='exact', **_):
    """Ordered nearest neighbours (vecchia.py:42-109): row i = {i} + the m nearest j < i, index-descending,
    -1 padded."""
    if method != 'exact':
        raise NotImplementedError("dgp_b200 implements the exact neighbour search only")
    xd = L.to_dev(x, np.float64)
    n, D = xd.shape
    m = min(int(m), n - 1)
    NN = L.empty((n, m + 1), "i8")
    L.check(L.load().dgpb_knn_ordered(L.ptr(xd), n, D, m, L.ptr(NN), L.stream()))
    return L.to_host(NN)


def _prep(X, y, NNarray, nugget_diag):
    Xd = L.to_dev(X, np.float64)
    yd = None if y is None else L.to_dev(np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1)))
    NNd = L.to_dev(NNarray, np.int64)
    nd = None if nugget_diag is None else L.to_dev(nugget_diag, np.float64)
    return Xd, yd, NNd, nd


def vecchia_llik(X, y, NNarray, scale, length, nugget, nugget_diag, name):
    """vecchia.py:164-180 (X, y in Vecchia order)."""
    Xd, yd, NNd, nd = _prep(X, y, NNarray, nugget_diag)
    larr, lptr = L.length_host(length)
    out = L.host_doubles(1)
    L.check(L.load().dgpb_vecchia_llik(L.ptr(Xd), L.ptr(yd), L.ptr(NNd), Xd.shape[0], Xd.shape[1], NNd.shape[1], lptr,
                                       len(larr), float(scale), float(nugget), L.ptr(nd), L.KIND[name], out,
                                       L.stream()))
    return out[0]


def vecchia_nllik(X, y, NNarray, scale, length, nugget, nugget_diag, name, scale_est, nugget_est):
    """vecchia.py:182-242 (no replicates).  Returns (nllik, grad, scale) without prior terms."""
    Xd, yd, NNd, nd = _prep(X, y, NNarray, nugget_diag)
    larr, lptr = L.length_host(length)
    P = len(larr) + (1 if nugget_est else 0)
    out = L.host_doubles(P + 2)
    L.check(L.load().dgpb_vecchia_nllik(L.ptr(Xd), L.ptr(yd), L.ptr(NNd), Xd.shape[0], Xd.shape[1], NNd.shape[1], lptr,
                                        len(larr), float(scale), float(nugget), L.ptr(nd), L.KIND[name],
                                        int(bool(scale_est)), int(bool(nugget_est)), out, L.stream()))
    return out[0], np.array(out[2:2 + P]), out[1]


def L_matrix(X, NNarray, length, nugget, name):
    """vecchia.py:409-424."""
    Xd, _, NNd, _ = _prep(X, None, NNarray, None)
    larr, lptr = L.length_host(length)
    out = L.empty((Xd.shape[0], NNd.shape[1]))
    L.check(L.load().dgpb_vecchia_Lmatrix(L.ptr(Xd), L.ptr(NNd), Xd.shape[0], Xd.shape[1], NNd.shape[1], lptr,
                                          len(larr), float(nugget), L.KIND[name], L.ptr(out), L.stream()))
    return L.to_host(out)


def fmvn_sp(X, NNarray, scale, length, nugget, name, z=None):
    """vecchia.py:133-140; `z` (standard normals) may be injected, else drawn from numpy's global RNG as the
    reference does."""
    Xd, _, NNd, _ = _prep(X, None, NNarray, None)
    n = Xd.shape[0]
    if z is None:
        z = np.random.randn(n)
    zd = L.to_dev(z, np.float64)
    larr, lptr = L.length_host(length)
    out = L.empty((n,))
    L.check(L.load().dgpb_vecchia_mvn_draw(L.ptr(Xd), L.ptr(NNd), n, Xd.shape[1], NNd.shape[1], lptr, len(larr),
                                           float(scale), float(nugget), L.KIND[name], L.ptr(zd), L.ptr(out),
                                           L.stream()))
    return L.to_host(out)
