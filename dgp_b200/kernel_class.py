"""`kernel` -- the GP node of a DGP hierarchy, with the reference's constructor, attributes and method
names (dgpsi/kernel_class.py:9-764) and every numeric method routed to libdgpb.so (sm_100a CUDA).

Host-side state stays in numpy on the object (so pickles and user code keep working, SURVEY.md section 5
"checkpoint/resume"); device tensors are caches.  Host-only pieces kept in Python exactly where the
reference has them: priors (:361-401), `compute_cl` (:207-225), `r2` (:227-243), the L-BFGS-B driver
(:516-579).  Not supported (out of the hot-path scope): replicate pooling (`rep`),
`ord_nn(pointer=True)` (Hetero exact posterior).
"""
from __future__ import annotations

import ctypes

import numpy as np
from numpy.linalg import LinAlgError, lstsq, matrix_rank
from scipy.optimize import Bounds, minimize

from . import _lib as L


class kernel:
    """GP node.  Arguments as in the reference (kernel_class.py:86)."""

    def __init__(self, length, scale=1., nugget=1e-6, name='sexp', prior_name='ga', prior_coef=None, bds=None,
                 nugget_est=False, scale_est=False, input_dim=None, connect=None):
        if name not in L.KIND:
            raise ValueError("name must be 'sexp' or 'matern2.5'")
        self.type = 'gp'
        self.length = length
        self.scale = np.atleast_1d(scale)
        self.nugget = np.atleast_1d(nugget)
        self.name = name
        self.prior_name = prior_name
        # stored prior coefficients follow kernel_class.py:93-110 (shape shifted by -1 / +1)
        if prior_name == 'ga':
            self.prior_coef = np.array([1.6, 0.3]) if prior_coef is None else prior_coef
            self.prior_coef[0] -= 1
        elif prior_name == 'inv_ga':
            self.prior_coef = np.array([1.6, 0.3]) if prior_coef is None else prior_coef
            self.prior_coef[0] += 1
        elif prior_name == 'ref':
            self.prior_coef = np.array([0.2]) if prior_coef is None else prior_coef
            self.cl = None
        self.nugget_est = nugget_est
        self.scale_est = scale_est
        self.input_dim = input_dim
        self.connect = connect
        self.para_path = None
        self.global_input = None
        self.input = None
        self.output = None
        self.rep = None
        self.rep_hetero = None
        self._Rinv = None
        self._Rinv_y = None
        self.R2sexp = None
        self.Psexp = None
        self.vecch = None
        self.D = None
        self.ord = None
        self.rev_ord = None
        self.m = None
        self.pred_m = None
        self.NNarray = None
        self.max_rep = None
        self.imp_NNarray = None
        self.imp_pointer_row = None
        self.imp_pointer_col = None
        self.nn_method = 'exact'
        self.ord_fun = None
        self.iter_count = 0
        self.target = 'dgp'
        self.bds = bds
        self.R2 = None
        self.loo_state = False
        self.sum_residual = None
        self.W_diag = None
        self._dcache = None

    # ---- R^-1 / R^-1 y live on the device; numpy views are produced on demand -------------------
    @property
    def Rinv(self):
        return _as_numpy(self._Rinv)

    @Rinv.setter
    def Rinv(self, value):
        self._Rinv = value

    @property
    def Rinv_y(self):
        return _as_numpy(self._Rinv_y)

    @Rinv_y.setter
    def Rinv_y(self, value):
        self._Rinv_y = value

    def __getstate__(self):
        state = dict(self.__dict__)
        state['Rinv'] = _as_numpy(state.pop('_Rinv'))
        state['Rinv_y'] = _as_numpy(state.pop('_Rinv_y'))
        for key in ('_dcache', '_batcher', '_vcache', '_frozen', '_Xcat', '_ycol', '_mid', '_nfev', '_r2_fresh'):
            state.pop(key, None)
        return state

    def __deepcopy__(self, memo):
        """`copy.deepcopy` of a node (emulator.__init__ keeps one copy of the hierarchy per imputation,
        emulation.py:44): numpy state is copied, R^-1 / R^-1 y stay the SAME device tensors -- `compute_stats` always
        binds fresh tensors, nothing updates them in place -- instead of a round trip of n^2 doubles through the host."""
        import copy as _copy
        new = type(self).__new__(type(self))
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            if key in ('_Rinv', '_Rinv_y'):
                new.__dict__[key] = val if not isinstance(val, np.ndarray) else val.copy()
            elif key in ('_dcache', '_batcher', '_vcache', '_Xcat', '_ycol'):
                new.__dict__[key] = None
            elif key == '_frozen':
                new.__dict__[key] = False
            else:
                new.__dict__[key] = _copy.deepcopy(val, memo)
        return new

    def __setstate__(self, state):
        state = dict(state)
        state['_Rinv'] = state.pop('Rinv', None)
        state['_Rinv_y'] = state.pop('Rinv_y', None)
        state['_dcache'] = None
        self.__dict__.update(state)

    # ---- small host helpers ------------------------------------------------------------------
    def _X(self):
        if self.global_input is not None:
            return np.concatenate((self.input, self.global_input), 1)
        return self.input

    def _X_pred(self):
        """`_X()` for predictions; a node marked `_frozen` (an emulator's imputation: its data no longer changes)
        keeps the concatenated array, so the emulator's upload pool recognises it by identity on every call."""
        if not getattr(self, '_frozen', False):
            return self._X()
        if getattr(self, '_Xcat', None) is None:
            self._Xcat = np.ascontiguousarray(self._X())
        return self._Xcat

    def _y_pred(self):
        if not getattr(self, '_frozen', False):
            return np.ascontiguousarray(self.output[:, 0])
        if getattr(self, '_ycol', None) is None:
            self._ycol = np.ascontiguousarray(self.output[:, 0])
        return self._ycol

    def _check_supported(self):
        if self.rep is not None:
            raise NotImplementedError("dgp_b200: replicate pooling (rep / W_diag) is outside the SI hot path")

    def compute_cl(self):
        """kernel_class.py:207-225 (reference-prior scale; host, O(n D))."""
        X = self._X()
        if len(self.length) == 1:
            if self.vecch:
                rng = np.max(X, axis=0) - np.min(X, axis=0)
                self.cl = np.sqrt(np.dot(rng, rng)) / len(self.output)
            else:
                from scipy.spatial.distance import pdist
                self.cl = np.max(pdist(X, metric="euclidean")) / len(self.output)
        else:
            rng = np.max(X, axis=0) - np.min(X, axis=0)
            self.cl = rng / len(self.output) ** (1 / len(self.length))

    def r2(self, overwritten=False):
        """R2 of the linear regression of `input` on `global_input` (kernel_class.py:227-243; host)."""
        if self.global_input is not None:
            X = np.concatenate((self.global_input, np.ones((len(self.global_input), 1))), axis=1)
            if matrix_rank(self.global_input) == matrix_rank(X):
                X = self.global_input
            N, D = X.shape
            if N == D:
                resids = np.zeros(self.input.shape[1], dtype=float)
            else:
                _, resids = lstsq(X, self.input, rcond=None)[:2]
            rsq = 1 - resids / (len(self.input) * np.var(self.input, axis=0))
            if overwritten or self.R2 is None:
                self.R2 = np.atleast_2d(rsq)
            else:
                self.R2 = np.vstack((self.R2, rsq))
            self._r2_fresh = True

    def ord_nn(self, ord=None, NNarray=None, pointer=False):
        """Vecchia ordering and ordered nearest neighbours (kernel_class.py:245-267); the neighbour search
        runs on the GPU (exact FP64 brute force, bit-exact with the reference's exact kNN)."""
        X = self._X() / self.length
        if ord is None:
            self.ord = np.random.permutation(self.input.shape[0]) if self.ord_fun is None else self.ord_fun(X)
        else:
            self.ord = ord
        self.rev_ord = np.argsort(self.ord)
        if NNarray is None:
            from .vecchia import nn
            self.NNarray = nn(X[self.ord], self.m)
        else:
            self.NNarray = NNarray
        if pointer:
            # conditioning sets of the latent-Vecchia draw of a mean process under a Hetero likelihood
            # (kernel_class.py:268-275): point i of the ordering -> [f_i (index i + n), y_i (index i), its m - 1
            # nearest other points: as latent values (index + n) when they come earlier in the ordering, as
            # observations when they come later]
            from .vecchia import get_pred_nn
            Xo = np.ascontiguousarray(X[self.ord])
            n = Xo.shape[0]
            NNs = get_pred_nn(Xo, Xo, self.m)[:, 1:]
            prev = NNs < np.arange(n)[:, None]
            NNs = np.where(prev, NNs + n, NNs)
            self.imp_NNarray = np.hstack((np.arange(n).reshape(-1, 1) + n, np.arange(n).reshape(-1, 1), NNs))
            # the reference also keeps COO pointers of the sparse U (imp_pointers, vecchia.py:462-476); here the
            # entries stay in imp_NNarray's own order, so only their presence is recorded
            self.imp_pointer_row = self.imp_pointer_col = True

    def log_t(self):
        if self.nugget_est:
            return np.log(np.concatenate((self.length, self.nugget)))
        return np.log(self.length)

    def update(self, log_theta):
        theta = np.exp(log_theta)
        if self.nugget_est:
            self.length = theta[0:-1]
            self.nugget = theta[[-1]]
        else:
            self.length = theta

    # ---- priors (host; kernel_class.py:361-401, functions.py:95-100) ---------------------------
    def gfod(self, x):
        if self.prior_name == 'ga':
            return self.prior_coef[0] - self.prior_coef[1] * x
        return -self.prior_coef[0] + self.prior_coef[1] / x

    def _g(self, x):
        a, b = self.prior_coef[0], self.prior_coef[1]
        if self.prior_name == 'ga':
            return np.sum(a * np.log(x) - b * x)
        return np.sum(-a * np.log(x) - b / x)

    def log_prior(self):
        if self.prior_name == 'ref':
            a, b = self.prior_coef[0], self.prior_coef[1]
            t = np.sum(self.cl / self.length) + self.nugget
            return a * np.log(t) - b * t
        lp = self._g(self.length)
        if self.nugget_est:
            lp += self._g(self.nugget)
        return lp

    def log_prior_fod(self):
        if self.prior_name == 'ref':
            a, b = self.prior_coef[0], self.prior_coef[1]
            t = np.sum(self.cl / self.length) + self.nugget
            fod = (b - a / t) * self.cl / self.length
            if self.nugget_est:
                fod = np.concatenate((fod, (a / t - b) * self.nugget))
            return fod
        fod = self.gfod(self.length)
        if self.nugget_est:
            fod = np.concatenate((fod, self.gfod(self.nugget)))
        return fod

    # ---- device descriptors ----------------------------------------------------------------------
    def _upload(self):
        """Upload this node's input (variable-major), global input and output; returns the tensors."""
        src = L.to_dev(np.ascontiguousarray(self.input.T))
        gsrc = L.to_dev(np.ascontiguousarray(self.global_input.T)) if self.global_input is not None else None
        out = L.to_dev(np.ascontiguousarray(self.output[:, 0])) if self.output is not None else None
        return src, gsrc, out

    def _node(self, bufs):
        src, gsrc, out = bufs
        node = L.DgpbNode()
        L.fill_node(node, kind=self.name, input_dim=np.arange(src.shape[0]),
                    connect=None if gsrc is None else np.arange(gsrc.shape[0]), length=self.length, scale=self.scale,
                    nugget=self.nugget, scale_est=self.scale_est, nugget_est=self.nugget_est, src=src, gsrc=gsrc,
                    output=out)
        return node

    # ---- 1. kernel matrix ------------------------------------------------------------------------
    def k_matrix(self, fod_eval=False):
        """Correlation matrix (and d/dlog-theta slices) -- kernel_class.py:304-359, built on the GPU."""
        self._check_supported()
        lib = L.load()
        X = L.to_dev(self._X())
        n, D = X.shape
        K = L.empty((n, n))
        P = len(self.length) + (1 if self.nugget_est else 0)
        dK = L.empty((P, n, n)) if fod_eval else None
        larr, lptr = L.length_host(self.length)
        L.check(lib.dgpb_kmatrix(L.ptr(X), n, D, lptr, len(larr), float(self.nugget[0]), None, L.KIND[self.name],
                                 int(bool(self.nugget_est)), L.ptr(K), L.ptr(dK), L.stream()))
        if fod_eval:
            return L.to_host(K), L.to_host(dK)
        return L.to_host(K)

    # ---- 2. dense likelihoods ----------------------------------------------------------------------
    def log_likelihood_func(self):
        """ESS log-likelihood -- kernel_class.py:481-492."""
        self._check_supported()
        bufs = self._dcache or self._upload()
        node = self._node(bufs)
        out = L.host_doubles(1)
        L.check(L.load().dgpb_loglik_dense(L.workspace(), ctypes.byref(node), self.input.shape[0], out, L.stream()))
        llik = out[0]
        if self.prior_name == 'ref':
            self.compute_cl()
            llik += self.log_prior()
        return llik

    def llik(self, x):
        """Negative log-likelihood and gradient wrt log-parameters -- kernel_class.py:403-449."""
        self._check_supported()
        self.update(x)
        bufs = self._dcache or self._upload()
        node = self._node(bufs)
        P = len(self.length) + (1 if self.nugget_est else 0)
        batcher = getattr(self, '_batcher', None)
        if batcher is not None:
            out = batcher.evaluate(node, self.input.shape[0], P, getattr(self, '_mid', 0))   # batched with the other nodes
        else:
            out = L.host_doubles(P + 2)
            L.check(L.load().dgpb_nllik_grad_dense(L.workspace(), ctypes.byref(node), self.input.shape[0], out,
                                                   L.stream()))
        neg_llik = np.array([out[0]])
        if self.scale_est:
            self.scale = np.array([out[1]])
        neg_St = np.array(out[2:2 + P])
        if self.prior_name is not None:
            neg_llik = neg_llik - self.log_prior()
            neg_St = neg_St - self.log_prior_fod()
        return neg_llik, neg_St

    def llik_vecch(self, x):
        """Vecchia negative log-likelihood and gradient -- kernel_class.py:451-479."""
        self._check_supported()
        self.update(x)
        from .vecchia import vecchia_nllik
        vc = getattr(self, '_vcache', None)
        if vc is not None and vc[0] is self.ord and vc[1] is self.NNarray:
            # ordered inputs / outputs / neighbour array already on the device (uploaded once per `maximise`)
            Xd, yd, NNd = vc[2]
            larr, lptr = L.length_host(self.length)
            P = len(larr) + (1 if self.nugget_est else 0)
            out = L.host_doubles(P + 2)
            L.check(L.load().dgpb_vecchia_nllik(L.ptr(Xd), L.ptr(yd), L.ptr(NNd), Xd.shape[0], Xd.shape[1],
                                                NNd.shape[1], lptr, len(larr), float(self.scale[0]),
                                                float(self.nugget[0]), None, L.KIND[self.name],
                                                int(bool(self.scale_est)), int(bool(self.nugget_est)), out,
                                                L.stream()))
            neg_llik, neg_St, scale = out[0], np.array(out[2:2 + P]), out[1]
        else:
            X = self._X()
            neg_llik, neg_St, scale = vecchia_nllik(X[self.ord], self.output[self.ord], self.NNarray, self.scale[0],
                                                    self.length, self.nugget[0], None, self.name, self.scale_est,
                                                    self.nugget_est)
        self.scale = np.array([scale])
        neg_llik = np.array([neg_llik])
        if self.prior_name is not None:
            neg_llik = neg_llik - self.log_prior()
            neg_St = neg_St - self.log_prior_fod()
        return neg_llik, neg_St

    def log_likelihood_func_vecch(self):
        """kernel_class.py:494-509."""
        self._check_supported()
        from .vecchia import vecchia_llik
        X = self._X()
        llik = vecchia_llik(X[self.ord], self.output[self.ord], self.NNarray, self.scale[0], self.length,
                            self.nugget[0], None, self.name)
        if self.prior_name == 'ref':
            self.compute_cl()
            llik += self.log_prior()
        return llik

    def callback(self, xk):
        self.iter_count += 1
        if self.iter_count & (self.iter_count - 1) == 0:
            self.ord_nn()

    def maximise(self, method='L-BFGS-B'):
        """M-step: L-BFGS-B over log-parameters with the reference's bounds and budgets
        (kernel_class.py:516-579).  Only (theta in, f and grad out: <= 34 doubles) crosses PCIe per
        evaluation; inputs/outputs are uploaded once for the whole optimisation."""
        x0 = self.log_t()
        npar = len(x0)
        fun = self.llik_vecch if self.vecch else self.llik
        budget = {'maxiter': 100, 'maxfun': np.max((30, 20 + 5 * self.D))}
        kwargs = {}
        reorder = self.vecch and self.target == 'gp' and len(self.length) != 1
        if reorder:  # plain Vecchia GP with ARD: re-order at power-of-two iterations (:537,:551,:560,:573)
            budget = {'maxfun': np.max((50, 20 + 5 * self.D))}
            kwargs['callback'] = self.callback
        nlen = npar - 1 if self.nugget_est else npar
        if self.bds is not None:
            with np.errstate(divide='ignore'):
                lb = np.log(self.bds[0]) * np.ones(nlen)
            ub = np.log(self.bds[1]) * np.ones(nlen)
        else:
            lb = -np.inf * np.ones(nlen)
            ub = (13. if self.prior_name == 'ref' else np.inf) * np.ones(nlen)
        if self.nugget_est:
            lb = np.concatenate((lb, np.log([1e-8])))
            ub = np.concatenate((ub, [np.inf]))
        bounded = self.nugget_est or self.bds is not None or self.prior_name == 'ref'
        if bounded:
            kwargs['bounds'] = Bounds(lb, ub)
        self._dcache = None if self.vecch else self._upload()
        self._vcache = None
        if self.vecch:
            # Vecchia: the ordered inputs, outputs and neighbour array go to the device once for the whole
            # optimisation (a re-ordering by the callback replaces `ord` / `NNarray`, which invalidates the cache)
            X = self._X()
            self._vcache = (self.ord, self.NNarray,
                            (L.to_dev(X[self.ord]), L.to_dev(np.ascontiguousarray(self.output[self.ord, 0])),
                             L.to_dev(self.NNarray, np.int64)))
        try:
            res = minimize(fun, x0, method=method, jac=True, options=budget, **kwargs)
            self._nfev = int(getattr(res, 'nfev', 0))   # next M-step's load balancing on several GPUs
        finally:
            self._dcache = None
            self._vcache = None
        if reorder:
            self.iter_count = 0
        self.add_to_path()

    def add_to_path(self):
        para = np.concatenate((self.scale, self.length, self.nugget))
        self.para_path = np.vstack((self.para_path, para))

    # ---- statistics for prediction -------------------------------------------------------------------
    def compute_stats(self):
        """R^-1 and R^-1 y (kernel_class.py:735-751).  Results stay on the device (`Rinv`/`Rinv_y` give numpy
        copies on demand).  The sexp tables R2sexp/Psexp (:752-764) are never built: the linked-GP kernel
        recomputes those terms on the fly."""
        self._check_supported()
        n = self.input.shape[0]
        bufs = self._upload()
        node = self._node(bufs)
        Rinv, Rinv_y = L.empty((n, n)), L.empty((n,))
        try:
            L.check(L.load().dgpb_compute_stats(L.workspace(), ctypes.byref(node), n, L.ptr(Rinv), L.ptr(Rinv_y),
                                                L.stream()))
            self._Rinv, self._Rinv_y = Rinv, Rinv_y
        except LinAlgError:
            # the reference's recovery branch: pseudo-inverse on the host (kernel_class.py:749-751)
            from scipy.linalg import pinvh
            R = self.k_matrix()
            Ri = pinvh(R, check_finite=False)
            self._Rinv, self._Rinv_y = L.to_dev(Ri), L.to_dev(np.dot(Ri, self.output).flatten())

    def _stats_dev(self):
        if self._Rinv is None:
            raise RuntimeError("compute_stats() has not been called on this GP node")
        if isinstance(self._Rinv, np.ndarray):
            self._Rinv, self._Rinv_y = L.to_dev(self._Rinv), L.to_dev(np.asarray(self._Rinv_y).flatten())
        return self._Rinv, self._Rinv_y

    # ---- 5. predictions --------------------------------------------------------------------------------
    def _nn_query(self, xq, w):
        """Neighbours of the query points among the training inputs, on coordinates divided by the
        length-scales (vecchia.py:20-40 called from kernel_class.py:640-668).  With ONE shared length-scale the
        division cannot change which points are nearest, so within a predict pass nodes that are asked about the
        same query tensor and the same training inputs with the same m reuse one search (first layer of a DGP:
        every node of every imputation).  Only an exact distance tie could be broken differently by another
        node's scaling, and the reference's own tie-break is library-defined."""
        from .vecchia import get_pred_nn_dev
        larr = np.atleast_1d(self.length)
        iso = len(larr) == 1
        pc = L.active_cache()
        if iso and pc is not None:
            for (xq0, w0, m0, NN0) in pc.nn:
                if xq0 is xq and w0 is w and m0 == self.pred_m:
                    return NN0[:, 1:].contiguous() if self.loo_state else NN0
        lt = L.to_dev(np.full(w.shape[1], larr[0]) if iso else larr)
        NN = get_pred_nn_dev(xq / lt, w / lt, self.pred_m)
        if iso and pc is not None:
            pc.nn.append((xq, w, self.pred_m, NN))
        return NN[:, 1:].contiguous() if self.loo_state else NN

    def _gp_prediction_dev(self, x, z):
        """Device tensors in / out; see `gp_prediction`."""
        lib = L.load()
        xq = L.cat_cols(x, z)
        W = L.to_dev_shared(self._X_pred())
        M, D = xq.shape
        mean, var = L.empty((M,)), L.empty((M,))
        larr, lptr = L.length_host(self.length)
        if self.vecch:
            NN = self._nn_query(xq, W)
            y = L.to_dev_shared(self._y_pred())
            L.check(lib.dgpb_gp_vecch(L.ptr(xq), M, L.ptr(W), L.ptr(y), W.shape[0], D, L.ptr(NN), NN.shape[1], lptr,
                                      len(larr), float(self.scale[0]), float(self.nugget[0]), None, L.KIND[self.name],
                                      L.ptr(mean), L.ptr(var), L.stream()))
        else:
            Rinv, Rinv_y = self._stats_dev()
            L.check(lib.dgpb_gp_predict(L.workspace(), L.ptr(xq), M, L.ptr(W), W.shape[0], D, L.ptr(Rinv),
                                        L.ptr(Rinv_y), lptr, len(larr), float(self.scale[0]), float(self.nugget[0]),
                                        L.KIND[self.name], L.ptr(mean), L.ptr(var), L.stream()))
        return mean, var

    def _linkgp_dev(self, m, v, z, w1, gw, loo=True):
        """Shared body of linkgp_prediction / linkgp_prediction_full on device tensors.  `loo=False`: keep the
        nearest neighbour even in LOO state (kernel_class.py:672-733: `linkgp_prediction_full` has no LOO slice)."""
        lib = L.load()
        torch = L.torch_mod()
        m, v = m.contiguous(), v.contiguous()
        z = None if z is None else z.contiguous()
        M, Dw = m.shape
        Dz = 0 if z is None else z.shape[1]
        mean, var = L.empty((M,)), L.empty((M,))
        larr, lptr = L.length_host(self.length)
        if self.vecch:
            xq = L.cat_cols(m, z)
            w = L.cat_cols(w1, gw)
            state, self.loo_state = self.loo_state, self.loo_state and loo
            try:
                NN = self._nn_query(xq, w)
            finally:
                self.loo_state = state
            y = L.to_dev_shared(self._y_pred())
            L.check(lib.dgpb_linkgp_vecch(L.ptr(m), L.ptr(v), L.ptr(z), M, L.ptr(w1), L.ptr(gw), L.ptr(y), w1.shape[0],
                                          Dw, Dz, L.ptr(NN), NN.shape[1], lptr, len(larr), float(self.scale[0]),
                                          float(self.nugget[0]), None, L.KIND[self.name], L.ptr(mean), L.ptr(var),
                                          L.stream()))
        else:
            Rinv, Rinv_y = self._stats_dev()
            L.check(lib.dgpb_linkgp_predict(L.workspace(), L.ptr(m), L.ptr(v), L.ptr(z), M, L.ptr(w1), L.ptr(gw),
                                            w1.shape[0], Dw, Dz, L.ptr(Rinv), L.ptr(Rinv_y), lptr, len(larr),
                                            float(self.scale[0]), float(self.nugget[0]), L.KIND[self.name],
                                            L.ptr(mean), L.ptr(var), L.stream()))
        return mean, var

    def _linkgp_prediction_dev(self, m, v, z):
        w1 = L.to_dev_shared(self.input)
        gw = L.to_dev_shared(self.global_input) if z is not None else None
        return self._linkgp_dev(m, v, z, w1, gw)

    def _linkgp_prediction_full_dev(self, m, v, m_z, v_z, z):
        torch = L.torch_mod()
        k1 = m_z.shape[1]
        m = torch.cat((m, m_z), 1)
        v = torch.cat((v, v_z), 1)
        w1 = L.to_dev(np.concatenate((self.input, self.global_input[:, :k1]), axis=1))
        gw = L.to_dev(self.global_input[:, k1:]) if z is not None else None
        return self._linkgp_dev(m, v, z, w1, gw, loo=False)

    def gp_prediction(self, x, z):
        """GP predictive mean/variance at deterministic inputs (kernel_class.py:587-625)."""
        m, v = self._gp_prediction_dev(L.to_dev(x), None if z is None else L.to_dev(z))
        return L.to_host(m), L.to_host(v)

    def linkgp_prediction(self, m, v, z):
        """Linked-GP moments for Gaussian inputs N(m, diag v) (kernel_class.py:627-670)."""
        mo, vo = self._linkgp_prediction_dev(L.to_dev(m), L.to_dev(v), None if z is None else L.to_dev(z))
        return L.to_host(mo), L.to_host(vo)

    def linkgp_prediction_full(self, m, v, m_z, v_z, z):
        """As `linkgp_prediction` with some connected inputs also Gaussian (kernel_class.py:672-733)."""
        mo, vo = self._linkgp_prediction_full_dev(L.to_dev(m), L.to_dev(v), L.to_dev(m_z), L.to_dev(v_z),
                                                  None if z is None else L.to_dev(z))
        return L.to_host(mo), L.to_host(vo)


def _as_numpy(v):
    if v is None or isinstance(v, np.ndarray):
        return v
    return L.to_host(v)


def combine(*layers):
    """Combine layers into one list (kernel_class.py:766-779)."""
    return [layer for layer in layers]
