"""`emulator` -- predictions from a trained DGP (dgpsi/emulation.py:14-44, 631-854) on the GPU.

`__init__` draws N imputations (ESS on the device) and computes each node's R^-1 / R^-1y with the
sliding-window factorisation; `predict(x, method='mean_var')` pushes the test inputs through the layers on
the device -- first layer `gp`, deeper layers `link_gp` -- keeps the per-imputation moments in HBM and
aggregates them with one kernel.  With a process group (one process per GPU) the test points are sharded
and the moments all-gathered over NCCL (`dgp_b200.parallel`).
`loo` re-uses the Vecchia prediction kernels with every point conditioned on the others;
`metric` (ALM / MICE / VIGF, same row) is host arithmetic on the per-imputation moments plus one dense inverse per
output node (MICE) and a nearest-training-point search (VIGF).
A final layer of likelihood nodes maps the latent moments through each likelihood's closed form;
`nllik` integrates the likelihood against them by Gauss-Hermite quadrature.  Process pools (`ppredict`, `pmetric`,
`ploo`) are the same calls: the GPU replaces the pool.
"""
from __future__ import annotations

import copy

import numpy as np

from . import _lib as L
from .imputation import imputer


class emulator:
    """Class to make predictions from the trained DGP model (arguments: emulation.py:24)."""

    group_first_layer = True   # one multi-node launch for first-layer Vecchia nodes that share inputs (see below)

    def __init__(self, all_layer, N=10, block=True):
        self.all_layer = all_layer
        self.n_layer = len(all_layer)
        self.vecch = bool(self.all_layer[0][0].vecch)
        self.imp = imputer(self.all_layer, block)
        if self.vecch:
            (self.imp).update_ord_nn()
            (self.imp).sample(burnin=20)
        else:
            (self.imp).sample(burnin=50)
        self.all_layer_set = []
        for _ in range(N):
            if self.vecch:
                (self.imp).update_ord_nn()
            (self.imp).sample()
            if not self.vecch:
                (self.imp).key_stats()
            (self.all_layer_set).append(copy.deepcopy(self.all_layer))

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_pool', None)   # device copies are rebuilt lazily
        return state

    def __setstate__(self, state):
        state.pop('all_layer_set_copy', None)
        state.pop('nb_parallel', None)
        state.setdefault('vecch', False)
        self.__dict__.update(state)

    def _frozen_pool(self):
        """The imputations of an emulator never change after construction: their training inputs and outputs are
        uploaded once (shared between nodes / imputations that hold identical arrays) and kept for every later
        predict / loo / metric call."""
        if getattr(self, '_pool', None) is None:
            self._pool = L.UploadPool()
        for one in self.all_layer_set:
            for layer in one:
                for kernel in layer:
                    if kernel.type == 'gp':
                        kernel._frozen = True
        return self._pool

    def to_vecchia(self):
        """emulation.py:62-74."""
        if self.vecch:
            raise Exception('The DGP emulator is already in Vecchia mode.')
        self.vecch = True
        for one in self.all_layer_set:
            for layer in one:
                for kernel in layer:
                    if kernel.type != 'gp':
                        continue
                    kernel.vecch = True
                    if kernel.ord is None:
                        kernel.m = min(25, kernel.input.shape[0] - 1) if kernel.m is None else kernel.m

    def remove_vecchia(self):
        """emulation.py:76-89."""
        if not self.vecch:
            raise Exception('The DGP emulator is already in non-Vecchia mode.')
        self.vecch = False
        for one in self.all_layer_set:
            for layer in one:
                for kernel in layer:
                    if kernel.type != 'gp':
                        continue
                    kernel.vecch = False
                    kernel.compute_stats()

    # ---- leave-one-out ----------------------------------------------------------------------------------
    def change_vecch_state(self):
        """Context of emulation.py:91-107: every GP node predicts in Vecchia form with itself removed from the
        conditioning set."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            for one in self.all_layer_set:
                for layer in one:
                    for kernel in layer:
                        if kernel.type != 'gp':
                            continue
                        if not self.vecch:
                            kernel.vecch = True
                        kernel.loo_state = True
            try:
                yield
            finally:
                for one in self.all_layer_set:
                    for layer in one:
                        for kernel in layer:
                            if kernel.type != 'gp':
                                continue
                            if not self.vecch:
                                kernel.vecch = False
                            kernel.loo_state = False

        return ctx()

    def loo(self, X, method=None, sample_size=50, m=30):
        """Leave-one-out cross-validation of the DGP emulator (emulation.py:109-144): predictions at the training
        inputs with each GP node conditioned on the m nearest OTHER training points (Vecchia emulators) or on all
        other points (dense emulators).  The conditioning blocks run on the Vecchia kernels, which hold at most 63
        conditioning points: dense emulators with more than 63 training points are not supported here."""
        if method is None:
            method = 'mean_var'
        n = len(self.all_layer[0][0].input)
        if len(X) != n:
            raise NotImplementedError("dgp_b200: replicated training inputs are outside the SI hot path")
        m_pred = m + 1 if self.vecch else X.shape[0]
        if m_pred - 1 > 63:
            raise NotImplementedError("dgp_b200: leave-one-out of a dense emulator conditions every point on all "
                                      "others; the block kernels hold at most 63 conditioning points "
                                      "(convert with to_vecchia() for larger designs)")
        with self.change_vecch_state():
            return self.predict(X, method=method, sample_size=sample_size, m=m_pred)

    def ploo(self, X, method=None, sample_size=50, m=30, core_num=None):
        """emulation.py:146-168: the process pool is replaced by the GPU."""
        return self.loo(X, method, sample_size, m)

    # ---- sequential-design criteria -----------------------------------------------------------------------
    def _layer_moments(self, x, m):
        """Per imputation, the (mean, var) device tensors of every layer at the inputs x."""
        xd = L.to_dev(x, np.float64)
        with L.predict_cache(self._frozen_pool()):
            return [self._predict_one_imputation(layers, xd, m, True)[2] for layers in self.all_layer_set]

    @staticmethod
    def _mice_var(x, x_extra, kern, nugget_s):
        """Smoothed predictive variance scale / diag(R^-1) of one output node on the candidate set
        (functions.py:244-256); R^-1 comes from the device factorisation."""
        from .kernel_class import kernel as ker
        kin = x[:, kern.input_dim]
        if kern.connect is not None:
            kin = np.concatenate((kin, x_extra[:, kern.connect]), 1)
        node = ker(length=np.array(kern.length, dtype=np.float64), scale=float(kern.scale[0]),
                   nugget=max(float(nugget_s), float(kern.nugget[0])), name=kern.name)
        node.input, node.global_input = np.ascontiguousarray(kin), None
        node.input_dim = np.arange(kin.shape[1])
        node.output = np.zeros((len(kin), 1))
        node.compute_stats()
        Rinv, _ = node._stats_dev()
        return float(kern.scale[0]) / L.to_host(L.torch_mod().diagonal(Rinv).contiguous())

    def metric(self, x_cand, method='ALM', obj=None, nugget_s=1., m=50, score_only=False):
        """ALM, MICE or VIGF criterion of sequential design at the candidate points (emulation.py:323-420).
        Returns the (M x D_out) scores when `score_only`, else (argmax rows, their scores) per output."""
        if x_cand.ndim == 1:
            raise Exception('The candidate design set has to be a numpy 2d-array.')
        lik = self.all_layer[-1][0].type == 'likelihood'
        last = self.n_layer - 2 if lik else self.n_layer - 1      # the last GP layer (emulation.py:344, 447, 533)
        if method == 'ALM':
            if lik:
                _, sigma2 = self.predict(x=x_cand, full_layer=True, m=m)
                score = sigma2[-2]
            else:
                _, score = self.predict(x=x_cand, m=m)
        elif method == 'MICE':
            if lik and self.n_layer == 2:   # emulation.py:362-372: one GP layer under the likelihood, no imputation
                xd = L.to_dev(x_cand, np.float64)
                with L.predict_cache():
                    _, var, _ = self._predict_one_imputation(self.all_layer[:1], xd, m, False)
                smooth = np.stack([self._mice_var(x_cand, x_cand, kern, nugget_s) for kern in self.all_layer[0]], 1)
                score = L.to_host(var) / smooth
            else:
                if last < 1:
                    raise Exception('The MICE criterion needs a DGP with at least two GP layers.')
                score = 0.
                for s, moments in enumerate(self._layer_moments(x_cand, m)):
                    pred_in, sigma2 = L.to_host(moments[last - 1][0]), L.to_host(moments[last][1])
                    smooth = np.stack([self._mice_var(pred_in, x_cand, kern, nugget_s)
                                       for kern in self.all_layer_set[s][last]], 1)
                    with np.errstate(divide='ignore'):
                        score = score + np.log(sigma2 / smooth)
                score = score / len(self.all_layer_set)
        elif method == 'VIGF':
            if obj is None:
                raise Exception('The dgp object that is used to build the emulator must be supplied to the argument '
                                '`obj` when VIGF criterion is chosen.')
            from .vecchia import get_pred_nn_dev
            # nearest training input of every candidate (emulation.py:396-401), searched on the device
            index = L.to_host(get_pred_nn_dev(L.to_dev(x_cand, np.float64), L.to_dev(obj.X, np.float64), 1)).ravel()
            bias, sigma2 = [], []
            for s, moments in enumerate(self._layer_moments(x_cand, m)):
                target = np.stack([kern.output[index, 0] for kern in self.all_layer_set[s][last]], 1)
                bias.append((L.to_host(moments[last][0]) - target) ** 2)
                sigma2.append(L.to_host(moments[last][1]))
            bias, sigma2 = np.asarray(bias), np.asarray(sigma2)
            E1 = np.mean(np.square(bias) + 6 * bias * sigma2 + 3 * np.square(sigma2), axis=0)
            E2 = np.mean(bias + sigma2, axis=0)
            score = E1 - E2 ** 2
        else:
            raise Exception("method must be 'ALM', 'MICE' or 'VIGF'")
        if score_only:
            return score
        idx = np.argmax(score, axis=0)
        return idx, score[idx, np.arange(score.shape[1])]

    def pmetric(self, x_cand, method='ALM', obj=None, nugget_s=1., m=50, score_only=False, chunk_num=None,
                core_num=None):
        """emulation.py:170-321: the process pool is replaced by the GPU."""
        return self.metric(x_cand, method, obj, nugget_s, m, score_only)

    # ---- prediction -------------------------------------------------------------------------------------
    def _predict_one_imputation(self, layers, xd, m, collect_layers):
        """Propagate device test inputs `xd` (M x d) through one imputed hierarchy; returns device tensors."""
        torch = L.torch_mod()
        mean = var = None
        per_layer = []
        for l, layer in enumerate(layers):
            ms, vs = [None] * len(layer), [None] * len(layer)
            if l == 0 and self.group_first_layer:
                for k, (mk, vk) in self._first_layer_vecchia_groups(layer, xd, m).items():
                    ms[k], vs[k] = mk, vk
            for k, kernel in enumerate(layer):
                if ms[k] is not None:
                    continue
                if kernel.type == 'likelihood':   # emulation.py:752-757: moments of the observable from the latent ones
                    if kernel.name == 'Categorical':   # :753-754 the latent moments are aggregated first, see predict
                        ms, vs = list(L.cols(mean, kernel.input_dim).T), list(L.cols(var, kernel.input_dim).T)
                        break
                    mk, vk = kernel.prediction(m=L.to_host(L.cols(mean, kernel.input_dim)),
                                               v=L.to_host(L.cols(var, kernel.input_dim)))
                    ms[k], vs[k] = L.to_dev(np.ascontiguousarray(mk)), L.to_dev(np.ascontiguousarray(vk))
                    continue
                kernel.pred_m = m
                z = L.cols(xd, kernel.connect) if kernel.connect is not None else None
                if l == 0:
                    mk, vk = kernel._gp_prediction_dev(L.cols(xd, kernel.input_dim), z)
                else:
                    mk, vk = kernel._linkgp_prediction_dev(L.cols(mean, kernel.input_dim), L.cols(var, kernel.input_dim), z)
                ms[k], vs[k] = mk, vk
            mean, var = torch.stack(ms, 1), torch.stack(vs, 1)
            if collect_layers:
                per_layer.append((mean, var))
        return mean, var, per_layer

    @staticmethod
    def _first_layer_vecchia_groups(layer, xd, m):
        """First-layer Vecchia nodes with a squared-exponential kernel, ONE length-scale and the same inputs share
        the test points, the training inputs and (see `kernel._nn_query`) the neighbour sets: they are predicted
        by one `dgpb_gp_vecch_multi` launch that forms each block's squared distances once.  Returns
        {node index: (mean, var)} for the nodes it handled (blocks of at most 32 points)."""
        import ctypes
        done = {}
        groups = {}
        for k, kern in enumerate(layer):
            if not (kern.vecch and kern.name == 'sexp' and len(np.atleast_1d(kern.length)) == 1 and kern.rep is None):
                continue
            key = (tuple(np.atleast_1d(kern.input_dim)), None if kern.connect is None else tuple(kern.connect))
            groups.setdefault(key, []).append(k)
        lib = L.load()
        torch = L.torch_mod()
        for key, idx in groups.items():
            if len(idx) < 2:
                continue
            first = layer[idx[0]]
            X0 = first._X_pred()
            if min(int(m), X0.shape[0]) - (1 if first.loo_state else 0) + 1 > 32:
                continue
            W = L.to_dev_shared(X0)
            # identical inputs share one device tensor in the upload pool: an identity test after the first call
            if any(L.to_dev_shared(layer[k]._X_pred()) is not W for k in idx[1:]):
                continue
            first.pred_m = m
            z = L.cols(xd, first.connect) if first.connect is not None else None
            xq = L.cat_cols(L.cols(xd, first.input_dim), z)
            NN = first._nn_query(xq, W)
            B, M = len(idx), xq.shape[0]
            Y = L.to_dev(np.ascontiguousarray(np.stack([layer[k].output[:, 0] for k in idx], 0)))
            par = np.ascontiguousarray([[float(np.atleast_1d(layer[k].length)[0]) for k in idx],
                                        [float(layer[k].scale[0]) for k in idx],
                                        [float(layer[k].nugget[0]) for k in idx]], dtype=np.float64)
            mean, var = L.empty((B, M)), L.empty((B, M))
            L.check(lib.dgpb_gp_vecch_multi(L.ptr(xq), M, L.ptr(W), L.ptr(Y), W.shape[0], W.shape[1], L.ptr(NN),
                                            NN.shape[1], B, par[0].ctypes.data_as(L.c_vp),
                                            par[1].ctypes.data_as(L.c_vp), par[2].ctypes.data_as(L.c_vp),
                                            L.ptr(mean), L.ptr(var), L.stream()))
            for b, k in enumerate(idx):
                done[k] = (mean[b], var[b])
        return done

    def predict(self, x, method='mean_var', full_layer=False, sample_size=50, m=50, aggregation=True, _device=False):
        """Predictions from the trained DGP (emulation.py:631-854).  `x`: (M x d) numpy array.
        method='mean_var' returns (mu, sigma2) exactly as the reference; method='sampling' draws
        `sample_size` normal samples per imputation from the final-layer moments."""
        if x.ndim == 1:
            raise Exception('The testing input has to be a numpy 2d-array')
        lib = L.load()
        torch = L.torch_mod()
        xd = L.to_dev(x, np.float64)
        S = len(self.all_layer_set)
        means, variances, layers_all = [], [], []
        lik_out = self.all_layer[-1][0].type == 'likelihood'
        if method == 'sampling' and lik_out:
            return self._sample_likelihood(xd, full_layer, sample_size, m)
        with L.predict_cache(self._frozen_pool()):
            for s in range(S):
                mean, var, per_layer = self._predict_one_imputation(self.all_layer_set[s], xd, m, full_layer)
                means.append(mean)
                variances.append(var)
                layers_all.append(per_layer)
        if method == 'sampling':
            # numpy's global generator in the reference's order (emulation.py:790-830): inner layers one (M x width)
            # block per (imputation, replicate), the final GP layer column by column
            def draw(mu, va, by_column):
                mu, va = L.to_host(mu), L.to_host(va)                     # S x M x width
                mu, va = np.repeat(mu, sample_size, 0), np.repeat(va, sample_size, 0)
                if by_column:
                    return np.random.normal(mu.transpose(0, 2, 1), np.sqrt(va.transpose(0, 2, 1))).transpose(1, 2, 0)
                return np.random.normal(mu, np.sqrt(va)).transpose(2, 1, 0)

            if full_layer:
                out = []
                for l in range(self.n_layer):
                    mu_l = torch.stack([layers_all[s][l][0] for s in range(S)], 0)
                    va_l = torch.stack([layers_all[s][l][1] for s in range(S)], 0)
                    out.append(list(draw(mu_l, va_l, l == self.n_layer - 1)))
                return out
            return list(draw(torch.stack(means, 0), torch.stack(variances, 0), True))
        if method != 'mean_var':
            raise Exception("method must be 'mean_var' or 'sampling'")

        def agg(ms, vs):
            ms, vs = torch.stack(ms, 0).contiguous(), torch.stack(vs, 0).contiguous()
            mu, s2 = torch.empty_like(ms[0]), torch.empty_like(ms[0])
            L.check(lib.dgpb_aggregate(L.ptr(ms), L.ptr(vs), ms.shape[0], ms[0].numel(), L.ptr(mu), L.ptr(s2),
                                       L.stream()))
            return L.to_host(mu), L.to_host(s2)

        # Categorical likelihood: class probabilities come from the AGGREGATED latent moments (emulation.py:831-834,
        # 841-844); without aggregation from each imputation's moments (:849-850)
        cat = self.all_layer[-1][0] if self.all_layer[-1][0].name == 'Categorical' else None
        if full_layer:
            mu, sigma2 = [], []
            for l in range(self.n_layer):
                a, b = agg([layers_all[s][l][0] for s in range(S)], [layers_all[s][l][1] for s in range(S)])
                if cat is not None and l == self.n_layer - 1:
                    a, b = cat.prediction(m=a, v=b)
                mu.append(a)
                sigma2.append(b)
            return mu, sigma2
        if aggregation and _device and cat is None:
            # moments stay on the device (dgp_b200.parallel gathers them over NCCL without a host bounce)
            ms, vs = torch.stack(means, 0).contiguous(), torch.stack(variances, 0).contiguous()
            mu, s2 = torch.empty_like(ms[0]), torch.empty_like(ms[0])
            L.check(lib.dgpb_aggregate(L.ptr(ms), L.ptr(vs), ms.shape[0], ms[0].numel(), L.ptr(mu), L.ptr(s2),
                                       L.stream()))
            return mu, s2
        if aggregation:
            mu, sigma2 = agg(means, variances)
            return cat.prediction(mu, sigma2) if cat is not None else (mu, sigma2)
        if cat is not None:
            pairs = [cat.prediction(L.to_host(a), L.to_host(b)) for a, b in zip(means, variances)]
            return [p[0] for p in pairs], [p[1] for p in pairs]
        return [L.to_host(t) for t in means], [L.to_host(t) for t in variances]

    def _sample_likelihood(self, xd, full_layer, sample_size, m):
        """method='sampling' of an emulator whose final layer holds likelihood nodes (emulation.py:765-810): draws
        of the latent layers from their Gaussian moments, pushed through each likelihood's sampler."""
        with L.predict_cache(self._frozen_pool()):
            moments = [self._predict_one_imputation(layers, xd, m, True)[2] for layers in self.all_layer_set]
        host = [[(L.to_host(mu), L.to_host(va)) for mu, va in per_layer] for per_layer in moments]
        last = self.all_layer[-1]

        def observe(latent):
            if last[0].name == 'Categorical':   # class probabilities, one column per class (emulation.py:794-795)
                return last[0].sampling(latent[:, last[0].input_dim])
            out = np.empty((latent.shape[0], len(last)))
            for count, kernel in enumerate(last):
                out[:, count] = kernel.sampling(latent[:, kernel.input_dim])
            return out

        if not full_layer:
            samples = []
            for per_layer in host:
                mu, va = per_layer[-2]
                for _ in range(sample_size):
                    samples.append(observe(np.random.normal(mu, np.sqrt(va))))
            return list(np.asarray(samples).transpose(2, 1, 0))
        out = []
        before = None
        for l in range(self.n_layer):
            if l == self.n_layer - 1:
                draws = [observe(latent) for latent in before]
            else:
                draws = [np.random.normal(per_layer[l][0], np.sqrt(per_layer[l][1]))
                         for per_layer in host for _ in range(sample_size)]
                if l == self.n_layer - 2:
                    before = draws
            out.append(list(np.asarray(draws).transpose(2, 1, 0)))
        return out

    @staticmethod
    def _expected_likelihood(pllik, mu, var, y, order=10):
        """E[p(y | f)] for f ~ N(mu, diag(var)) per test point by a tensor-product Gauss-Hermite rule
        (`ghdiag`, functions.py:233-241): mu, var (M x Q), y (M x 1) -> (M x 1)."""
        nodes, weights = np.polynomial.hermite.hermgauss(order)
        Q = mu.shape[1]
        grid = np.stack([a.ravel() for a in np.meshgrid(*([nodes] * Q), indexing='ij')], 1)        # (order^Q, Q)
        logw = np.sum(np.stack([a.ravel() for a in np.meshgrid(*([np.log(weights)] * Q), indexing='ij')], 1), 1)
        f = np.sqrt(2.0) * np.sqrt(var)[:, None, :] * grid[None, :, :] + mu[:, None, :]            # (M, order^Q, Q)
        ll = pllik(y[:, None], f)                                                                  # (M, order^Q, 1)
        return np.sum(np.exp((logw - 0.5 * Q * np.log(np.pi))[None, :, None] + ll), axis=1)

    def nllik(self, x, y, m=50):
        """Negative predicted log-likelihood of test outputs under a DGP with ONE likelihood node on top
        (emulation.py:855-911): returns (average, per test point)."""
        if len(self.all_layer[-1]) != 1 or self.all_layer[-1][0].type != 'likelihood':
            raise Exception('The method is only applicable to a DGP with the final layer formed by only ONE node, '
                            'which must be a likelihood node.')
        # repeated test inputs are predicted once (emulation.py:872-874).  Without repeats the reference still
        # indexes the predictions by np.unique's inverse map (:909), which scores y against other points'
        # predictions unless x is already sorted; here every output is scored against its own input.
        X0, indices = np.unique(x, return_inverse=True, axis=0)
        if len(X0) != len(x):
            x = X0
        else:
            indices = np.arange(len(x))
        lik = self.all_layer[-1][0]
        predicted = []
        for moments in self._layer_moments(x, m):
            mu, var = L.to_host(moments[-2][0]), L.to_host(moments[-2][1])
            predicted.append(self._expected_likelihood(lik.pllik, mu[indices][:, lik.input_dim],
                                                       var[indices][:, lik.input_dim], y))
        nllik = -np.log(np.mean(predicted, axis=0)).flatten()
        return np.mean(nllik), nllik

    def ppredict(self, x, method='mean_var', full_layer=False, sample_size=50, m=50, chunk_num=None, core_num=None):
        """The reference splits test points over a process pool (emulation.py:578-629); one GPU replaces the
        pool, so this is `predict`.  Use dgp_b200.parallel.predict_sharded for several GPUs."""
        return self.predict(x, method, full_layer, sample_size, m, True)
