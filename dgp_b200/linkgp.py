"""`container` and `lgp` -- linked systems of (D)GP emulators (dgpsi/linkgp.py:12-608) on the GPU.

Every emulator of the system is evaluated on device tensors: first-layer emulators see the deterministic
global input (`gp`), later ones the Gaussian moments of their feeders (`link_gp`, and
`linkgp_prediction_full` when a DGP's internal node re-connects to Gaussian inputs).  Per-imputation
moments stay in HBM and are mixed by one aggregation kernel (linkgp.py:493-494).
"""
from __future__ import annotations

import copy

import numpy as np

from . import _lib as L
from .imputation import imputer


class container:
    """Trained GP or DGP emulator of one computer model (arguments: linkgp.py:38)."""

    def __init__(self, structure, local_input_idx=None, block=True):
        if len(structure) == 1:
            self.type = 'gp'
            self.structure = structure[0]
            self.vecch = bool(self.structure.vecch)
        else:
            self.type = 'dgp'
            self.structure = structure
            self.vecch = bool(self.structure[0][0].vecch)
            self.imp = imputer(self.structure, block)
            if self.vecch:
                (self.imp).update_ord_nn()
            self.imp.sample(burnin=50)
        self.local_input_idx = local_input_idx

    def __setstate__(self, state):
        state.setdefault('vecch', False)
        self.__dict__.update(state)

    def _kernels(self):
        if self.type == 'gp':
            return [self.structure]
        return [k for layer in self.structure for k in layer if k.type == 'gp']

    def to_vecchia(self):
        """linkgp.py:64-76."""
        if not self.vecch:
            self.vecch = True
            for k in self._kernels():
                k.vecch = True
                if k.m is None:
                    k.m = min(25, k.input.shape[0] - 1)
                if k.ord is None:
                    k.ord_nn()

    def remove_vecchia(self):
        """linkgp.py:78-90."""
        if self.vecch:
            self.vecch = False
            for k in self._kernels():
                k.vecch = False
            if self.type == 'gp':
                self.structure.compute_stats()

    def set_local_input(self, idx, new=False):
        """linkgp.py:92-116."""
        if new:
            cp = copy.copy(self)
            cp.local_input_idx = idx
            return cp
        self.local_input_idx = idx

    def __copy__(self):
        new = type(self).__new__(self.__class__)
        new.type, new.structure, new.vecch = self.type, self.structure, self.vecch
        if self.type == 'dgp':
            new.imp = self.imp
        new.local_input_idx = copy.copy(self.local_input_idx)
        return new


class lgp:
    """System of GP and DGP emulators (arguments: linkgp.py:140)."""

    def __init__(self, all_layer, N=10):
        self.L = len(all_layer)
        self.all_layer = all_layer
        self.num_model = [len(all_layer[l]) for l in range(1, self.L)]
        if not any(cont.type == 'dgp' for layer in all_layer for cont in layer):
            N = 1
        self.all_layer_set = []
        work = copy.deepcopy(self.all_layer)
        for _ in range(N):
            one = []
            for layer in work:
                row = []
                for cont in layer:
                    if cont.type == 'dgp':
                        if cont.vecch:
                            (cont.imp).update_ord_nn()
                        (cont.imp).sample()
                        if not cont.vecch:
                            (cont.imp).key_stats()
                    row.append(copy.deepcopy(cont))
                one.append(row)
            self.all_layer_set.append(one)

    def __setstate__(self, state):
        state.pop('nb_parallel', None)
        self.__dict__.update(state)

    def set_vecchia(self, mode):
        """linkgp.py:180-212."""
        if not isinstance(mode, list):
            mode = [[mode for _ in layer] for layer in self.all_layer]
        elif [len(r) for r in mode] != [len(r) for r in self.all_layer]:
            raise Exception('mode has a different shape as all_layer.')
        for system in [self.all_layer] + self.all_layer_set:
            for layer, mode_layer in zip(system, mode):
                for cont, on in zip(layer, mode_layer):
                    if on:
                        cont.to_vecchia()
                    else:
                        cont.remove_vecchia()
                        if cont.type == 'dgp' and system is not self.all_layer:
                            (cont.imp).key_stats()

    # ---- device-level emulator evaluation -----------------------------------------------------------
    @staticmethod
    def _gp_pred_dev(x, m, v, z, structure, m_pred):
        """linkgp.py:503-515."""
        structure.pred_m = m_pred
        if x is None:
            mo, vo = structure._linkgp_prediction_dev(m, v, z)
        else:
            mo, vo = structure._gp_prediction_dev(x, z)
        return mo.reshape(-1, 1), vo.reshape(-1, 1)

    @staticmethod
    def _dgp_pred_dev(x, m, v, z, structure, pred_m):
        """Moments of a DGP emulator whose input is deterministic (`x`) or Gaussian (`m`, `v`, plus optional
        deterministic external input `z`) -- linkgp.py:517-608."""
        torch = L.torch_mod()
        nl = len(structure)
        internal_idx = structure[0][0].input_dim
        external_idx = structure[0][0].connect
        mean = var = None
        for l, layer in enumerate(structure):
            ms, vs = [], []
            for kernel in layer:
                if kernel.type == 'likelihood':   # linkgp.py:569-571: a DGP + likelihood emulator inside the system
                    mk, vk = kernel.prediction(m=L.to_host(L.cols(mean, kernel.input_dim)),
                                               v=L.to_host(L.cols(var, kernel.input_dim)))
                    ms.append(L.to_dev(np.ascontiguousarray(np.asarray(mk).reshape(-1))))
                    vs.append(L.to_dev(np.ascontiguousarray(np.asarray(vk).reshape(-1))))
                    continue
                kernel.pred_m = pred_m
                if l == 0:
                    if x is None:
                        mk, vk = kernel._linkgp_prediction_dev(m, v, z)
                    else:
                        mk, vk = kernel._gp_prediction_dev(x, z)
                else:
                    mi, vi = L.cols(mean, kernel.input_dim), L.cols(var, kernel.input_dim)
                    if kernel.connect is None:
                        mk, vk = kernel._linkgp_prediction_dev(mi, vi, None)
                    elif x is not None:
                        mk, vk = kernel._linkgp_prediction_dev(mi, vi, L.cols(x, kernel.connect))
                    else:
                        # connected inputs that are themselves Gaussian (idx1) vs deterministic external (idx2)
                        if l == nl - 1:
                            idx1 = np.where(kernel.connect[:, None] == np.atleast_1d(internal_idx)[None, :])[1]
                            idx2 = (np.where(kernel.connect[:, None] == np.atleast_1d(external_idx)[None, :])[1]
                                    if external_idx is not None else np.array([], dtype=int))
                        else:
                            Dm = m.shape[1]
                            idx1 = kernel.connect[kernel.connect <= (Dm - 1)]
                            idx2 = kernel.connect[kernel.connect > (Dm - 1)] - Dm
                        if idx1.size == 0:
                            mk, vk = kernel._linkgp_prediction_dev(mi, vi, L.cols(z, idx2))
                        elif idx2.size == 0:
                            mk, vk = kernel._linkgp_prediction_full_dev(mi, vi, L.cols(m, idx1), L.cols(v, idx1), None)
                        else:
                            mk, vk = kernel._linkgp_prediction_full_dev(mi, vi, L.cols(m, idx1), L.cols(v, idx1),
                                                                        L.cols(z, idx2))
                ms.append(mk)
                vs.append(vk)
            mean, var = torch.stack(ms, 1), torch.stack(vs, 1)
        return mean, var

    def _eval(self, model, x, m, v, z, m_pred):
        if model.type == 'gp':
            return self._gp_pred_dev(x, m, v, z, model.structure, m_pred)
        return self._dgp_pred_dev(x, m, v, z, model.structure, m_pred)

    def predict(self, x, method='mean_var', full_layer=False, sample_size=50, m=50):
        """Predictions from the linked (D)GP model (linkgp.py:285-501)."""
        with L.predict_cache(self._frozen_pool()):
            return self._predict(x, method, full_layer, sample_size, m)

    def _frozen_pool(self):
        """The emulators of a linked system never change after construction: their data is uploaded once and kept
        for every later call (arrays with identical contents share one device tensor)."""
        if getattr(self, '_pool', None) is None:
            self._pool = L.UploadPool()
        for one in self.all_layer_set:
            for layer in one:
                for cont in layer:
                    for k in cont._kernels():
                        k._frozen = True
        return self._pool

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_pool', None)
        return state

    def _predict(self, x, method, full_layer, sample_size, m):
        torch = L.torch_mod()
        lib = L.load()
        if isinstance(x, list) and len(x) != self.L:
            raise Exception('When test input is given as a list, it must contain global inputs to the all layers '
                            '(even with no global inputs to internal layers). Set None as the global input to the '
                            'internal models if they have no global inputs.')
        if not isinstance(x, list):
            if x.ndim == 1:
                raise Exception('The testing input has to be a numpy 2d-array.')
            x = [x] + [[None] * num for num in self.num_model]
        if method not in ('mean_var', 'sampling'):
            raise Exception("method must be 'mean_var' or 'sampling'")
        x0 = L.to_dev(x[0], np.float64)
        xext = [None] + [[None if a is None else L.to_dev(a, np.float64) for a in x[l]] for l in range(1, self.L)]
        per_imp = []  # [imputation][layer][emulator] -> (mean, var) device tensors
        for one in self.all_layer_set:
            outs, feeds_m, feeds_v = [], [], []
            for l, layer in enumerate(one):
                res = []
                for k, model in enumerate(layer):
                    if l == 0:
                        if isinstance(model.local_input_idx, list):
                            raise Exception('When an emulator is in the first layer, local_input_idx must be a 1d-array.')
                        res.append(self._eval(model, L.cols(x0, model.local_input_idx), None, None, None, m))
                    else:
                        if isinstance(model.local_input_idx, list):
                            if len(model.local_input_idx) != l:
                                raise Exception('local_input_idx should be a list that has length of %i.' % l)
                            lidx = model.local_input_idx
                        else:
                            lidx = [None] * (l - 1) + [model.local_input_idx]
                        mi = [L.cols(feeds_m[i], lidx[i]) for i in range(l) if lidx[i] is not None]
                        vi = [L.cols(feeds_v[i], lidx[i]) for i in range(l) if lidx[i] is not None]
                        res.append(self._eval(model, None, torch.cat(mi, 1), torch.cat(vi, 1), xext[l][k], m))
                outs.append(res)
                feeds_m.append(torch.cat([r[0] for r in res], 1))
                feeds_v.append(torch.cat([r[1] for r in res], 1))
            per_imp.append(outs)
        S = len(per_imp)

        layers = range(self.L) if full_layer else [self.L - 1]
        if method == 'sampling':
            # draws in the reference's order (linkgp.py:381-384, 415-417): imputation by imputation, layer by layer,
            # emulator by emulator, one (sample_size x M x D_out) block of normals each
            blocks = {}
            for s in range(S):
                for l in layers:
                    for k in range(len(self.all_layer[l])):
                        mu_k, va_k = L.to_host(per_imp[s][l][k][0]), L.to_host(per_imp[s][l][k][1])
                        draw = np.random.normal(mu_k, np.sqrt(va_k), size=(sample_size,) + mu_k.shape)
                        blocks.setdefault((l, k), []).append(draw.transpose(2, 1, 0))
            out = [[np.concatenate(blocks[(l, k)], axis=2) for k in range(len(self.all_layer[l]))] for l in layers]
            return out if full_layer else out[0]

        def agg(l, k):
            ms = torch.stack([per_imp[s][l][k][0] for s in range(S)], 0).contiguous()
            vs = torch.stack([per_imp[s][l][k][1] for s in range(S)], 0).contiguous()
            mu, s2 = torch.empty_like(ms[0]), torch.empty_like(ms[0])
            L.check(lib.dgpb_aggregate(L.ptr(ms), L.ptr(vs), S, ms[0].numel(), L.ptr(mu), L.ptr(s2), L.stream()))
            return L.to_host(mu), L.to_host(s2)

        mus, s2s = [], []
        for l in layers:
            pairs = [agg(l, k) for k in range(len(self.all_layer[l]))]
            mus.append([p[0] for p in pairs])
            s2s.append([p[1] for p in pairs])
        if full_layer:
            return mus, s2s
        return mus[0], s2s[0]

    def ppredict(self, x, method='mean_var', full_layer=False, sample_size=50, m=50, chunk_num=None, core_num=None):
        """Process-pool variant of the reference (linkgp.py:214-283); one GPU replaces the pool."""
        return self.predict(x, method, full_layer, sample_size, m)
