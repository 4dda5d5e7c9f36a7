"""`gp` -- single Gaussian-process emulator (dgpsi/gp.py:12-60, 211-222, 412-453), GPU-backed through
`kernel`.  In scope as a member of linked systems: construction, `train`,
`predict`, `export`; `loo`, `metric` (ALM / MICE / VIGF) and `update_xy` follow SURVEY.md 8f-2 / 8f-4."""
from __future__ import annotations

import copy

import numpy as np


class gp:
    """Class for Gaussian process emulation (arguments: gp.py:26)."""

    def __init__(self, X, Y, kernel, check_rep=True, vecchia=False, m=25, ord_fun=None):
        if Y.ndim == 1 or X.ndim == 1:
            raise Exception('The input and output data have to be numpy 2d-arrays.')
        self.check_rep = check_rep
        self.indices = None
        if self.check_rep and len(np.unique(X, axis=0)) != len(X):
            raise NotImplementedError("dgp_b200: repeated input rows (replicates) are outside the SI hot path")
        self.X, self.Y = X, Y
        self.kernel = kernel
        self.vecch = vecchia
        self.n_data = self.X.shape[0]
        self.m = min(m, self.n_data - 1)
        self.ord_fun = ord_fun
        self.initialize()
        if self.vecch:
            self.kernel.ord_nn()
        else:
            self.kernel.compute_stats()

    def __setstate__(self, state):
        for key, val in (('vecch', False), ('nn_method', 'exact'), ('m', 25), ('ord_fun', None), ('indices', None),
                         ('check_rep', False)):
            state.setdefault(key, val)
        state.setdefault('n_data', state['X'].shape[0])
        self.__dict__.update(state)
        self.kernel.target = 'gp'

    def initialize(self):
        """Assign input/output data to the kernel for training (gp.py:79-115)."""
        k = self.kernel
        if k.input_dim is not None:
            k.input = self.X[:, k.input_dim]
        else:
            k.input = (self.X).copy()
            k.input_dim = np.arange(np.shape(self.X)[1])
        if k.connect is not None:
            if len(np.intersect1d(k.connect, k.input_dim)) != 0:
                raise Exception('The local input and global input should not have any overlap. Change input_dim or '
                                'connect so they do not have any common indices.')
            k.global_input = self.X[:, k.connect]
        k.output = (self.Y).copy()
        k.D = np.shape(k.input)[1] + (len(k.connect) if k.connect is not None else 0)
        k.para_path = np.atleast_2d(np.concatenate((k.scale, k.length, k.nugget)))
        k.vecch, k.m, k.target = self.vecch, self.m, 'gp'
        if self.ord_fun is not None:
            k.ord_fun = self.ord_fun
        if k.prior_name == 'ref':
            p = k.D
            b = 1 / len(k.output) ** (1 / p) * (k.prior_coef + p)
            k.prior_coef = np.concatenate((k.prior_coef, b))
            k.compute_cl()

    def to_vecchia(self, m=25, ord_fun=None):
        if self.vecch:
            raise Exception('The GP emulator is already in Vecchia mode.')
        self.vecch = True
        self.m = min(m, self.n_data - 1)
        self.ord_fun = ord_fun
        self.kernel.vecch, self.kernel.m, self.kernel.ord_fun = True, self.m, ord_fun
        self.kernel.ord_nn()

    def remove_vecchia(self):
        if not self.vecch:
            raise Exception('The GP emulator is already in non-Vecchia mode.')
        self.vecch = False
        self.kernel.vecch = False
        self.kernel.compute_stats()

    def update_xy(self, X, Y, reset=False):
        """Update the trained GP emulator with new input and output data (gp.py:144-181)."""
        if Y.ndim == 1 or X.ndim == 1:
            raise Exception('The input and output data have to be numpy 2d-arrays.')
        self.indices = None
        if self.check_rep and len(np.unique(X, axis=0)) != len(X):
            raise NotImplementedError("dgp_b200: repeated input rows (replicates) are outside the SI hot path")
        self.X, self.Y = X, Y
        self.n_data = self.X.shape[0]
        self.m = min(self.m, self.n_data - 1)
        self.update_kernel(reset_lengthscale=reset)
        if self.vecch:
            self.kernel.ord_nn()
        else:
            self.kernel.compute_stats()

    def update_kernel(self, reset_lengthscale):
        """Assign new input/output data to the kernel (gp.py:183-209)."""
        k = self.kernel
        k.rep = None
        k.input = self.X[:, k.input_dim]
        if k.connect is not None:
            if len(np.intersect1d(k.connect, k.input_dim)) != 0:
                raise Exception('The local input and global input should not have any overlap. Change input_dim or '
                                'connect so they do not have any common indices.')
            k.global_input = self.X[:, k.connect]
        k.output = self.Y.copy()
        k.m = self.m
        if reset_lengthscale:
            hyp = k.para_path[0, :]
            k.scale, k.length, k.nugget = hyp[[0]], hyp[1:-1], hyp[[-1]]
        if k.prior_name == 'ref':
            k.compute_cl()

    def metric(self, x_cand, method='MICE', nugget_s=1., m=50, score_only=False):
        """ALM, MICE or VIGF criterion at the candidate points (gp.py:271-324)."""
        if x_cand.ndim == 1:
            raise Exception('The candidate design set has to be a numpy 2d-array.')
        mu, sigma2 = self.predict(x=x_cand, m=m)
        if method == 'ALM':
            score = sigma2
        elif method == 'MICE':
            from .emulation import emulator
            score = sigma2 / emulator._mice_var(x_cand, x_cand, self.kernel, nugget_s).reshape(-1, 1)
        elif method == 'VIGF':
            from .vecchia import get_pred_nn
            index = get_pred_nn(x_cand, self.X, 1).flatten()     # nearest training input, searched on the device
            bias = (mu - self.Y[index, :]) ** 2
            score = 4 * sigma2 * bias + 2 * sigma2 ** 2
        else:
            raise Exception("method must be 'ALM', 'MICE' or 'VIGF'")
        if score_only:
            return score
        idx = np.argmax(score, axis=0)
        return idx, score[idx, 0]

    def pmetric(self, x_cand, method='MICE', nugget_s=1., m=50, score_only=False, chunk_num=None, core_num=None):
        """gp.py:224-269: the process pool is replaced by the GPU."""
        return self.metric(x_cand, method, nugget_s, m, score_only)

    def ppredict(self, x, method='mean_var', sample_size=50, m=50, chunk_num=None, core_num=None):
        """gp.py:373-410: the process pool is replaced by the GPU."""
        return self.predict(x, method, sample_size, m)

    def train(self):
        """Train the GP model (gp.py:211-216)."""
        self.kernel.maximise()
        if not self.vecch:
            self.kernel.compute_stats()

    def export(self):
        """Export the trained GP (gp.py:218-222)."""
        return [copy.deepcopy(self.kernel)]

    def predict(self, x, method='mean_var', sample_size=50, m=50):
        """gp.py:412-453."""
        if x.ndim == 1:
            raise Exception('The testing input has to be a numpy 2d-array')
        k = self.kernel
        z = x[:, k.connect] if k.connect is not None else None
        k.pred_m = m
        mu, sigma2 = k.gp_prediction(x=x[:, k.input_dim], z=z)
        if method == 'mean_var':
            return mu.reshape(-1, 1), sigma2.reshape(-1, 1)
        if method == 'sampling':
            return np.random.normal(mu, np.sqrt(sigma2), size=(sample_size, len(x))).T
        raise Exception("method must be 'mean_var' or 'sampling'")

    def loo(self, method='mean_var', sample_size=50, m=30):
        """Leave-one-out cross-validation of the GP (gp.py:326-371).  Dense: the closed form from the diagonal of
        R^-1 and R^-1 y already on the device; Vecchia: `gp_vecch` of every training point conditioned on its m
        nearest OTHER training points (loo_gp_vecch, vecchia.py:657-673)."""
        from . import _lib as L
        k = self.kernel
        if self.vecch:
            k.pred_m, k.loo_state = m + 1, True
            try:
                z = self.X[:, k.connect] if k.connect is not None else None
                mu, sigma2 = k.gp_prediction(x=self.X[:, k.input_dim], z=z)
            finally:
                k.loo_state = False
            mu, sigma2 = mu.reshape(-1, 1), sigma2.reshape(-1, 1)
        else:
            Rinv, Rinv_y = k._stats_dev()
            torch = L.torch_mod()
            s2 = 1.0 / torch.diagonal(Rinv)
            sigma2 = L.to_host(s2).reshape(-1, 1)
            mu = self.Y - L.to_host(Rinv_y).reshape(-1, 1) * sigma2
            sigma2 = k.scale * sigma2
        if method == 'mean_var':
            return mu, sigma2
        if method == 'sampling':
            return np.random.normal(mu.flatten(), np.sqrt(sigma2.flatten()), size=(sample_size, len(mu))).T
        raise Exception("method must be 'mean_var' or 'sampling'")
