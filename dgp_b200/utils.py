"""Host utilities with the reference's names (dgpsi/utils.py): pickling, seeding, thread shims and the
Nystrom kernel-PCA used once at initialisation when a latent layer narrows (dgp.py:568-571; stays on the CPU,
SURVEY.md section 2 row 9)."""
from __future__ import annotations

import pickle

import numpy as np

from .imputation import nb_seed  # noqa: F401  (re-exported: `from dgp_b200 import nb_seed`)

try:
    from dill import dump, load
except Exception:  # pragma: no cover - dill is optional
    dump, load = pickle.dump, pickle.load

_threads = [1]


def write(emu, pkl_file):
    """Save an emulator to `<pkl_file>.pkl` (utils.py:18-27); device caches are dropped by the objects'
    __getstate__ so the file holds numpy only."""
    with open(pkl_file + ".pkl", "wb") as f:
        dump(emu, f)


def read(pkl_file):
    """Load `<pkl_file>.pkl` (utils.py:30-42)."""
    with open(pkl_file + ".pkl", "rb") as f:
        return load(f)


def get_thread():
    """The reference uses the numba thread count to choose serial vs parallel njit kernels
    (kernel_class.py:601-602); on the GPU path it has no effect and is kept for API compatibility."""
    return _threads[0]


def set_thread(value):
    _threads[0] = int(value)


def summary(obj, tablefmt='fancy_grid'):
    """Plain-text summary of the GP nodes of a dgp / emulator / gp / list structure (utils.py:69-190, reduced
    to the columns that exist on this path)."""
    layers = getattr(obj, 'all_layer', None)
    if layers is None and hasattr(obj, 'kernel'):
        layers = [[obj.kernel]]
    if layers is None:
        layers = obj
    rows = []
    for l, layer in enumerate(layers):
        for k, node in enumerate(layer):
            node = getattr(node, 'structure', node)
            if isinstance(node, list):
                continue
            if getattr(node, 'type', 'gp') == 'likelihood':   # utils.py:119-123 prints NA for these columns
                rows.append([f"Likelihood{l + 1}.{k + 1}", node.name, 'NA', 'NA', 'NA',
                             None if node.input_dim is None else list(np.atleast_1d(node.input_dim)), 'NA'])
                continue
            rows.append([f"GP{l + 1}.{k + 1}", node.name, np.array2string(np.atleast_1d(node.length), precision=3),
                         float(np.atleast_1d(node.scale)[0]), float(np.atleast_1d(node.nugget)[0]),
                         None if node.input_dim is None else list(np.atleast_1d(node.input_dim)),
                         None if node.connect is None else list(np.atleast_1d(node.connect))])
    header = ["node", "kernel", "length", "scale", "nugget", "input_dim", "connect"]
    try:
        from tabulate import tabulate
        text = tabulate(rows, headers=header, tablefmt=tablefmt)
    except Exception:  # pragma: no cover
        text = "\n".join(str(r) for r in [header] + rows)
    print(text)
    return text


class NystromKPCA:
    """Sigmoid-kernel PCA through an m-point Nystrom approximation (utils.py:203-269)."""

    def __init__(self, n_components, m=200):
        self.m = m
        self.n_components = n_components
        self.basis_inds = None

    @staticmethod
    def _pinv(K, sqrt=False):
        U, S, Vt = np.linalg.svd(K)
        S = np.maximum(S, 1e-12)
        return np.dot(U / (np.sqrt(S) if sqrt else S), Vt)

    def fit_transform(self, X):
        from sklearn.metrics.pairwise import pairwise_kernels
        n = X.shape[0]
        self.m = min(n, self.m)
        self.basis_inds = np.random.permutation(n)[:self.m]
        K_nm = pairwise_kernels(X, X[self.basis_inds], metric='sigmoid', filter_params=True)
        K_mm = K_nm[self.basis_inds]
        # double centring in feature space
        col_mean = K_nm.sum(0) / n
        m0 = self._pinv(K_mm) @ col_mean[:, None]
        M1 = np.tile(col_mean, (n, 1))
        M3 = col_mean @ m0
        K_nm_c = K_nm - M1 - np.tile(K_nm @ m0, (1, self.m)) + M3
        M1 = M1[:self.m]
        K_mm_c = K_mm - M1 - M1.T + M3
        Kis = self._pinv(K_mm_c, sqrt=True)
        _, U = np.linalg.eigh(Kis @ K_nm_c.T @ K_nm_c @ Kis / n)
        U = U[:, ::-1]
        scores = K_nm_c @ (Kis @ U[:, :self.n_components])
        flip = (scores.min(0) + scores.max(0)) / 2 < 0
        return scores @ np.diag(1 - 2 * flip)
