"""`dgp` -- deep Gaussian process trained by stochastic-imputation EM, API as dgpsi/dgp.py:26-129,1364-1541.

Python keeps the object graph, the SEM loop, the L-BFGS-B driver and the parameter traces; the I-step
(ESS sweeps) and every likelihood / gradient evaluation of the M-step run in libdgpb.so.  Scope follows
SURVEY.md section 2 row 5: `__init__`, the generic branch of `initialize` (dgp.py:565-691), `train`,
`estimate`, Vecchia switches and the restart logic; widened by likelihood final layers with
their warm starts (dgp.py:163-203, 279-296, 327-372, 411-459, 526-532) and `update_xy*`.  Replicated inputs and
plotting are not part of the SI hot path (NotImplementedError); `ptrain` is `train`.
"""
from __future__ import annotations

import copy
import threading
import time

import numpy as np

from .imputation import imputer
from .kernel_class import combine
from .kernel_class import kernel as ker
from .utils import NystromKPCA

try:  # tqdm is optional plumbing for the progress bar
    from tqdm import tqdm, trange
except Exception:  # pragma: no cover
    tqdm = None

    def trange(a, b, disable=False):
        class _R:
            def __iter__(self):
                return iter(range(a, b))

            def set_description(self, *_):
                pass

            def close(self):
                pass

        return _R()


_tls = threading.local()
_MSTEP_POOL = None      # persistent optimiser threads of the batched M-step
_MSTEP_POOL_SIZE = 0


def _single_threaded_blas():
    """The host side of an M-step is many SMALL LAPACK calls (`kernel.r2`: two ranks and a least-squares fit of
    n x 9 matrices per node, from one optimiser thread per node).  A multi-threaded BLAS spends 30-100 ms per call
    spinning up its pool for 1-2 ms of arithmetic (measured: lstsq 112 ms with 8 threads, 2.4 ms with one), so the
    pool is limited to one thread for the duration."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:  # pragma: no cover
        import contextlib
        return contextlib.nullcontext()


class _GradBatcher:
    """Rendezvous of the M-step's optimiser threads: a request blocks until every thread that is still optimising
    has one pending, then the last arrival runs the whole batch through `dgpb_nllik_grad_dense_batch` (on the
    workspace of the thread that started the M-step) and wakes the others."""

    def __init__(self, nworkers, ws, dev, shard=None):
        import threading
        self.cond = threading.Condition()
        self.active = nworkers
        self.pending = []       # [node, n, P, result slot, request id]
        self.ws, self.dev = ws, dev
        # one chain on several GPUs: (rank, world, torch.distributed).  Every rank runs every optimiser (identical
        # numbers in, identical decisions out); the matrices of a round are dealt over the ranks and the results
        # exchanged, so the ranks stay balanced however unevenly the optimisers finish.
        self.shard = shard
        # batched calls, matrices, seconds inside the library, seconds between rounds (optimiser threads on the host),
        # seconds before the first round: bench.py's M-step breakdown
        self.stats = [0, 0, 0.0, 0.0, 0.0, 0.0, 0.0]   # ... + whole rounds, tail after the last round
        self._t_last = time.perf_counter()
        self._first = True
        self._step = None          # matrices per batched call, fixed at the first round (cudaMemGetInfo is not free)

    MAX_BATCH = 32   # MAXB of the library's batched launches (dense.cuh)

    def _chunk(self, n):
        """Matrices per batched call: at most MAXB, and no more than the augmented (2n+1)^2 layouts that fit in
        about half of the free device memory (a DGP may have any number of GP nodes; the reference has no limit)."""
        from . import _lib as L
        per = (2 * (n + 64) + 8) ** 2 * 8
        try:
            free = L.torch_mod().cuda.mem_get_info(self.dev)[0]
        except Exception:  # pragma: no cover
            free = 32 << 30
        held = L.load().dgpb_ws_bytes(self.ws)   # scratch the workspace already owns is reused, not added
        return int(max(1, min(self.MAX_BATCH, (free // 2 + held) // per)))

    def _flush_locked(self):
        from . import _lib as L
        reqs, self.pending = self.pending, []
        reqs.sort(key=lambda r: r[4])     # the same order on every rank, whatever the thread timing
        t_in = time.perf_counter()
        self.stats[4 if self._first else 3] += t_in - self._t_last
        self._first = False
        try:
            n = reqs[0][1]
            if any(r[1] != n for r in reqs):
                raise RuntimeError("nodes of one M-step must share the number of training points")
            L.torch_mod().cuda.set_device(self.dev)
            ldo = max(r[2] for r in reqs) + 2
            out = np.zeros((len(reqs), ldo + 1))          # last column: status
            rank, world = (self.shard[0], self.shard[1]) if self.shard else (0, 1)
            mine = [i for i in range(len(reqs)) if i % world == rank]
            if self._step is None:
                self._step = self._chunk(n)
            step = self._step
            msg = ""
            for lo in range(0, len(mine), step):
                part = mine[lo:lo + step]
                B = len(part)
                arr = (L.DgpbNode * B)(*[reqs[i][0] for i in part])
                res = np.zeros((B, ldo))
                status = np.zeros(B, dtype=np.int32)
                t0 = time.perf_counter()
                rc = L.load().dgpb_nllik_grad_dense_batch(self.ws, arr, B, n, res.ctypes.data_as(L.c_vp), ldo,
                                                          status.ctypes.data_as(L.c_vp), L.stream())
                self.stats[0] += 1
                self.stats[1] += B
                self.stats[2] += time.perf_counter() - t0
                if rc != L.DGPB_OK:
                    msg = L.load().dgpb_last_error().decode("utf-8", "replace")
                for b, i in enumerate(part):
                    out[i, :ldo] = res[b]
                    out[i, ldo] = rc if rc != L.DGPB_OK else int(status[b])
            if world > 1:   # rows of the other ranks are zero here: one sum all-reduce hands every result to every rank
                torch = L.torch_mod()
                t = torch.from_numpy(out).to(L.device())
                self.shard[2].all_reduce(t)
                out = t.cpu().numpy()
            for i, r in enumerate(reqs):
                r[3].extend([int(out[i, ldo]), out[i, :ldo].copy(),
                             msg or "matrix %d of the batch is not positive definite" % i])
        except Exception as exc:  # never leave the other optimiser threads waiting
            for r in reqs:
                if not r[3]:
                    r[3].extend([L.DGPB_CUDA_ERROR, None, str(exc)])
        self._t_last = time.perf_counter()
        self.stats[5] += self._t_last - t_in
        self.cond.notify_all()

    def evaluate(self, node, n, P, rid=0):
        from . import _lib as L
        slot = []
        with self.cond:
            self.pending.append([node, n, P, slot, rid])
            if len(self.pending) >= self.active:
                self._flush_locked()
            else:
                while not slot:
                    self.cond.wait()
        rc, out, msg = slot
        if rc == L.DGPB_NOT_PD:
            raise np.linalg.LinAlgError(msg or "matrix is not positive definite")
        if rc != L.DGPB_OK:
            raise RuntimeError("dgp_b200: " + msg)
        return out

    def retire(self):
        """An optimiser thread has finished: the others no longer wait for it."""
        with self.cond:
            self.active -= 1
            if self.pending and len(self.pending) >= self.active:
                self._flush_locked()


class dgp:
    """Deep GP hierarchy for stochastic imputation inference (arguments: dgp.py:71)."""

    def __init__(self, X, Y, all_layer=None, check_rep=True, block=True, vecchia=False, m=25, ord_fun=None):
        self.Y = Y
        if isinstance(self.Y, list):
            if len(self.Y) == 1:
                self.Y = self.Y[0]
            else:
                raise Exception('Y has to be a numpy 2d-array rather than a list. The list version of Y (for linked '
                                'emulation) has been reduced. Please use the dedicated lgp class for linked emulation.')
        if (self.Y).ndim == 1 or X.ndim == 1:
            raise Exception('The input and output data have to be numpy 2d-arrays.')
        self.check_rep = check_rep
        self.indices = None
        self.counts = None
        if self.check_rep:
            X0 = np.unique(X, axis=0)
            if len(X0) != len(X):
                raise NotImplementedError("dgp_b200: repeated input rows (replicates) are outside the SI hot path")
        self.X = X
        self.vecch = vecchia
        self.n_data = self.X.shape[0]
        self.nn_method = 'exact'
        self.m = min(m, self.n_data - 1)
        self.ord_fun = ord_fun
        if all_layer is None:  # default input-connected two-layer DGP (dgp.py:105-109)
            D, Y_D = np.shape(self.X)[1], np.shape(self.Y)[1]
            layer1 = [ker(length=np.array([1.])) for _ in range(D)]
            layer2 = [ker(length=np.array([1.]), scale_est=True, connect=np.arange(D)) for _ in range(Y_D)]
            all_layer = combine(layer1, layer2)
        self.all_layer = all_layer
        self.n_layer = len(self.all_layer)
        for l, layer in enumerate(self.all_layer):
            for node in layer:
                if getattr(node, 'type', 'gp') != 'gp':
                    from . import _lib as L
                    if l != self.n_layer - 1 or l == 0 or any(nd.type == 'gp' for nd in layer):
                        raise NotImplementedError("dgp_b200: likelihood nodes are supported as a final layer made of "
                                                  "likelihood nodes only")
                    if node.name not in L.LIK_KIND:
                        raise NotImplementedError("dgp_b200: likelihood '%s' is outside the SI hot path" % node.name)
        top = self.all_layer[-1][0]
        if getattr(top, 'name', None) == 'Categorical':   # dgp.py:112-121
            from sklearn.preprocessing import LabelEncoder
            top.class_encoder = LabelEncoder()
            self.Y = top.class_encoder.fit_transform(self.Y.flatten()).reshape(-1, 1)
            if top.num_classes is None:
                top.num_classes = len(top.class_encoder.classes_)
            if top.link is None:
                top.link = "logit" if top.num_classes == 2 else "softmax"
        self.initialize()
        self.block = block
        with self.change_init_scale():
            self.imp = imputer(self.all_layer, self.block)
            (self.imp).sample(burnin=10)
            self.compute_r2()
        self.N = 0
        self.burnin = None
        self.timing = {'i_step': 0.0, 'm_step': 0.0}

    def __setstate__(self, state):
        for key, val in (('block', True), ('vecch', False), ('nn_method', 'exact'), ('m', 25), ('ord_fun', None),
                         ('counts', None)):
            state.setdefault(key, val)
        state.setdefault('n_data', state['X'].shape[0])
        self.__dict__.update(state)

    # ---- wiring of inputs / outputs (generic branch of dgp.py:154-691 and :1097-1362) ---------------
    def change_init_scale(self):
        """Context of dgp.py:1575-1586: during the first burn-in the GP nodes that feed a Categorical likelihood and
        estimate their scale run with scale 40 (the latent start values are +-2 sqrt(40))."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old = []
            cat = getattr(self.all_layer[-1][0], 'name', None) == 'Categorical'
            if cat:
                for kernel in self.all_layer[-2]:
                    old.append(kernel.scale)
                    if kernel.scale_est:
                        kernel.scale = np.array([40.0])
            try:
                yield
            finally:
                if cat:
                    for scale, kernel in zip(old, self.all_layer[-2]):
                        kernel.scale = scale

        return ctx()

    def _latent_init_likelihood(self, In, l):
        """Warm start of the GP layer that feeds a single likelihood node (dgp.py:163-203, 327-330, 526-532, branches
        without replicates); None when layer l is not such a layer."""
        if l != self.n_layer - 2 or len(self.all_layer[l + 1]) != 1:
            return None
        lik = self.all_layer[l + 1][0]
        if getattr(lik, 'type', 'gp') != 'likelihood':
            return None
        width = len(self.all_layer[l])
        y = self.Y.flatten()
        if lik.name == 'Categorical':   # dgp.py:279-296
            if width != lik.n_inputs:
                raise Exception('You need %d GP node(s) to feed the Categorical likelihood node.' % lik.n_inputs)
            c = 2 * np.sqrt(40)
            if lik.num_classes == 2:
                return np.where(self.Y == 1, c, -c)
            Out = -c * np.ones((self.n_data, lik.num_classes))
            Out[np.arange(self.n_data), self.Y.ravel()] = c
            return Out
        if lik.name in ('ZIP', 'ZINB'):   # dgp.py:337-372, 411-459
            N = len(y)
            Out = np.empty((np.shape(In)[0], width))
            Out[:, 0] = np.log(np.maximum(y + 0.5, 1e-6) + 1e-12)
            if lik.name == 'ZINB':
                sigma = (y.var(ddof=1) - y.mean()) / (y.mean() ** 2 + 1e-8) if N > 1 else 1.0
                Out[:, 1] = np.log(min(max(sigma, 1e-3), 10.0))
            p0 = ((y == 0).sum() + 0.5) / (N + 1.0)          # smoothed share of zeros
            mu = y.mean()
            if mu <= 0:
                pi0 = p0
            else:
                q0 = np.exp(-max(mu, 1e-6))                  # zero probability of the count part
                pi0 = 0.0 if q0 >= 1.0 - 1e-8 else np.clip((p0 - q0) / (1.0 - q0), 0.0, 0.99)
            pi0 = np.clip(pi0, 1e-4, 1.0 - 1e-4)
            Out[:, -1] = np.log(pi0 / (1.0 - pi0))
            return Out
        if lik.name == 'Poisson':
            return np.log(self.Y + .5 + 1e-12)
        if lik.name == 'NegBin':
            Out = np.empty((np.shape(In)[0], width))
            Out[:, 0] = np.log(y + .5 + 1e-12)
            # the reference leaves the dispersion column unset without replicates (dgp.py:527-532); start it from the
            # method-of-moments value its replicate branch uses globally (dgp.py:534-539)
            sigma = max((y.var(ddof=1) - y.mean()) / (y.mean() ** 2 + 1e-8), 1e-3)
            Out[:, 1:] = np.log(sigma)
            return Out
        if lik.name == 'Hetero' and width == 2:
            from .gp import gp
            D = self.X.shape[1]
            Out = np.empty((np.shape(In)[0], width))
            Out[:, 0] = y
            fit = gp(self.X, y.reshape(-1, 1), ker(length=np.ones(D), name=self.all_layer[-2][0].name, scale_est=True,
                                                  nugget_est=True, prior_name='ref', nugget=1e-2),
                     vecchia=self.vecch, m=self.m, ord_fun=self.ord_fun)
            fit.train()
            mean_mu = fit.loo()[0].flatten()
            z = np.log(np.maximum((y - mean_mu) ** 2, 1e-12) + 1e-12)
            fit = gp(self.X, z.reshape(-1, 1), ker(length=np.ones(D), name=self.all_layer[-2][1].name, scale_est=True,
                                                  nugget_est=True, prior_name='ref', nugget=1e-2),
                     vecchia=self.vecch, m=self.m, ord_fun=self.ord_fun)
            fit.train()
            mean_lv, var_lv = fit.loo()
            mean_lv = mean_lv.flatten()
            var_lv = np.maximum((var_lv - fit.kernel.nugget * fit.kernel.scale).flatten(), 1e-12)
            sd_lv = np.sqrt(var_lv)
            z_init = np.random.normal(loc=mean_lv, scale=sd_lv)
            Out[:, 1] = np.clip(z_init, mean_lv - 2.576 * sd_lv, mean_lv + 2.576 * sd_lv)
            return Out
        return None

    def _latent_init(self, In, width):
        """Warm start of a latent layer: copy of its input, Nystrom kernel-PCA when it narrows, column
        resampling when it widens (dgp.py:565-576)."""
        d = np.shape(In)[1]
        if d == width:
            return copy.copy(In)
        if d > width:
            if self.vecch or self.n_data >= 500:
                return NystromKPCA(n_components=width).fit_transform(In)
            from sklearn.decomposition import KernelPCA
            return KernelPCA(n_components=width, kernel='sigmoid').fit_transform(In)
        return np.concatenate((In, In[:, np.random.choice(d, width - d)]), 1)

    def _needs_pointer(self, l, k):
        """dgp.py:632-639: the node is the mean process of a Hetero likelihood (its exact conditional draw under the
        Vecchia approximation needs the latent-Vecchia conditioning sets, kernel.ord_nn(pointer=True))."""
        if l != self.n_layer - 2:
            return False
        linked = [lk for lk in self.all_layer[l + 1] if lk.input_dim is None or k in lk.input_dim]
        if len(linked) != 1 or linked[0].type != 'likelihood' or linked[0].exact_post_idx is None:
            return False
        idx = np.where(np.asarray(linked[0].input_dim) == k)[0] if linked[0].input_dim is not None else np.array([k])
        return bool(np.isin(idx, linked[0].exact_post_idx).all()) and len(idx) > 0

    def _share_or_draw_ord(self, layer, k, pointer=False):
        kernel = layer[k]
        for j in range(k):
            same = np.array_equal(kernel.input_dim, layer[j].input_dim) and np.array_equal(kernel.connect,
                                                                                           layer[j].connect)
            if len(kernel.length) == 1 and same and len(layer[j].length) == 1:
                kernel.ord_nn(ord=layer[j].ord, NNarray=layer[j].NNarray, pointer=pointer)
                return
            if len(kernel.length) != 1 and same and np.array_equal(kernel.length, layer[j].length):
                kernel.ord_nn(ord=layer[j].ord.copy(), NNarray=layer[j].NNarray.copy(), pointer=pointer)
                return
        kernel.ord_nn(pointer=pointer)

    def _wire(self, first_time, reset_row=None):
        In = self.X
        for l, layer in enumerate(self.all_layer):
            last = l == self.n_layer - 1
            if not last:
                Out = self._latent_init_likelihood(In, l)
                if Out is None:
                    Out = self._latent_init(In, len(layer))
            for k, kernel in enumerate(layer):
                if kernel.input_dim is None:
                    kernel.input_dim = np.arange(np.shape(In)[1])
                kernel.input = In[:, kernel.input_dim].copy()
                if kernel.type == 'likelihood':  # dgp.py:584-590, 665-667
                    if len(kernel.input_dim) != kernel.n_inputs:
                        raise Exception('You need %d and only %d GP node(s) to feed the %s likelihood node.'
                                        % (kernel.n_inputs, kernel.n_inputs, kernel.name))
                    kernel.output = self.Y[:, [k]]
                    continue
                if kernel.connect is not None:
                    if l == 0 and len(np.intersect1d(kernel.connect, kernel.input_dim)) != 0:
                        raise Exception('The local input and global input should not have any overlap. Change '
                                        'input_dim or connect so they do not have any common indices.')
                    kernel.global_input = self.X[:, kernel.connect]
                kernel.m = self.m
                if first_time:
                    kernel.vecch, kernel.nn_method = self.vecch, self.nn_method
                    if self.ord_fun is not None:
                        kernel.ord_fun = self.ord_fun
                    kernel.D = np.shape(kernel.input)[1] + (len(kernel.connect) if kernel.connect is not None else 0)
                elif reset_row is not None:
                    hyp = kernel.para_path[reset_row, :]
                    kernel.scale, kernel.length, kernel.nugget = hyp[[0]], hyp[1:-1], hyp[[-1]]
                if kernel.vecch:
                    self._share_or_draw_ord(layer, k, self._needs_pointer(l, k))
                kernel.output = self.Y[:, [k]] if last else Out[:, [k]].copy()
                if kernel.prior_name == 'ref':
                    if first_time:
                        p = kernel.D
                        b = 1 / len(kernel.output) ** (1 / p) * (kernel.prior_coef + p)
                        kernel.prior_coef = np.concatenate((kernel.prior_coef, b))
                    kernel.compute_cl()
                if first_time:
                    kernel.para_path = np.atleast_2d(np.concatenate((kernel.scale, kernel.length, kernel.nugget)))
            if not last:
                In = copy.copy(Out)

    def initialize(self):
        """Initialise all_layer attribute for training (dgp.py:154)."""
        self._wire(first_time=True)

    def reinit_all_layer(self, reset_lengthscale, row=0):
        """Re-initialise the latent layers (and optionally the hyper-parameters) after a failed run
        (dgp.py:1097-1362, generic branch)."""
        self._wire(first_time=False, reset_row=row if reset_lengthscale else None)

    # ---- warm start with new data (sequential design loops, dgp.py:824-1095) -----------------------------
    def update_xy(self, X, Y, reset=False):
        """Update the trained DGP with new input and output data (dgp.py:824-884).  Without `reset` the latent
        layers are carried over: rows kept when the new design is a subset of the old one, GP conditional means at
        the added rows when it is a superset, a fresh warm start otherwise."""
        self.Y = Y
        if isinstance(self.Y, list):
            if len(self.Y) == 1:
                self.Y = self.Y[0]
            else:
                raise Exception('Y has to be a numpy 2d-array rather than a list. The list version of Y (for linked '
                                'emulation) has been reduced. Please use the dedicated lgp class for linked emulation.')
        if (self.Y).ndim == 1 or X.ndim == 1:
            raise Exception('The input and output data have to be numpy 2d-arrays.')
        if self.check_rep and len(np.unique(X, axis=0)) != len(X):
            raise NotImplementedError("dgp_b200: repeated input rows (replicates) are outside the SI hot path")
        if getattr(self.all_layer[-1][0], 'name', None) == 'Categorical':   # dgp.py:845-846
            self.Y = self.all_layer[-1][0].class_encoder.transform(self.Y.flatten()).reshape(-1, 1)
        self.indices = None
        origin_X = (self.X).copy()
        self.X = X
        self.n_data = self.X.shape[0]
        self.m = min(self.m, self.n_data - 1)
        if reset:
            self.reinit_all_layer(reset_lengthscale=True)
            burnin = 10
        elif (self.X[:, None] == origin_X).all(-1).any(-1).all():
            self.update_all_layer_smaller(np.where((origin_X == self.X[:, None]).all(-1))[1])
            burnin = 50
        elif (origin_X[:, None] == self.X).all(-1).any(-1).all():
            self.update_all_layer_larger(np.where((self.X == origin_X[:, None]).all(-1))[1])
            burnin = 50
        else:
            self.reinit_all_layer(reset_lengthscale=False)
            burnin = 200
        self.imp = imputer(self.all_layer, self.block)
        (self.imp).sample(burnin=burnin)
        self.compute_r2()

    def _rewire_node(self, layer, k, last):
        """Per-node tail shared by the two carry-over updates: global inputs, Vecchia ordering, final outputs."""
        kernel = layer[k]
        if kernel.type == 'likelihood':
            kernel.output = (self.Y[:, [k]]).copy()
            return
        if kernel.connect is not None:
            kernel.global_input = (self.X[:, kernel.connect]).copy()
        kernel.m = self.m
        if kernel.vecch:
            self._share_or_draw_ord(layer, k, kernel.imp_pointer_row is not None)
        if last:
            kernel.output = (self.Y[:, [k]]).copy()
        if kernel.prior_name == 'ref':
            kernel.compute_cl()

    def update_all_layer_smaller(self, sub_idx):
        """The new design is a subset of the old one: keep those rows of every latent layer (dgp.py:1014-1095)."""
        for l, layer in enumerate(self.all_layer):
            last = l == self.n_layer - 1
            for k, kernel in enumerate(layer):
                kernel.input = kernel.input[sub_idx, :]
                if not last:
                    kernel.output = (kernel.output[sub_idx, :]).copy()
                if kernel.connect is not None and l == 0 and len(np.intersect1d(kernel.connect, kernel.input_dim)) != 0:
                    raise Exception('The local input and global input should not have any overlap. Change '
                                    'input_dim or connect so they do not have any common indices.')
                self._rewire_node(layer, k, last)

    def update_all_layer_larger(self, sub_idx):
        """The old design is a subset of the new one: latent values at the added rows are the conditional means of
        the current GP nodes given their imputed outputs (dgp.py:886-1012; `cond_mean` functions.py:301-309,
        `cond_mean_vecch` vecchia.py:624-633 with 50 neighbours), computed by the prediction kernels."""
        from . import _lib as L
        In = (self.X).copy()
        mask = np.zeros(len(self.X), dtype=bool)
        mask[sub_idx] = True
        new_rows = np.where(~mask)[0]
        for l, layer in enumerate(self.all_layer):
            last = l == self.n_layer - 1
            if not last:
                Out = np.empty((len(In), len(layer)))
            for k, kernel in enumerate(layer):
                if not last:
                    if len(new_rows):
                        if kernel.vecch:
                            kernel.pred_m = 50
                        else:
                            kernel.compute_stats()
                        x_new = L.to_dev(np.ascontiguousarray(In[new_rows][:, kernel.input_dim]), np.float64)
                        z_new = None if kernel.connect is None else \
                            L.to_dev(np.ascontiguousarray(self.X[new_rows][:, kernel.connect]), np.float64)
                        mu, _ = kernel._gp_prediction_dev(x_new, z_new)
                        Out[new_rows, k] = L.to_host(mu)
                        kernel._Rinv = kernel._Rinv_y = None
                    Out[sub_idx, k] = kernel.output.flatten()
                    kernel.output = Out[:, [k]].copy()
                kernel.input = (In[:, kernel.input_dim]).copy()
                self._rewire_node(layer, k, last)
            if not last:
                In = Out.copy()

    def to_vecchia(self, m=25, ord_fun=None):
        """Convert the DGP structure to the Vecchia mode (dgp.py:693-746)."""
        if self.vecch:
            raise Exception('The DGP structure is already in Vecchia mode.')
        self.vecch = True
        self.m = min(m, self.n_data - 1)
        self.ord_fun = ord_fun
        for l, layer in enumerate(self.all_layer):
            for k, kernel in enumerate(layer):
                if kernel.type != 'gp':
                    continue
                kernel.vecch, kernel.m, kernel.ord_fun = self.vecch, self.m, self.ord_fun
                self._share_or_draw_ord(layer, k, self._needs_pointer(l, k))

    def remove_vecchia(self):
        """Remove the Vecchia mode from the DGP structure (dgp.py:748-758)."""
        if not self.vecch:
            raise Exception('The DGP structure is already in non-Vecchia mode.')
        self.vecch = False
        for layer in self.all_layer:
            for kernel in layer:
                if kernel.type == 'gp':
                    kernel.vecch = self.vecch

    # ---- stochastic EM ----------------------------------------------------------------------------------
    def train(self, N=500, ess_burn=10, disable=False):
        """Train the DGP model by stochastic EM (dgp.py:1364-1412): per iteration an I-step of `ess_burn`+1
        ESS sweeps and an M-step (L-BFGS-B) over every GP node; up to three restarts on LinAlgError."""
        N0 = self.N
        restarts, max_restarts = 0, 3
        pgb = None
        while True:
            try:
                pgb = trange(1, N + 1, disable=disable)
                for i in pgb:
                    t0 = time.perf_counter()
                    if i == 1:   # dgp.py:1383-1385: the first I-step of a call runs under the initial scale
                        with self.change_init_scale():
                            (self.imp).sample(burnin=ess_burn)
                    else:
                        (self.imp).sample(burnin=ess_burn)
                    if self.vecch and (self.N + i & (self.N + i - 1)) == 0 and self.N + i > 1:
                        (self.imp).update_ord_nn()
                    t1 = time.perf_counter()
                    self._m_step()
                    # wall-clock split (both phases end with a device-to-host read): instrumentation for bench.py
                    if not hasattr(self, 'timing'):
                        self.timing = {'i_step': 0.0, 'm_step': 0.0}
                    self.timing['i_step'] += t1 - t0
                    self.timing['m_step'] += time.perf_counter() - t1
                    pgb.set_description('Iteration %i' % i)
                self.N += N
                return
            except (np.linalg.LinAlgError, SystemError):
                restarts += 1
                if pgb is not None:
                    pgb.close()
                if restarts > max_restarts:
                    raise RuntimeError(f"Training failed after {max_restarts} restarts.")
                if not disable and tqdm is not None:
                    tqdm.write(f"Restart {restarts}/{max_restarts}:")
                self.N = N0
                self.reinit_all_layer(reset_lengthscale=True, row=self.N)

    def _m_step(self):
        """M-step over every GP node (dgp.py:1391-1398).  Given the imputation the nodes are independent, so their
        L-BFGS-B runs proceed side by side (one host thread each) and every round of objective/gradient requests
        is served by ONE batched sliding-window factorisation (`dgpb_nllik_grad_dense_batch`): the batch fills the
        GPU where a single n = 5000 node leaves the tensor pipe waiting on its serial panel chain.  Each node sees
        exactly the numbers it would see alone (the batched kernels treat matrices independently), so parameter
        paths do not depend on the scheduling.  Vecchia nodes keep the one-node-at-a-time path."""
        from . import parallel

        nodes = [(l, kernel) for l in range(self.n_layer) for kernel in self.all_layer[l] if kernel.type == 'gp']
        ch = parallel.chain()
        with _single_threaded_blas():
            self._m_step_sharded(nodes, ch)

    def _m_step_sharded(self, nodes, ch):
        import os

        from . import parallel

        if ch is None:
            self._m_step_nodes(nodes)
            return
        # One chain on several GPUs (the reference deals the nodes of a layer to a process pool, dgp.py:1463-1467).
        # Dense nodes: every rank runs every optimiser and the EVALUATIONS of each round are dealt over the ranks
        # (`_GradBatcher.shard`), so no rank waits for a slow optimiser it does not hold.  Other nodes (Vecchia): the
        # nodes themselves are dealt, then the results are exchanged.
        dense = [it for it in nodes if not it[1].vecch]
        # dealing evaluations keeps two ranks balanced (M-step 492 -> 370 ms on config 3); on more ranks every rank
        # would run all optimiser threads (their Python side is serialised by the interpreter lock) through the
        # global maximum of rounds, which costs more than it balances (4 GPUs: 444 against 323 ms), so the nodes
        # themselves are dealt there, longest optimiser first
        if (len(dense) > 1 and self.n_data >= 128 and ch["device"] and ch["world"] == 2
                and os.environ.get('DGPB_MSTEP_BATCH', '1') != '0'):
            self._m_step_nodes(dense, shard=(ch["rank"], ch["world"], ch["dist"]))
            nodes = [it for it in nodes if it[1].vecch]
            if not nodes:
                return
        costs = [getattr(k, '_nfev', 0) for _, k in nodes]
        mine = parallel.mstep_share(len(nodes), ch["rank"], ch["world"], costs if all(c > 0 for c in costs) else None)
        for _, kernel in nodes:
            kernel._r2_fresh = False
        failed = None
        try:
            self._m_step_nodes([nodes[i] for i in mine])
        except np.linalg.LinAlgError as exc:
            failed = exc
        mine_set = set(mine)
        for i, (l, kernel) in enumerate(nodes):   # cheap host state the owner updated inside its share
            if i not in mine_set and kernel.prior_name == 'ref':
                kernel.compute_cl()
        if parallel.sync_params([k for _, k in nodes], mine, failed is not None):
            raise failed if failed is not None else np.linalg.LinAlgError(
                "a GP node optimised on another rank is not positive definite")

    def _m_step_nodes(self, nodes, shard=None):
        """The M-step of the nodes in `nodes` ((layer index, kernel) pairs) on this GPU; with `shard` = (rank, world,
        dist) every rank calls this with the SAME dense nodes and the evaluations are dealt over the ranks."""
        import os
        from concurrent.futures import ThreadPoolExecutor

        from . import _lib as L

        if not nodes:
            return
        torch = L.torch_mod()
        dev = L.device()
        dense = [it for it in nodes if not it[1].vecch]
        # small models: an evaluation is a handful of microsecond kernels, the rendezvous of one host thread per node
        # would cost more than it saves; the nodes are optimised one after the other (same numbers either way)
        batch = shard is not None or (len(dense) > 1 and os.environ.get('DGPB_MSTEP_BATCH', '1') != '0' and
                                      (self.n_data >= 128 or os.environ.get('DGPB_MSTEP_BATCH') == '1'))
        if not batch:
            dense = []
        for l, kernel in nodes:   # nodes outside the batch
            if any(kernel is k for _, k in dense):
                continue
            if kernel.prior_name == 'ref':
                kernel.compute_cl()
            if l != 0:
                kernel.r2()
            kernel.maximise()
        if not dense:
            return
        torch.cuda.current_stream().synchronize()
        batcher = _GradBatcher(len(dense), L.workspace(), dev, shard)
        for rid, (_, kernel) in enumerate(dense):
            kernel._mid = rid        # request id: the rounds are ordered by it on every rank

        def work(item):
            l, kernel = item
            torch.cuda.set_device(dev)
            try:
                if kernel.prior_name == 'ref':
                    kernel.compute_cl()
                if l != 0:
                    kernel.r2()
                kernel._batcher = batcher
                kernel.maximise()
            finally:
                kernel._batcher = None
                batcher.retire()

        # one long-lived pool: the library keeps per-thread CUDA streams and events (look-ahead streams of the
        # factorisation), so the optimiser threads must not be re-created every iteration
        global _MSTEP_POOL, _MSTEP_POOL_SIZE
        if _MSTEP_POOL is None or _MSTEP_POOL_SIZE < len(dense):
            if _MSTEP_POOL is not None:
                _MSTEP_POOL.shutdown(wait=True)
            _MSTEP_POOL_SIZE = max(len(dense), _MSTEP_POOL_SIZE)
            _MSTEP_POOL = ThreadPoolExecutor(max_workers=_MSTEP_POOL_SIZE, thread_name_prefix="dgpb-mstep")
        futures = [_MSTEP_POOL.submit(work, item) for item in dense]
        errors = []
        for fut in futures:   # wait for every optimiser before reporting a failure (dgp.train restarts on LinAlgError)
            try:
                fut.result()
            except BaseException as exc:  # noqa: BLE001
                errors.append(exc)
        batcher.stats[6] = time.perf_counter() - batcher._t_last
        if hasattr(self, 'timing'):
            for key, val in zip(('m_batched_calls', 'm_batched_matrices', 'm_batched_s', 'm_round_gap_s',
                                 'm_first_round_s', 'm_rounds_s', 'm_tail_s'), batcher.stats):
                self.timing[key] = self.timing.get(key, 0) + val
        if errors:
            raise errors[0]

    def ptrain(self, N=500, ess_burn=10, disable=False, core_num=None):
        """dgp.py:1414-1467 optimises the GP nodes of a layer in a process pool; here the nodes of ALL layers already
        share one batched factorisation per optimiser round (`_m_step`), and with `dgp_b200.parallel.enable` they are
        dealt over the GPUs of the box (one process per GPU), so this is `train`."""
        return self.train(N, ess_burn, disable)

    def compute_r2(self):
        for l in range(1, self.n_layer):
            for kernel in self.all_layer[l]:
                if kernel.type == 'gp':
                    kernel.r2(overwritten=True)

    def aggregate_r2(self, burnin=0.75, agg='median'):
        """dgp.py:1481-1515."""
        if burnin < 0 or burnin > 1:
            raise Exception('burnin must be between 0 and 1.')
        if agg not in ('mean', 'median'):
            raise Exception("agg must be either 'median' or 'mean'.")
        f = np.mean if agg == 'mean' else np.median
        res = []
        for layer in self.all_layer:
            res.append([None if getattr(k, 'R2', None) is None else f(k.R2[int(len(k.R2) * burnin):, :], axis=0)
                        for k in layer])
        return res

    def estimate(self, burnin=None):
        """Point estimates = mean of the parameter traces after burn-in (dgp.py:1517-1541)."""
        self.burnin = int(self.N * (3 / 4)) if burnin is None else burnin
        final_struct = copy.deepcopy(self.all_layer)
        for layer in final_struct:
            for kernel in layer:
                if kernel.type != 'gp':
                    continue
                point_est = np.mean(kernel.para_path[self.burnin:, :], axis=0)
                kernel.scale = np.atleast_1d(point_est[0])
                kernel.length = np.atleast_1d(point_est[1:-1])
                kernel.nugget = np.atleast_1d(point_est[-1])
        return final_struct
