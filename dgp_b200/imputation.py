"""`imputer` -- ESS-within-Gibbs imputation of the latent layers (dgpsi/imputation.py:6-262) with the whole
I-step resident on the GPU: latent layers are uploaded once per `sample()` call, every block update runs in
libdgpb.so (`dgpb_ess_block`), and the imputed layers are written back to the nodes' numpy attributes at the
end, so the object graph looks exactly as the reference leaves it (`kernel.output`, `kernel.input`).
A final layer of likelihood nodes is handled by `dgpb_ess_block_lik` (elementwise log-likelihoods in the ESS
threshold) and, for Hetero, the exact conditional draw of the mean (`Hetero.posterior_dev`).

Randomness: standard normals for dense prior draws come from the module RNG seeded by `nb_seed` (the
reference draws them from numba's RNG inside `fmvn`, functions.py:118), Vecchia prior draws and all uniforms
from numpy's global RNG (vecchia.py:137, imputation.py:79-119).  The number of uniforms consumed per block
update equals the reference's (threshold, first angle, one per rejection).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib as L

_nb_rng = np.random.RandomState()


def nb_seed(value):
    """Seed the generator behind the dense prior draws (the role numba's RNG plays in the reference,
    utils.py:51-55)."""
    _nb_rng.seed(value)


class _DeviceLayers:
    """Device image of a DGP hierarchy during one I-step."""

    def __init__(self, all_layer):
        self.all_layer = all_layer
        self.n = len(all_layer[0][0].output)
        self.F, self.nodes, self.keep = [], [], []
        self.threshold = {}   # layer pair -> sum of upper log-likelihoods at the current state (last pair only)
        # stored factors belong to the previous hyper-parameters / inputs
        L.check(L.load().dgpb_cache_clear(L.workspace()))
        n = self.n
        self.liks = None      # descriptors of a final likelihood layer
        for l, layer in enumerate(all_layer):
            for kern in layer:
                if kern.rep is not None:
                    raise NotImplementedError("dgp_b200: replicate pooling is outside the SI hot path")
                if l > 0 and kern.type == 'gp' and kern.prior_name == 'ref':
                    # the reference adds log_prior() -- whose scale `cl` follows the proposed inputs -- to the ESS
                    # likelihood of such a node (kernel_class.py:489-491, 506-508); the device loop has no such term
                    raise NotImplementedError("dgp_b200: the reference prior ('ref') on a GP node above the first "
                                              "layer changes the ESS acceptance rule and is not built")
            if any(kern.type != 'gp' for kern in layer):
                if l != len(all_layer) - 1 or l == 0 or any(kern.type != 'likelihood' for kern in layer):
                    raise NotImplementedError("dgp_b200: likelihood nodes are supported as a final layer made of "
                                              "likelihood nodes only")
                self.liks = []
                for kern in layer:
                    if kern.name not in L.LIK_KIND:
                        raise NotImplementedError("dgp_b200: likelihood '%s' is outside the SI hot path" % kern.name)
                    y = L.to_dev(np.ascontiguousarray(kern.output[:, 0], dtype=np.float64))
                    self.keep.append(y)
                    self.liks.append((kern, kern._descriptor(kern.input_dim, y), y))
                continue
            Fl = L.to_dev(np.ascontiguousarray(np.stack([k.output[:, 0] for k in layer], 0)))
            self.F.append(Fl)
            arr = (L.DgpbNode * len(layer))()
            for k, kern in enumerate(layer):
                if l == 0:
                    src = L.to_dev(np.ascontiguousarray(kern.input.T))
                    input_dim = np.arange(kern.input.shape[1])
                else:
                    src = self.F[l - 1]
                    input_dim = kern.input_dim
                gsrc = L.to_dev(np.ascontiguousarray(kern.global_input.T)) if kern.global_input is not None else None
                ordd = nnd = None
                if kern.vecch:
                    ordd, nnd = L.to_dev(kern.ord, np.int64), L.to_dev(kern.NNarray, np.int64)
                self.keep += [src, gsrc, ordd, nnd]
                L.fill_node(arr[k], kind=kern.name, input_dim=input_dim,
                            connect=None if gsrc is None else np.arange(gsrc.shape[0]), length=kern.length,
                            scale=kern.scale, nugget=kern.nugget, src=src, gsrc=gsrc, output=None, ord=ordd,
                            NNarray=nnd, m=(kern.NNarray.shape[1] - 1) if kern.vecch else 0, vecch=bool(kern.vecch))
                arr[k].output = Fl.data_ptr() + k * n * 8
            self.nodes.append(arr)

    @staticmethod
    def _key(l, k):
        return l * 4096 + k

    def ess_call(self, l, tks, uks, z, u, reuse=True):
        """One `dgpb_ess_block_cached` call with explicit draws: z (len(tks) x n) standard normals, u uniforms in
        the reference's consumption order.  Returns (proposals evaluated, angles tried).
        `reuse`: keep Cholesky factors / the accepted likelihood across block updates of this I-step (results
        are unchanged: the matrices being re-factored by the reference are identical)."""
        lib = L.load()
        n = self.n
        targets = (L.DgpbNode * len(tks))(*[self.nodes[l][k] for k in tks])
        uppers = (L.DgpbNode * len(uks))(*[self.nodes[l + 1][j] for j in uks])
        rows = np.ascontiguousarray(tks, dtype=np.int32)
        zd = L.to_dev(np.ascontiguousarray(z, dtype=np.float64))
        u = np.ascontiguousarray(u, dtype=np.float64)
        theta = np.zeros(len(u))
        nprop = ctypes.c_int(0)
        tkeys = ukeys = thr = None
        if reuse:
            tkeys = np.ascontiguousarray([self._key(l, k) for k in tks], dtype=np.int32)
            ukeys = np.ascontiguousarray([self._key(l + 1, j) for j in uks], dtype=np.int32)
            whole = len(tks) == len(self.nodes[l]) and len(uks) == len(self.nodes[l + 1])
            if whole and l + 2 == len(self.all_layer):  # outputs of the uppers are the fixed training targets
                thr = ctypes.c_double(self.threshold.get(l, float("nan")))
        status = lib.dgpb_ess_block_cached(
            L.workspace(), targets, len(tks), rows.ctypes.data_as(L.c_vp), L.ptr(self.F[l]), self.F[l].shape[0], uppers,
            len(uks), n, L.ptr(zd), u.ctypes.data_as(L.c_vp), len(u), ctypes.byref(nprop),
            theta.ctypes.data_as(L.c_vp), None if tkeys is None else tkeys.ctypes.data_as(L.c_vp),
            None if ukeys is None else ukeys.ctypes.data_as(L.c_vp), None if thr is None else ctypes.byref(thr),
            L.stream())
        self.last_nprop = nprop.value
        if status != L.DGPB_OK:
            self.threshold.pop(l, None)
        L.check(status)
        if thr is not None:
            self.threshold[l] = thr.value
        else:
            self.threshold.pop(l, None)
        return nprop.value, theta[:nprop.value]

    def lik_call(self, l, tks, lks, z, u):
        """One `dgpb_ess_block_lik` call: targets `tks` of the last GP layer, likelihood nodes `lks`."""
        lib = L.load()
        n = self.n
        targets = (L.DgpbNode * len(tks))(*[self.nodes[l][k] for k in tks])
        liks = (L.DgpbLik * len(lks))(*[self.liks[j][1] for j in lks])
        rows = np.ascontiguousarray(tks, dtype=np.int32)
        zd = L.to_dev(np.ascontiguousarray(z, dtype=np.float64))
        u = np.ascontiguousarray(u, dtype=np.float64)
        theta = np.zeros(len(u))
        nprop = ctypes.c_int(0)
        tkeys = np.ascontiguousarray([self._key(l, k) for k in tks], dtype=np.int32)
        status = lib.dgpb_ess_block_lik(
            L.workspace(), targets, len(tks), rows.ctypes.data_as(L.c_vp), L.ptr(self.F[l]), self.F[l].shape[0], liks,
            len(lks), n, L.ptr(zd), u.ctypes.data_as(L.c_vp), len(u), ctypes.byref(nprop),
            theta.ctypes.data_as(L.c_vp), tkeys.ctypes.data_as(L.c_vp), L.stream())
        self.last_nprop = nprop.value
        L.check(status)
        return nprop.value, theta[:nprop.value]

    def hetero_update(self, l, k, j, sd=None):
        """Exact conditional draw of the mean process feeding Hetero node j (imputation.py:141-164, dense branch);
        `sd` (n x 2 standard normals) defaults to the reference's own draw."""
        kern, desc, y = self.liks[j]
        row_var = int(kern.input_dim[1])
        target = self.all_layer[l][k]
        if target.vecch:   # latent-Vecchia draw (imputation.py:141-158); the reference draws np.random.randn(n)
            if target.imp_NNarray is None:
                raise RuntimeError("dgp_b200: the mean node of a Hetero likelihood needs ord_nn(pointer=True)")
            if sd is None:
                sd = np.random.randn(self.n)               # likelihood_class.py:178
            # the node's inputs may be latent themselves: take them from the device image of the layer below
            if l > 0:
                target.input = np.ascontiguousarray(L.to_host(self.F[l - 1])[np.atleast_1d(target.input_dim)].T)
            f = kern.posterior_vecch_dev(target, self.n, self.F[l][row_var], y, sd)
        else:
            if sd is None:
                sd = np.random.randn(self.n, 2)            # likelihood_class.py:200
            f = kern.posterior_dev(self.nodes[l][k], self.n, self.F[l][row_var], y, sd)
        self.F[l][k].copy_(f)
        L.check(L.load().dgpb_cache_output_changed(L.workspace(), self._key(l, k)))

    def block_update(self, l, tks, uks, max_u=64):
        """One ESS update of the target nodes `tks` of layer l given the upper nodes `uks` of layer l+1,
        drawing from the RNG streams described in the module docstring."""
        n = self.n
        layer = self.all_layer[l]
        z = np.empty((len(tks), n))
        for i, k in enumerate(tks):
            # fmvn_sp draws from numpy's global RNG, fmvn from the (numba-role) module RNG
            z[i] = np.random.randn(n) if layer[k].vecch else _nb_rng.standard_normal(n)
        while True:
            state = np.random.get_state()
            u = np.random.uniform(size=max_u)
            try:
                if self.liks is not None and l + 2 == len(self.all_layer):
                    nprop, _ = self.lik_call(l, tks, uks, z, u)
                else:
                    nprop, _ = self.ess_call(l, tks, uks, z, u)
            except ValueError:
                if self.last_nprop + 1 < max_u:
                    raise
                np.random.set_state(state)  # ran out of uniforms: redo the same update with a longer array
                max_u *= 4
                continue
            np.random.set_state(state)
            np.random.uniform(size=1 + nprop)  # consume exactly what the reference would have
            return nprop

    SMALL_N = 64   # training points up to which the whole I-step runs in one kernel (csrc/ess_small.cu)

    def small_ok(self, block):
        """The whole I-step can run in ONE launch (`dgpb_ess_sweeps_small`): a dense GP hierarchy with block updates,
        at most 64 training points and 8 nodes per layer, one process (no shared chain)."""
        import os
        from . import parallel
        if not block or self.liks is not None or self.n > self.SMALL_N or parallel.chain() is not None:
            return False
        if os.environ.get('DGPB_ESS_SMALL', '1') == '0':
            return False
        for l, layer in enumerate(self.all_layer):
            if len(layer) > 8 or any(k.vecch or k.type != 'gp' for k in layer):
                return False
        return len(self.all_layer) <= 8

    def sweeps_small(self, sweeps, z=None, u=None, max_u=48):
        """`sweeps` block-update sweeps over all layer pairs in one kernel launch.  Draws as in `block_update`: n
        normals per target node and update from the module generator, uniforms from numpy's global generator, which
        is advanced by exactly the number the reference would have consumed.  `z` / `u` inject explicit draws
        (tests).  Returns the number of proposals evaluated."""
        n, Lg = self.n, len(self.all_layer)
        widths = np.ascontiguousarray([len(layer) for layer in self.all_layer], dtype=np.int32)
        draws = int(sweeps * widths[:-1].sum())
        if z is None:
            z = _nb_rng.standard_normal((draws, n))
        zd = L.to_dev(np.ascontiguousarray(z, dtype=np.float64))
        flat = (L.DgpbNode * int(widths.sum()))(*[nd for arr in self.nodes for nd in arr])
        ptrs = (L.c_vp * Lg)(*[L.c_vp(F.data_ptr()) for F in self.F])
        counts = np.zeros(3, dtype=np.int32)
        injected = u is not None
        updates = sweeps * (Lg - 1)
        while True:
            state = None if injected else np.random.get_state()
            uu = np.ascontiguousarray(u if injected else np.random.uniform(size=updates * max_u), dtype=np.float64)
            saved = [F.clone() for F in self.F[:-1]]
            status = L.load().dgpb_ess_sweeps_small(L.workspace(), flat, widths.ctypes.data_as(L.c_vp), Lg, ptrs, n, sweeps,
                                                    L.ptr(zd), zd.shape[0], uu.ctypes.data_as(L.c_vp), len(uu),
                                                    counts.ctypes.data_as(L.c_vp), L.stream())
            if status == L.DGPB_BAD_ARG and not injected and updates * max_u < 1 << 22:
                # ran out of uniforms: the same sweeps again from the same state with a longer block
                np.random.set_state(state)
                for F, old in zip(self.F, saved):
                    F.copy_(old)
                max_u *= 4
                continue
            if not injected:
                np.random.set_state(state)
                np.random.uniform(size=int(counts[0]))   # consume exactly what the reference would have
            L.check(status)
            self.last_uniforms = int(counts[0])
            return int(counts[1])

    def write_back(self):
        """Imputed layers -> `kernel.output` of their nodes and `kernel.input` of the nodes they feed."""
        for l in range(len(self.all_layer) - 1):
            Fl = L.to_host(self.F[l])
            for k, kern in enumerate(self.all_layer[l]):
                kern.output[:, 0] = Fl[k]
            for kern in self.all_layer[l + 1]:
                kern.input = np.ascontiguousarray(Fl[np.atleast_1d(kern.input_dim)].T)


class imputer:
    """Class to implement imputation of latent variables (imputation.py:6-20)."""

    def __init__(self, all_layer, block=True):
        self.all_layer = all_layer
        self.block = block
        self.n_proposals = 0      # instrumentation for the roofline accounting
        self.n_block_updates = 0

    def __setstate__(self, state):
        state.setdefault('block', True)
        state.setdefault('n_proposals', 0)
        state.setdefault('n_block_updates', 0)
        self.__dict__.update(state)

    def sample(self, burnin=0):
        """ESS-within-Gibbs: burnin+1 sweeps over the layer pairs (imputation.py:22-42)."""
        n_layer = len(self.all_layer)
        if n_layer < 2:
            return
        dev = _DeviceLayers(self.all_layer)
        if dev.small_ok(self.block):   # the whole I-step in one launch (n <= 64)
            self.n_proposals += dev.sweeps_small(burnin + 1)
            self.n_block_updates += (burnin + 1) * (n_layer - 1)
            dev.write_back()
            return
        for _ in range(burnin + 1):
            for l in range(n_layer - 1):
                layer, linked = self.all_layer[l], self.all_layer[l + 1]
                exact = any(kern.type == 'likelihood' and kern.exact_post_idx is not None for kern in linked)
                if exact:   # imputation.py:34-42: node-wise updates, closed-form draw where the likelihood has one
                    for k in range(len(layer)):
                        uks = [j for j, kern in enumerate(linked) if k in kern.input_dim]
                        if len(uks) == 1 and linked[uks[0]].exact_post_idx is not None:
                            idx = np.where(np.asarray(linked[uks[0]].input_dim) == k)[0]
                            if idx in linked[uks[0]].exact_post_idx:
                                dev.hetero_update(l, k, uks[0])
                                continue
                        self.n_proposals += dev.block_update(l, [k], uks)
                        self.n_block_updates += 1
                elif self.block:
                    self.n_proposals += dev.block_update(l, list(range(len(layer))), list(range(len(linked))))
                    self.n_block_updates += 1
                else:
                    for k in range(len(layer)):
                        uks = [j for j, kern in enumerate(linked) if k in kern.input_dim]
                        self.n_proposals += dev.block_update(l, [k], uks)
                        self.n_block_updates += 1
        dev.write_back()
        import os
        if os.environ.get('DGPB_CHAIN_CHECK') == '1':
            from . import parallel
            if parallel.chain() is not None:   # debug aid: the ranks of a shared chain hold identical layers
                parallel.assert_in_step([float(np.sum(k.output)) for layer in self.all_layer for k in layer
                                         if k.type == 'gp'] + [float(self.n_proposals)], "latent layers after an I-step")

    def key_stats(self):
        """Compute and store key statistics used in predictions (imputation.py:223-231)."""
        for layer in self.all_layer:
            for kernel in layer:
                if kernel.type == 'gp':
                    kernel.compute_stats()

    def update_ord_nn(self):
        """Re-draw the Vecchia ordering / neighbours of every node, sharing them between nodes of a layer
        that see the same inputs and length-scales (imputation.py:233-262)."""
        for layer in self.all_layer:
            for k, kernel in enumerate(layer):
                if kernel.type != 'gp':
                    continue
                match = None
                for j in range(k):
                    same_in = np.array_equal(kernel.input_dim, layer[j].input_dim) and np.array_equal(
                        kernel.connect, layer[j].connect)
                    if len(kernel.length) == 1:
                        ok = same_in and len(layer[j].length) == 1
                    else:
                        ok = same_in and np.array_equal(kernel.length, layer[j].length)
                    if ok:
                        match = layer[j]
                        break
                pointer = kernel.imp_pointer_row is not None      # imputation.py:241
                if match is None:
                    kernel.ord_nn(pointer=pointer)
                elif len(kernel.length) == 1:
                    kernel.ord_nn(ord=match.ord, NNarray=match.NNarray, pointer=pointer)
                else:
                    kernel.ord_nn(ord=match.ord.copy(), NNarray=match.NNarray.copy(), pointer=pointer)
