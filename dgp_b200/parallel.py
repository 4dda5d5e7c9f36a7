"""Multi-GPU plumbing: one process per GPU.

* ONE chain on several GPUs (`enable`): every rank holds the same model and draws the same random numbers; the
  candidate angles of every ESS wave are dealt over the ranks inside libdgpb.so (its own NCCL communicator,
  `dgpb_comm_init`) and the GP nodes of the M-step are dealt over the ranks here (`mstep_share` / `sync_params`,
  one all-reduce of <= 36 doubles per node).  The reference's counterpart is `ptrain`'s process pool over the
  nodes of a layer (dgp.py:1414-1472).
* Prediction shards by test points (`predict_sharded`: each point is independent, one all-gather of the moments).
"""
from __future__ import annotations

import numpy as np

_chain = None   # {"dist": torch.distributed, "rank": r, "world": G, "device": bool} once `enable` was called


def enable(dist, seed=None):
    """Share ONE chain between the ranks of the initialised default process group.  Every rank must then build
    the same model from the same data and call the same methods in the same order; `seed` (optional) seeds
    numpy's global generator and the generator behind the dense prior draws identically on every rank.  With the
    "nccl" backend the library's own communicator is created on this thread's workspace (ESS waves); with "gloo"
    (CPU tests) only the host-side exchanges are active."""
    global _chain
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        _chain = None
        return None
    import torch

    from . import _lib as L
    from .imputation import nb_seed

    rank, world = dist.get_rank(), dist.get_world_size()
    on_device = dist.get_backend() == "nccl"
    if on_device:
        lib = L.load()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (L.ctypes.c_char * 128)()
            L.check(lib.dgpb_comm_unique_id(L.ctypes.cast(buf, L.c_vp)))
            ident = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        ident = ident.to(L.device())
        dist.broadcast(ident, 0)
        raw = bytes(ident.cpu().numpy().tobytes())
        buf = L.ctypes.create_string_buffer(raw, 128)
        L.check(lib.dgpb_comm_init(L.workspace(), rank, world, L.ctypes.cast(buf, L.c_vp)))
    if seed is not None:
        np.random.seed(seed)
        nb_seed(seed)
    _chain = {"dist": dist, "rank": rank, "world": world, "device": on_device}
    return _chain


def disable():
    """Back to one independent chain per process."""
    global _chain
    if _chain is not None and _chain["device"]:
        from . import _lib as L
        L.check(L.load().dgpb_comm_destroy(L.workspace()))
    _chain = None


def chain():
    """The shared-chain state set by `enable`, or None."""
    return _chain


def mstep_share(n_nodes, rank, world, costs=None):
    """Indices of the GP nodes whose L-BFGS-B run this rank carries in an M-step (given the imputation the nodes
    are independent; the reference deals the nodes of one layer to a process pool).  Without `costs`: round-robin
    over all layers.  With `costs` (one number per node: its objective evaluations in the previous M-step, known to
    every rank): longest-processing-time-first, so the optimiser that needs the most rounds does not share its rank
    with other slow ones.  Every rank computes the same assignment."""
    if costs is None or len(costs) != n_nodes:
        return list(range(rank, n_nodes, world))
    # A rank's M-step is a sequence of rounds (one batched evaluation each): as many as its slowest optimiser needs,
    # and a round costs a fixed part plus a part per matrix (measured at n = 5000: ~6.5 ms + ~3 ms per matrix).  Greedy
    # on that estimate, longest optimiser first; ties go to the lowest rank so every rank computes the same answer.
    FIXED, PER = 6.5, 3.0
    top, tot, mine = [0.0] * world, [0.0] * world, []
    for i in sorted(range(n_nodes), key=lambda i: (-float(costs[i]), i)):
        c = max(1.0, float(costs[i]))
        r = min(range(world), key=lambda r: (FIXED * max(top[r], c) + PER * (tot[r] + c), r))
        top[r], tot[r] = max(top[r], c), tot[r] + c
        if r == rank:
            mine.append(i)
    return sorted(mine)


def _allreduce_sum(arr):
    """Element-wise sum of a float64 numpy array over the ranks of the chain (NCCL on the device, gloo on the host)."""
    import torch

    ch = _chain
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
    if ch["device"]:
        from . import _lib as L
        t = t.to(L.device())
    ch["dist"].all_reduce(t)
    return t.cpu().numpy()


_PAR_W = 5 + 32 + 32   # scale, nugget, len(length), objective evaluations, len(R2 row), length[<= 32], R2 row[<= 32]


def sync_params(kernels, mine, failed):
    """After the ranks optimised their share of the GP nodes: hand every node's (scale, length, nugget) to every
    rank.  `mine` = indices this rank optimised, `failed` = this rank hit a LinAlgError.  Rows of the other ranks'
    nodes are zero here, so one sum all-reduce moves each value unchanged.  Returns True if ANY rank failed (every
    rank then raises, which keeps dgp.train's restart logic in step)."""
    buf = np.zeros((len(kernels) + 1, _PAR_W))
    for i in mine:
        k = kernels[i]
        buf[i, 0], buf[i, 1], buf[i, 2] = k.scale[0], k.nugget[0], len(k.length)
        buf[i, 3] = getattr(k, '_nfev', 0)
        buf[i, 5:5 + len(k.length)] = k.length
        r2 = getattr(k, 'R2', None)
        if r2 is not None and getattr(k, '_r2_fresh', False):   # the row the owner appended in this M-step
            row = np.atleast_2d(r2)[-1]
            buf[i, 4] = len(row)
            buf[i, 37:37 + len(row)] = row
    buf[-1, 0] = 1.0 if failed else 0.0
    out = _allreduce_sum(buf)
    if out[-1, 0] > 0:
        return True
    mine = set(mine)
    for i, k in enumerate(kernels):
        if i in mine:
            continue
        nl = int(round(out[i, 2]))
        k.scale = np.array([out[i, 0]])
        k.nugget = np.array([out[i, 1]])
        k.length = out[i, 5:5 + nl].copy()
        k._nfev = int(round(out[i, 3]))
        nr = int(round(out[i, 4]))
        if nr:   # R2 of the regression on the global input (kernel.r2): computed by the owner, appended here
            row = out[i, 37:37 + nr].copy()
            k.R2 = np.atleast_2d(row) if k.R2 is None else np.vstack((k.R2, row))
        k.add_to_path()
    return False


def assert_in_step(values, what="chain state"):
    """Debug aid (DGPB_CHAIN_CHECK=1): the ranks of a shared chain must hold identical numbers."""
    v = np.atleast_1d(np.asarray(values, dtype=np.float64)).ravel()
    G, r = _chain["world"], _chain["rank"]
    buf = np.zeros((G, len(v)))
    buf[r] = v
    out = _allreduce_sum(buf)
    if not all(np.array_equal(out[0], out[g], equal_nan=True) for g in range(1, G)):
        raise RuntimeError(f"dgp_b200: the ranks of the shared chain diverged ({what}): {out.tolist()}")


def shard_bounds(M, rank, world):
    """Contiguous row range [lo, hi) of rank `rank` out of `world` (np.array_split boundaries, the same split the
    reference's ppredict uses, emulation.py:607)."""
    base, extra = divmod(M, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def predict_sharded(emu, x, dist=None, **kwargs):
    """`emu.predict(x, method='mean_var')` with the rows of `x` sharded over the ranks of the default process
    group; every rank returns the full (mu, sigma2).  `emu` is an `emulator` or an `lgp` (whose predict returns one
    array per emulator of the last layer); `dist` is the initialised torch.distributed module or None (single
    process).  With the "nccl" backend the shard's moments stay on the device from the prediction kernels to ONE
    all-gather and come to the host once; with "gloo" (CPU tests) the same logic runs on host tensors."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return emu.predict(x, **kwargs)
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    M = x.shape[0]
    lo, hi = shard_bounds(M, rank, world)
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    try:
        mu, var = emu.predict(x[lo:hi], _device=True, **kwargs) if use_cuda else emu.predict(x[lo:hi], **kwargs)
    except TypeError:   # a predictor without the device-resident return (lgp, stand-ins of the tests)
        mu, var = emu.predict(x[lo:hi], **kwargs)
    as_list = isinstance(mu, (list, tuple))
    mus, vars_ = (list(mu), list(var)) if as_list else ([mu], [var])

    def tens(a):
        return a.to(dev) if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    parts = [tens(a).reshape(hi - lo, -1) for a in mus + vars_]
    widths = [p.shape[1] for p in parts]
    maxrows = shard_bounds(M, 0, world)[1]  # the first shard is the largest
    buf = torch.zeros((maxrows, sum(widths)), dtype=torch.float64, device=dev)
    if hi > lo:
        buf[: hi - lo] = torch.cat(parts, 1)
    out = torch.empty((world * maxrows, sum(widths)), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, maxrows, sum(widths))
    full = np.concatenate([out[r, : shard_bounds(M, r, world)[1] - shard_bounds(M, r, world)[0]] for r in range(world)], 0)
    cols = np.cumsum([0] + widths)
    pieces = [np.ascontiguousarray(full[:, cols[i]:cols[i + 1]]) for i in range(len(widths))]
    k = len(mus)
    if as_list:
        return pieces[:k], pieces[k:]
    return pieces[0], pieces[1]
