"""Multi-GPU plumbing: one process per GPU, `torch.distributed` (NCCL over NVLink) for the only exchange the
path has -- gathering the per-shard predictive moments.  The prediction of a test point
depends on no other test point, so rows are split across ranks with no collective on the data path.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(M, rank, world):
    """Contiguous row range [lo, hi) of rank `rank` out of `world` (np.array_split boundaries, the same split the
    reference's ppredict uses, emulation.py:607)."""
    base, extra = divmod(M, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def predict_sharded(emu, x, dist=None, **kwargs):
    """`emu.predict(x, method='mean_var')` with the rows of `x` sharded over the ranks of the default process
    group; every rank returns the full (mu, sigma2).  `dist` is the initialised torch.distributed module or
    None (single process).  Works with the `gloo` backend on CPU tensors (tests) and `nccl` on CUDA tensors."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return emu.predict(x, **kwargs)
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    M = x.shape[0]
    lo, hi = shard_bounds(M, rank, world)
    mu, var = emu.predict(x[lo:hi], **kwargs)
    D = mu.shape[1]
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    maxrows = shard_bounds(M, 0, world)[1]  # the first shard is the largest
    buf = torch.zeros((maxrows, 2 * D), dtype=torch.float64, device=dev)
    buf[: hi - lo, :D] = torch.from_numpy(np.ascontiguousarray(mu)).to(dev)
    buf[: hi - lo, D:] = torch.from_numpy(np.ascontiguousarray(var)).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    mus, vars_ = [], []
    for r, p in enumerate(parts):
        a, b = shard_bounds(M, r, world)
        p = p[: b - a].cpu().numpy()
        mus.append(p[:, :D])
        vars_.append(p[:, D:])
    return np.concatenate(mus, 0), np.concatenate(vars_, 0)
