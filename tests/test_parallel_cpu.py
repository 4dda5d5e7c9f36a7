"""World-size-2 `gloo` test of the only multi-GPU exchange on the path: test points are sharded over ranks and
the predictive moments all-gathered (dgp_b200/parallel.py).  Runs on CPU with a stand-in predictor so that the
host-side sharding / gather logic is covered without a GPU."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402

from dgp_b200.parallel import predict_sharded, shard_bounds  # noqa: E402


class _FakeEmulator:
    def predict(self, x, **kw):
        return np.stack([x.sum(1), x[:, 0] * 2], 1), np.stack([np.abs(x[:, 0]), x.var(1)], 1)


class _FakeLinkedSystem:
    """`lgp.predict` returns one array per emulator of the last layer."""

    def predict(self, x, **kw):
        return [x.sum(1, keepdims=True), x[:, :2] * 3], [np.abs(x[:, :1]), x[:, :2] ** 2]


def _worker(rank, world, port, M, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(0).uniform(size=(M, 3))
    mu, var = predict_sharded(_FakeEmulator(), x, dist)
    np.save(os.path.join(out, f"mu{rank}.npy"), mu)
    np.save(os.path.join(out, f"var{rank}.npy"), var)
    mus, vars_ = predict_sharded(_FakeLinkedSystem(), x, dist)
    assert isinstance(mus, list) and len(mus) == 2 and len(vars_) == 2
    np.save(os.path.join(out, f"lmu{rank}.npy"), np.concatenate(mus, 1))
    np.save(os.path.join(out, f"lvar{rank}.npy"), np.concatenate(vars_, 1))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for M in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(M, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == M
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("M", [11, 64])
def test_sharded_predict_matches_single_process(tmp_path, M):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, M, str(tmp_path)), nprocs=2, join=True)
    x = np.random.default_rng(0).uniform(size=(M, 3))
    mu, var = _FakeEmulator().predict(x)
    lmu, lvar = _FakeLinkedSystem().predict(x)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"mu{r}.npy"), mu)
        assert np.array_equal(np.load(tmp_path / f"var{r}.npy"), var)
        assert np.array_equal(np.load(tmp_path / f"lmu{r}.npy"), np.concatenate(lmu, 1))
        assert np.array_equal(np.load(tmp_path / f"lvar{r}.npy"), np.concatenate(lvar, 1))


# ---- one chain on several ranks: the host-side exchanges (M-step shares, parameter hand-over, wave plan) -----------
class _FakeKernel:
    def __init__(self, i):
        self.scale, self.nugget = np.array([1.0]), np.array([1e-6])
        self.length = np.ones(1 + i % 3)
        self.path = []
        self.R2 = None

    def add_to_path(self):
        self.path.append(np.concatenate((self.scale, self.length, self.nugget)))


def _chain_worker(rank, world, port, out, fail_rank):
    import torch.distributed as dist

    from dgp_b200 import parallel

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ch = parallel.enable(dist, seed=7)
    assert ch["world"] == world and not ch["device"]
    kernels = [_FakeKernel(i) for i in range(7)]
    mine = parallel.mstep_share(len(kernels), rank, world)
    for i in mine:   # "optimise" my share
        kernels[i].scale = np.array([10.0 + i])
        kernels[i].length = np.arange(1, 2 + i % 3) * (0.5 + i)
        kernels[i].nugget = np.array([1e-3 * (i + 1)])
        kernels[i].add_to_path()
    failed = parallel.sync_params(kernels, mine, failed=(rank == fail_rank))
    parallel.assert_in_step([k.scale[0] for k in kernels] if not failed else [0.0])
    np.save(os.path.join(out, f"chain{rank}.npy"),
            np.concatenate([[float(failed)], np.random.uniform(size=2)] + [k.path[-1] for k in kernels if k.path]))
    parallel.disable()
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_rank", [-1, 1])
def test_mstep_shares_and_parameter_handover(tmp_path, fail_rank):
    from dgp_b200.parallel import mstep_share

    for n_nodes in (1, 7, 18):
        for world in (1, 2, 8):
            shares = [mstep_share(n_nodes, r, world) for r in range(world)]
            assert sorted(i for s in shares for i in s) == list(range(n_nodes))
            assert max(map(len, shares)) - min(map(len, shares)) <= 1
            # with the evaluation counts of the previous M-step: every node exactly once, loads within one node's cost
            costs = [3 + (7 * i) % 11 for i in range(n_nodes)]
            shares = [mstep_share(n_nodes, r, world, costs) for r in range(world)]
            assert sorted(i for s in shares for i in s) == list(range(n_nodes))
            est = [6.5 * max([costs[i] for i in s] or [0]) + 3.0 * sum(costs[i] for i in s) for s in shares]
            # list-scheduling bound under the cost model: slowest optimiser's rounds + (average + one node) of matrices
            assert max(est) <= 6.5 * max(costs) + 3.0 * (sum(costs) / world + max(costs)) + 1e-9
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_chain_worker, args=(2, port, str(tmp_path), fail_rank), nprocs=2, join=True)
    a, b = np.load(tmp_path / "chain0.npy"), np.load(tmp_path / "chain1.npy")
    assert np.array_equal(a[:3], b[:3])  # same verdict, same random stream on both ranks
    if fail_rank >= 0:
        # a failure on one rank is a failure on all; nothing was handed over (each rank only has its own share)
        assert a[0] == 1.0 and len(a) < 3 + 7 * 3 and len(b) < 3 + 7 * 3
    else:
        assert np.array_equal(a, b)      # same parameters everywhere
        assert a[0] == 0.0
        want = np.concatenate([np.concatenate(([10.0 + i], np.arange(1, 2 + i % 3) * (0.5 + i), [1e-3 * (i + 1)]))
                               for i in range(7)])
        assert np.array_equal(a[3:], want)


def test_wave_plan_follows_the_bracket_rule():
    """dgpb_ess_plan_wave (host-only entry of libdgpb.so, the same routine the device loop uses): the angles of a wave
    are those the reference's shrink rule (imputation.py:111-119) would draw one rejection at a time, and the items
    are dealt round-robin over the ranks."""
    import ctypes

    from dgp_b200 import _lib as L

    lib = L.load()
    rng = np.random.default_rng(3)
    for world, cap, first in ((1, 8, 0), (2, 4, 1), (8, 1, 0), (8, 4, 1), (3, 2, 0)):
        u = rng.uniform(size=40)
        theta0 = 2 * np.pi * rng.uniform()
        tmin, tmax = theta0 - 2 * np.pi, theta0
        for nu_left in (0, 1, 5, 40):
            thetas = np.zeros(64)
            ranks, slots = np.zeros(65, np.int32), np.zeros(65, np.int32)
            S = ctypes.c_int(0)
            L.check(lib.dgpb_ess_plan_wave(theta0, tmin, tmax, u.ctypes.data_as(L.c_vp), nu_left, first, cap, world,
                                           ctypes.byref(S), thetas.ctypes.data_as(L.c_vp),
                                           ranks.ctypes.data_as(L.c_vp), slots.ctypes.data_as(L.c_vp)))
            S = S.value
            assert S == max(1, min(cap * world, 1 + nu_left))
            th, lo, hi, want = theta0, tmin, tmax, []
            for s in range(S):   # the reference loop under "every proposal rejected"
                want.append(th)
                if th < 0:
                    lo = th
                else:
                    hi = th
                th = lo + (hi - lo) * u[s] if s < len(u) else th
            assert np.array_equal(thetas[:S], want)
            items = first + S
            assert np.array_equal(ranks[:items], np.arange(items) % world)
            assert np.array_equal(slots[:items], np.arange(items) // world)
            # every rank's slots are 0..L-1 without gaps and fit cap (+1 for the threshold item on rank 0)
            for r in range(world):
                mine = slots[:items][ranks[:items] == r]
                assert np.array_equal(mine, np.arange(len(mine))) and len(mine) <= cap + (1 if r == 0 else 0)
