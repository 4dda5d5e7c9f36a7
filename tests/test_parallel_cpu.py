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


def _worker(rank, world, port, M, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = np.random.default_rng(0).uniform(size=(M, 3))
    mu, var = predict_sharded(_FakeEmulator(), x, dist)
    np.save(os.path.join(out, f"mu{rank}.npy"), mu)
    np.save(os.path.join(out, f"var{rank}.npy"), var)
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
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"mu{r}.npy"), mu)
        assert np.array_equal(np.load(tmp_path / f"var{r}.npy"), var)
