"""Worker of tests/test_multigpu.py: one rank of a torchrun job (or a plain single process) that
  (a) replays the reference's ESS sweeps from tests/golden/ess_replay.npz, and
  (b) trains a small dense DGP for a few SEM iterations,
with the chain shared between the ranks (dgp_b200.parallel.enable).  Rank 0 writes what it ended up with; the test
compares the multi-rank run with the single-process run bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path):
    import torch

    import dgp_b200 as D
    from dgp_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        parallel.enable(dist)
        assert parallel.chain()["world"] == world
    os.environ["DGPB_CHAIN_CHECK"] = "1"
    res = {}

    # ---- (a) the reference's ESS sweeps, replayed with its own draws (decisions must be the reference's)
    from conftest import load_golden
    from test_gpu_parity import _ess_replay
    from dgp_b200.imputation import _DeviceLayers
    _ess_replay(load_golden("ess_replay"), _DeviceLayers)
    res["replay_ok"] = np.array(1)

    # ---- (b) a few SEM iterations of a small 3-layer DGP
    seed = 1234
    np.random.seed(seed)
    D.nb_seed(seed)
    rng = np.random.default_rng(seed)
    n = 384
    X = rng.uniform(0, 1, (n, 3))
    Y = np.stack([np.sin(3 * X.sum(1)), X[:, 0] * X[:, 1] - X[:, 2] ** 2], 1)
    l1 = [D.kernel(length=np.array([1.0]), name="sexp") for _ in range(3)]
    l2 = [D.kernel(length=np.array([1.0]), name="sexp", connect=np.arange(3)) for _ in range(3)]
    l3 = [D.kernel(length=np.array([1.0]), name="sexp", scale_est=True, connect=np.arange(3)) for _ in range(2)]
    model = D.dgp(X, Y, [l1, l2, l3])
    model.train(3, disable=True)
    for l, layer in enumerate(model.all_layer):
        for k, node in enumerate(layer):
            res[f"path_{l}_{k}"] = node.para_path
            res[f"out_{l}_{k}"] = node.output
    res["nprop"] = np.array(model.imp.n_proposals)
    res["rng"] = np.random.uniform(size=3)   # the generators must be in step too
    if rank == 0:
        np.savez(out_path, **res)
    if dist is not None:
        dist.barrier()
        parallel.disable()
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
