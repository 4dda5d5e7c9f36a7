"""One chain on two GPUs: the candidate angles of every ESS wave are dealt over the ranks (libdgpb's
own NCCL communicator) and the GP nodes of the M-step over the ranks (dgp_b200/parallel.py).  The shared chain must
reproduce the single-GPU chain BIT FOR BIT -- same angles, same uniforms consumed, same latent layers, same
hyper-parameter paths -- and replay the reference's ESS decisions from the golden fixture on every rank.
Needs a box with >= 2 GPUs (`gpurun --gpus 2`); skipped elsewhere."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_chain_worker.py")


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2])
def test_shared_chain_equals_single_gpu_chain(tmp_path, world):
    one, many = str(tmp_path / "one.npz"), str(tmp_path / "many.npz")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, WORKER, one], check=True, env=env, timeout=900)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                    "--master-addr", "127.0.0.1", "--master-port", "29533", WORKER, many], check=True, env=env,
                   timeout=900)
    a, b = np.load(one), np.load(many)
    assert sorted(a.files) == sorted(b.files)
    for key in a.files:
        assert np.array_equal(a[key], b[key]), key
