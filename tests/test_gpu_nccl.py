"""The NCCL transport of the sharded build, one process per GPU: needs a box with at least two GPUs
(`gpurun --gpus 2`); skipped on the single-GPU test box, where tests/test_gpu_sharded.py covers the same exchange
logic through the in-process transport."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("direct", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_build_over_nccl_equals_one_gpu(world, direct):
    """direct = 1: bulk exchange through mapped peer memory (default); 0: everything through NCCL messages."""
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "nccl_worker.py")]
    env = dict(os.environ, ESA_COMM_DIRECT=direct)
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert pr.returncode == 0, pr.stdout[-3000:] + pr.stderr[-3000:]
    assert f"NCCL-SHARDED-OK world={world}" in pr.stdout
