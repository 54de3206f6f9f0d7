"""The sharded provers INSIDE the library (ripp_comm_*, ripp_*_sharded_dev: NCCL over NVLink, one process per GPU)
against the single-GPU entry points, byte for byte.  Needs at least two GPUs (NCCL refuses two ranks on one device):
skipped on a one-GPU box, run with `gpurun --gpus 2` (profiles/ keeps the log)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_library_sharded_provers_match_single_gpu(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + world), os.path.join(ROOT, "tests", "mgpu", "worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MGPU_OK world=%d" % world in out.stdout
