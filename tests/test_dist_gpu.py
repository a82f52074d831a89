"""Multi-rank GPU test of the one collective of the hot path: FusedActorTrainer under torchrun with
NCCL on 2 GPUs -- the real kernels, the overlapped bucketed all-reduce, CUDA-graph capture of the collective and the
process-group teardown.  Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_fused_trainer_two_ranks_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dist_trainer_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + "\n" + res.stderr[-6000:]
    assert res.stdout.count(": ok") == 2
