"""-m gpu, needs >= 2 GPUs (``gpurun --gpus 2``): a 2-rank NCCL run through the CUDA path reproduces the 1-GPU loss and
gradients, for segment-row sharding and for sequence sharding (tests/mp_nccl_worker.py)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_run_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "mp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
