"""pnerf_peer_allreduce (gradient all-reduce over NVLink peer memory) vs NCCL, world size 2: needs two GPUs on the box
(skipped otherwise; `gpurun --gpus 2 -- python -m pytest tests/test_peer_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.gpu
def test_peer_allreduce_matches_nccl_world2(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "peer_allreduce_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "PEER_ALLREDUCE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
