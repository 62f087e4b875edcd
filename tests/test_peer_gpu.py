"""Multi-GPU kernels over NVLink peer memory, world size 2 — needs two GPUs on the box (skipped otherwise;
`gpurun --gpus 2 -- python -m pytest tests/test_peer_gpu.py -m gpu`):
  * pnerf_peer_allreduce (gradient all-reduce of the DP training step) vs NCCL;
  * ShardedView: one view split in interleaved tiles over the ranks, every rank's renderer storing its rays straight into
    rank 0's image — bit-identical to the single-GPU view;
  * the data-parallel density-grid refresh: ranks sweep disjoint tiles with a shared seed, merge with one all-reduce(max)
    and end with identical grids and bitfields, equal to the single-rank refresh;
  * the data-parallel training step: table gradients scattered into the bucket, their all-reduce on a side stream under
    the weight-gradient kernel — same parameters after three steps as with the NCCL bucket, eager, graphed, smooth loss."""
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


@pytest.mark.gpu
def test_sharded_view_equals_single_gpu_view_bit_for_bit_world2(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "sharded_view_worker.py")]
    env = dict(os.environ)
    env.pop("PNERF_RENDER_KERNEL", None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and "SHARDED_VIEW_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.gpu
def test_data_parallel_density_refresh_ranks_agree_world2(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29553", os.path.join(ROOT, "tests", "density_dp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DENSITY_DP_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.gpu
def test_data_parallel_train_step_with_early_table_allreduce_world2(cuda):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29563", os.path.join(ROOT, "tests", "dp_train_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DP_TRAIN_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]


@pytest.mark.gpu
def test_data_parallel_stage1_train_step_with_early_table_allreduce_world2(cuda):
    """the same for the stage-1 (NeRF) model: its fused backward scatters the density table's gradient into the bucket and
    starts that region's all-reduce under the weight-gradient kernel (csrc/nerf_train.cu, fused_nerf_train.py)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29564", os.path.join(ROOT, "tests", "dp_train_worker.py"), "--stage-nerf"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DP_TRAIN_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
