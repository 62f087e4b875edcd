"""Host-side multi-GPU logic on the CPU: world_size-2 gloo process groups (the NCCL path on the box runs the same code).
Covers ray sharding (bands and interleaved tiles), the single flat gradient all-reduce with the found-inf flag riding in
the bucket and unused parameters excluded, and the sharded density-grid merge."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from palettenerf_b200 import distributed as D


def test_ray_shard_partitions_every_ray_exactly_once():
    for n, ws, tile in ((640000, 8, 0), (1089480, 8, 0), (1000, 3, 0), (5, 8, 0), (640000, 8, 1024), (1001, 4, 32), (0, 2, 0)):
        seen = torch.zeros(n, dtype=torch.int32)
        for r in range(ws):
            s = D.ray_shard(n, ws, r, tile)
            seen[s] += 1
        assert bool((seen == 1).all()), (n, ws, tile)
    with pytest.raises(ValueError):
        D.ray_shard(10, 2, 2)


def test_cell_shard_covers_the_grid():
    for n, ws in ((128 ** 3, 8), (128 ** 3, 3), (7, 8)):
        spans = [D.cell_shard(n, ws, r) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)
        # a model with one parameter that never receives a gradient (like the sigma grid in the palette stage)
        used = [torch.nn.Parameter(torch.randn(1000, 2)), torch.nn.Parameter(torch.randn(64, 31))]
        unused = torch.nn.Parameter(torch.randn(10))
        frozen = torch.nn.Parameter(torch.randn(4, 3), requires_grad=False)
        x = torch.full((1000, 2), float(rank + 1))
        loss = (used[0] * x).sum() + (used[1] * (rank + 1)).sum()
        loss.backward()
        bucket = D.GradBucket(used + [unused, frozen])
        flag = bucket.all_reduce(found_inf=1.0 if rank == 1 else 0.0)
        ok = bucket.signature() == [(1000, 2), (64, 31)]
        ok &= torch.allclose(used[0].grad, torch.full((1000, 2), (1 + ws) / 2))       # mean of rank+1 over ranks
        ok &= torch.allclose(used[1].grad, torch.full((64, 31), (1 + ws) / 2))
        ok &= unused.grad is None and float(flag) == 1.0                               # rank 1's inf is seen everywhere
        flag2 = bucket.all_reduce(found_inf=0.0, average=False)
        ok &= float(flag2) == 0.0 and torch.allclose(used[0].grad, torch.full((1000, 2), (1 + ws) / 2 * ws))

        # inference: each rank "renders" its shard; the gathered map equals the single-process result
        n = 1001
        truth = torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 3)
        for tile in (0, 64):
            sh = D.ray_shard(n, ws, rank, tile)
            full = D.gather_maps(truth[sh].clone(), n, sh, tile=tile)
            ok &= torch.equal(full, truth)

        # density refresh: ranks evaluate disjoint cell ranges, merged grid identical everywhere
        cells = 4096
        fresh = -torch.ones(2, cells)
        lo, hi = D.cell_shard(cells, ws, rank)
        fresh[:, lo:hi] = torch.arange(lo, hi, dtype=torch.float32)
        merged = D.merge_density(fresh)
        ok &= torch.equal(merged, torch.arange(cells, dtype=torch.float32)[None].repeat(2, 1))
        seed = D.shared_seed()
        seeds = [None] * ws
        dist.all_gather_object(seeds, seed)
        ok &= len(set(seeds)) == 1
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_data_parallel_bucket_sharded_render_and_density_merge_gloo_world2():
    ws = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(ws, port, ret), nprocs=ws, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_grad_bucket_single_process_semantics():
    """world size 1 (no process group): pack once, gradients re-pointed at the bucket, flag forms, repeated calls"""
    import torch
    from palettenerf_b200.distributed import GradBucket
    ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
    ps[2].requires_grad_(False)
    b = GradBucket(ps)
    for step in range(3):
        ps[0].grad, ps[1].grad = torch.full((5, 3), 1.0 + step), torch.arange(7.0) * (step + 1)
        want0, want1 = ps[0].grad.clone(), ps[1].grad.clone()
        flag = b.all_reduce(found_inf=torch.tensor(float(step == 1)), average=True)
        assert float(flag) == float(step == 1)
        assert torch.equal(ps[0].grad, want0) and torch.equal(ps[1].grad, want1)
        assert ps[0].grad.shape == (5, 3) and ps[0].grad.data_ptr() == b.flat.data_ptr()            # a view of the bucket
        assert ps[1].grad.data_ptr() == b.flat[15:].data_ptr() and b.flat.numel() == 15 + 7 + 1
    assert b.signature() == [(5, 3), (7,)]
    # a parameter that stops receiving gradients changes the layout: the bucket is rebuilt
    ps[1].grad = None
    ps[0].grad = torch.ones(5, 3)
    b.all_reduce()
    assert b.flat.numel() == 16 and b.signature() == [(5, 3)]


def test_grad_bucket_puts_large_tensors_first_and_hands_out_no_slots_without_peer_memory():
    """layout of the bucket: tensors of >= 2^20 elements (the hash tables) in front of the small ones — the region the peer
    path all-reduces early. On the CPU / NCCL path there is no peer memory: slot() must return None and early() must do
    nothing, so that the fused backward keeps its ordinary gradient tensors."""
    import torch
    from palettenerf_b200.distributed import GradBucket
    small = torch.nn.Parameter(torch.zeros(64, 3))
    big = torch.nn.Parameter(torch.zeros(GradBucket.BIG // 2 + 8, 2))
    tail = torch.nn.Parameter(torch.zeros(9))
    b = GradBucket([small, big, tail])
    assert b.slot(big) is None                                            # nothing laid out yet
    small.grad, big.grad, tail.grad = torch.ones(64, 3), torch.full(big.shape, 2.0), torch.arange(9.0)
    flag = b.all_reduce(average=True)
    assert float(flag) == 0.0
    assert b.signature() == [tuple(big.shape), (64, 3), (9,)]
    assert big.grad.data_ptr() == b.flat.data_ptr()                       # the large tensor leads the bucket
    assert small.grad.data_ptr() == b.flat[big.numel():].data_ptr()
    assert torch.equal(big.grad, torch.full(big.shape, 2.0)) and torch.equal(tail.grad, torch.arange(9.0))
    assert b.slot(big) is None and b.early([big]) is False and b.early_count == 0
    # a gradient that already lives in its view is not copied again
    big.grad.mul_(3.0)
    b.all_reduce(average=True)
    assert torch.equal(big.grad, torch.full(big.shape, 6.0))
