"""Worker of tests/test_peer_gpu.py (run under torchrun, one rank per GPU): pnerf_peer_allreduce through GradBucket(peer=True)
against dist.all_reduce (NCCL) on the same gradients; eager and replayed from a CUDA graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from palettenerf_b200.distributed import GradBucket  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    shapes = [(1 << 20, 2), (64, 32), (16, 64), (4, 3), (1000003,)]          # odd sizes: padding path
    params = [torch.nn.Parameter(torch.zeros(*s, device=dev)) for s in shapes]

    def fill(step):
        for p in params:
            p.grad = torch.randn(p.shape, device=dev, generator=g) * (1.0 + step)

    peer, nccl = GradBucket(params, peer=True), GradBucket(params, peer=False)
    for step in range(3):
        fill(step)
        mine = [p.grad.clone() for p in params]
        flag = peer.all_reduce(found_inf=torch.tensor(1.0 if (step == 1 and rank == world - 1) else 0.0, device=dev), average=True)
        got = [p.grad.clone() for p in params]
        for p, m in zip(params, mine):
            p.grad = m.clone()
        nccl.all_reduce(average=True)
        for a, p in zip(got, params):
            err = (a - p.grad).abs().max().item()
            assert err <= 1e-5 * (1.0 + step), f"rank {rank} step {step}: peer vs nccl {err}"
        assert (flag.item() != 0.0) == (step == 1), "found-inf flag must reach every rank"
        # bit-identical on every rank (one owner per element)
        chk = torch.stack([a.double().sum() for a in got])
        ref = chk.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(chk, ref), "ranks hold different reduced gradients"
    assert peer._pm is not None, "the peer path must be the one that ran"

    # CUDA graph: pack + barrier + kernel + barrier captured; replays follow new gradients written to the same buffers
    static = [torch.zeros_like(p) for p in params]
    for p, s in zip(params, static):
        p.grad = s
    gb = GradBucket(params, peer=True)
    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        for p, s in zip(params, static):
            p.grad = s
        gb.all_reduce(average=True)
    torch.cuda.current_stream().wait_stream(s_)
    for p, s in zip(params, static):
        p.grad = s
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        gb.all_reduce(average=True)
    outs = [p.grad for p in params]                     # views of the bucket
    for it in range(3):
        for s in static:
            s.copy_(torch.randn(s.shape, device=dev, generator=g))
        exp = [s.clone() for s in static]
        for e in exp:
            dist.all_reduce(e)
            e.div_(world)
        graph.replay()
        torch.cuda.synchronize()
        for o, e in zip(outs, exp):
            assert (o - e).abs().max().item() <= 1e-5, f"rank {rank}: graph replay {it}"
    dist.barrier()
    if rank == 0:
        print("PEER_ALLREDUCE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
