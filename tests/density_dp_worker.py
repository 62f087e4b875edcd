"""Worker of tests/test_peer_gpu.py (torchrun, one rank per GPU): the data-parallel density-grid refresh. Ranks evaluate
disjoint tiles of the sweep with a shared seed, merge the temporary grid with one all-reduce(max), and must end with
IDENTICAL density grids and bitfields — equal to what one rank computes alone with the same seed."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from palettenerf_b200 import fused_nerf, synthetic as S  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    for it0 in (0, 16):                                    # full sweep, then the partial refresh
        m = S.build_nerf_model(dev, seed=9, table_scale=0.5)
        m.iter_density = it0
        m.update_extra_state()                              # world 2: tiles split over the ranks, shared seed
        assert m._last_update_schedule == "fused"
        grids = [torch.empty_like(m.density_grid) for _ in range(world)]
        bits = [torch.empty_like(m.density_bitfield) for _ in range(world)]
        dist.all_gather(grids, m.density_grid.contiguous())
        dist.all_gather(bits, m.density_bitfield.contiguous())
        for r in range(1, world):
            assert torch.equal(grids[0], grids[r]) and torch.equal(bits[0], bits[r]), f"ranks disagree (iter_density {it0})"
        # the same refresh done by ONE rank alone with the seed the ranks shared: bit-identical
        seed = torch.tensor([0], dtype=torch.int64, device=dev)
        if rank == 0:
            seed[0] = 123456789
        dist.broadcast(seed, src=0)
        a = S.build_nerf_model(dev, seed=9, table_scale=0.5)
        a.iter_density = it0
        fused_nerf.update_density_grid(a, seed=int(seed.item()))          # sharded (process group is up)
        b = S.build_nerf_model(dev, seed=9, table_scale=0.5)
        b.iter_density = it0
        import palettenerf_b200.distributed as D
        real_world = D.world
        D.world = lambda: (1, 0)                            # pretend to be alone: this rank sweeps every tile itself
        try:
            fused_nerf.update_density_grid(b, seed=int(seed.item()))
        finally:
            D.world = real_world
        assert torch.equal(a.density_grid, b.density_grid) and torch.equal(a.density_bitfield, b.density_bitfield), \
            f"sharded refresh differs from the single-rank refresh (iter_density {it0})"
        assert b.density_bitfield.count_nonzero().item() > 100
    dist.barrier()
    if rank == 0:
        print("DENSITY_DP_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
