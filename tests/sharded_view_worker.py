"""Worker of tests/test_peer_gpu.py (torchrun, one rank per GPU): ONE view rendered by all ranks through
palettenerf_b200.distributed.ShardedView — every rank's persistent kernel stores its rays straight into rank 0's image over
NVLink peer memory — must equal the single-GPU render of the same view BIT FOR BIT, for every output map."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from palettenerf_b200 import synthetic as S  # noqa: E402
from palettenerf_b200.distributed import ShardedView  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    side = 200
    for clip, gui in ((False, False), (True, False), (False, True)):
        model = S.build_palette_model(dev, seed=5, pred_clip=clip, table_scale=0.5)
        model.eval()
        model.density_scale = 25.0            # rays terminate: both the early-out and the full-length path are exercised
        o, d = S.camera_rays(side, side, azimuth_deg=70.0)
        o, d = o.to(dev), d.to(dev)
        view = ShardedView(model, side * side, gui_mode=gui, tile=64)
        got = None
        for _ in range(2):                    # twice: the owner's maps are re-zeroed between views
            got = view.render(o, d, bg_color=1)
        torch.cuda.synchronize()
        if rank == 0:
            import os as _os
            _os.environ["PNERF_RENDER_KERNEL"] = "tc"
            # bit for bit against the REPRODUCIBLE single-GPU render (every ray's windows start at its own first sample, which
            # is also what a sharded view uses); the default render shares windows between queue neighbours and agrees to
            # the fp32 re-association of the per-ray sums
            model.fused_reproducible = True
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                ref = model.render(o[None], d[None], staged=True, bg_color=1, perturb=False, gui_mode=gui)
            model.fused_reproducible = False
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                fast = model.render(o[None], d[None], staged=True, bg_color=1, perturb=False, gui_mode=gui)
            for k, v in got.items():
                err = (v.reshape(-1) - fast[k].float().reshape(-1)).abs().max().item()
                assert err <= 2e-5 * max(1.0, v.abs().max().item()), f"clip={clip} gui={gui} {k}: vs the default render {err:.3e}"
            for k, v in got.items():
                a, b = v.reshape(-1), ref[k].float().reshape(-1)
                assert torch.equal(a, b), f"clip={clip} gui={gui} {k}: sharded view differs from the 1-GPU view " \
                                          f"(max abs {(a - b).abs().max().item():.3e})"
            assert got["weights_sum"].max().item() > 0.9
        dist.barrier()
    if rank == 0:
        print("SHARDED_VIEW_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
