"""Ray generation (ref nerf/utils.py:52-151): oracle and CUDA kernel against the reference's own
get_rays (fixtures tests/golden/ref_python.npz, written by tests/golden/make_golden_cpu.py from the reference source).

Tolerance: directions are unit vectors rotated by a 3x3 matmul whose accumulation order belongs to the BLAS the
reference happened to run on; fp32, max-abs 4e-7 (< 4 ulp at 1.0) on rays_d, rays_o bit-exact, pixel indices bit-exact.
"""
import os

import numpy as np
import pytest

from conftest import ROOT

TOL_D = 4e-7
CASES = ["full", "rand", "patch", "pair", "err"]


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_python.npz"), allow_pickle=True)


@pytest.mark.parametrize("case", CASES)
def test_oracle_get_rays_vs_reference(gp, case):
    import oracle
    H, W = (int(v) for v in gp["rays_HW"])
    inds = gp[f"rays_{case}_inds"]
    o, d = oracle.get_rays(gp["rays_poses"], gp["rays_intrinsics"], H, W, None if case == "full" else inds)
    assert o.shape == gp[f"rays_{case}_o"].shape
    assert np.array_equal(o, gp[f"rays_{case}_o"])
    assert np.abs(d - gp[f"rays_{case}_d"]).max() < TOL_D
    assert np.abs(np.linalg.norm(d.astype(np.float64), axis=-1) - 1).max() < 1e-6


def test_get_rays_needs_cuda_tensors():
    import torch
    from palettenerf_b200.nerf.utils import get_rays
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        get_rays(torch.eye(4)[None], [100.0, 100.0, 8.0, 8.0], 16, 16, N=-1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_get_rays_vs_reference(cuda, gp, case):
    """same seeded torch RNG calls as the reference => the same pixels; then the kernel's rays vs the reference's"""
    import torch
    from palettenerf_b200.nerf.utils import get_rays, rays_from_indices
    H, W = (int(v) for v in gp["rays_HW"])
    poses = torch.from_numpy(gp["rays_poses"]).to(cuda)
    intr = gp["rays_intrinsics"]
    # (1) the kernel on the reference's own indices
    inds = torch.from_numpy(gp[f"rays_{case}_inds"]).to(cuda)
    o, d, _, _ = rays_from_indices(poses, intr, H, W, None if case == "full" else inds)
    assert np.array_equal(o.cpu().numpy(), gp[f"rays_{case}_o"])
    assert np.abs(d.cpu().numpy() - gp[f"rays_{case}_d"]).max() < TOL_D
    # (2) the drop-in entry point: signature, result keys, shapes, index ranges (the CUDA generator differs from the CPU
    # generator the fixtures were drawn with, so the indices themselves are checked structurally)
    kw = {"full": dict(N=-1), "rand": dict(N=64), "patch": dict(N=64, patch_size=4), "pair": dict(N=64, random_size=3),
          "err": dict(N=64, error_map=torch.from_numpy(gp["rays_error_map"]).to(cuda))}[case]
    torch.manual_seed(11)
    r = get_rays(poses, intr, H, W, **kw)
    n = H * W if case == "full" else 64
    assert r["rays_o"].shape == (2, n, 3) and r["rays_d"].shape == (2, n, 3) and r["inds"].shape == (2, n)
    assert int(r["inds"].min()) >= 0 and int(r["inds"].max()) < H * W
    assert ("inds_coarse" in r) == (case == "err")
    import oracle
    oo, od = oracle.get_rays(gp["rays_poses"], intr, H, W, r["inds"].cpu().numpy())
    assert np.array_equal(r["rays_o"].cpu().numpy(), oo)
    assert np.abs(r["rays_d"].cpu().numpy() - od).max() < TOL_D
    if case == "patch":   # 4x4 patches: consecutive groups of 16 indices form a patch
        p = r["inds"][0].view(-1, 4, 4).cpu().numpy()
        assert np.all(p[:, 1:, :] - p[:, :-1, :] == W) and np.all(p[:, :, 1:] - p[:, :, :-1] == 1)


@pytest.mark.gpu
def test_cuda_get_rays_full_view_with_fused_near_far(cuda):
    """800x800 view (BASELINE config 3's rays): kernel vs oracle; the fused slab test is bit-identical to
    near_far_from_aabb run on the kernel's own rays; odd ray counts exercise the unaligned store path"""
    import torch
    import oracle
    import palettenerf_b200.raymarching as rm
    from palettenerf_b200 import synthetic as S
    from palettenerf_b200.nerf.utils import get_rays, rays_from_indices
    H = W = 800
    pose = S.lookat_pose(S.LEGO["radius"], 35.0)
    import math
    f = 0.5 * W / math.tan(0.5 * S.LEGO["camera_angle_x"])
    intr = [f, f, W / 2, H / 2]
    aabb = torch.tensor([-2, -2, -2, 2, 2, 2], dtype=torch.float32, device=cuda)
    r = get_rays(pose[None].to(cuda), intr, H, W, N=-1, aabb=aabb, min_near=0.2)
    oo, od = oracle.get_rays(pose[None].numpy(), intr, H, W)
    assert np.array_equal(r["rays_o"].cpu().numpy(), oo)
    assert np.abs(r["rays_d"].cpu().numpy() - od).max() < TOL_D
    so, sd = S.camera_rays(H, W, azimuth_deg=35.0)      # the torch program the benchmark used so far
    assert np.abs(r["rays_d"][0].cpu().numpy() - sd.numpy()).max() < TOL_D
    nears, fars = rm.near_far_from_aabb(r["rays_o"][0], r["rays_d"][0], aabb, 0.2)
    assert torch.equal(nears, r["nears"][0]) and torch.equal(fars, r["fars"][0])
    # ragged: B = 3 cameras x 1001 rays (per-camera base offsets not 16-byte aligned), per-camera index rows
    poses = torch.stack([S.lookat_pose(S.LEGO["radius"], a) for a in (10.0, 130.0, 250.0)]).to(cuda)
    inds = torch.randint(0, H * W, (3, 1001), device=cuda)
    o3, d3, n3, f3 = rays_from_indices(poses, intr, H, W, inds, aabb, 0.2)
    oo, od = oracle.get_rays(poses.cpu().numpy(), intr, H, W, inds.cpu().numpy())
    assert np.array_equal(o3.cpu().numpy(), oo) and np.abs(d3.cpu().numpy() - od).max() < TOL_D
    on, of = oracle.near_far_from_aabb(o3.view(-1, 3).cpu().numpy(), d3.view(-1, 3).cpu().numpy(), aabb.cpu().numpy(), 0.2)
    assert np.array_equal(n3.view(-1).cpu().numpy(), on) and np.array_equal(f3.view(-1).cpu().numpy(), of)
    # empty
    e = rays_from_indices(poses, intr, H, W, inds[:, :0])
    assert e[0].shape == (3, 0, 3)


@pytest.mark.parametrize("case", ["rand", "patch", "pair", "err"])
def test_index_sampling_draws_the_reference_pixels(gp, case):
    """host logic (CPU): with the same torch seed the samplers make the reference's RNG calls in the reference's order, so
    they draw exactly the pixels the reference's get_rays drew when the fixtures were generated (CPU generator, seed 11)"""
    import torch
    from palettenerf_b200.nerf import utils as U
    H, W = (int(v) for v in gp["rays_HW"])
    kw = {"rand": {}, "patch": dict(patch_size=4), "pair": dict(random_size=3),
          "err": dict(error_map=torch.from_numpy(gp["rays_error_map"]))}[case]
    torch.manual_seed(11)
    inds, extra = U._draw_indices(2, H, W, 64, kw.get("error_map"), kw.get("patch_size", 1), kw.get("random_size", 0),
                                  torch.device("cpu"))
    assert np.array_equal(inds.contiguous().numpy(), gp[f"rays_{case}_inds"])
    if case == "err":
        assert np.array_equal(extra["inds_coarse"].numpy(), gp["rays_err_inds_coarse"])


@pytest.mark.gpu
def test_collate_gathers_training_pixels_in_the_ray_kernel(cuda):
    """f1, second half (ref: palette/provider.py:377-403): ground-truth colours and semantic features at the sampled pixels,
    gathered by the launch that generates the rays == torch.gather of the reference's collate"""
    import torch
    from palettenerf_b200.nerf.utils import collate
    B, H, W, N = 2, 37, 53, 500
    g = torch.Generator().manual_seed(3)
    poses = torch.eye(4).repeat(B, 1, 1)
    poses[:, :3, 3] = torch.randn(B, 3, generator=g)
    images = torch.rand(B, H, W, 4, generator=g).to(cuda)
    feats = torch.randn(B, H, W, 16, generator=g).to(cuda)
    intr = [61.7, 59.3, W / 2, H / 2]
    torch.manual_seed(5)
    res = collate(poses.to(cuda), intr, H, W, N, images=images, feat_images=feats)
    inds = res["inds"]
    assert res["images"].shape == (B, N, 4) and res["feat_images"].shape == (B, N, 16)
    want_i = torch.gather(images.view(B, -1, 4), 1, torch.stack(4 * [inds], -1))
    want_f = torch.gather(feats.view(B, -1, 16), 1, torch.stack(16 * [inds], -1))
    assert torch.equal(res["images"], want_i) and torch.equal(res["feat_images"], want_f)
    torch.manual_seed(5)
    plain = collate(poses.to(cuda), intr, H, W, N)                     # same rays without the gathers
    assert torch.equal(plain["rays_d"], res["rays_d"]) and torch.equal(plain["inds"], inds) and "images" not in plain
    ev = collate(poses.to(cuda), intr, H, W, -1, images=images, training=False)
    assert ev["images"] is images and ev["rays_o"].shape == (B, H * W, 3)
