"""Stage-1 model and density-grid refresh on the tensor-core kernels (csrc/field_tc.cu model_kind = 1, csrc/density_tc.cu).

  * fused_nerf.density vs NeRFNetwork.density (torch fp32 on the stand-alone kernels): relative 1e-2 on sigma (fp16 MLP);
  * the persistent renderer vs the host-loop schedule of NeRFRenderer.run_cuda (1e-3) — the reference-pinned comparison of the
    same path is tests/test_golden_palette_gpu.py::test_nerf_stage_forward_render_and_train;
  * update_extra_state: the three-kernel refresh against a torch restatement of nerf/renderer.py:476-553 fed the SAME jitter
    (full sweep), and the invariants of the partial refresh (bitfield == packbits(grid, threshold), threshold == min(mean,
    density_thresh), only legitimate cells touched, no host synchronisation besides mean_count)."""
import numpy as np
import pytest
import torch

import oracle
from palettenerf_b200 import fused_nerf, synthetic as S
import palettenerf_b200.raymarching as rm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nerf(cuda):
    m = S.build_nerf_model(cuda, seed=4, table_scale=0.5)
    m.eval()
    return m


def test_density_tc_matches_torch_density(cuda, nerf):
    g = torch.Generator(device=cuda).manual_seed(0)
    x = (torch.rand(100003, 3, device=cuda, generator=g) * 2 - 1) * nerf.bound
    x[:7] = torch.tensor([nerf.bound, -nerf.bound, 0.0], device=cuda)          # on the boundary: still in range
    with torch.no_grad():
        ref = nerf.density(x)["sigma"].float()
    got = fused_nerf.density(nerf, x)
    rel = ((got - ref).abs() / ref.abs().clamp(min=1e-3)).max().item()
    assert rel < 1e-2, rel
    assert ref.std().item() > 1e-2
    pm = S.build_palette_model(cuda, seed=2, table_scale=0.5)                   # the palette model's density() is the same sub-net
    with torch.no_grad():
        ref = pm.density(x)["sigma"].float()
    rel = ((fused_nerf.density(pm, x) - ref).abs() / ref.abs().clamp(min=1e-3)).max().item()
    assert rel < 1e-2, rel


@pytest.mark.parametrize("ds", [1.0, 40.0])
def test_nerf_fused_render_matches_loop(cuda, nerf, ds):
    o, d = S.camera_rays(48, 48)
    o, d = o.to(cuda)[None], d.to(cuda)[None]
    nerf.density_scale = ds
    try:
        with torch.no_grad():
            loop = nerf.render(o, d, staged=True, bg_color=1, perturb=False, fused=False)
            with torch.autocast("cuda", dtype=torch.float16):
                fus = nerf.render(o, d, staged=True, bg_color=1, perturb=False)
        assert nerf._last_schedule == "fused"
    finally:
        nerf.density_scale = 1.0
    for k in ("image", "depth", "weights_sum"):
        err = (loop[k].float() - fus[k].float()).abs().max().item()
        assert err < 1e-3, (k, err)
    assert fus["weights_sum"].max().item() > (0.99 if ds > 1 else 0.2)
    q = nerf._last_queue.cpu().numpy()
    assert q[1] > 0 and q[0] >= q[2] > 0


def _torch_full_sweep(model, jitter, decay):
    """nerf/renderer.py:476-553 (full sweep) in torch fp32 with explicit jitter numbers [C*H^3, 3] in x-major cell order"""
    C, H = model.cascade, model.grid_size
    H3 = H ** 3
    dev = model.density_grid.device
    g = torch.arange(H, dtype=torch.int32, device=dev)
    coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    idx = rm.morton3D(coords).long()
    unit = 2 * coords.float() / (H - 1) - 1
    tmp = -torch.ones(C, H3, device=dev)
    for cas in range(C):
        b = min(2 ** cas, model.bound)
        half = b / H
        pts = unit * (b - half) + (jitter[cas * H3:(cas + 1) * H3] * 2 - 1) * half
        sig = torch.cat([model.density(pts[i:i + (1 << 20)])["sigma"].float() for i in range(0, H3, 1 << 20)])
        tmp[cas, idx] = sig * model.density_scale
    grid = model.density_grid.clone()
    ok = (grid >= 0) & (tmp >= 0)
    grid[ok] = torch.maximum(grid[ok] * decay, tmp[ok])
    return grid


def test_update_extra_state_full_sweep_matches_torch_restatement(cuda):
    m = S.build_nerf_model(cuda, seed=6, table_scale=0.5)
    m.eval()
    with torch.no_grad():
        m.density_grid[0, :1000] = -1.0                       # "untrained" cells stay untouched (and never set a bit)
    C, H3 = m.cascade, m.grid_size ** 3
    jit = torch.rand(C * H3, 3, device=cuda, generator=torch.Generator(device=cuda).manual_seed(3))
    with torch.no_grad():
        want = _torch_full_sweep(m, jit, 0.95)
    stats = fused_nerf.update_density_grid(m, decay=0.95, jitter=jit.contiguous())
    got = m.density_grid
    assert torch.equal(got[0, :1000], torch.full((1000,), -1.0, device=cuda))
    rel = ((got - want).abs() / want.abs().clamp(min=1e-2)).max().item()
    assert rel < 1e-2, rel                                     # fp16 MLP vs fp32 torch
    mean, thresh = stats.tolist()
    assert abs(mean - got.clamp(min=0).mean().item()) <= 1e-5 * max(1.0, mean)
    assert abs(thresh - min(mean, m.density_thresh)) <= 1e-7
    want_bits = rm.packbits(got, thresh)
    assert torch.equal(m.density_bitfield, want_bits)
    assert m.density_bitfield.count_nonzero().item() > 100


def test_update_extra_state_schedule_partial_refresh_and_no_cpu_fallback(cuda):
    m = S.build_nerf_model(cuda, seed=7, table_scale=0.5)
    m.eval()
    H3 = m.grid_size ** 3
    m.iter_density = 16                                        # -> partial refresh: H^3/4 uniform + H^3/4 occupied cells per cascade
    before = m.density_grid.clone()
    occupied_before = before > 0
    m.update_extra_state()
    assert m._last_update_schedule == "fused" and m.iter_density == 17
    after = m.density_grid
    changed = after != before
    n_changed = changed.sum(dim=1)
    # every occupied cell decays or is refreshed only if it was selected; about 1 - exp(-1/4) of all cells get a uniform sample
    assert (n_changed > 0.15 * H3).all() and (n_changed < 0.45 * H3).all(), n_changed
    # a changed cell was either occupied before (decay / refresh) or received a fresh positive density
    assert (after[changed & ~occupied_before] > 0).all()
    st = m._density_scratch["stats"].tolist()
    assert torch.equal(m.density_bitfield, rm.packbits(after, st[1]))
    assert abs(m.mean_density - after.clamp(min=0).mean().item()) <= 1e-5 * max(1.0, st[0])
    assert (m._density_scratch["tmp"] == -1).all()            # the temporary grid is reset for the next refresh
    # the occupied list the sweep sampled from: ascending, exactly the cells with density > 0 BEFORE the refresh
    cnt = m._density_scratch["occ_count"].tolist()
    for cas in range(m.cascade):
        want = torch.nonzero(occupied_before[cas]).squeeze(-1).int()
        assert cnt[cas] == want.numel() and torch.equal(m._density_scratch["occ_list"][cas, :cnt[cas]], want)
    # render still works on the refreshed grid, and the torch schedule stays selectable
    m2 = S.build_nerf_model(cuda, seed=7, table_scale=0.5)
    m2.update_extra_state(fused=False)
    assert m2._last_update_schedule == "torch" and m2.iter_density == 1
