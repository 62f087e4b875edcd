"""GPU parity of the fused field kernel and the persistent fused renderer (csrc/fused.cu).

Oracle: the CPU restatement of PaletteNetwork.forward / the cuda_ray schedule (oracle/cpu_render.py, fp32) and the
unfused torch path of this repo. Tolerance: the fused MLPs run fp16 tensor-core math (fp16 activations between
layers, fp32 accumulation) like the reference under `-O` (fp16 autocast): max-abs 5e-3 on O(1) outputs
(north_star: 1e-3 for fp16 per op; the field chains up to 9 fp16 layers), relative 1e-2 on sigma."""
import numpy as np
import pytest
import torch

import oracle
from oracle import cpu_render
from palettenerf_b200 import fused, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[True, False], ids=["clip", "noclip"])
def model(request, cuda):
    m = S.build_palette_model(cuda, seed=1, pred_clip=request.param, table_scale=0.5)
    m.eval()
    return m


def _samples(model, cuda, n_side=48):
    import palettenerf_b200.raymarching as rm
    o, d = S.camera_rays(n_side, n_side)
    o, d = o.to(cuda), d.to(cuda)
    nears, fars = rm.near_far_from_aabb(o, d, model.aabb_infer, model.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, model.bound, model.density_bitfield, model.cascade, model.grid_size,
                                                   nears, fars, counter, -1, False, -1, True, 0.0, 1024)
    return xyzs, dirs


def test_fused_field_matches_torch_path_and_oracle(cuda, model):
    assert fused.supported(model)
    xyzs, dirs = _samples(model, cuda)
    M = xyzs.shape[0]
    assert M > 20000
    out = fused.field_forward(model, xyzs, dirs)
    with torch.no_grad():
        ref = model(xyzs, dirs)                      # unfused fp32 torch path on the stand-alone kernels
    names = ["sigma", "clip", "omega", "offsets_radiance", "view_dep", "diffuse"]
    for n, a, b in zip(names, out, ref):
        a, b = a.float(), b.float().reshape(a.shape)
        if n == "sigma":
            rel = ((a - b).abs() / b.abs().clamp(min=1e-3)).max().item()
            assert rel < 1e-2, f"sigma rel err {rel}"
        else:
            err = (a - b).abs().max().item()
            assert err < 5e-3, f"{n}: max-abs {err}"
    assert out[2].sum(-1).sub(1).abs().max().item() < 1e-5          # omega is normalised
    assert out[0].std().item() > 1e-3 and out[4].std().item() > 1e-3  # not vacuous
    # CPU oracle on a subset
    idx = torch.randperm(M, generator=torch.Generator().manual_seed(0))[:3000].to(cuda)
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    oref = cpu_render.palette_forward(params, xyzs[idx].cpu(), dirs[idx].cpu(), model.bound, model.encoder.per_level_scale,
                                      model.opt.pred_clip)
    for n, a, b in zip(names, out, oref):
        a = a[idx].float().cpu()
        if n == "sigma":
            assert ((a - b).abs() / b.abs().clamp(min=1e-3)).max().item() < 1e-2
        else:
            assert (a - b.reshape(a.shape)).abs().max().item() < 5e-3, n


def test_fused_field_ragged_and_empty(cuda, model):
    xyzs, dirs = _samples(model, cuda, 16)
    full = fused.field_forward(model, xyzs, dirs)
    for m in (1, 31, 33, 257):
        part = fused.field_forward(model, xyzs[:m], dirs[:m])
        for a, b in zip(part, full):
            assert torch.equal(a, b[:m])              # a sample's result does not depend on its tile neighbours
    empty = fused.field_forward(model, xyzs[:0], dirs[:0])
    assert empty[0].numel() == 0


@pytest.mark.parametrize("gui_mode", [False, True])
def test_fused_render_matches_loop_and_oracle(cuda, model, gui_mode):
    o, d = S.camera_rays(40, 40)
    oc, dc = o.to(cuda)[None], d.to(cuda)[None]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        loop = model.render(oc, dc, staged=True, bg_color=1, perturb=False, gui_mode=gui_mode, fused=False)
        fus = model.render(oc, dc, staged=True, bg_color=1, perturb=False, gui_mode=gui_mode)   # default policy
    assert model._last_schedule == "fused"
    assert set(loop.keys()) == set(fus.keys())
    for k in loop:
        err = (loop[k].float() - fus[k].float()).abs().max().item()
        scale = max(1.0, loop[k].float().abs().max().item())     # depth_origin is in scene units (up to ~5)
        assert err < 5e-3 * scale, f"{k}: fused vs loop max-abs {err}"
    params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = cpu_render.render_cuda_ray(params, o, d, model.density_bitfield.cpu(), pred_clip=model.opt.pred_clip, gui_mode=gui_mode)
    q = model._last_queue.cpu().numpy()
    # every hit ray consumed; same samples shaded (fp16 sigma noise may move an early termination by a sample)
    assert q[0] >= q[2] > 0 and abs(int(q[1]) - ref["n_samples"]) <= 0.002 * ref["n_samples"]
    # with termination disabled the sample sets must be identical: exact count
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model.render(oc, dc, staged=True, bg_color=1, perturb=False, gui_mode=gui_mode, T_thresh=-1.0)
    ref_all = cpu_render.render_cuda_ray(params, o, d, model.density_bitfield.cpu(), pred_clip=model.opt.pred_clip, gui_mode=True,
                                         T_thresh=-1.0)
    assert int(model._last_queue.cpu().numpy()[1]) == ref_all["n_samples"]
    for k in ["image", "depth", "weights_sum", "clip_feat"] + ([] if gui_mode else ["direct_rgb", "view_dep_rgb", "basis_rgb",
                                                                                    "unscaled_basis_rgb", "basis_acc"]):
        a = fus[k].float().cpu().numpy().reshape(ref[k].shape)
        assert np.abs(a - ref[k]).max() < 5e-3 * max(1.0, np.abs(ref[k]).max()), f"{k}: fused vs oracle {np.abs(a - ref[k]).max()}"
    assert fus["weights_sum"].max().item() > 0.2


def test_fused_render_early_termination_and_step_budget(cuda):
    """dense field (sigma shifted up): rays terminate on T < T_thresh; and a max_steps budget smaller than the ray's
    sample count truncates it. Checked against the oracle's schedule with the same T_thresh."""
    m = S.build_palette_model(cuda, seed=2, pred_clip=False, table_scale=0.5)
    m.eval()
    m.density_scale = 400.0           # alpha ~ 0.75 per sample -> T < 1e-2 within a handful of samples
    o, d = S.camera_rays(32, 32)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        fus = m.render(o.to(cuda)[None], d.to(cuda)[None], staged=True, bg_color=1, perturb=False, gui_mode=True, T_thresh=1e-2)
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = cpu_render.render_cuda_ray(params, o, d, m.density_bitfield.cpu(), gui_mode=True, T_thresh=1e-2, density_scale=400.0)
    q = m._last_queue.cpu().numpy()
    # fp16 sigma noise can move a termination by one sample on a few rays
    assert abs(int(q[1]) - ref["n_samples"]) <= 0.02 * ref["n_samples"] + 8
    assert q[1] < 0.25 * 17664 * (32 * 32) / (24 * 24)                  # far fewer samples than the unterminated march
    assert np.abs(fus["weights_sum"].cpu().numpy() - ref["weights_sum"]).max() < 1e-2
    assert np.abs(fus["image"][0].cpu().numpy() - ref["image"]).max() < 1e-2


def test_fused_render_follows_in_place_bitfield_updates(cuda):
    """the ray walks are clipped at the bounds of the occupied cells, cached per bitfield tensor: an in-place update of
    `density_bitfield` (what update_extra_state does) must invalidate them — the fused render after the update equals
    the loop schedule (no clip anywhere on that path) on the updated grid, and its sample count changes accordingly"""
    m = S.build_palette_model(cuda, seed=3, pred_clip=False, table_scale=0.5)
    m.eval()
    o, d = S.camera_rays(40, 40)
    o, d = o.to(cuda)[None], d.to(cuda)[None]

    def both():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            f = m.render(o, d, staged=True, bg_color=1, perturb=False, gui_mode=True, fused=True)
            n = int(m._last_queue[1].item())
            l = m.render(o, d, staged=True, bg_color=1, perturb=False, gui_mode=True, fused=False)
        return f, l, n

    f0, l0, n0 = both()
    assert (f0["image"] - l0["image"]).abs().max().item() < 2e-3
    # grow the occupancy far outside the old bounds, in place: a slab of cascade-0 cells near +x
    H = m.grid_size
    coords = torch.stack(torch.meshgrid(torch.arange(118, 124), torch.arange(40, 90), torch.arange(40, 90), indexing="ij"), -1)
    codes = torch.from_numpy(oracle.morton3D(coords.reshape(-1, 3).int().numpy())).long().to(cuda)
    bf = m.density_bitfield
    for b in range(8):      # index_put with duplicates is not an OR: set one bit plane at a time
        sel = codes[(codes % 8) == b] // 8
        bf[sel] = bf[sel] | (1 << b)
    f1, l1, n1 = both()
    assert n1 > n0, "the new cells lie outside the old occupied bounds: a stale clip would not see them"
    assert (f1["image"] - l1["image"]).abs().max().item() < 2e-3
    assert (f1["weights_sum"] - l1["weights_sum"]).abs().max().item() < 2e-3


@pytest.mark.parametrize("gui_mode", [False, True])
def test_shared_windows_equal_reproducible_render(cuda, gui_mode):
    """The tensor-core renderer fills the free lanes of a ray's last 32-sample window with the head of the next ray in the
    queue (default) — exercised only when there are many more rays than resident warps (148 x 16), hence the 256 x 256
    view. PNERF_RENDER_REPRODUCIBLE (model.fused_reproducible) starts every ray's windows at its own first sample. Same
    samples, same maps up to the fp32 re-association of the per-ray sums; the reproducible render is bit-identical from
    run to run; the shared one evaluates fewer tiles."""
    m = S.build_palette_model(cuda, seed=3, pred_clip=False, table_scale=0.5)
    m.eval()
    o, d = S.camera_rays(256, 256)
    oc, dc = o.to(cuda)[None], d.to(cuda)[None]

    def render(reproducible, **kw):
        m.fused_reproducible = reproducible
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            out = m.render(oc, dc, staged=True, bg_color=1, perturb=False, gui_mode=gui_mode, **kw)
        assert m._last_schedule == "fused"
        return {k: v.float().clone() for k, v in out.items()}, m._last_queue.cpu().numpy().copy()

    for kw, tol, count_tol in (({}, 2e-5, 0.0), (dict(T_thresh=1e-2), 2e-4, 1e-3)):
        if kw:
            m.density_scale = 12.0        # some rays terminate on T < T_thresh, some of them inside a shared window
        rep, q_rep = render(True, **kw)
        rep2, q_rep2 = render(True, **kw)
        sh, q_sh = render(False, **kw)
        assert q_rep[2] > 4 * 148 * 16, "the view must give every warp several rays"
        for k in rep:
            assert torch.equal(rep[k], rep2[k]), f"{k}: reproducible render differs between runs"
            scale = max(1.0, rep[k].abs().max().item())
            err = (rep[k] - sh[k]).abs().max().item()
            assert err <= tol * scale, f"{k}: shared vs reproducible windows {err}"
        assert int(q_rep[1]) == int(q_rep2[1]) and int(q_rep[2]) == int(q_sh[2])
        assert abs(int(q_rep[1]) - int(q_sh[1])) <= count_tol * int(q_rep[1])
        fill_rep, fill_sh = q_rep[1] / (32.0 * q_rep[3]), q_sh[1] / (32.0 * q_sh[3])
        if not kw:      # (with early termination most rays end inside their first, full window: nothing to share)
            assert fill_sh > fill_rep + 0.03 and fill_sh > 0.9, (fill_rep, fill_sh)
    m.fused_reproducible = False


def test_weight_image_product_layers_kernel_equals_torch_restatement(cuda, model):
    """the two product layers of the tcgen05 weight image (field_tc.cuh layer table): pnerf_field_cache_merge (the kernel the
    cache refresh runs) against fused_train.tc_merge_layers (torch matmuls, the path for non-contiguous parameters) and against
    the products formed in fp64; everything else in the image is a pure gather"""
    from palettenerf_b200 import fused, fused_train as FT
    cache = fused._cache(model)
    cache.invalidate()
    cache.get()
    img = cache.buf["wpack_tc"].clone()
    alt = torch.zeros_like(img)
    FT.tc_merge_layers(model, alt)
    sd = {k: v.detach().double() for k, v in model.named_parameters()}
    d0 = sd["diff_net.0.weight"] @ sd["sigma_net.1.weight"][1:16]
    b1 = torch.zeros(32, 64, dtype=torch.float64, device=cuda)
    b1[0:13] = sd["offsets_radiance_net.weight"] @ sd["basis_net.1.weight"]
    b1[13:17] = sd["omega_net.0.weight"] @ sd["basis_net.1.weight"]
    for name, W in (("d0", d0), ("b1", b1)):
        n, k = W.shape
        off = FT.tc_layer_offset(name)
        want = W.reshape(n, k // 8, 8).permute(1, 0, 2).reshape(-1)
        got, got_alt = img[off:off + n * k].double(), alt[off:off + n * k].double()
        tol = 2.0 ** -10 * want.abs().clamp(min=1e-4)                    # fp16 rounding of an fp32 product sum
        assert ((got - want).abs() <= tol).all(), name
        assert ((got_alt - want).abs() <= tol).all(), name
        assert want.abs().max() > 1e-3
    # the unused pad layer stays zero
    off = FT.tc_layer_offset("h")
    assert img[off:off + 256].abs().max().item() == 0.0
