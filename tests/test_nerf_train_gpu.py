"""Fused stage-1 training step (csrc/nerf_train.cu, palettenerf_b200/fused_nerf_train.py).

  * the fused field (forward, data gradients, weight gradients, hash-grid gradient) against the module's own torch-op field
    in fp32 (NeRFNetwork.forward on the stand-alone kernels, itself pinned to the reference in test_golden_palette_gpu.py):
    outputs <= 1e-3 (sigma relative), gradients <= 2e-2 relative L2 (fp16 tensor-core arithmetic, the bar of the palette
    stage's fused step);
  * the one-pass compositor against spread_ray_to_sample + composite_rays_train twice (the reference's schedule on kernels
    that are bit-compared with the reference's): <= 1e-5 forward, <= 1e-4 relative on the gradients;
  * the whole run_cuda training branch, fused vs torch schedule, incl. capacity overflow with NaN-poisoned scratch;
  * the step inside a CUDA graph with FusedAdam.
The reference-pinned comparison of the same step is tests/test_golden_palette_gpu.py::test_nerf_stage_forward_render_and_train
(its f16 case takes the fused schedule)."""
import numpy as np
import pytest
import torch

from palettenerf_b200 import fused_nerf_train as FT, synthetic as S
import palettenerf_b200.raymarching as rm

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


@pytest.fixture()
def nerf(cuda):
    m = S.build_nerf_model(cuda, seed=4, table_scale=0.5)
    m.train()
    return m


@pytest.mark.parametrize("counted", [False, True])
def test_fused_field_matches_torch_field(cuda, nerf, counted):
    g = torch.Generator(device=cuda).manual_seed(1)
    M = 5000
    x = (torch.rand(M, 3, device=cuda, generator=g) * 2 - 1) * nerf.bound * 0.98
    d = torch.nn.functional.normalize(torch.randn(M, 3, device=cuda, generator=g), dim=-1)
    gs = torch.randn(M, device=cuda, generator=g) * 0.1
    gc = torch.randn(M, 3, device=cuda, generator=g)
    nerf.density_scale = 2.0
    valid = 4321 if counted else M
    try:
        sig_r, rgb_r = nerf(x[:valid], d[:valid])
        sig_r = nerf.density_scale * sig_r
        (sig_r * gs[:valid]).sum().backward(retain_graph=True)
        (rgb_r * gc[:valid]).sum().backward()
        ref = {n: p.grad.clone() for n, p in nerf.named_parameters() if p.grad is not None}
        for p in nerf.parameters():
            p.grad = None
        count = torch.tensor([valid], dtype=torch.int32, device=cuda) if counted else None
        if counted:                      # rows behind the count must be neither read nor written
            x = x.clone(); x[valid:] = float("nan")
        sig, rgb = FT.field(nerf, x, d, count=count)
        (sig[:valid] * gs[:valid]).sum().backward(retain_graph=True)
        (rgb[:valid] * gc[:valid]).sum().backward()
    finally:
        nerf.density_scale = 1.0
    assert ((sig[:valid] - sig_r).abs() / sig_r.abs().clamp(min=1e-3)).max().item() < 5e-3
    assert (rgb[:valid] - rgb_r).abs().max().item() < 1e-3
    assert set(ref) == {n for n, p in nerf.named_parameters() if p.grad is not None}
    for n, p in nerf.named_parameters():
        assert torch.isfinite(p.grad).all(), n
        assert _rel(p.grad, ref[n]) < 2e-2, (n, _rel(p.grad, ref[n]))


def _march(m, n_rays, cuda, static=False, mean_count=-1, force=True):
    o, d = S.training_rays(n_rays, H=200, W=200, seed=3, n_views=2)
    o, d = o.to(cuda), d.to(cuda)
    nears, fars = rm.near_far_from_aabb(o, d, m.aabb_train, m.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    out = rm.march_rays_train(o, d, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter, mean_count, False,
                              128, force, 0.0, 1024, static)
    return out


@pytest.mark.parametrize("with_gt", [True, False])
def test_one_pass_compositor_matches_reference_schedule(cuda, nerf, with_gt):
    xyzs, dirs, deltas, rays = _march(nerf, 777, cuda)
    M, N = xyzs.shape[0], rays.shape[0]
    g = torch.Generator(device=cuda).manual_seed(2)
    sig = (torch.rand(M, device=cuda, generator=g) * 30).requires_grad_()
    rgb = torch.rand(M, 3, device=cuda, generator=g).requires_grad_()
    gt = torch.rand(N, 3, device=cuda, generator=g) if with_gt else None
    g_img, g_ws, g_err = (torch.randn(N, 3, device=cuda, generator=g), torch.randn(N, device=cuda, generator=g),
                          torch.randn(N, device=cuda, generator=g))
    # the reference's schedule (nerf/renderer.py:301-327)
    ws_r, depth_r, img_r = rm.composite_rays_train(sig, rgb, deltas, rays, 1e-4)
    if with_gt:
        spread = torch.zeros_like(xyzs)
        rm.spread_ray_to_sample(gt, rays, spread)
        err = ((spread - rgb) ** 2).sum(-1, keepdim=True).repeat(1, 3)
        err_r = rm.composite_rays_train(sig, err, deltas, rays, 1e-4)[2].mean(-1)
    else:
        err_r = torch.zeros_like(ws_r)
    ((img_r * g_img).sum() + (ws_r * g_ws).sum() + (err_r * g_err).sum()).backward()
    ref = (sig.grad.clone(), rgb.grad.clone())
    sig.grad = rgb.grad = None
    ws, depth, img, e = FT.composite(sig, rgb, deltas, rays, gt, 1e-4)
    ((img * g_img).sum() + (ws * g_ws).sum() + (e * g_err).sum()).backward()
    for a, b, name in ((ws, ws_r, "ws"), (depth, depth_r, "depth"), (img, img_r, "image"), (e, err_r, "err")):
        assert (a - b).abs().max().item() < 1e-5, name
    assert (ws > 0.5).sum() > 10
    assert _rel(sig.grad, ref[0]) < 1e-4 and _rel(rgb.grad, ref[1]) < 1e-5


def _step(m, o, d, gt, fused, **kw):
    for p in m.parameters():
        p.grad = None
    with torch.autocast("cuda", dtype=torch.float16):
        out = m.render(o, d, rays_gt=gt, staged=False, bg_color=1, perturb=False, fused=fused, **kw)
        loss = ((out["image"] - gt) ** 2).mean() + 0.1 * out["rgb_norm"].mean()
    (loss * 128.0).backward()
    return out, loss, {n: p.grad.detach().float() / 128.0 for n, p in m.named_parameters() if p.grad is not None}


def test_run_cuda_training_branch_fused_vs_torch_schedule(cuda, nerf):
    o, d = S.training_rays(1024, H=200, W=200, seed=5, n_views=3)
    o, d = o.to(cuda)[None], d.to(cuda)[None]
    gt = torch.rand(1, 1024, 3, device=cuda)
    nerf.density_scale = 10.0                          # opaque enough for rays to terminate
    try:
        out_t, loss_t, g_t = _step(nerf, o, d, gt, False, force_all_rays=True)
        assert nerf._last_train_schedule == "torch"
        out_f, loss_f, g_f = _step(nerf, o, d, gt, None, force_all_rays=True)
        assert nerf._last_train_schedule == "fused"
    finally:
        nerf.density_scale = 1.0
    for k in ("image", "depth", "weights_sum", "rgb_norm"):
        assert (out_f[k].float() - out_t[k].float()).abs().max().item() < 2e-3, k
    assert abs(loss_f.item() - loss_t.item()) < 1e-3 * max(1.0, abs(loss_t.item()))
    assert set(g_f) == set(g_t)
    for n in g_t:
        assert _rel(g_f[n], g_t[n]) < 3e-2, (n, _rel(g_f[n], g_t[n]))   # both sides are fp16 runs of the same step


def test_fused_step_survives_capacity_overflow_with_poisoned_scratch(cuda, nerf):
    from palettenerf_b200.arena import ARENA
    o, d = S.training_rays(1024, H=200, W=200, seed=6, n_views=3)
    o, d = o.to(cuda)[None], d.to(cuda)[None]
    gt = torch.rand(1, 1024, 3, device=cuda)
    full, _, _ = _step(nerf, o, d, gt, None, force_all_rays=True)
    total = int(nerf.step_counter[(nerf.local_step - 1) % 16, 0].item())
    nerf.mean_count = total // 2
    cap = nerf.mean_count + (128 - nerf.mean_count % 128)
    ARENA.clear()
    for nm, w in (("march_xyzs", 3), ("march_dirs", 3), ("march_deltas", 2), ("nerf_rgb", 3), ("nerf_g_rgbs", 3)):
        ARENA.get(nm, (cap, w), torch.float32, cuda).fill_(float("nan"))
    for nm in ("nerf_sigma", "nerf_g_sigmas"):
        ARENA.get(nm, (cap,), torch.float32, cuda).fill_(float("nan"))
    out, loss, grads = _step(nerf, o, d, gt, None, force_all_rays=False)
    assert nerf._last_train_schedule == "fused" and torch.isfinite(loss)
    for n, gv in grads.items():
        assert torch.isfinite(gv).all(), n
    kept = out["weights_sum"] > 0
    assert kept.sum() > 10 and (~kept).sum() > 10
    assert torch.allclose(out["image"][0][kept], full["image"][0][kept], atol=1e-6)


def test_fused_step_in_cuda_graph_trains(cuda):
    from palettenerf_b200.graphs import GraphedStep
    from palettenerf_b200.optim import FusedAdam
    m = S.build_nerf_model(cuda, seed=7, table_scale=0.5)
    m.train()
    m.density_scale = 10.0
    opt = FusedAdam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda", init_scale=128.0)
    o, d = S.training_rays(1024, H=200, W=200, seed=8, n_views=3)
    o, d = o.to(cuda)[None].contiguous(), d.to(cuda)[None].contiguous()
    gt = torch.full((1, 1024, 3), 0.25, device=cuda)

    def step_fn():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = m.render(o, d, rays_gt=gt, staged=False, bg_color=1, perturb=True, force_all_rays=True)
            loss = ((out["image"] - gt) ** 2).mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        return loss.detach()
    gs = GraphedStep(step_fn, warmup=3)
    losses = [float(gs.replay().item()) for _ in range(60)]
    assert m._last_train_schedule == "fused"
    # most rays miss the object (white background against the 0.25 target): only the object's share of the loss can fall
    assert np.isfinite(losses).all() and losses[-1] < losses[0] - 5e-3, losses
