"""GPU parity of the fused TRAINING field (csrc/fused_train.cu: forward, hand-written backward, weight gradients).

Reference: the per-op training field of this repository (`PaletteRenderer._train_field_torch`, the reference's
palette/renderer.py:333-385 on torch fp32 autograd + the stand-alone kernels, themselves parity-checked against the
reference's extensions). Tolerances: the fused path keeps fp16 activations and fp16 pre-activation gradients between
layers (fp32 accumulation), like the reference under fp16 autocast:
  forward  max-abs 5e-3 on O(1) outputs (same bar as tests/test_fused_gpu.py)
  backward relative L2 error per parameter gradient <= 3e-2 (fp16 rounding of every layer's dY, 2^-11 relative each,
           accumulated over <= 5 chained layers and ~10^5 samples), cosine similarity >= 0.999
"""
import numpy as np
import pytest
import torch

from palettenerf_b200 import fused_train, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[True, False], ids=["clip", "noclip"])
def model(request, cuda):
    m = S.build_palette_model(cuda, seed=2, pred_clip=request.param, table_scale=0.5)
    m.train()
    return m


def _samples(model, cuda, n_side=40, max_steps=1024):
    import palettenerf_b200.raymarching as rm
    o, d = S.camera_rays(n_side, n_side)
    o, d = o.to(cuda), d.to(cuda)
    nears, fars = rm.near_far_from_aabb(o, d, model.aabb_train, model.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=cuda)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, model.bound, model.density_bitfield, model.cascade, model.grid_size,
                                                   nears, fars, counter, -1, False, -1, True, 0.0, max_steps)
    return xyzs, dirs, deltas, rays


def _palette(model, requires_grad):
    p = model.basis_color[None].clamp(0, 1)
    return p if requires_grad else p.detach()


def _torch_field(model, xyzs, dirs, palette):
    """sigma, rgb, channels of the unfused path WITHOUT the composite (dummy one-ray layout)"""
    M = xyzs.shape[0]
    rays = torch.tensor([[0, 0, M]], dtype=torch.int32, device=xyzs.device)
    deltas = torch.full((M, 2), 1e-3, device=xyzs.device)
    sig, rgb, ch, _, _, _ = model._train_field_torch(xyzs, dirs, deltas, rays, palette, 1e-4)
    return sig, rgb, ch


def test_fused_train_forward_matches_torch_field(cuda, model):
    assert fused_train.supported(model)
    xyzs, dirs, _, _ = _samples(model, cuda)
    M = xyzs.shape[0]
    assert M > 10000
    with torch.no_grad():
        sig, rgb, ch = fused_train.field(model, xyzs, dirs, _palette(model, False)[0])
        rsig, rrgb, rch = _torch_field(model, xyzs, dirs, _palette(model, False))
    rel = ((sig - rsig).abs() / rsig.abs().clamp(min=1e-3)).max().item()
    assert rel < 1e-2, f"sigma rel err {rel}"
    assert (rgb - rrgb).abs().max().item() < 5e-3
    cd = model.opt.clip_dim
    names = ["omega_sparsity", "view_dep_norm", "offsets_norm", "smooth_norm"] + ["view_dep"] * 3 + ["direct"] * 3 + \
            ["diffuse"] * 3 + ["clip"] * cd + ["omega"] * 4
    err = (ch - rch).abs().max(dim=0).values.cpu().numpy()
    scale = rch.abs().max(dim=0).values.clamp(min=1.0).cpu().numpy()
    for i, n in enumerate(names):
        assert err[i] <= 5e-3 * scale[i] + (2e-2 if n in ("omega_sparsity", "offsets_norm") else 0), (i, n, err[i], scale[i])
    assert rch[:, 4:13].std().item() > 1e-2 and rch[:, 13 + cd:].std().item() > 1e-3     # not vacuous
    if model.opt.pred_clip:
        assert rch[:, 13:13 + cd].std().item() > 1e-3
    else:
        assert ch[:, 13:13 + cd].abs().max().item() == 0


def _grads(model, fn):
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    for p in model.parameters():
        p.grad = None
    fn().backward()
    return {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters() if n in names}


def test_fused_train_backward_matches_torch_autograd(cuda, model):
    xyzs, dirs, _, _ = _samples(model, cuda, n_side=32)
    M = xyzs.shape[0]
    model.fused_basis_net_grad = True        # off by default: the reference never steps basis_net (dead work)
    nflex = 13 + model.opt.clip_dim + 4
    g = torch.Generator(device=cuda).manual_seed(5)
    w_rgb = torch.randn(M, 3, device=cuda, generator=g)
    w_ch = torch.randn(M, nflex, device=cuda, generator=g)
    w_ch[:, 3] = 0                                    # smooth_norm: identically zero on both paths

    def loss_fused():
        _, rgb, ch = fused_train.field(model, xyzs, dirs, _palette(model, True)[0])
        return (rgb * w_rgb).sum() + (ch * w_ch).sum()

    def loss_torch():
        _, rgb, ch = _torch_field(model, xyzs, dirs, _palette(model, True))
        return (rgb * w_rgb).sum() + (ch * w_ch).sum()

    gf, gt = _grads(model, loss_fused), _grads(model, loss_torch)
    expect = ["encoder_palette.embeddings", "diff_net.0.weight", "diff_net.1.weight", "diff_net.2.weight", "color_net.0.weight",
              "color_net.1.weight", "color_net.2.weight", "basis_net.0.weight", "basis_net.1.weight",
              "offsets_radiance_net.weight", "offsets_radiance_net.bias", "omega_net.0.weight", "basis_color"]
    if model.opt.pred_clip:
        expect += ["encoder_clip.embeddings", "clip_net.0.weight", "clip_net.1.weight"]
    for n in expect:
        a, b = gf[n], gt[n]
        assert a is not None and b is not None, n
        a, b = a.double().reshape(-1), b.double().reshape(-1)
        assert torch.isfinite(a).all(), n
        rel = (a - b).norm() / b.norm().clamp(min=1e-12)
        cos = torch.dot(a, b) / (a.norm() * b.norm()).clamp(min=1e-30)
        assert b.norm() > 0, n
        assert rel.item() < 3e-2 and cos.item() > 0.999, f"{n}: rel {rel.item():.3e} cos {cos.item():.6f}"
    # constants of the palette stage receive no gradient on either path
    for n in ("encoder.embeddings", "sigma_net.0.weight", "sigma_net.1.weight"):
        assert gf[n] is None or gf[n].abs().max().item() == 0, n
        assert gt[n] is None or gt[n].abs().max().item() == 0, n
    if not model.opt.pred_clip:
        assert gf["encoder_clip.embeddings"] is None
    model.fused_basis_net_grad = False
    g0 = _grads(model, loss_fused)
    assert g0["basis_net.0.weight"] is None and g0["basis_net.1.weight"] is None     # default: no gradient at all
    assert torch.equal(g0["diff_net.1.weight"], gf["diff_net.1.weight"]) or \
        torch.allclose(g0["diff_net.1.weight"], gf["diff_net.1.weight"], rtol=1e-4, atol=1e-7)


def test_fused_train_ragged_and_empty(cuda, model):
    xyzs, dirs, _, _ = _samples(model, cuda, n_side=16)
    pal = _palette(model, False)[0]
    with torch.no_grad():
        full = fused_train.field(model, xyzs, dirs, pal)
    for m in (1, 31, 33, 257):
        with torch.no_grad():
            part = fused_train.field(model, xyzs[:m].contiguous(), dirs[:m].contiguous(), pal)
        for a, b in zip(part, full):
            assert torch.equal(a, b[:m])
    # gradient through a ragged batch must be finite and non-zero
    _, rgb, ch = fused_train.field(model, xyzs[:45].contiguous(), dirs[:45].contiguous(), pal)
    for p in model.parameters():
        p.grad = None
    (rgb.sum() + ch.sum()).backward()
    gpal = model.encoder_palette.embeddings.grad
    assert gpal is not None and torch.isfinite(gpal).all() and gpal.abs().sum().item() > 0
    with torch.no_grad():
        e = fused_train.field(model, xyzs[:0].contiguous(), dirs[:0].contiguous(), pal)
    assert e[0].shape == (0,) and e[1].shape == (0, 3)


def test_training_render_fused_vs_torch_schedule_and_step(cuda):
    """model.render in training mode: same maps from both schedules, and an optimizer step on the fused path under
    fp16 autocast + GradScaler lowers the loss"""
    model = S.build_palette_model(cuda, seed=3, pred_clip=False, table_scale=0.3)
    model.train()
    o, d = S.training_rays(1024, seed=1)
    o, d = o.to(cuda), d.to(cuda)
    kw = dict(staged=False, bg_color=1, perturb=False, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
    with torch.autocast("cuda", dtype=torch.float16):
        a = model.render(o[None], d[None], fused=True, **kw)
        assert model._last_train_schedule == "fused"
    b = model.render(o[None], d[None], fused=False, **kw)
    for k in ("image", "direct_rgb", "diffuse_rgb", "view_dep_rgb", "basis_acc", "weights_sum", "omega_sparsity", "offsets_norm",
              "view_dep_norm"):
        err = (a[k].float() - b[k].float()).abs().max().item()
        ref = b[k].float().abs().max().item()
        assert err < 5e-3 * max(1.0, ref) + (1e-2 if k in ("omega_sparsity", "offsets_norm") else 0), (k, err, ref)

    gt = torch.rand(1, 1024, 3, device=cuda, generator=torch.Generator(device=cuda).manual_seed(0))
    hit = b["weights_sum"].detach() > 0.05          # rays that miss the solid render the background: constant loss
    assert hit.sum().item() > 100
    opt = torch.optim.Adam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    losses = []
    for _ in range(12):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(o[None], d[None], **kw)
            loss = ((out["image"] - gt) ** 2)[0, hit].mean() + ((out["direct_rgb"] - gt) ** 2)[0, hit].mean() \
                + 2e-4 * out["omega_sparsity"].mean()
        assert model._last_train_schedule == "fused"
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < 0.9 * losses[0], losses


def test_cuda_graph_captured_step_matches_eager_steps(cuda):
    """the whole step (static-capacity march, fused field fwd/bwd/wgrad, one-pass compositor, loss, GradScaler, Adam) is
    captured in ONE CUDA graph; replaying it must train exactly like the eager schedule"""
    from palettenerf_b200.graphs import GraphedStep, make_palette_train_step
    o, d = S.training_rays(512, seed=4)
    o, d = o.to(cuda)[None].contiguous(), d.to(cuda)[None].contiguous()
    gt = torch.rand(1, 512, 3, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))

    def loss_fn(out):
        return ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean() \
            + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean()

    def make():
        m = S.build_palette_model(cuda, seed=7, pred_clip=False, table_scale=0.3)
        m.train()
        opt = torch.optim.Adam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True)
        sc = torch.amp.GradScaler("cuda", init_scale=1024.0)
        return m, make_palette_train_step(m, opt, sc, o, d, loss_fn, render_kwargs=dict(perturb=False))

    m_eager, step_eager = make()
    m_graph, step_graph = make()
    n_warm, n_replay = 2, 4
    g = GraphedStep(step_graph, warmup=n_warm)          # n_warm eager steps + 1 captured (not executed) step
    losses_g = [float(g.replay()) for _ in range(n_replay)]
    losses_e = [float(step_eager()) for _ in range(n_warm + n_replay)][n_warm:]
    assert m_graph._last_train_schedule == "fused"
    np.testing.assert_allclose(losses_g, losses_e, rtol=2e-3, atol=1e-5)    # fp32 atomics reorder sums between runs
    assert losses_g[-1] < losses_g[0]
    for (n1, p1), (_, p2) in zip(m_graph.named_parameters(), m_eager.named_parameters()):
        if p1.requires_grad and p1.numel() < 100000:
            assert torch.allclose(p1, p2, rtol=5e-2, atol=5e-3), n1


def test_fused_adam_mirror_trains_like_torch_adam_with_table_refresh(cuda):
    """FusedAdam writes the fp16 copy of the trained hash table (the interleaved table the forward kernel reads) in its own
    pass; torch's Adam leaves that to the per-step refresh copy. Same model, same rays: the two must train alike — eager and
    from a CUDA graph (whose capture then contains no refresh copy at all) — and the mirror must equal fp16(parameter)."""
    from palettenerf_b200.graphs import GraphedStep, make_palette_train_step
    from palettenerf_b200.optim import FusedAdam
    from palettenerf_b200.palette.losses import palette_loss
    o, d = S.training_rays(512, seed=5)
    o, d = o.to(cuda)[None].contiguous(), d.to(cuda)[None].contiguous()
    gt = torch.rand(1, 512, 3, device=cuda, generator=torch.Generator(device=cuda).manual_seed(2))

    def make(fused_adam):
        m = S.build_palette_model(cuda, seed=8, pred_clip=False, table_scale=0.3)
        m.train()
        params = m.get_params(1e-2)
        opt = FusedAdam(params, betas=(0.9, 0.99), eps=1e-15) if fused_adam else \
            torch.optim.Adam(params, betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True)
        sc = torch.amp.GradScaler("cuda", init_scale=1024.0)
        return m, make_palette_train_step(m, opt, sc, o, d, lambda out: palette_loss(out, gt, 2e-4, 0.03, 0.1)[0],
                                          render_kwargs=dict(perturb=False))

    m_ref, step_ref = make(False)
    m_new, step_new = make(True)
    m_gr, step_gr = make(True)
    l_ref = [float(step_ref()) for _ in range(6)]
    l_new = [float(step_new()) for _ in range(6)]
    g = GraphedStep(step_gr, warmup=2)
    l_gr = [float(g.replay()) for _ in range(4)]
    np.testing.assert_allclose(l_new, l_ref, rtol=2e-3, atol=1e-5)
    np.testing.assert_allclose(l_gr, l_ref[2:], rtol=2e-3, atol=1e-5)
    assert l_new[-1] < l_new[0]
    for m in (m_new, m_gr):
        p = m.encoder_palette.embeddings
        buf, off, stride = p._pnerf_half_mirror
        assert p._pnerf_mirror_version == p._version and (off, stride) == (4, 8)
        assert torch.equal(buf[:, 1, :], p.detach().to(torch.float16))          # the mirror IS fp16(parameter)
        assert torch.equal(buf[:, 0, :], m.encoder.embeddings.detach().to(torch.float16))
    for (n1, p1), (_, p2) in zip(m_new.named_parameters(), m_ref.named_parameters()):
        if p1.requires_grad and p1.numel() < 100000:
            assert torch.allclose(p1, p2, rtol=5e-2, atol=5e-3), n1


@pytest.mark.parametrize("stage", ["palette", "nerf"])
def test_eager_steps_reuse_the_arena_without_the_cyclic_gc(cuda, stage):
    """the static-capacity step takes its large buffers from the arena, which hands a buffer out again as soon as nothing
    references it. An autograd ctx that holds one of its own OUTPUTS forms a reference cycle only the cyclic GC breaks: with the
    collector off (what a long eager training loop looks like between two gen-2 collections) the arena must still stay at the
    working set of ONE step"""
    import gc
    from palettenerf_b200.arena import ARENA
    if stage == "palette":
        m = S.build_palette_model(cuda, seed=0, pred_clip=False)
    else:
        m = S.build_nerf_model(cuda, seed=0)
    m.train()
    o, d = S.training_rays(512, H=200, W=200, seed=0, n_views=2)
    o, d = o.to(cuda)[None], d.to(cuda)[None]
    gt = torch.rand(1, 512, 3, device=cuda)

    def step():
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.float16):
            out = m.render(o, d, staged=False, bg_color=1, perturb=True, force_all_rays=True, max_steps=256)
            loss = ((out["image"] - gt) ** 2).mean()
        loss.backward()
    ARENA.clear()
    gc.collect()
    gc.disable()
    try:
        step(); step()
        base = ARENA.bytes()
        for _ in range(5):
            step()
        assert ARENA.bytes() == base, (ARENA.bytes(), base)
    finally:
        gc.enable()
    assert m._last_train_schedule == "fused" and base > 0
