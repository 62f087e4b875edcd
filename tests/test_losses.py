"""Fused per-ray losses (ref palette/utils.py:486-567): `pnerf_palette_loss` vs the torch restatement
of the reference's train_step arithmetic (oracle/cpu_render.py::palette_train_loss, fp64 on the CPU), values and
gradients through autograd. fp32 sums of 4096-12288 O(1) terms: rel 2e-6 on the value, 1e-6 abs on O(1/N) gradients
scaled by N."""
import numpy as np
import pytest
import torch

from oracle import cpu_render


def _fake_outputs(N, cd, nb, dev, seed=0, views=True, dtype=torch.float32):
    """a render dict with the training branch's layout: maps [N, 13+cd+nb] and column views (palette/renderer.py:232-242)"""
    g = torch.Generator().manual_seed(seed)            # drawn in fp32 so that the fp64 oracle sees the same numbers
    maps = torch.rand(N, 13 + cd + nb, generator=g).to(dtype).to(dev).requires_grad_(True)
    image = torch.rand(1, N, 3, generator=g).to(dtype).to(dev).requires_grad_(True)
    ws = torch.rand(N, generator=g).to(dtype).to(dev)
    m = maps * 1.0 if views else maps                   # a non-leaf base, like the composite's output
    pick = (lambda a, b: m[..., a:b]) if views else (lambda a, b: m[..., a:b].clone())
    out = {"image": image, "weights_sum": ws,
           "omega_sparsity": pick(0, 1).view(1, N), "view_dep_norm": pick(1, 2).view(1, N), "offsets_norm": pick(2, 3).view(1, N),
           "smooth_norm": pick(3, 4).view(1, N), "direct_rgb": (m[..., 7:10] + (1 - ws).unsqueeze(-1)).view(1, N, 3),
           "clip_feat": pick(13, 13 + cd).view(1, N, cd), "basis_acc": pick(13 + cd, 13 + cd + nb).view(1, N, nb)}
    return out, maps, image


CASES = [dict(N=4096, clip=False, weights=False, palette=False, smooth=0.0),
         dict(N=4096, clip=True, weights=True, palette=True, smooth=0.3),
         dict(N=1, clip=True, weights=False, palette=True, smooth=0.0),
         dict(N=5000, clip=False, weights=True, palette=False, smooth=0.0)]     # not a multiple of the CTA size


def test_losses_module_needs_cuda():
    from palettenerf_b200.palette.losses import palette_loss
    out, _, _ = _fake_outputs(8, 16, 4, "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        palette_loss(out, torch.rand(1, 8, 3), 2e-4, 0.03, 0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("views", [True, False], ids=["column-views", "gathered"])
@pytest.mark.parametrize("case", CASES)
def test_fused_palette_loss_vs_oracle(cuda, case, views):
    from palettenerf_b200.palette.losses import palette_loss, TERMS
    N, cd, nb = case["N"], 16, 4
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(1, N, 3, generator=g)
    gt_clip = torch.randn(1, N, cd, generator=g) if case["clip"] else None
    gt_w = torch.rand(1, N, nb, generator=g) if case["weights"] else None
    bc = torch.rand(nb, 3, generator=g) if case["palette"] else None
    bc0 = torch.rand(nb, 3, generator=g) if case["palette"] else None
    lam = dict(lambda_sparsity=2e-4, lambda_offsets=0.03, lambda_view_dep=0.1, lambda_smooth=case["smooth"], lambda_weight=0.05,
               lambda_palette=0.001)
    # oracle: fp64 on the CPU
    o64, maps64, img64 = _fake_outputs(N, cd, nb, "cpu", dtype=torch.float64)
    bc64 = bc.double().requires_grad_(True) if bc is not None else None
    ref, rd, ref_per_ray = cpu_render.palette_train_loss(
        o64, gt.double(), gt_clip_feat=None if gt_clip is None else gt_clip.double(), gt_weights=None if gt_w is None else gt_w.double(),
        basis_color=bc64, basis_color_origin=None if bc0 is None else bc0.double(), **lam)
    scale = 1024.0                                       # the GradScaler's scale arrives as the upstream gradient
    (ref * scale).backward()
    # product
    oc, maps, img = _fake_outputs(N, cd, nb, cuda, views=views)
    bcc = bc.to(cuda).requires_grad_(True) if bc is not None else None
    with torch.autocast("cuda", dtype=torch.float16):
        loss, terms, per_ray = palette_loss(oc, gt.to(cuda), gt_clip_feat=None if gt_clip is None else gt_clip.to(cuda),
                                            gt_weights=None if gt_w is None else gt_w.to(cuda), basis_color=bcc,
                                            basis_color_origin=None if bc0 is None else bc0.to(cuda), **lam)
    assert loss.dtype == torch.float32 and loss.dim() == 0 and terms.shape == (len(TERMS),)
    (loss * scale).backward()
    assert abs(loss.item() - ref.item()) <= 2e-6 * abs(ref.item())
    t = dict(zip(TERMS, terms.cpu().tolist()))
    for k in ("rgb", "direct", "clip_feat", "sparsity", "offsets", "view_dep", "smooth", "weight", "palette"):
        assert abs(t[k] - float(rd[k])) <= 2e-6 * abs(float(rd[k])) + 1e-12, k
    assert abs(t["total"] - sum(t[k] for k in TERMS[1:])) <= 1e-6
    assert np.abs(per_ray.cpu().numpy() - ref_per_ray.reshape(-1).numpy()).max() < 1e-6
    for name, a, b in (("maps", maps.grad, maps64.grad), ("image", img.grad, img64.grad)):
        assert a is not None, name
        err = (a.double().cpu() - b).abs().max().item() * N / scale
        assert err < 1e-5, f"d loss / d {name}: {err}"
    assert maps.grad.abs().sum().item() > 0 and (maps.grad[:, 4:7] == 0).all()     # view_dep_rgb columns feed no term
    if bc is not None:
        assert (bcc.grad.double().cpu() - bc64.grad).abs().max().item() / scale < 1e-7


@pytest.mark.gpu
def test_fused_loss_on_a_real_training_render_matches_torch_loss(cuda):
    """the column-view fast path on the renderer's real output dict; gradients reach the model parameters"""
    from palettenerf_b200 import synthetic as S
    from palettenerf_b200.palette.losses import palette_loss
    gt = torch.rand(1, 1024, 3, device=cuda)
    o, d = S.training_rays(1024, seed=3)
    o, d = o.to(cuda)[None].contiguous(), d.to(cuda)[None].contiguous()
    grads = []
    for use_fused_loss in (False, True):
        torch.manual_seed(0)
        m = S.build_palette_model(cuda, seed=4, pred_clip=False, table_scale=0.5)
        m.train()
        with torch.autocast("cuda", dtype=torch.float16):
            out = m.render(o, d, staged=False, bg_color=1, perturb=False, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
            if use_fused_loss:
                from palettenerf_b200.palette import losses
                assert losses._column_view(out["omega_sparsity"], 1) is not None      # the fast path is what runs
                loss, _, _ = palette_loss(out, gt, 2e-4, 0.03, 0.1)
            else:
                loss = ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean() \
                    + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean()
        (loss * 128.0).backward()
        grads.append((loss.item(), m.encoder_palette.embeddings.grad.clone(), m.basis_color.grad.clone()))
    (l0, g0, b0), (l1, g1, b1) = grads
    assert abs(l0 - l1) < 2e-6 * abs(l0)
    assert (g0 - g1).abs().max().item() <= 1e-3 * g0.abs().max().item() + 1e-9     # fp32 atomics reorder + fp16 activations
    assert (b0 - b1).abs().max().item() <= 1e-3 * b0.abs().max().item() + 1e-9
