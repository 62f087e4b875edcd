"""Parity against the REFERENCE's own PaletteNetwork / run_cuda / train_step / RegionEdit / Stylizer / NeRFNetwork.

tests/golden/ref_palette.npz holds outputs of the unmodified reference Python running on the reference's own CUDA
kernels on a B200 (tests/golden/make_golden_palette.py). Inputs are rebuilt from seeds (tests/golden/palette_cases.py).

CPU half (this file, not marked gpu): the oracle restatements (oracle/cpu_render.py) against those outputs ->
the oracle for the field, the blend, the renderers and the loss is PINNED by the reference, not by itself.
GPU half: tests/test_golden_palette_gpu.py runs the CUDA path against the same vectors.

Tolerances (stated per output, two-sided):
  vs the reference's fp32 run : 5e-5 abs on O(1) outputs (both sides fp32; exp / sums reassociate), 2e-4 relative on sigma
  vs the reference's fp16 run : 1e-3 abs (north_star's fp16 bar) — the reference's own fp16 rounding is the difference
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import palette_cases as PC  # noqa: E402
from oracle import cpu_render  # noqa: E402

GOLD = os.path.join(HERE, "golden", "ref_palette.npz")
FWD_NAMES = ["sigma", "clip", "omega", "offsets_radiance", "view_dep", "diffuse"]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module", params=list(PC.MODEL_CASES), ids=list(PC.MODEL_CASES))
def case(request):
    name = request.param
    m = PC.build_model(name, "cpu")
    params = {k: v.detach() for k, v in m.state_dict().items()}
    return name, PC.MODEL_CASES[name], m, params


def _close(a, r, tol, what, rel=False):
    a = np.asarray(a, np.float64).reshape(r.shape)
    err = np.abs(a - r)
    if rel:
        err = err / np.maximum(np.abs(r), 1e-3)
    assert err.max() <= tol, f"{what}: max {'rel' if rel else 'abs'} err {err.max():.3e} > {tol:g}"


def test_field_oracle_matches_reference_forward(gold, case):
    name, cfg, m, params = case
    out = cpu_render.palette_forward(params, gold[f"fwd_{name}_x"], gold[f"fwd_{name}_d"], 2.0, m.encoder.per_level_scale,
                                     cfg["pred_clip"])
    for n, t in zip(FWD_NAMES, out):
        _close(t.numpy(), gold[f"fwd_{name}_fp32_{n}"], 2e-4 if n == "sigma" else 5e-5, f"{name}/{n} vs ref fp32", rel=n == "sigma")
        _close(t.numpy(), gold[f"fwd_{name}_f16_{n}"], 1e-3, f"{name}/{n} vs ref f16", rel=n == "sigma")
    assert gold[f"fwd_{name}_fp32_view_dep"].std() > 1e-3 and gold[f"fwd_{name}_fp32_offsets_radiance"].std() > 1e-2   # not vacuous


EVAL_KEYS = ["image", "depth", "depth_origin", "weights_sum", "clip_feat", "direct_rgb", "view_dep_rgb", "basis_rgb",
             "unscaled_basis_rgb", "basis_acc"]


@pytest.mark.parametrize("ds", [1, 40])
def test_render_oracle_matches_reference_run_cuda(gold, case, ds):
    name, cfg, m, params = case
    o, d = PC.eval_rays()
    out = cpu_render.render_cuda_ray(params, o, d, m.density_bitfield, pred_clip=cfg["pred_clip"], density_scale=float(ds),
                                     **PC.RENDER_KW)
    for k in EVAL_KEYS:
        r = gold[f"eval_{name}_ds{ds}_fp32_{k}"]
        scale = max(1.0, np.abs(r).max())
        _close(out[k], r, 5e-5 * scale, f"{name}/ds{ds}/{k} vs ref fp32")
        _close(out[k], gold[f"eval_{name}_ds{ds}_f16_{k}"], 1e-3 * scale, f"{name}/ds{ds}/{k} vs ref f16")
    ws = gold[f"eval_{name}_ds{ds}_fp32_weights_sum"]
    assert ws.max() > (0.99 if ds == 40 else 0.2)          # ds 40: rays terminate inside the solid


def _edit_dict(gold, name, clip):
    return dict(delta_hsv=gold[f"edit_{name}_delta_hsv"], mean_xyz=np.array([0.2, 0.1, 0.0], np.float32),
                mean_clip=np.linspace(-0.2, 0.2, 16, dtype=np.float32) if clip else None, std_xyz=0.15, std_clip=0.5)


def _style_dict(gold, name):
    return dict(dI=np.array([0.1, -0.05, 0.2, 0.0], np.float32), ddelta=gold[f"style_{name}_ddelta"],
                dP=np.array([[[0.05, -0.1, 0.0], [0.1, 0.1, 0.1], [-0.2, 0.0, 0.05], [0.0, 0.15, -0.05]]], np.float32))


def test_render_oracle_matches_reference_with_region_edit_and_stylizer(gold, case):
    name, cfg, m, params = case
    o, d = PC.eval_rays()
    out = cpu_render.render_cuda_ray(params, o, d, m.density_bitfield, pred_clip=cfg["pred_clip"], density_scale=40.0,
                                     edit=_edit_dict(gold, name, cfg["pred_clip"]), **PC.RENDER_KW)
    plain = gold[f"eval_{name}_ds40_fp32_image"]
    assert np.abs(gold[f"edit_{name}_fp32_image"] - plain).max() > 0.05          # the edit recolours the view
    for k in EVAL_KEYS:
        r = gold[f"edit_{name}_fp32_{k}"]
        # the reference's HSV kernels are built with -use_fast_math (palette/setup.py): 3e-4 abs on the recoloured maps
        _close(out[k], r, 3e-4 * max(1.0, np.abs(r).max()), f"{name}/edit/{k} vs ref fp32")
    out = cpu_render.render_cuda_ray(params, o, d, m.density_bitfield, pred_clip=cfg["pred_clip"], density_scale=40.0,
                                     stylizer=_style_dict(gold, name), gui_mode=True, **PC.RENDER_KW)
    assert np.abs(gold[f"style_{name}_fp32_image"] - plain).max() > 0.05
    for k in ["image", "depth", "weights_sum", "clip_feat"]:
        r = gold[f"style_{name}_fp32_{k}"]
        _close(out[k], r, 5e-5 * max(1.0, np.abs(r).max()), f"{name}/style/{k} vs ref fp32")


def test_region_edit_and_stylizer_oracle_match_reference_modules(gold):
    edit = _edit_dict(gold, "clip", True)
    out = cpu_render.region_edit(edit, gold["regionedit_in_final"], gold["regionedit_in_xyz"], gold["regionedit_in_clip"])
    _close(out.numpy(), gold["regionedit_out"], 3e-4, "RegionEdit.forward")
    wm = cpu_render.region_edit(dict(edit, weight_mode=True), gold["regionedit_in_final"], gold["regionedit_in_xyz"],
                                gold["regionedit_in_clip"])
    _close(wm.numpy(), gold["regionedit_out_weight_mode"], 1e-6, "RegionEdit.forward weight_mode")
    m = PC.build_model("clip", "cpu")
    pal = m.basis_color.detach().clamp(0, 1)[None]
    out = cpu_render.stylize(_style_dict(gold, "clip"), gold["stylizer_in_radiance"], torch.from_numpy(gold["stylizer_in_omega"]),
                             pal, gold["stylizer_in_offsets"], gold["stylizer_in_view_dep"])
    _close(out.numpy(), gold["stylizer_out"], 2e-6, "Stylizer.forward")


TRAIN_KEYS = ["image", "depth", "weights_sum", "omega_sparsity", "view_dep_norm", "offsets_norm", "smooth_norm", "view_dep_rgb",
              "direct_rgb", "diffuse_rgb", "clip_feat", "basis_acc"]


def oracle_train_outputs(name, cfg, m, params, smooth):
    """the oracle's training forward on the golden case -> dict with the reference's result keys (palette/renderer.py:415-429)"""
    o, d = PC.train_rays()
    with PC.FixedRandom():
        noises = torch.rand(PC.TRAIN_RAYS).numpy()                # what march_rays_train draws with perturb=True
    res = cpu_render.train_forward_cuda_ray(params, o, d, m.density_bitfield, pred_clip=cfg["pred_clip"], noises=noises,
                                            smooth=bool(smooth), jitter_fn=PC.hash_uniform, **PC.RENDER_KW)
    maps, cd = res["maps"], 16
    ws = res["weights_sum"]
    return {"image": res["image"], "depth": res["depth"], "weights_sum": ws, "omega_sparsity": maps[:, 0], "view_dep_norm": maps[:, 1],
            "offsets_norm": maps[:, 2], "smooth_norm": maps[:, 3], "view_dep_rgb": maps[:, 4:7],
            "direct_rgb": maps[:, 7:10] + (1 - ws)[:, None] * 1.0, "diffuse_rgb": maps[:, 10:13], "clip_feat": maps[:, 13:13 + cd],
            "basis_acc": maps[:, 13 + cd:13 + cd + 4]}


@pytest.mark.parametrize("smooth", [0, 1])
def test_train_oracle_and_loss_match_reference_train_step(gold, case, smooth):
    name, cfg, m, params = case
    out = oracle_train_outputs(name, cfg, m, params, smooth)
    tag = f"train_{name}_s{smooth}"
    for k in TRAIN_KEYS:
        r = gold[f"{tag}_fp32_{k}"]
        scale = max(1.0, np.abs(r).max())
        _close(out[k], r, 5e-5 * scale, f"{tag}/{k} vs ref fp32")
        _close(out[k], gold[f"{tag}_f16_{k}"], 1e-3 * scale, f"{tag}/{k} vs ref f16")
    if smooth:
        assert gold[f"{tag}_fp32_smooth_norm"].max() > 1e-6          # the smooth branch is live
    # the loss block of PaletteTrainer.train_step on the REFERENCE's own maps -> its loss and loss_dict
    outs = {k: torch.from_numpy(gold[f"{tag}_fp32_{k}"]) for k in TRAIN_KEYS}
    gt, feat = PC.train_targets(cfg["pred_clip"])
    bc = m.basis_color.detach()
    lam = PC.LAMBDAS
    loss, terms, _ = cpu_render.palette_train_loss(
        outs, gt, lambda_sparsity=lam["lambda_sparsity"], lambda_offsets=lam["lambda_offsets"],
        lambda_view_dep=lam["lambda_view_dep"], lambda_smooth=lam["lambda_smooth"] if smooth else 0.0, gt_clip_feat=feat,
        basis_color=bc, basis_color_origin=bc * 0.9 + 0.03, lambda_palette=lam["lambda_palette"])
    assert abs(float(loss) - float(gold[f"{tag}_fp32_loss"])) <= 2e-6 * max(1.0, abs(float(gold[f"{tag}_fp32_loss"])))
    for k_ref, k in (("loss_sparsity", "sparsity"), ("loss_offsets", "offsets"), ("loss_view_dep", "view_dep"),
                     ("loss_smooth", "smooth"), ("loss_palette", "palette"), ("loss_direct", "direct")):
        assert abs(float(terms[k]) - float(gold[f"{tag}_fp32_{k_ref}"])) <= 1e-6 + 1e-5 * abs(float(gold[f"{tag}_fp32_{k_ref}"])), k
    if cfg["pred_clip"]:
        assert abs(float(terms["clip_feat"]) - float(gold[f"{tag}_fp32_loss_clip_feat"])) <= 1e-6


def test_nerf_oracle_matches_reference_forward(gold):
    from palettenerf_b200 import synthetic as S
    m = S.build_nerf_model("cpu", seed=4, table_scale=0.5)
    params = {k: v.detach() for k, v in m.state_dict().items()}
    sigma, rgb = cpu_render.nerf_forward(params, gold["nerf_fwd_x"], gold["nerf_fwd_d"], 2.0, m.encoder.per_level_scale)
    _close(sigma.numpy(), gold["nerf_fwd_fp32_sigma"], 2e-4, "nerf sigma vs ref fp32", rel=True)
    _close(rgb.numpy(), gold["nerf_fwd_fp32_color"], 5e-5, "nerf colour vs ref fp32")
    _close(rgb.numpy(), gold["nerf_fwd_f16_color"], 1e-3, "nerf colour vs ref f16")


def test_nerf_train_oracle_matches_reference_step_outputs_and_gradients(gold):
    """stage-1 training branch (nerf/renderer.py:282-330): the oracle's maps and — through the oracle's composite backward and
    torch autograd of the restated field — the gradients of the reference's own step (`nerf_train_fp32_*`)"""
    from palettenerf_b200 import synthetic as S
    m = S.build_nerf_model("cpu", seed=4, table_scale=0.5)
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = [n for n, _ in m.named_parameters()]
    for n in names:
        params[n].requires_grad_()
    o, d = PC.train_rays()
    gt, _ = PC.train_targets(False)
    with PC.FixedRandom():
        noises = torch.rand(PC.TRAIN_RAYS).numpy()
    gt_flat = gt.reshape(-1, 3)
    maps, grads = cpu_render.nerf_train_step_cuda_ray(
        params, o, d, gt, m.density_bitfield, per_level_scale=m.encoder.per_level_scale, noises=noises,
        loss_fn=lambda r: ((r["image"] - gt_flat) ** 2).mean(), **PC.RENDER_KW)
    for k in ("image", "depth", "weights_sum", "rgb_norm"):
        r = gold[f"nerf_train_fp32_{k}"].reshape(maps[k].shape)
        _close(maps[k], r, 5e-5, f"nerf train {k} vs ref fp32")
        _close(maps[k], gold[f"nerf_train_f16_{k}"].reshape(maps[k].shape), 1e-3, f"nerf train {k} vs ref f16")
    checked = 0
    for n in names:
        g = grads[n].double().numpy()
        if f"nerf_train_fp32_grad_{n}" in gold.files:
            a, r = g.reshape(-1), gold[f"nerf_train_fp32_grad_{n}"].astype(np.float64).reshape(-1)
        else:
            idx = PC.table_grad_indices(g.shape[0]).numpy()
            a, r = g[idx].reshape(-1), gold[f"nerf_train_fp32_gradrows_{n}"].astype(np.float64).reshape(-1)
            assert abs(np.linalg.norm(g) - gold[f"nerf_train_fp32_gradnorm_{n}"][0]) <= 1e-3 * gold[f"nerf_train_fp32_gradnorm_{n}"][0]
        rel = np.linalg.norm(a - r) / np.linalg.norm(r)
        # the sigma gradient of the compositor is a difference of nearly equal sums (T c_i - remaining colour): the 1e-6 the
        # CPU matmuls differ from the GPU's by comes out as ~1e-3 on everything upstream of sigma (measured: table 1.0e-3,
        # sigma_net.0 3.8e-4; the colour net, which sees no cancellation, 7e-5 and below). The CUDA fp32 path, which shares the
        # reference's GEMM kernels, is held to 2e-4 in test_golden_palette_gpu.py.
        assert rel <= (3e-3 if n.startswith(("encoder", "sigma_net")) else 3e-4), (n, rel)
        checked += 1
    assert checked == 6


# ------------------------------------------------------------------------------------------------------------------
# checkpoint layout (a20 / f4): key set + shapes == the reference's, and a reference-layout dict loads
# ------------------------------------------------------------------------------------------------------------------
def test_state_dict_keys_and_shapes_equal_the_reference(gold, case):
    name, cfg, m, _ = case
    ref = json.loads(str(gold[f"keys_{name}"]))
    m.initialize_palette()                                   # the reference records its keys after initialize_palette()
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert ours == ref
    assert list(ours) == list(ref)                           # same order, too (torch.save keeps it)


def test_nerf_state_dict_equals_reference_and_loads_into_the_palette_model(gold):
    from palettenerf_b200 import synthetic as S
    nerf = S.build_nerf_model("cpu", seed=4, table_scale=0.5)
    ref = json.loads(str(gold["keys_nerf"]))
    assert {k: list(v.shape) for k, v in nerf.state_dict().items()} == ref
    # stage hand-off (palette/utils.py:1306-1318): non-strict load of the NeRF checkpoint, NO unexpected keys allowed
    pal = PC.build_model("noclip", "cpu")
    before = pal.encoder_palette.embeddings.detach().clone()
    missing, unexpected = pal.load_state_dict(nerf.state_dict(), strict=False)
    assert not unexpected
    assert any(k.startswith("encoder_palette") for k in missing) and "basis_color" in missing
    assert torch.equal(pal.encoder.embeddings, nerf.encoder.embeddings)
    assert torch.equal(pal.sigma_net[0].weight, nerf.sigma_net[0].weight)
    assert torch.equal(pal.color_net[2].weight, nerf.color_net[2].weight)
    assert torch.equal(pal.density_bitfield, nerf.density_bitfield)
    assert torch.equal(pal.encoder_palette.embeddings, before)          # untouched by the hand-off


def test_reference_layout_dict_round_trips_strict(gold, case):
    name, cfg, m, _ = case
    ref = json.loads(str(gold[f"keys_{name}"]))
    g = torch.Generator().manual_seed(0)
    sd = {}
    for k, shp in ref.items():
        like = m.state_dict().get(k)
        dt = like.dtype if like is not None else torch.float32
        sd[k] = torch.randint(0, 100, shp, generator=g).to(dt) if not dt.is_floating_point else torch.rand(shp, generator=g).to(dt)
    m2 = PC.build_model(name, "cpu")
    m2.initialize_palette()
    res = m2.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m2.state_dict().items():
        # initialize_palette() makes basis_color_origin an alias of basis_color's storage (palette/renderer.py:258, kept):
        # the later key of the dict wins for both, in the reference exactly as here
        assert torch.equal(v, sd["basis_color_origin" if k == "basis_color" else k]), k


def test_render_oracle_matches_reference_on_rays_of_the_800x800_view():
    """tests/golden/ref_view800.npz: the reference's own render of the full 800 x 800 view; the oracle renders the pinned
    rays (every 131st; rays are independent) — a second, larger set of rays than the 32 x 32 view, same bars."""
    import make_golden_view800 as V
    g = np.load(os.path.join(HERE, "golden", "ref_view800.npz"))
    stride = int(g["stride"])
    m = PC.build_model("noclip", "cpu")
    params = {k: v.detach() for k, v in m.state_dict().items()}
    o, d = V.view_rays()
    o, d = o[::stride].contiguous(), d[::stride].contiguous()
    for ds in (40,):            # terminating rays: ~6 s on one core (ds 1 marches every ray to its end: covered on the GPU)
        out = cpu_render.render_cuda_ray(params, o, d, m.density_bitfield, pred_clip=False, density_scale=float(ds), **PC.RENDER_KW)
        for k in EVAL_KEYS:
            r = g[f"ds{ds}_fp32_{k}_rows"]
            scale = max(1.0, np.abs(r).max())
            _close(out[k], r, 5e-5 * scale, f"view800/ds{ds}/{k} vs ref fp32")
            _close(out[k], g[f"ds{ds}_f16_{k}_rows"], 1e-3 * scale, f"view800/ds{ds}/{k} vs ref f16")
