#!/usr/bin/env python
"""Golden vectors from the reference's own PaletteNetwork / PaletteRenderer.run_cuda / PaletteTrainer.train_step /
RegionEdit / Stylizer / NeRFNetwork, run UNMODIFIED on the reference's own CUDA kernels. TEST INFRASTRUCTURE.

Runs on the GPU box (the kernels need a device):
    gpurun -- 'python tests/golden/make_golden_palette.py'      ->  gpurun_out/ref_palette.npz  (then committed as
                                                                    tests/golden/ref_palette.npz)
The reference modules are imported by oracle/ref_python.py from oracle/_ref/py (staged by oracle/stage_ref_py.py) on top
of oracle/_ref/_ref_*.so (oracle/build_ref.py); third-party packages the image lacks come from compat/.
The models are this repository's synthetic models (tests/golden/palette_cases.py) loaded into the reference classes
with load_state_dict(strict=True) — which also pins the checkpoint layout.

What is stored (per model case `noclip` / `clip`; fp32 = no autocast, f16 = torch.autocast(float16) like `-O`):
  keys_<case>                       reference state_dict keys and shapes (JSON)
  fwd_<case>_{x,d}                  sample positions / directions fed to PaletteNetwork.forward (palette/network.py:156-185)
  fwd_<case>_<prec>_<name>          its six outputs
  eval_<case>_ds<scale>_<prec>_<k>  every key of the inference dict of run_cuda (palette/renderer.py:531-550), gui_mode=False
  edit_/style_<...>                 the same view with a RegionEdit / Stylizer active (palette/renderer.py:474-483)
  train_<case>_s<smooth>_<prec>_<k> every key of the training dict (:415-429), the loss and loss_dict of train_step
                                    (palette/utils.py:447-603, perturb=True with the shared noise stream), and the
                                    gradients of all small parameters + sampled rows of the hash-table gradients
  regionedit_*, stylizer_*          RegionEdit.forward / Stylizer.forward on fixed inputs (palette/renderer.py:121-183)
  nerf_*                            NeRFNetwork.forward and NeRFRenderer.run_cuda, eval and train (nerf/network.py:95-124,
                                    nerf/renderer.py:258-393)
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import palette_cases as PC  # noqa: E402
from oracle import ref_python  # noqa: E402


def np_(t):
    return t.detach().float().cpu().numpy()


def autocast(prec):
    return torch.autocast("cuda", dtype=torch.float16, enabled=(prec == "f16"))


def ref_palette_model(R, ours, pred_clip, density_scale=1.0):
    opt = PC.make_opt(pred_clip)
    m = R.network.PaletteNetwork(opt, bound=2.0, cuda_ray=True, min_near=0.2, density_thresh=10.0, density_scale=density_scale)
    missing = m.load_state_dict(ours.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.cuda()


def samples_for_forward(R, model, dev):
    rm = R.raymarching
    o, d = PC.eval_rays()
    o, d = o.to(dev), d.to(dev)
    nears, fars = rm.near_far_from_aabb(o, d, model.aabb_infer, model.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, model.bound, model.density_bitfield, model.cascade, model.grid_size,
                                                   nears, fars, counter, -1, False, 128, True, 0.0, 1024)
    n = int(counter[0].item())
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(5))[: PC.FWD_SAMPLES].to(dev)
    return xyzs[sel].contiguous(), dirs[sel].contiguous()


class FakeTrainer:
    """the attributes PaletteTrainer.train_step reads (palette/utils.py:447-603)"""

    def __init__(self, model, opt, smooth):
        self.model, self.opt = model, opt
        self.criterion = torch.nn.MSELoss(reduction="none")
        self.require_smooth_loss = smooth
        self.lambda_palette = opt.lambda_palette
        self.lambda_weight = 0.0
        self.error_map = None
        self.device = next(model.parameters()).device


def main():
    dev = torch.device("cuda:0")
    R = ref_python.load()
    out = {}
    for case, cfg in PC.MODEL_CASES.items():
        ours = PC.build_model(case, "cpu")
        clip = cfg["pred_clip"]
        m = ref_palette_model(R, ours, clip)
        # initialize_palette() (palette/renderer.py:248-262) adds basis_color_origin as an ALIAS of basis_color's storage
        # (nn.Parameter(self.basis_color.data)); the key list is recorded with it, the loss term is exercised de-aliased
        m.initialize_palette()
        out[f"keys_{case}"] = np.array(json.dumps({k: list(v.shape) for k, v in m.state_dict().items()}))
        m.basis_color_origin = torch.nn.Parameter(m.basis_color.data.clone() * 0.9 + 0.03, requires_grad=False)

        # ---- PaletteNetwork.forward ----
        m.eval()
        x, d = samples_for_forward(R, m, dev)
        out[f"fwd_{case}_x"], out[f"fwd_{case}_d"] = np_(x), np_(d)
        names = ["sigma", "clip", "omega", "offsets_radiance", "view_dep", "diffuse"]
        for prec in ("fp32", "f16"):
            with torch.no_grad(), autocast(prec):
                res = m(x, d)
            for n, t in zip(names, res):
                out[f"fwd_{case}_{prec}_{n}"] = np_(t)

        # ---- run_cuda, inference ----
        o, dd = PC.eval_rays()
        o, dd = o.to(dev)[None], dd.to(dev)[None]
        for ds in PC.DENSITY_SCALES:
            m.density_scale = ds
            for prec in ("fp32", "f16"):
                with torch.no_grad(), autocast(prec):
                    res = m.render(o, dd, staged=True, bg_color=1, perturb=False, gui_mode=False, **PC.RENDER_KW)
                for k, t in res.items():
                    out[f"eval_{case}_ds{int(ds)}_{prec}_{k}"] = np_(t)
        # GUI edits (ds = 40 so that the recoloured surface is what the image shows)
        m.density_scale = 40.0
        edit = R.renderer.RegionEdit(m.opt)
        rgb_orig = m.basis_color.detach().clamp(0, 1)
        rgb_new = torch.tensor([[0.2, 0.7, 0.3], [0.25, 0.2, 0.3], [0.9, 0.6, 0.7], [0.1, 0.3, 0.8]], device=dev)
        edit.update_delta_hsv(rgb_orig, rgb_new)
        edit.update_cent(mean_xyz=torch.tensor([0.2, 0.1, 0.0], device=dev),
                         mean_clip=(torch.linspace(-0.2, 0.2, 16, device=dev) if clip else None))
        edit.update_std(std_xyz=0.15, std_clip=0.5)
        out[f"edit_{case}_delta_hsv"] = np_(edit.delta_hsv)
        m.edit = edit
        for prec in ("fp32", "f16"):
            with torch.no_grad(), autocast(prec):
                res = m.render(o, dd, staged=True, bg_color=1, perturb=False, gui_mode=False, **PC.RENDER_KW)
            for k, t in res.items():
                out[f"edit_{case}_{prec}_{k}"] = np_(t)
        m.edit = None
        sty = R.renderer.Stylizer(m.opt).to(dev)
        with torch.no_grad():
            sty.dI.copy_(torch.tensor([0.1, -0.05, 0.2, 0.0]))
            sty.dP.copy_(torch.tensor([[[0.05, -0.1, 0.0], [0.1, 0.1, 0.1], [-0.2, 0.0, 0.05], [0.0, 0.15, -0.05]]]))
            sty.ddelta.add_(0.1 * torch.randn(4, 3, 3, generator=torch.Generator().manual_seed(3)).to(dev))
        out[f"style_{case}_ddelta"] = np_(sty.ddelta)
        m.stylizer = sty
        for prec in ("fp32", "f16"):
            with torch.no_grad(), autocast(prec):
                res = m.render(o, dd, staged=True, bg_color=1, perturb=False, gui_mode=True, **PC.RENDER_KW)
            for k, t in res.items():
                out[f"style_{case}_{prec}_{k}"] = np_(t)
        m.stylizer = None
        m.density_scale = 1.0

        # ---- run_cuda (training) through PaletteTrainer.train_step ----
        m.train()
        to, td = PC.train_rays()
        gt, feat = PC.train_targets(clip)
        data = {"rays_o": to.to(dev)[None], "rays_d": td.to(dev)[None], "images": gt.to(dev)}
        if clip:
            data["feat_images"] = feat.to(dev)
        gidx = PC.table_grad_indices(m.encoder.embeddings.shape[0]).to(dev)
        for smooth in (0, 1):
            m.require_smooth_loss = bool(smooth)
            tr = FakeTrainer(m, m.opt, bool(smooth))
            for prec in ("fp32", "f16"):
                for p in m.parameters():
                    p.grad = None
                captured = {}
                render = m.render

                def spy(*a, _render=render, **k):
                    captured["out"] = _render(*a, **k)
                    return captured["out"]
                m.render = spy
                with PC.FixedRandom(), autocast(prec):
                    _, _, loss, loss_dict = R.utils.PaletteTrainer.train_step(tr, dict(data))
                del m.render
                (loss * PC.GRAD_SCALE).backward()
                tag = f"train_{case}_s{smooth}_{prec}"
                for k, t in captured["out"].items():
                    out[f"{tag}_{k}"] = np_(t)
                out[f"{tag}_loss"] = np.float32(loss.item())
                for k, v in loss_dict.items():
                    out[f"{tag}_{k}"] = np.float32(float(v))
                for n, p in m.named_parameters():
                    if p.grad is None:
                        continue
                    g = p.grad.detach().float() / PC.GRAD_SCALE
                    if p.numel() <= 8192:
                        out[f"{tag}_grad_{n}"] = np_(g)
                    else:
                        out[f"{tag}_gradrows_{n}"] = np_(g[gidx])
                        out[f"{tag}_gradnorm_{n}"] = np.array([g.double().norm().item(), g.double().abs().sum().item(),
                                                               float((g != 0).any(dim=1).sum().item())])
        m.require_smooth_loss = False

        # ---- RegionEdit.forward / Stylizer.forward on fixed inputs ----
        if case == "clip":
            g = torch.Generator().manual_seed(21)
            M = 512
            final = torch.rand(M, 4, 3, generator=g).to(dev) * 1.2
            xyz = (torch.rand(M, 3, generator=g).to(dev) - 0.5)
            cf = (torch.randn(M, 16, generator=g) * 0.2).to(dev)
            out["regionedit_in_final"], out["regionedit_in_xyz"], out["regionedit_in_clip"] = np_(final), np_(xyz), np_(cf)
            with torch.no_grad():
                out["regionedit_out"] = np_(edit(final, xyz, cf))
                edit.weight_mode = True
                out["regionedit_out_weight_mode"] = np_(edit(final, xyz, cf))
                edit.weight_mode = False
                rad = torch.randn(M, 1, generator=g).to(dev)
                om = torch.softmax(torch.randn(M, 4, generator=g), -1).to(dev)
                off = (torch.randn(M, 4, 3, generator=g) * 0.2).to(dev)
                vd = torch.rand(M, 3, generator=g).to(dev) * 0.3
                out["stylizer_in_radiance"], out["stylizer_in_omega"], out["stylizer_in_offsets"] = np_(rad), np_(om), np_(off)
                out["stylizer_in_view_dep"] = np_(vd)
                out["stylizer_out"] = np_(sty(rad.reshape(M, 1, 1), om.reshape(M, 4, 1), rgb_orig[None], off, vd))
        del m
        torch.cuda.empty_cache()

    # ---- stage-1 model: NeRFNetwork / NeRFRenderer.run_cuda ----
    from palettenerf_b200 import synthetic as S
    ours = S.build_nerf_model("cpu", seed=4, table_scale=0.5)
    nm = R.nerf_network.NeRFNetwork(encoding="hashgrid", bound=2.0, cuda_ray=True, density_scale=1, min_near=0.2,
                                    density_thresh=10.0, bg_radius=-1)
    nm.load_state_dict(ours.state_dict(), strict=True)
    nm = nm.cuda()
    out["keys_nerf"] = np.array(json.dumps({k: list(v.shape) for k, v in nm.state_dict().items()}))
    nm.eval()
    x, d = samples_for_forward(R, nm, dev)
    out["nerf_fwd_x"], out["nerf_fwd_d"] = np_(x), np_(d)
    o, dd = PC.eval_rays()
    o, dd = o.to(dev)[None], dd.to(dev)[None]
    for prec in ("fp32", "f16"):
        with torch.no_grad(), autocast(prec):
            s, c = nm(x, d)
        out[f"nerf_fwd_{prec}_sigma"], out[f"nerf_fwd_{prec}_color"] = np_(s), np_(c)
        for ds in PC.DENSITY_SCALES:
            nm.density_scale = ds
            with torch.no_grad(), autocast(prec):
                res = nm.render(o, dd, staged=True, bg_color=1, perturb=False, **PC.RENDER_KW)
            for k, t in res.items():
                out[f"nerf_eval_ds{int(ds)}_{prec}_{k}"] = np_(t)
    nm.density_scale = 1.0
    nm.train()
    to, td = PC.train_rays()
    gt, _ = PC.train_targets(False)
    for prec in ("fp32", "f16"):
        for p in nm.parameters():
            p.grad = None
        with PC.FixedRandom(), autocast(prec):
            res = nm.render(to.to(dev)[None], td.to(dev)[None], rays_gt=gt.to(dev), staged=False, bg_color=1, perturb=True,
                            force_all_rays=True, **PC.RENDER_KW)
            loss = ((res["image"] - gt.to(dev)) ** 2).mean()
        (loss * PC.GRAD_SCALE).backward()
        for k, t in res.items():
            out[f"nerf_train_{prec}_{k}"] = np_(t)
        out[f"nerf_train_{prec}_loss"] = np.float32(loss.item())
        gidx = PC.table_grad_indices(nm.encoder.embeddings.shape[0]).to(dev)
        for n, p in nm.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach().float() / PC.GRAD_SCALE
            if p.numel() <= 8192:
                out[f"nerf_train_{prec}_grad_{n}"] = np_(g)
            else:
                out[f"nerf_train_{prec}_gradrows_{n}"] = np_(g[gidx])
                out[f"nerf_train_{prec}_gradnorm_{n}"] = np.array([g.double().norm().item(), g.double().abs().sum().item(),
                                                                   float((g != 0).any(dim=1).sum().item())])

    dst = os.environ.get("PNERF_GOLDEN_OUT", os.path.join(ROOT, "gpurun_out", "ref_palette.npz"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print(f"[make_golden_palette] {len(out)} arrays, {os.path.getsize(dst) / 1e6:.2f} MB -> {dst}")


if __name__ == "__main__":
    main()
