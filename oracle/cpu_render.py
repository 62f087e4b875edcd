"""CPU restatement (plain torch fp32) of the reference's PyTorch-level path. TEST INFRASTRUCTURE / CPU BASELINE ONLY.

  * torch_grid_encode      — torch transliteration of gridencoder.cu:54-72,97-175 (multi-threaded via torch ops)
  * torch_sh               — SH basis via the shared derivation in tools/gen_sh.py (cf. testing/test_shencoder.py:8-89)
  * palette_forward        — PaletteNetwork.forward / color / density, palette/network.py:156-280
  * nerf_forward           — NeRFNetwork.forward, nerf/network.py:95-124
  * blend                  — the palette blend of palette/renderer.py:470-494 (inference) / :333-352,357-359 (train)
  * render_sampler         — the pure-torch renderer of NeRFRenderer.run, nerf/renderer.py:127-255, with num_steps=512,
                             upsample_steps=0 (main_palette.py:33-34) — BASELINE config 1 (the "reference PyTorch path")
  * render_cuda_ray        — the cuda_ray inference schedule of palette/renderer.py:430-552 driven by the C oracle kernels
  * train_forward_cuda_ray — the cuda_ray training branch of palette/renderer.py:322-429 on the C oracle kernels
  * region_edit / stylize  — RegionEdit.forward / Stylizer.forward, palette/renderer.py:121-147, 166-183
  * random_palette_params  — a random-init state_dict of the architecture built WITHOUT the product package (bench.py's
                             reference arm must not map libpnerf_b200.so)
All take a `params` dict = the model's state_dict (numpy or torch CPU tensors, fp32).

PARITY PINNING: palette_forward, blend, render_cuda_ray (incl. RegionEdit / Stylizer), train_forward_cuda_ray (incl. the
smooth-loss branch) and palette_train_loss are checked against outputs of the REFERENCE ITSELF — its unmodified Python on
its own CUDA kernels, run on a B200 by tests/golden/make_golden_palette.py -> tests/golden/ref_palette.npz — in
tests/test_golden_palette.py (fp32 reference run: <= 5e-5; the reference's fp16-autocast run: <= 1e-3).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import cpu_oracle as O

_PRIMES = [1, 2654435761, 805459861]


def _t(a):
    return a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))


def torch_grid_encode(x01, emb, offsets, per_level_scale, H=16):
    """x01 [B,3] in [0,1] fp32; emb [n,2] fp32; hash gridtype, align_corners False -> [B, L*2] fp32"""
    x01, emb = _t(x01).float(), _t(emb).float()
    offsets = [int(v) for v in offsets]
    L = len(offsets) - 1
    S = np.float32(np.log2(per_level_scale))
    B = x01.shape[0]
    outs = []
    oob = ((x01 < 0) | (x01 > 1)).any(dim=1)
    for level in range(L):
        size = offsets[level + 1] - offsets[level]
        e = np.exp2(np.float32(np.float32(level) * S)).astype(np.float32)
        scale = np.float32(np.float64(e) * float(H) - 1.0)
        res = int(np.ceil(scale)) + 1
        pos = x01 * float(scale) + 0.5
        pg = torch.floor(pos)
        fr = pos - pg
        pg = pg.long()
        dense = (res + 1) ** 3 <= size
        acc = torch.zeros(B, emb.shape[1])
        for c in range(8):
            w = torch.ones(B)
            loc = []
            for dd in range(3):
                bit = (c >> dd) & 1
                w = w * (fr[:, dd] if bit else (1 - fr[:, dd]))
                loc.append(pg[:, dd] + bit)
            if dense:
                idx = loc[0] + loc[1] * (res + 1) + loc[2] * (res + 1) ** 2
            else:
                idx = (loc[0] * _PRIMES[0]) ^ ((loc[1] * _PRIMES[1]) & 0xFFFFFFFF) ^ ((loc[2] * _PRIMES[2]) & 0xFFFFFFFF)
            idx = (idx % size) + offsets[level]
            acc = acc + w[:, None] * emb[idx]
        outs.append(acc)
    out = torch.cat(outs, dim=1)
    out[oob] = 0
    return out


def torch_sh(d, degree=4):
    return torch.from_numpy(O.sh_encode(_t(d).double().numpy(), degree)).float()


def _mlp(params, prefix, n, h, act=F.relu):
    for i in range(n):
        h = F.linear(h, _t(params[f"{prefix}.{i}.weight"]).float())
        if i != n - 1:
            h = act(h)
    return h


def _encode(params, name, x, bound, per_level_scale):
    return torch_grid_encode((x + bound) / (2 * bound), params[f"{name}.embeddings"], _t(params[f"{name}.offsets"]).tolist(),
                             per_level_scale)


def palette_forward(params, x, d, bound, per_level_scale, pred_clip, clip_dim=16):
    """-> (sigma, clip_feat, omega, offsets_radiance, view_dep, diffuse); palette/network.py:156-185,223-280"""
    x, d = _t(x).float(), _t(d).float()
    h = _mlp(params, "sigma_net", 2, _encode(params, "encoder", x, bound, per_level_scale))
    sigma = torch.exp(h[..., 0])
    geo = h[..., 1:]
    if pred_clip:
        clip = _mlp(params, "clip_net", 2, _encode(params, "encoder_clip", x, bound, per_level_scale))
    else:
        clip = torch.zeros(x.shape[0], clip_dim)
    diffuse = torch.sigmoid(_mlp(params, "diff_net", 3, geo))
    view_dep = torch.sigmoid(_mlp(params, "color_net", 3, torch.cat([torch_sh(d, 4), geo], dim=-1)))
    hp = torch.cat([_encode(params, "encoder_palette", x, bound, per_level_scale), diffuse], dim=-1)
    hp = _mlp(params, "basis_net", 2, hp, act=F.elu)
    offsets_radiance = F.linear(hp, _t(params["offsets_radiance_net.weight"]).float(), _t(params["offsets_radiance_net.bias"]).float())
    omega = F.softplus(F.linear(hp, _t(params["omega_net.0.weight"]).float())) + 0.05
    omega = omega / omega.sum(dim=-1, keepdim=True)
    return sigma, clip, omega, offsets_radiance, view_dep, diffuse


def nerf_forward(params, x, d, bound, per_level_scale):
    """-> (sigma, rgb); nerf/network.py:95-124"""
    x, d = _t(x).float(), _t(d).float()
    h = _mlp(params, "sigma_net", 2, _encode(params, "encoder", x, bound, per_level_scale))
    sigma = torch.exp(h[..., 0])
    rgb = torch.sigmoid(_mlp(params, "color_net", 3, torch.cat([torch_sh(d, 4), h[..., 1:]], dim=-1)))
    return sigma, rgb


def blend(params, omega, offsets_radiance, view_dep, num_basis=4, offsets_weight=1.0, view_dep_weight=1.0, edit=None,
          xyzs=None, clip=None):
    """palette/renderer.py:470-494: -> rgbs [M,3], basis_rgb [M,Nb,3], unscaled_basis_rgb [M,Nb,3]"""
    M = omega.shape[0]
    offsets = offsets_radiance[..., :-1].reshape(M, num_basis, 3)
    radiance = offsets_radiance[..., -1:].reshape(M, 1, 1)
    palette = _t(params["basis_color"]).float()[None].clamp(0, 1)
    final = F.softplus(radiance) * (palette + offsets_weight * offsets)          # :477
    if edit is not None:
        final = region_edit(edit, final, xyzs, clip)                              # :481-482
    basis_rgb = omega.reshape(M, num_basis, 1) * final
    return basis_rgb.sum(dim=-2) + view_dep_weight * view_dep, basis_rgb, (palette + offsets)


def region_edit(edit, rgbs, xyz=None, clip_feat=None):
    """RegionEdit.forward, palette/renderer.py:121-147. edit: dict(delta_hsv [Nb,3], mean_xyz [3]|None, mean_clip [cd]|None,
    std_xyz, std_clip, weight_mode). rgbs [M,Nb,3] -> [M,Nb,3]"""
    rgbs = _t(rgbs).float()
    M, nb, _ = rgbs.shape
    hsv = torch.from_numpy(O.rgb_to_hsv(rgbs.reshape(-1, 3).numpy())).reshape(M, nb, 3)
    dh = _t(edit["delta_hsv"]).float()
    weight = torch.ones(M, 1)
    if xyz is not None and edit.get("mean_xyz") is not None:
        weight = weight * torch.exp(-((_t(xyz).float() - _t(edit["mean_xyz"]).float()[None]) ** 2.).sum(-1, keepdim=True)
                                    / edit.get("std_xyz", 1))
    if clip_feat is not None and edit.get("mean_clip") is not None:
        weight = weight * torch.exp(-((_t(clip_feat).float() - _t(edit["mean_clip"]).float()[None]) ** 2.).sum(-1, keepdim=True)
                                    / edit.get("std_clip", 1))
    new = torch.stack([torch.fmod(hsv[..., 0] + dh[..., 0] + 360, 360), torch.clip(hsv[..., 1] * dh[..., 1], 0),
                       torch.clip(hsv[..., 2] * dh[..., 2], 0)], dim=-1)
    rgb_new = torch.from_numpy(O.hsv_to_rgb(new.reshape(-1, 3).numpy())).reshape(M, nb, 3)
    if edit.get("weight_mode"):
        return weight[..., None].repeat(1, nb, 3)
    return torch.lerp(rgbs, rgb_new, weight[..., None])


def stylize(sty, radiance, omega, palette, offsets, view_dep=None):
    """Stylizer.forward, palette/renderer.py:166-183. sty: dict(dI [Nb], dP [1,Nb,3], ddelta [Nb,3,3])"""
    nb = omega.shape[-1] if omega.dim() == 2 else omega.shape[-2]
    radiance, omega = _t(radiance).float().reshape(-1, 1, 1), _t(omega).float().reshape(-1, nb, 1)
    palette = _t(palette).float().reshape(-1, nb, 3) + _t(sty["dP"]).float()
    offsets = torch.einsum("npi,pij->npj", _t(offsets).float().reshape(-1, nb, 3), _t(sty["ddelta"]).float())
    gain = (F.softplus(radiance).repeat(1, nb, 1) + _t(sty["dI"]).float()[None, :, None]).clamp(0)
    rgbs = (omega * (gain * (palette + offsets)).clamp(0, 1)).sum(dim=-2)
    if view_dep is not None:
        rgbs = rgbs + _t(view_dep).float()
    return rgbs


def render_sampler(params, rays_o, rays_d, bound=2.0, min_near=0.2, per_level_scale=2 ** (8 / 15), num_steps=512,
                   pred_clip=False, bg_color=1.0, chunk=4096, density_scale=1.0):
    """BASELINE config 1: uniform sampler + full field + weights compositing (nerf/renderer.py:127-255, staged in
    max_ray_batch=4096 chunks like :564-599), palette field + blend instead of NeRFNetwork.color."""
    rays_o, rays_d = _t(rays_o).float().reshape(-1, 3), _t(rays_d).float().reshape(-1, 3)
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    image_all, depth_all, ws_all = [], [], []
    for h0 in range(0, rays_o.shape[0], chunk):
        o, d = rays_o[h0:h0 + chunk], rays_d[h0:h0 + chunk]
        N = o.shape[0]
        nears, fars = O.near_far_from_aabb(o.numpy(), d.numpy(), aabb, min_near)
        nears, fars = torch.from_numpy(nears)[:, None], torch.from_numpy(fars)[:, None]
        z = nears + (fars - nears) * torch.linspace(0.0, 1.0, num_steps)[None]
        sample_dist = (fars - nears) / num_steps
        xyz = o[:, None, :] + d[:, None, :] * z[..., None]
        xyz = torch.min(torch.max(xyz, torch.from_numpy(aabb[:3])), torch.from_numpy(aabb[3:]))
        dirs = d[:, None, :].expand_as(xyz)
        sigma, _, omega, off_rad, view_dep, _ = palette_forward(params, xyz.reshape(-1, 3), dirs.reshape(-1, 3), bound,
                                                               per_level_scale, pred_clip)
        rgbs, _, _ = blend(params, omega, off_rad, view_dep)
        deltas = torch.cat([z[:, 1:] - z[:, :-1], sample_dist], dim=-1)
        alphas = 1 - torch.exp(-deltas * density_scale * sigma.view(N, num_steps))
        trans = torch.cumprod(torch.cat([torch.ones(N, 1), 1 - alphas + 1e-15], dim=-1), dim=-1)[:, :-1]
        w = alphas * trans
        ws = w.sum(-1)
        depth = (w * ((z - nears) / (fars - nears)).clamp(0, 1)).sum(-1)
        image = (w[..., None] * rgbs.view(N, num_steps, 3)).sum(-2) + (1 - ws)[:, None] * bg_color
        image_all.append(image); depth_all.append(depth); ws_all.append(ws)
    return torch.cat(image_all), torch.cat(depth_all), torch.cat(ws_all)


def render_cuda_ray(params, rays_o, rays_d, bitfield, bound=2.0, C=2, H=128, min_near=0.2, per_level_scale=2 ** (8 / 15),
                    dt_gamma=0.0, max_steps=1024, T_thresh=1e-4, pred_clip=False, bg_color=1.0, gui_mode=False,
                    density_scale=1.0, num_basis=4, clip_dim=16, edit=None, stylizer=None, offsets_weight=1.0,
                    view_dep_weight=1.0):
    """palette/renderer.py:430-552 (inference schedule incl. n_step growth and alive-list compaction), oracle kernels.
    edit: dict(delta_hsv, mean_xyz, mean_clip, std_xyz, std_clip) -> RegionEdit (:481-482); stylizer: dict(dI, dP, ddelta)
    -> Stylizer replaces the blend (:474-475)"""
    rays_o = np.ascontiguousarray(_t(rays_o).float().reshape(-1, 3).numpy())
    rays_d = np.ascontiguousarray(_t(rays_d).float().reshape(-1, 3).numpy())
    N = rays_o.shape[0]
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    bitfield = _t(bitfield).numpy()
    z = lambda *s: np.zeros(s, np.float32)  # noqa: E731
    acc = dict(weights_sum=z(N), depth=z(N), image=z(N, 3), clip_feat=z(N, clip_dim), direct_rgb=z(N, 3),
               view_dep_rgb=z(N, 3), basis_acc=z(N, num_basis), basis_rgb=z(N, 3 * num_basis),
               unscaled_basis_rgb=z(N, 3 * num_basis))
    alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    step = 0
    n_samples = 0
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        M = n_alive * n_step
        M += 128 - (M % 128)
        xyzs, dirs, deltas = O.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars,
                                          np.zeros(n_alive, np.float32), dt_gamma, max_steps, M=M)
        n_samples += int((deltas[:, 0] > 0).sum())
        sigma, clip, omega, off_rad, view_dep, diffuse = palette_forward(params, xyzs, dirs, bound, per_level_scale, pred_clip,
                                                                         clip_dim)
        if stylizer is not None:
            rgbs = stylize(stylizer, off_rad[..., -1:], omega, _t(params["basis_color"]).float()[None].clamp(0, 1),
                           off_rad[..., :-1].reshape(M, num_basis, 3), view_dep)
            basis_rgb = unscaled = None
        else:
            rgbs, basis_rgb, unscaled = blend(params, omega, off_rad, view_dep, num_basis, offsets_weight=offsets_weight,
                                              view_dep_weight=view_dep_weight, edit=edit, xyzs=_t(xyzs).float(), clip=clip)
        sig = (density_scale * sigma).numpy()
        if not gui_mode and basis_rgb is not None:
            for name, val in (("direct_rgb", diffuse + view_dep), ("view_dep_rgb", view_dep), ("basis_acc", omega),
                              ("basis_rgb", basis_rgb.reshape(M, -1)),
                              ("unscaled_basis_rgb", unscaled.expand(M, num_basis, 3).reshape(M, -1))):
                acc[name] = O.composite_rays_flex(n_alive, n_step, alive, sig, val.numpy(), deltas, acc["weights_sum"], acc[name],
                                                  T_thresh)
        acc["clip_feat"] = O.composite_rays_flex(n_alive, n_step, alive, sig, clip.numpy(), deltas, acc["weights_sum"],
                                                 acc["clip_feat"], T_thresh)
        alive, rays_t, acc["weights_sum"], acc["depth"], acc["image"] = O.composite_rays(
            n_alive, n_step, alive, rays_t, sig, rgbs.numpy(), deltas, acc["weights_sum"], acc["depth"], acc["image"], T_thresh)
        alive = alive[alive >= 0]
        step += n_step
    ws = acc["weights_sum"]
    out = dict(acc)
    out["depth_origin"] = acc["depth"].copy()
    out["image"] = acc["image"] + (1 - ws)[:, None] * bg_color
    out["direct_rgb"] = acc["direct_rgb"] + (1 - ws)[:, None] * bg_color
    out["depth"] = np.clip(acc["depth"] - nears, 0, None) / (fars - nears)
    out["n_samples"] = n_samples
    return out


def train_forward_cuda_ray(params, rays_o, rays_d, bitfield, bound=2.0, C=2, H=128, min_near=0.2,
                           per_level_scale=2 ** (8 / 15), dt_gamma=0.0, max_steps=1024, T_thresh=1e-4, pred_clip=False,
                           bg_color=1.0, density_scale=1.0, num_basis=4, clip_dim=16, noises=None, smooth=False,
                           jitter_fn=None, smooth_sigma_xyz=0.005, smooth_sigma_color=0.2, smooth_sigma_clip=0.0):
    """palette/renderer.py:322-429 forward, oracle kernels; returns the result maps + sample count.
    smooth=True: the smooth-loss branch (:360-381); jitter_fn(xyzs) -> U[0,1) noise [M,3] stands for torch.rand_like"""
    rays_o = np.ascontiguousarray(_t(rays_o).float().reshape(-1, 3).numpy())
    rays_d = np.ascontiguousarray(_t(rays_d).float().reshape(-1, 3).numpy())
    N = rays_o.shape[0]
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    noises = np.zeros(N, np.float32) if noises is None else noises
    # first call sizes M like raymarching.py:196-226 does (N*max_steps, then slice to m + pad)
    xyzs, dirs, deltas, rays, counter = O.march_rays_train(rays_o, rays_d, _t(bitfield).numpy(), bound, dt_gamma, max_steps, C,
                                                           H, N * max_steps if N * max_steps < (1 << 24) else (1 << 24),
                                                           nears, fars, noises)
    m = int(counter[0])
    m += 128 - m % 128
    xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    sigma, clip, omega, off_rad, view_dep, diffuse = palette_forward(params, xyzs, dirs, bound, per_level_scale, pred_clip,
                                                                     clip_dim)
    rgbs, _, _ = blend(params, omega, off_rad, view_dep, num_basis)
    sig = (density_scale * sigma).numpy()
    ws, depth, image = O.composite_rays_train_forward(sig, rgbs.numpy(), deltas, rays, T_thresh)
    offsets = off_rad[..., :-1].reshape(m, num_basis, 3)
    sparsity = omega.sum(-1, keepdim=True) / ((omega ** 2).sum(-1, keepdim=True) + 1e-6) - 1
    offsets_norm = (offsets ** 2).sum(-1).sum(-1, keepdim=True)
    view_dep_norm = (view_dep ** 2).sum(-1, keepdim=True)
    smooth_norm = torch.zeros_like(sparsity)
    if smooth:
        x = torch.from_numpy(np.ascontiguousarray(xyzs))
        xj = (x + jitter_fn(x) * bound * 0.03).clamp(-bound, bound)                                          # :362
        _, clip_j, omega_j, _, _, diffuse_j = palette_forward(params, xj, dirs, bound, per_level_scale, pred_clip, clip_dim)
        k_xyz = (x - xj).norm(dim=-1, keepdim=True) ** 2 / bound ** 2 / smooth_sigma_xyz                      # :368
        k_rgb = (diffuse - diffuse_j).norm(dim=-1, keepdim=True) ** 2 / smooth_sigma_color                   # :369
        k_clip = (clip - clip_j).norm(dim=-1, keepdim=True) / smooth_sigma_clip if (pred_clip and smooth_sigma_clip > 0) else 0
        gate = torch.exp(-k_xyz - k_rgb - k_clip)                                                             # :375
        smooth_norm = ((omega_j - omega) ** 2).sum(dim=-1, keepdim=True) * gate                               # :376
        if pred_clip:
            smooth_norm = smooth_norm + ((clip_j - clip) ** 2).sum(dim=-1, keepdim=True) * gate               # :378
    buf = torch.cat([sparsity, view_dep_norm, offsets_norm, smooth_norm, view_dep, diffuse + view_dep, diffuse,
                     clip, omega], dim=-1)
    maps = O.composite_rays_flex_train_forward(sig, buf.numpy(), deltas, rays, T_thresh)
    return dict(image=image + (1 - ws)[:, None] * bg_color, depth=np.clip(depth - nears, 0, None) / (fars - nears),
                weights_sum=ws, maps=maps, n_samples=int(counter[0]))


def nerf_train_step_cuda_ray(params, rays_o, rays_d, rays_gt, bitfield, bound=2.0, C=2, H=128, min_near=0.2,
                             per_level_scale=2 ** (8 / 15), dt_gamma=0.0, max_steps=1024, T_thresh=1e-4, bg_color=1.0,
                             density_scale=1.0, noises=None, loss_fn=None):
    """stage-1 training branch of NeRFRenderer.run_cuda (nerf/renderer.py:282-330) on the oracle kernels, and — with `loss_fn`
    (maps dict -> scalar torch loss) — the gradients of the step by the oracle's composite backward + torch autograd of the
    field. `params`: state dict; tensors that should receive gradients must have requires_grad. Returns (maps, grads | None);
    maps = image / depth / weights_sum / rgb_norm as the reference returns them."""
    rays_o = np.ascontiguousarray(_t(rays_o).float().reshape(-1, 3).numpy())
    rays_d = np.ascontiguousarray(_t(rays_d).float().reshape(-1, 3).numpy())
    gt = np.ascontiguousarray(_t(rays_gt).float().reshape(-1, 3).numpy())
    N = rays_o.shape[0]
    aabb = np.array([-bound] * 3 + [bound] * 3, np.float32)
    nears, fars = O.near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    noises = np.zeros(N, np.float32) if noises is None else noises
    xyzs, dirs, deltas, rays, counter = O.march_rays_train(rays_o, rays_d, _t(bitfield).numpy(), bound, dt_gamma, max_steps, C,
                                                           H, N * max_steps if N * max_steps < (1 << 24) else (1 << 24),
                                                           nears, fars, noises)
    m = int(counter[0])
    m += 128 - m % 128
    xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    sigma_t, rgb_t = nerf_forward(params, xyzs, dirs, bound, per_level_scale)
    sig_t = density_scale * sigma_t                                                                  # :299
    sig, rgbs = sig_t.detach().numpy(), rgb_t.detach().numpy()
    gt_s = O.spread_ray_to_sample(gt, rays, m)                                                       # :303
    err_t = ((torch.from_numpy(gt_s) - rgb_t) ** 2).sum(-1, keepdim=True).repeat(1, 3)              # :304
    err = err_t.detach().numpy()
    ws, depth, image = O.composite_rays_train_forward(sig, rgbs, deltas, rays, T_thresh)            # :325
    _, _, err_img = O.composite_rays_train_forward(sig, err, deltas, rays, T_thresh)                # :326
    maps = dict(image=image + (1 - ws)[:, None] * bg_color, depth=np.clip(depth - nears, 0, None) / (fars - nears),
                weights_sum=ws, rgb_norm=err_img.mean(-1), n_samples=int(counter[0]))
    if loss_fn is None:
        return maps, None
    # per-ray gradients of the loss by autograd on the maps, through the compositor by the oracle's backward kernels, through
    # the field by autograd again
    leaf = {k: torch.from_numpy(np.ascontiguousarray(maps[k])).requires_grad_() for k in ("image", "weights_sum", "rgb_norm")}
    loss = loss_fn(leaf)
    loss.backward()
    z = lambda t, shape: np.zeros(shape, np.float32) if t.grad is None else t.grad.numpy()     # noqa: E731
    g_img, g_ws_direct, g_err = z(leaf["image"], (N, 3)), z(leaf["weights_sum"], (N,)), z(leaf["rgb_norm"], (N,))
    # image = composite + (1 - ws) * bg: d/d ws picks up -bg . g_img
    g_ws = g_ws_direct - (g_img * bg_color).sum(-1)
    gs1, gr1 = O.composite_rays_train_backward(g_ws, g_img, sig, rgbs, deltas, rays, ws, image, T_thresh)
    g_err3 = np.repeat(g_err[:, None] / 3.0, 3, axis=1).astype(np.float32)                           # mean over 3 equal channels
    gs2, ge2 = O.composite_rays_train_backward(np.zeros(N, np.float32), g_err3, sig, err, deltas, rays,
                                               np.zeros(N, np.float32) + O.composite_rays_train_forward(sig, err, deltas, rays, T_thresh)[0],
                                               err_img, T_thresh)
    grads_in = [torch.from_numpy(gs1 + gs2), torch.from_numpy(gr1), torch.from_numpy(ge2)]
    leaves = [v for v in params.values() if torch.is_tensor(v) and v.requires_grad]
    g = torch.autograd.grad([sig_t, rgb_t, err_t], leaves, grads_in, allow_unused=True)
    names = [k for k, v in params.items() if torch.is_tensor(v) and v.requires_grad]
    return maps, {k: gi for k, gi in zip(names, g) if gi is not None}


# ---------------------------------------------------------------------------------------------------------------
# per-ray losses of the palette training step (TEST INFRASTRUCTURE; restates palette/utils.py:486-567 in torch)
# ---------------------------------------------------------------------------------------------------------------
def palette_train_loss(outputs, gt_rgb, lambda_sparsity=0.0, lambda_offsets=0.0, lambda_view_dep=0.0, lambda_smooth=0.0,
                       gt_clip_feat=None, gt_weights=None, lambda_weight=0.0, basis_color=None, basis_color_origin=None,
                       lambda_palette=0.0):
    """PaletteTrainer.train_step after model.render, expression by expression (criterion = MSELoss(reduction='none'),
    palette/utils.py:334): returns (loss, loss_dict, per_ray) where per_ray is `loss` before the final .mean() minus the
    broadcast scalar terms, i.e. the rgb error the error map stores. Pinned against the reference's own train_step
    (tests/golden/ref_palette.npz: loss and every loss_dict entry, tests/test_golden_palette.py)."""
    pred_rgb = outputs["image"]
    loss = ((pred_rgb - gt_rgb) ** 2).mean(-1)                                   # :486  [B, N]
    per_ray = loss.detach().clone()
    d = {}
    d["direct"] = ((outputs["direct_rgb"] - gt_rgb) ** 2).mean()                 # :487
    d["clip_feat"] = ((outputs["clip_feat"] - gt_clip_feat) ** 2).mean() if gt_clip_feat is not None else 0.0   # :490-492
    d["sparsity"] = lambda_sparsity * outputs["omega_sparsity"].mean()           # :521-522, 546
    d["offsets"] = lambda_offsets * outputs["offsets_norm"].mean()               # :524-525, 549
    d["view_dep"] = lambda_view_dep * outputs["view_dep_norm"].mean()            # :527-528, 552
    d["smooth"] = lambda_smooth * outputs["smooth_norm"].mean() if lambda_smooth != 0.0 else 0.0   # :538-542, 555
    d["weight"] = lambda_weight * ((gt_weights - outputs["basis_acc"]) ** 2).mean() if gt_weights is not None else 0.0  # :532-536
    d["palette"] = lambda_palette * ((basis_color - basis_color_origin) ** 2).sum(dim=-1).mean() \
        if basis_color is not None else 0.0                                      # :544, 561
    for k in ("sparsity", "offsets", "view_dep", "smooth", "palette", "weight", "direct", "clip_feat"):   # order of :546-571
        loss = loss + d[k]
    d["rgb"] = per_ray.mean()
    return loss.mean(), d, per_ray                                               # :598


# ---------------------------------------------------------------------------------------------------------------
# mark_untrained_grid (TEST INFRASTRUCTURE; restates NeRFRenderer.mark_untrained_grid, nerf/renderer.py:395-465)
# ---------------------------------------------------------------------------------------------------------------
def mark_untrained_grid(density_grid, poses, intrinsic, cascade, grid_size, bound, min_near, filter_close_point=False,
                        cam_chunk=16, margin=False):
    """-> (new density grid [C, H^3] in Morton order, fp64 decision margins or None). The reference's tensor program
    (:421-456) on the CPU in float64, cell by cell in chunks: world = (2 c/(H-1) - 1)(bound_k - bound_k/H) (:427, :431-434),
    cam = (world - t) @ R (:443-444), the three frustum tests (:447-450), the min_near tests (:452-453), count /
    too_close accumulation over all cameras (:455-458) and the final marking (:462-463).
    margin=True additionally returns, per cell, the smallest |lhs - rhs| over every comparison made for it: cells whose
    margin is below fp32 resolution may legitimately flip between an fp32 and an fp64 evaluation."""
    import oracle
    H = grid_size
    fx, fy, cx, cy = (float(v) for v in intrinsic)
    poses = torch.as_tensor(poses, dtype=torch.float64).reshape(-1, 4, 4)
    g = torch.arange(H, dtype=torch.int32)
    coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    indices = torch.from_numpy(oracle.morton3D(coords.numpy())).long()
    unit = 2 * coords.double() / (H - 1) - 1
    out = density_grid.clone()
    margins = torch.full((cascade, H ** 3), float("inf"), dtype=torch.float64) if margin else None
    for cas in range(cascade):
        b = min(2 ** cas, bound)
        half = b / H
        world = unit * (b - half)
        count = torch.zeros(H ** 3, dtype=torch.int64)
        close = torch.zeros(H ** 3, dtype=torch.int64)
        mg = torch.full((H ** 3,), float("inf"), dtype=torch.float64)
        for h in range(0, poses.shape[0], cam_chunk):
            R, t = poses[h:h + cam_chunk, :3, :3], poses[h:h + cam_chunk, :3, 3]
            cam = (world.unsqueeze(0) - t.unsqueeze(1)) @ R
            z = cam[..., 2]
            lim_x, lim_y = cx / fx * z + half * 2, cy / fy * z + half * 2
            vis = (z > 0) & (cam[..., 0].abs() < lim_x) & (cam[..., 1].abs() < lim_y)
            count += vis.sum(0)
            close += (vis & (z < min_near)).sum(0)
            if filter_close_point:
                close += (cam.norm(dim=-1) < min_near).sum(0)
            if margin:
                m = torch.minimum(z.abs(), torch.minimum((cam[..., 0].abs() - lim_x).abs(), (cam[..., 1].abs() - lim_y).abs()))
                m = torch.minimum(m, (z - min_near).abs())
                if filter_close_point:
                    m = torch.minimum(m, (cam.norm(dim=-1) - min_near).abs())
                mg = torch.minimum(mg, m.min(0).values)
        mark = (count * (close == 0)) == 0
        out[cas, indices[mark]] = -1
        if margin:
            margins[cas, indices] = mg
    return out, margins


# ---------------------------------------------------------------------------------------------------------------
# random-init parameters of the palette architecture, built without the product package (bench.py reference arm)
# ---------------------------------------------------------------------------------------------------------------
def random_palette_params(seed=0, pred_clip=False, bound=2.0, num_basis=4, clip_dim=16):
    """state_dict-shaped dict: nn.Linear default init U(-1/sqrt(in), 1/sqrt(in)) (kaiming_uniform with
    a = sqrt(5)), hash tables U(-1e-4, 1e-4) (gridencoder/grid.py:131-133), basis_color 0.5"""
    g = torch.Generator().manual_seed(seed)
    per_level_scale = float(np.exp2(np.log2(2048 * bound / 16) / 15))
    offsets = O.grid_offsets(3, 16, 16, 19, per_level_scale)
    n = int(offsets[-1])

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    def table():
        return (torch.rand(n, 2, generator=g) * 2 - 1) * 1e-4
    p = {}
    for name in ("encoder", "encoder_palette", "encoder_clip"):
        p[f"{name}.embeddings"] = table()
        p[f"{name}.offsets"] = torch.as_tensor(np.asarray(offsets, dtype=np.int32))
    p["sigma_net.0.weight"], p["sigma_net.1.weight"] = lin(64, 32), lin(16, 64)
    p["color_net.0.weight"], p["color_net.1.weight"], p["color_net.2.weight"] = lin(64, 31), lin(64, 64), lin(3, 64)
    p["diff_net.0.weight"], p["diff_net.1.weight"], p["diff_net.2.weight"] = lin(64, 15), lin(64, 64), lin(3, 64)
    p["basis_net.0.weight"], p["basis_net.1.weight"] = lin(64, 35), lin(15, 64)
    p["offsets_radiance_net.weight"] = lin(3 * num_basis + 1, 15)
    p["offsets_radiance_net.bias"] = (torch.rand(3 * num_basis + 1, generator=g) * 2 - 1) / math.sqrt(15)
    p["omega_net.0.weight"] = lin(num_basis, 15)
    if pred_clip:
        p["clip_net.0.weight"], p["clip_net.1.weight"] = lin(64, 32), lin(clip_dim, 64)
    p["basis_color"] = torch.full((num_basis, 3), 0.5)
    return p, per_level_scale
