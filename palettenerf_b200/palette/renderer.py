"""PaletteRenderer — palette-mode volume renderer, drop-in for palette/renderer.py:185-572 of the reference
(`render`, `run_cuda`, `initialize_palette`, `reset_extra_state`, the result-dict keys of :415-429 / :531-550),
plus the GUI-time edit modules RegionEdit and Stylizer (:83-183).

run_cuda has two inference schedules:
  * `fused=False` (compatibility): the reference's host loop — march n_step samples for the alive rays, evaluate
    the field, blend, composite the six auxiliary maps and the image, compact the alive list — on the new kernels;
  * `fused=True` (default when the fused extension is loaded): see palettenerf_b200/fused.py.
The training branch follows palette/renderer.py:322-429: one march, one field evaluation, the palette blend, one
rgb composite and ONE n-channel composite carrying all per-sample regularisers and auxiliary maps.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import raymarching
from ..nerf.renderer import OccupancyState, mix_background, normalise_depth, render_tail  # noqa: F401
from .backend import rgb_to_hsv, hsv_to_rgb


def srgb_to_linear(x):
    return torch.where(x < 0.04045, x / 12.92, ((x + 0.055) / 1.055) ** 2.4)


class RegionEdit(nn.Module):
    """regional recolouring controller (ref: palette/renderer.py:83-147): per-basis HSV shift/scale, optionally
    gated by distance to a picked 3-D point and/or semantic feature"""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.mean_xyz = None
        self.mean_clip = None
        self.std_xyz = 1
        self.std_clip = 1
        self.weight_mode = False
        self.delta_hsv = torch.zeros(opt.num_basis, 3)
        self.delta_hsv[..., 1:3] = 1

    def update_cent(self, mean_xyz=None, mean_clip=None):
        self.mean_xyz = None if mean_xyz is None else mean_xyz[None, ...]
        self.mean_clip = None if mean_clip is None else mean_clip[None, ...]

    def update_std(self, std_xyz=None, std_clip=None):
        if std_xyz is not None:
            self.std_xyz = std_xyz
        if std_clip is not None:
            self.std_clip = std_clip

    def update_delta_hsv(self, rgb_orig, rgb_new):
        if rgb_orig.device != self.delta_hsv.device:
            self.delta_hsv = self.delta_hsv.type_as(rgb_orig)
        nb = self.opt.num_basis
        hsv = rgb_to_hsv(torch.cat([rgb_orig, rgb_new], dim=0))
        old, new = hsv[:nb], hsv[nb:]
        self.delta_hsv[:, 0] = torch.fmod(new[:, 0] - old[:, 0] + 360, 360)
        self.delta_hsv[:, 1] = new[:, 1] / old[:, 1] + 1e-9
        self.delta_hsv[:, 2] = new[:, 2] / old[:, 2] + 1e-9

    def forward(self, rgbs, xyz=None, clip_feat=None):
        hsv = rgb_to_hsv(rgbs)
        if rgbs.device != self.delta_hsv.device:
            self.delta_hsv = self.delta_hsv.type_as(rgbs)
        weight = torch.ones_like(rgbs[..., 0:1, 0])
        if xyz is not None and self.mean_xyz is not None:
            weight = weight * torch.exp(-((xyz - self.mean_xyz) ** 2.).sum(dim=-1, keepdim=True) / self.std_xyz)
        if clip_feat is not None and self.mean_clip is not None:
            weight = weight * torch.exp(-((clip_feat - self.mean_clip) ** 2.).sum(dim=-1, keepdim=True) / self.std_clip)
        edited = torch.stack([torch.fmod(hsv[..., 0] + self.delta_hsv[..., 0] + 360, 360),
                              torch.clip(hsv[..., 1] * self.delta_hsv[..., 1], 0),
                              torch.clip(hsv[..., 2] * self.delta_hsv[..., 2], 0)], dim=-1)
        if self.weight_mode:
            return weight[..., None].repeat(1, self.opt.num_basis, 3)
        return torch.lerp(rgbs, hsv_to_rgb(edited), weight[..., None])


class Stylizer(nn.Module):
    """photorealistic style-transfer solver (ref: palette/renderer.py:151-183): learnable per-basis intensity,
    palette shift and 3x3 offset transform"""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.dI = nn.Parameter(torch.zeros(opt.num_basis))
        self.dP = nn.Parameter(torch.zeros(1, opt.num_basis, 3))
        self.ddelta = nn.Parameter(torch.eye(3)[None].repeat(opt.num_basis, 1, 1))

    def ARAP_loss(self):
        eye = torch.eye(3, dtype=torch.float32, device=self.ddelta.device)[None]
        return ((torch.bmm(self.ddelta, self.ddelta.transpose(1, 2)) - eye) ** 2).sum()

    def forward(self, radiance, omega, palette, offsets, view_dep=None):
        nb = self.opt.num_basis
        prefix = offsets.shape[:-2]
        radiance, omega = radiance.reshape(-1, 1, 1), omega.reshape(-1, nb, 1)
        palette = palette.reshape(-1, nb, 3) + self.dP
        offsets = torch.einsum("npi,pij->npj", offsets.reshape(-1, nb, 3), self.ddelta)
        gain = (F.softplus(radiance).repeat(1, nb, 1) + self.dI[None, :, None]).clamp(0)
        rgbs = (omega * (gain * (palette + offsets)).clamp(0, 1)).sum(dim=-2)
        if view_dep is not None:
            rgbs = rgbs + view_dep.detach()
        return rgbs.reshape(*prefix, 3)


class PaletteRenderer(nn.Module, OccupancyState):
    def __init__(self, opt, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1):
        super().__init__()
        self.opt = opt
        self._init_occupancy(bound, cuda_ray, min_near, density_thresh, density_scale, bg_radius)
        self.num_basis = opt.num_basis
        self.freeze_basis_color = opt.use_initialization_from_rgbxy
        self.require_smooth_loss = False
        self.color_weight = 0
        self.edit = None
        self.stylizer = None
        self.view_dep_weight = 1
        self.offsets_weight = 1
        if opt.test or not opt.use_initialization_from_rgbxy:
            self.basis_color = nn.Parameter(torch.zeros([self.num_basis, 3]) + 0.5, requires_grad=True)
        else:
            self.basis_color = None  # set by initialize_palette() from the extracted palette

    def initialize_palette(self, color_list=None, hist_weights=None):
        if color_list is None:
            if self.basis_color is None:
                self.basis_color = nn.Parameter(torch.zeros([self.num_basis, 3]) + 0.5, requires_grad=True)
        else:
            colors = torch.zeros([self.num_basis, 3])
            for i, c in enumerate(color_list):
                c = torch.as_tensor(c, dtype=torch.float32)
                colors[i] = srgb_to_linear(c) if self.opt.color_space == "linear" else c
            self.basis_color = nn.Parameter(colors, requires_grad=True)
        self.basis_color_origin = nn.Parameter(self.basis_color.data, requires_grad=False)
        if hist_weights is not None:
            hw = torch.from_numpy(hist_weights).float().permute(3, 0, 1, 2).unsqueeze(0)
            self.hist_weights = nn.Parameter(hw, requires_grad=False)

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def run(self, *args, **kwargs):
        raise ValueError("Pure pytorch version is not available for now.")  # same as the reference (:292-294)

    def _background(self, rays_o, rays_d, bg_color):
        if self.bg_radius > 0:
            return self.background(raymarching.sph_from_ray(rays_o, rays_d, self.bg_radius), rays_d)
        return 1 if bg_color is None else bg_color

    def _train_field_torch(self, xyzs, dirs, deltas, rays, palette, T_thresh):
        """the reference's per-op schedule of the training field (palette/renderer.py:333-385) on the new kernels"""
        nb, cd = self.num_basis, self.opt.clip_dim
        M = xyzs.shape[0]
        sigmas, clip_feat, omega, offsets_radiance, view_dep, diffuse = self(xyzs, dirs)
        sigmas = (self.density_scale * sigmas).detach()      # geometry is frozen in the palette stage (ref :334-335)
        offsets = offsets_radiance[..., :-1].reshape(M, nb, 3)
        radiance = offsets_radiance[..., -1:].reshape(M, 1, 1)
        omega = omega.reshape(M, nb, 1)
        view_dep, diffuse, clip_feat = view_dep.reshape(M, 3), diffuse.reshape(M, 3), clip_feat.reshape(M, cd)

        rgbs = (omega * (F.softplus(radiance) * (palette + offsets))).sum(dim=-2) + view_dep.detach()
        direct_rgb = diffuse + view_dep
        weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)

        w = omega[..., 0]
        omega_sparsity = w.sum(dim=-1, keepdim=True) / ((w ** 2).sum(dim=-1, keepdim=True) + 1e-6) - 1
        offsets_norm = (offsets ** 2).sum(dim=-1).sum(dim=-1, keepdim=True)
        view_dep_norm = (view_dep ** 2).sum(dim=-1, keepdim=True)
        if self.require_smooth_loss:
            jitter = (xyzs + torch.rand_like(xyzs) * self.bound * 0.03).clamp(-self.bound, self.bound)
            _, clip_j, omega_j, _, _, diffuse_j = self(jitter, dirs)
            omega_j, diffuse_j = omega_j.reshape(M, nb, 1), diffuse_j.reshape(M, 3)
            k_xyz = (xyzs - jitter).norm(dim=-1, keepdim=True) ** 2 / self.bound ** 2 / self.opt.smooth_sigma_xyz
            k_rgb = (diffuse - diffuse_j).norm(dim=-1, keepdim=True) ** 2 / self.opt.smooth_sigma_color
            k_clip = 0
            if self.opt.pred_clip and self.opt.smooth_sigma_clip > 0:
                k_clip = (clip_feat - clip_j).norm(dim=-1, keepdim=True) / self.opt.smooth_sigma_clip
            gate = torch.exp(-k_xyz - k_rgb - k_clip).detach()
            smooth_norm = ((omega_j - omega)[..., 0] ** 2).sum(dim=-1, keepdim=True) * gate
            if self.opt.pred_clip:
                smooth_norm = smooth_norm + ((clip_j - clip_feat) ** 2).sum(dim=-1, keepdim=True) * gate
        else:
            smooth_norm = torch.zeros_like(omega_sparsity)

        channels = torch.cat([omega_sparsity, view_dep_norm, offsets_norm, smooth_norm, view_dep, direct_rgb, diffuse,
                              clip_feat, w], dim=-1)
        return sigmas, rgbs, channels, weights_sum, depth, image

    def _smooth_channels(self, fused_train, xyzs, dirs, palette, valid, channels):
        """smooth-loss branch of the training field (ref: palette/renderer.py:360-381) on the fused kernels: a SECOND fused
        forward on the jittered points (its backward runs through the same hand-written kernels: the reference lets the
        gradient of smooth_norm flow into both evaluations), the gate / norm arithmetic as one kernel forward and one
        backward over the valid rows (fused_train.smooth_gate), written into column 3 of the channel buffer the one-pass
        compositor consumes.
        Rows beyond `valid` (static capacity) hold garbage on both sides and are never composited."""
        nb, cd = self.num_basis, self.opt.clip_dim
        b = self.bound
        jitter = (xyzs + torch.rand_like(xyzs) * b * 0.03).clamp(-b, b)
        _, _, ch_j = fused_train.field(self, jitter, dirs, palette, count=valid)
        return fused_train.smooth_gate(self, channels, ch_j, xyzs, jitter, valid)

    # ------------------------------------------------------------------------------------------------
    def _train_branch(self, rays_o, rays_d, nears, fars, bg_color, prefix, dt_gamma, perturb, force_all_rays, max_steps,
                      T_thresh, fused=None):
        nb, cd = self.num_basis, self.opt.clip_dim
        use_fused = self._fused_train_available() if fused is None else bool(fused)
        self._last_train_schedule = "fused" if use_fused else "torch"
        palette = self.basis_color[None].clamp(0, 1)
        if self.freeze_basis_color:
            palette = palette.detach()
        counter = self._next_counter()
        if use_fused:
            # Static-capacity schedule: the march leaves the sample count on the device (no D2H sync, shapes independent
            # of the data -> the whole step can be captured in a CUDA graph); ONE forward kernel (hash grids + MLPs +
            # blend + regulariser channels) with a hand-written backward; ONE compositing pass for rgb + all channels.
            from .. import fused_train
            xyzs, dirs, deltas, rays, valid = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
                self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps, True)
            # `valid` (not counter[0]): when the capacity overflows, rows behind the first dropped ray are uninitialised
            sigmas, rgbs, channels = fused_train.field(self, xyzs, dirs, palette[0], count=valid)
            if self.require_smooth_loss:
                channels = self._smooth_channels(fused_train, xyzs, dirs, palette[0], valid, channels)
            if channels.shape[1] == 33:
                weights_sum, depth, image, maps = fused_train.composite(sigmas, rgbs, channels, deltas, rays, T_thresh)
            else:
                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
                maps = raymarching.composite_rays_flex_train(sigmas, channels, deltas, rays, T_thresh)
        else:
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size, nears, fars, counter,
                self.mean_count, perturb, 128, force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs, channels, weights_sum, depth, image = self._train_field_torch(xyzs, dirs, deltas, rays, palette,
                                                                                        T_thresh)
            # all auxiliary channels ride through ONE n-channel composite: [M, 13 + clip_dim + Nb]
            maps = raymarching.composite_rays_flex_train(sigmas, channels, deltas, rays, T_thresh)

        # depth normalisation + both background mixes: one launch forward, one backward (csrc/tail.cu)
        depth_n, image_mixed, direct_mixed = render_tail(depth, nears, fars, image, weights_sum, bg_color, maps, 7)
        out = {
            "depth": depth_n.view(*prefix),
            "image": image_mixed.view(*prefix, 3),
            "weights_sum": weights_sum,
            "omega_sparsity": maps[..., 0:1].view(*prefix),
            "view_dep_norm": maps[..., 1:2].view(*prefix),
            "offsets_norm": maps[..., 2:3].view(*prefix),
            "smooth_norm": maps[..., 3:4].view(*prefix),
            "view_dep_rgb": maps[..., 4:7].view(*prefix, 3),
            "direct_rgb": direct_mixed.view(*prefix, 3),
            "diffuse_rgb": maps[..., 10:13].view(*prefix, 3),
            "clip_feat": maps[..., 13:13 + cd].view(*prefix, cd),
            "basis_acc": maps[..., 13 + cd:13 + cd + nb].view(*prefix, nb),
        }
        return out

    def _shade(self, xyzs, dirs):
        """field evaluation + palette blend for inference samples -> dict of per-sample quantities"""
        nb, cd = self.num_basis, self.opt.clip_dim
        M = xyzs.shape[0]
        sigmas, clip_feat, omega, offsets_radiance, view_dep, diffuse = self(xyzs, dirs)
        offsets = offsets_radiance[..., :-1].reshape(M, nb, 3)
        radiance = offsets_radiance[..., -1:].reshape(M, 1, 1)
        omega = omega.reshape(M, nb, 1)
        view_dep, diffuse, clip_feat = view_dep.reshape(M, 3), diffuse.reshape(M, 3), clip_feat.reshape(M, cd)
        palette = self.basis_color[None].clamp(0, 1)
        s = {"sigmas": self.density_scale * sigmas, "clip_feat": clip_feat, "omega": omega, "view_dep": view_dep,
             "diffuse": diffuse}
        if self.stylizer is not None:
            s["rgbs"] = self.stylizer(radiance, omega, palette, offsets, view_dep)
            return s
        final = F.softplus(radiance) * (palette + self.offsets_weight * offsets)
        if self.edit is not None:
            final = self.edit(final, xyzs, clip_feat)
        s["basis_rgb"] = omega * final
        s["unscaled_basis_rgb"] = (palette + offsets).expand(M, nb, 3)
        s["rgbs"] = s["basis_rgb"].sum(dim=-2) + self.view_dep_weight * view_dep
        return s

    def _infer_loop(self, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode):
        """the reference's host-driven schedule (palette/renderer.py:430-523) on the new kernels"""
        nb, cd = self.num_basis, self.opt.clip_dim
        N, dev = rays_o.shape[0], rays_o.device
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)  # noqa: E731
        acc = {"weights_sum": z(N), "depth": z(N), "image": z(N, 3), "clip_feat": z(N, cd)}
        if not gui_mode:
            acc.update(direct_rgb=z(N, 3), view_dep_rgb=z(N, 3), basis_acc=z(N, nb), basis_rgb=z(N, 3 * nb),
                       unscaled_basis_rgb=z(N, 3 * nb))
        rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
        rays_t = nears.clone()
        step = 0
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                        self.density_bitfield, self.cascade, self.grid_size, nears, fars,
                                                        128, perturb if step == 0 else False, dt_gamma, max_steps)
            M = xyzs.shape[0]
            s = self._shade(xyzs, dirs)

            def flex(name, value, ch):
                raymarching.composite_rays_flex(n_alive, n_step, ch, rays_alive, rays_t, s["sigmas"], value, deltas,
                                                acc["weights_sum"], acc[name], T_thresh)
            if not gui_mode and "basis_rgb" in s:
                flex("direct_rgb", s["diffuse"] + s["view_dep"], 3)
                flex("view_dep_rgb", s["view_dep"], 3)
                flex("basis_acc", s["omega"].reshape(M, nb), nb)
                flex("basis_rgb", s["basis_rgb"].reshape(M, 3 * nb), 3 * nb)
                flex("unscaled_basis_rgb", s["unscaled_basis_rgb"].reshape(M, 3 * nb), 3 * nb)
            flex("clip_feat", s["clip_feat"], cd)
            # must run last: it advances weights_sum / rays_t and kills rays (ref :517-519)
            raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, s["sigmas"], s["rgbs"], deltas,
                                       acc["weights_sum"], acc["depth"], acc["image"], T_thresh)
            rays_alive = rays_alive[rays_alive >= 0]
            step += n_step
        return acc

    def run_cuda(self, rays_o, rays_d, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False, max_steps=1024,
                 T_thresh=1e-4, gui_mode=False, fused=None, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        bg_color = self._background(rays_o, rays_d, bg_color)
        if self.training:
            return self._train_branch(rays_o, rays_d, nears, fars, bg_color, prefix, dt_gamma, perturb, force_all_rays,
                                      max_steps, T_thresh, fused=fused)

        nb, cd = self.num_basis, self.opt.clip_dim
        use_fused = self._fused_available(gui_mode) if fused is None else fused
        if use_fused:
            acc = self._infer_fused(rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode)
        else:
            acc = self._infer_loop(rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode)
        self._last_schedule = "fused" if use_fused else "loop"
        self._last_queue = acc.get("_queue")
        ws = acc["weights_sum"]
        depth_n, image_mixed, direct_mixed = render_tail(acc["depth"], nears, fars, acc["image"], ws, bg_color,
                                                         None if gui_mode else acc["direct_rgb"], 0)
        out = {
            "depth": depth_n.view(*prefix),
            "depth_origin": acc["depth"].view(*prefix),     # the accumulator itself (nothing writes it afterwards)
            "image": image_mixed.view(*prefix, 3),
            "weights_sum": ws,
            "clip_feat": acc["clip_feat"].view(*prefix, cd),
        }
        if not gui_mode:
            out["direct_rgb"] = direct_mixed.view(*prefix, 3)
            out["view_dep_rgb"] = acc["view_dep_rgb"].view(*prefix, 3)
            out["basis_rgb"] = acc["basis_rgb"].view(*prefix, nb * 3)
            out["unscaled_basis_rgb"] = acc["unscaled_basis_rgb"].view(*prefix, nb * 3)
            out["basis_acc"] = acc["basis_acc"].view(*prefix, nb)
        return out

    # -- fused schedule (csrc/fused.cu): one persistent kernel instead of the host loop -------------------------
    def _fused_available(self, gui_mode):
        """default policy: use the fused renderer under fp16 autocast (its MLPs run fp16 tensor-core math) whenever
        the architecture is the one it implements. RegionEdit / Stylizer are evaluated inside the renderer's blend (csrc/
        field_tc.cu); a Stylizer renders without the debug maps, so it takes the fused path in gui_mode only — exactly the
        combination the reference supports (palette/renderer.py:474-475 leaves basis_rgb undefined otherwise)."""
        from .. import fused
        if self.stylizer is not None and not gui_mode:
            return False
        return torch.is_autocast_enabled() and fused.supported(self)

    def _fused_train_available(self):
        """fused training field: same policy as inference (the smooth-loss branch included, see _smooth_channels)"""
        from .. import fused
        return (torch.is_autocast_enabled() and self.edit is None and self.stylizer is None and fused.supported(self))

    def _infer_fused(self, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode):
        from .. import fused
        if not fused.supported(self):
            raise RuntimeError("fused render path does not cover this model architecture")
        return fused.render(self, rays_o.float(), rays_d.float(), nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode)

    def render(self, rays_o, rays_d, staged=False, max_ray_batch=4096, test_mode=False, gui_mode=False, **kwargs):
        run = self.run_cuda if self.cuda_ray else self.run
        return run(rays_o, rays_d, gui_mode=gui_mode, **kwargs)
