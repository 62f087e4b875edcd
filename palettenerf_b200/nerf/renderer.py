"""NeRFRenderer — occupancy-grid volume renderer (stage-1 model), drop-in for nerf/renderer.py:61-603 of the
reference: same constructor, buffers (`aabb_train/infer`, `density_grid`, `density_bitfield`, `step_counter`),
`render / run_cuda / update_extra_state / mark_untrained_grid / reset_extra_state` and result-dict keys.

Only the cuda_ray path exists (the product has no CPU fallback); `run()` — the reference's pure-torch sampler — is
restated in oracle/ as the CPU baseline instead. The density-grid maintenance is reorganised: one Morton-ordered
coordinate table is built once and cached, the full sweep evaluates whole cascades in a few large batches, and
`mark_untrained_grid` projects all cells of a cascade against a batch of cameras per launch.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib as _L
from .. import raymarching


def _ceil_log2(x):
    return math.ceil(math.log2(x))


class OccupancyState:
    """mixin holding the cuda_ray state shared by the NeRF and palette renderers"""

    def _init_occupancy(self, bound, cuda_ray, min_near, density_thresh, density_scale, bg_radius):
        self.bound = bound
        self.cascade = 1 + _ceil_log2(bound)
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius
        box = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32)
        self.register_buffer("aabb_train", box)
        self.register_buffer("aabb_infer", box.clone())
        self.cuda_ray = cuda_ray
        if cuda_ray:
            H3 = self.grid_size ** 3
            self.register_buffer("density_grid", torch.zeros(self.cascade, H3))
            self.register_buffer("density_bitfield", torch.zeros(self.cascade * H3 // 8, dtype=torch.uint8))
            self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
            self.__dict__["_mean_density"], self.__dict__["_mean_density_dev"] = 0, None
            self.iter_density = 0
            self.mean_count = 0
            self.local_step = 0
        self._cell_table = None

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.step_counter.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.mean_count = 0
        self.local_step = 0

    # -- helpers ------------------------------------------------------------------------------------
    def _cells(self, device):
        """(coords [H^3,3] int32 in x-major order, morton indices [H^3] int64), cached per device"""
        if self._cell_table is None or self._cell_table[0].device != device:
            g = torch.arange(self.grid_size, dtype=torch.int32, device=device)
            coords = torch.stack(torch.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).contiguous()
            self._cell_table = (coords, raymarching.morton3D(coords).long())
        return self._cell_table

    def _cascade_extent(self, cas):
        b = min(2 ** cas, self.bound)
        return b, b / self.grid_size

    def _next_counter(self):
        counter = self.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        return counter

    # -- density grid maintenance (ref: nerf/renderer.py:395-561) -----------------------------------------
    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """cells seen by no training camera (or closer than min_near to one) get density -1 and never set a bit.
        One kernel over (cascade, cell) with the cameras staged in shared memory (csrc/raymarch.cu::k_mark_untrained)
        instead of the reference's 5-deep Python loop (nerf/renderer.py:395-465); `S` (its chunk size) is ignored."""
        if not self.cuda_ray:
            return
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        dev = self.density_bitfield.device
        _L.require_cuda(self.density_grid)
        poses = poses.to(dev).to(torch.float32).reshape(-1, 4, 4).contiguous()
        fx, fy, cx, cy = (float(v) for v in intrinsic)
        if not self.density_grid.is_contiguous():
            raise RuntimeError("density_grid must be contiguous")
        n_marked = torch.zeros(1, dtype=torch.int32, device=dev)
        _L.call("pnerf_mark_untrained_grid", _L.ptr(poses), poses.shape[0], fx, fy, cx, cy, self.cascade, self.grid_size,
                float(self.bound), float(self.min_near), int(bool(getattr(self, "filter_close_point", False))),
                _L.ptr(self.density_grid), _L.ptr(n_marked), _L.stream())
        torch.autograd.graph.increment_version(self.density_grid)     # written through a raw pointer
        print(f"[mark untrained grid] {int(n_marked.item())} from {self.grid_size ** 3 * self.cascade}")

    @property
    def mean_density(self):
        """mean of clamp(density_grid, 0) after the last refresh. The kernel path keeps it on the device (the threshold of
        packbits never visits the host); reading this attribute is what synchronises."""
        dev_val = self.__dict__.get("_mean_density_dev")
        if dev_val is not None:
            return float(dev_val[0].item())
        return self.__dict__.get("_mean_density", 0)

    @mean_density.setter
    def mean_density(self, v):
        self.__dict__["_mean_density"] = v
        self.__dict__["_mean_density_dev"] = None

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, fused=None):
        """EMA-max refresh of the density grid + packbits + mean sample count (ref: nerf/renderer.py:467-561).
        fused (default: whenever the model is the architecture the tensor-core field covers): three kernels — cell selection
        + jitter + density + scatter, EMA-max + partial sums, mean / threshold / packbits — see fused_nerf.update_density_grid;
        the only host read left is the mean sample count (it sizes the next steps' sample buffers)."""
        if not self.cuda_ray:
            return
        from .. import fused_nerf
        use_fused = (fused_nerf.supported(self) and self.density_grid.is_cuda) if fused is None else bool(fused)
        if use_fused:
            stats = fused_nerf.update_density_grid(self, decay)
            self.__dict__["_mean_density_dev"] = stats
            self.iter_density += 1
            self._last_update_schedule = "fused"
            steps = min(16, self.local_step)
            if steps > 0:
                self.mean_count = int(self.step_counter[:steps, 0].sum().item() / steps)
            self.local_step = 0
            return
        self._last_update_schedule = "torch"
        dev = self.density_bitfield.device
        H = self.grid_size
        fresh = -torch.ones_like(self.density_grid)
        if self.iter_density < 16:  # full sweep
            coords, indices = self._cells(dev)
            unit = 2 * coords.float() / (H - 1) - 1
            chunk = H ** 3 // 4
            for cas in range(self.cascade):
                b, half = self._cascade_extent(cas)
                for c0 in range(0, unit.shape[0], chunk):
                    pts = unit[c0:c0 + chunk] * (b - half)
                    pts = pts + (torch.rand_like(pts) * 2 - 1) * half
                    sig = self.density(pts)["sigma"].reshape(-1).detach().float() * self.density_scale
                    fresh[cas, indices[c0:c0 + chunk]] = sig
        else:  # quarter of the cells at random + as many random occupied cells
            n = H ** 3 // 4
            for cas in range(self.cascade):
                b, half = self._cascade_extent(cas)
                rnd = torch.randint(0, H, (n, 3), device=dev)
                occupied = torch.nonzero(self.density_grid[cas] > 0).squeeze(-1)
                occupied = occupied[torch.randint(0, occupied.shape[0], [n], dtype=torch.long, device=dev)]
                ind = torch.cat([raymarching.morton3D(rnd).long(), occupied])
                cells = torch.cat([rnd, raymarching.morton3D_invert(occupied)])
                pts = (2 * cells.float() / (H - 1) - 1) * (b - half)
                pts = pts + (torch.rand_like(pts) * 2 - 1) * half
                fresh[cas, ind] = self.density(pts)["sigma"].reshape(-1).detach().float() * self.density_scale
        ok = (self.density_grid >= 0) & (fresh >= 0)
        self.density_grid[ok] = torch.maximum(self.density_grid[ok] * decay, fresh[ok])
        self.mean_density = torch.mean(self.density_grid.clamp(min=0)).item()
        self.iter_density += 1
        self.density_bitfield = raymarching.packbits(self.density_grid, min(self.mean_density, self.density_thresh),
                                                    self.density_bitfield)
        steps = min(16, self.local_step)
        if steps > 0:
            self.mean_count = int(self.step_counter[:steps, 0].sum().item() / steps)
        self.local_step = 0


def mix_background(image, weights_sum, bg_color):
    return image + (1 - weights_sum).unsqueeze(-1) * bg_color


def normalise_depth(depth, nears, fars):
    return torch.clamp(depth - nears, min=0) / (fars - nears)


class _RenderTail(Function):
    """depth normalisation + background mixing of run_cuda's epilogue in one launch (csrc/tail.cu); backward in one."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, depth, nears, fars, image, weights_sum, bg, maps, col_direct):
        N = weights_sum.shape[0]
        dev = image.device
        image, weights_sum = image.contiguous(), weights_sum.contiguous()
        depth_n = torch.empty(N, dtype=torch.float32, device=dev)
        image_out = torch.empty(N, 3, dtype=torch.float32, device=dev)
        direct_out = torch.empty(N, 3, dtype=torch.float32, device=dev) if maps is not None else None
        bg_stride = 0 if bg.numel() == 3 else 3
        stride = maps.shape[1] if maps is not None else 0
        _L.call("pnerf_render_tail_forward", N, _L.ptr(depth.contiguous()), _L.ptr(nears.contiguous()), _L.ptr(fars.contiguous()),
                _L.ptr(image), _L.ptr(weights_sum), None if maps is None else maps.data_ptr() + 4 * col_direct, stride,
                _L.ptr(bg), bg_stride, _L.ptr(depth_n), _L.ptr(image_out), _L.ptr(direct_out), _L.stream())
        ctx.save_for_backward(bg)
        ctx.dims = (N, bg_stride, stride, col_direct, maps is not None)
        ctx.mark_non_differentiable(depth_n)      # the compositors ignore d depth (raymarching.py:275), so does this
        if maps is None:
            return depth_n, image_out
        return depth_n, image_out, direct_out

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, _g_depth, g_image, g_direct=None):
        (bg,) = ctx.saved_tensors
        N, bg_stride, stride, col, has_maps = ctx.dims
        dev = bg.device
        g_image = None if g_image is None else g_image.contiguous().float()
        g_direct = None if g_direct is None else g_direct.contiguous().float()
        g_ws = torch.empty(N, dtype=torch.float32, device=dev)
        g_maps = torch.empty(N, stride, dtype=torch.float32, device=dev) if has_maps else None
        _L.call("pnerf_render_tail_backward", N, _L.ptr(g_image), _L.ptr(g_direct), _L.ptr(bg), bg_stride, stride, col,
                _L.ptr(g_ws), _L.ptr(g_maps), _L.stream())
        return None, None, None, g_image, g_ws, None, g_maps, None


_BG_CACHE = {}


def render_tail(depth, nears, fars, image, weights_sum, bg_color, maps=None, col_direct=0):
    """-> (depth_n [N], image [N,3], direct [N,3] | None): normalise_depth + mix_background (twice with `maps`, whose
    columns [col_direct, col_direct+3) hold the un-mixed direct colour). One kernel when everything is a plain fp32 CUDA
    tensor and the background is a constant / per-ray colour without gradient; the tensor expressions otherwise."""
    N = weights_sum.shape[0]
    bg = bg_color
    if not torch.is_tensor(bg):
        key = (float(bg), str(image.device))
        bg = _BG_CACHE.get(key)
        if bg is None:
            bg = _BG_CACHE[key] = torch.full((3,), float(bg_color), dtype=torch.float32, device=image.device)
    ok = (image.is_cuda and image.dtype == torch.float32 and weights_sum.dtype == torch.float32 and N > 0
          and not bg.requires_grad and bg.is_cuda and bg.numel() in (3, 3 * N)
          and (maps is None or (maps.dim() == 2 and maps.is_contiguous() and maps.dtype == torch.float32)))
    if not ok:
        direct = None if maps is None else mix_background(maps[..., col_direct:col_direct + 3], weights_sum, bg_color)
        return normalise_depth(depth, nears, fars), mix_background(image, weights_sum, bg_color), direct
    bg = bg.detach().to(torch.float32).contiguous()
    out = _RenderTail.apply(depth, nears, fars, image, weights_sum, bg, maps, col_direct)
    return (out[0], out[1], out[2]) if maps is not None else (out[0], out[1], None)


class NeRFRenderer(nn.Module, OccupancyState):
    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01, bg_radius=-1,
                 filter_close_point=False):
        super().__init__()
        self.filter_close_point = filter_close_point
        self._init_occupancy(bound, cuda_ray, min_near, density_thresh, density_scale, bg_radius)

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def run(self, *args, **kwargs):
        raise RuntimeError("palettenerf_b200 implements the cuda_ray path only (no CPU / pure-torch sampler); "
                           "construct the model with cuda_ray=True")

    def _background(self, rays_o, rays_d, bg_color):
        if self.bg_radius > 0:
            sph = raymarching.sph_from_ray(rays_o, rays_d, self.bg_radius)
            return self.background(sph, rays_d)
        return 1 if bg_color is None else bg_color

    def _fused_infer_available(self, fused=None):
        """default policy like the palette model's: the fused renderer under fp16 autocast for the architecture it covers"""
        from .. import fused_nerf
        if fused is not None:
            return bool(fused)
        return torch.is_autocast_enabled() and fused_nerf.supported_color(self)

    def _fused_train_available(self, fused=None):
        """default policy like the palette stage's: the fused step under fp16 autocast for the architecture it covers"""
        from .. import fused_nerf_train
        if fused is not None:
            return bool(fused)
        return torch.is_autocast_enabled() and fused_nerf_train.supported(self) and self.bg_radius <= 0

    def run_cuda(self, rays_o, rays_d, rays_gt=None, dt_gamma=0, bg_color=None, perturb=False, force_all_rays=False,
                 max_steps=1024, T_thresh=1e-4, **kwargs):
        """rays_o, rays_d: [B, N, 3] (B == 1) -> dict(image [B,N,3], depth [B,N], rgb_norm, weights_sum)"""
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        bg_color = self._background(rays_o, rays_d, bg_color)
        C, H = self.cascade, self.grid_size

        if self.training and self._fused_train_available(kwargs.get("fused")):
            # Fused stage-1 step: static-capacity march (sample count stays on the device: no D2H sync), ONE forward kernel
            # for hash grid + sigma net + SH + colour net with hand-written backward kernels, ONE compositing pass for
            # rgb / depth / weights and the per-ray error channel (fused_nerf_train).
            from .. import fused_nerf_train
            self._last_train_schedule = "fused"
            xyzs, dirs, deltas, rays, valid = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, C, H, nears, fars, self._next_counter(), self.mean_count,
                perturb, 128, force_all_rays, dt_gamma, max_steps, True)
            sigmas, rgbs = fused_nerf_train.field(self, xyzs, dirs, count=valid)      # density_scale applied in the kernel
            gt = None if rays_gt is None else rays_gt.contiguous().view(-1, 3)
            weights_sum, depth, image, err_map = fused_nerf_train.composite(sigmas, rgbs, deltas, rays, gt, T_thresh, static=True)
            depth, image, _ = render_tail(depth, nears, fars, image, weights_sum, bg_color)
            image, depth = image.view(*prefix, 3), depth.view(*prefix)
            rgb_norm = err_map.view(*prefix)
        elif self.training:
            self._last_train_schedule = "torch"
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(
                rays_o, rays_d, self.bound, self.density_bitfield, C, H, nears, fars, self._next_counter(), self.mean_count,
                perturb, 128, force_all_rays, dt_gamma, max_steps)
            sigmas, rgbs = self(xyzs, dirs)
            sigmas = self.density_scale * sigmas
            if rays_gt is not None:
                gt = torch.zeros_like(xyzs)
                raymarching.spread_ray_to_sample(rays_gt.contiguous().view(-1, 3), rays, gt)
                err = ((gt - rgbs) ** 2).sum(-1, keepdim=True).repeat(1, 3)
            else:
                err = torch.zeros_like(rgbs)
            if sigmas.dim() == 2:  # CCNeRF-style stack of K residual fields
                images, depths = [], []
                for k in range(sigmas.shape[0]):
                    weights_sum, depth, image = raymarching.composite_rays_train(sigmas[k], rgbs[k], deltas, rays, T_thresh)
                    images.append(mix_background(image, weights_sum, bg_color).view(*prefix, 3))
                    depths.append(normalise_depth(depth, nears, fars).view(*prefix))
                image, depth = torch.stack(images, 0), torch.stack(depths, 0)
                rgb_norm = torch.zeros_like(depth[0])
            else:
                weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
                _, _, err_map = raymarching.composite_rays_train(sigmas, err, deltas, rays, T_thresh)
                depth, image, _ = render_tail(depth, nears, fars, image, weights_sum, bg_color)
                image, depth = image.view(*prefix, 3), depth.view(*prefix)
                rgb_norm = err_map.mean(dim=-1).view(*prefix)
        elif self._fused_infer_available(kwargs.get("fused")):
            # ONE persistent kernel (warp-per-ray, field on tcgen05) instead of the host loop below
            from .. import fused_nerf
            acc = fused_nerf.render(self, rays_o.float(), rays_d.float(), nears, fars, perturb, dt_gamma, max_steps, T_thresh)
            self._last_schedule, self._last_queue = "fused", acc["_queue"]
            weights_sum = acc["weights_sum"]
            depth, image, _ = render_tail(acc["depth"], nears, fars, acc["image"], weights_sum, bg_color)
            image, depth = image.view(*prefix, 3), depth.view(*prefix)
            rgb_norm = torch.zeros_like(image[..., 0])
        else:
            self._last_schedule = "loop"
            weights_sum = torch.zeros(N, dtype=torch.float32, device=dev)
            depth = torch.zeros(N, dtype=torch.float32, device=dev)
            image = torch.zeros(N, 3, dtype=torch.float32, device=dev)
            rays_alive = torch.arange(N, dtype=torch.int32, device=dev)
            rays_t = nears.clone()
            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound,
                                                            self.density_bitfield, C, H, nears, fars, 128,
                                                            perturb if step == 0 else False, dt_gamma, max_steps)
                sigmas, rgbs = self(xyzs, dirs)
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, self.density_scale * sigmas, rgbs, deltas,
                                           weights_sum, depth, image, T_thresh)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step
            depth, image, _ = render_tail(depth, nears, fars, image, weights_sum, bg_color)
            image, depth = image.view(*prefix, 3), depth.view(*prefix)
            rgb_norm = torch.zeros_like(image[..., 0])

        return {"depth": depth, "image": image, "rgb_norm": rgb_norm, "weights_sum": weights_sum}

    def render(self, rays_o, rays_d, rays_gt=None, staged=False, max_ray_batch=4096, **kwargs):
        if not self.cuda_ray:
            return self.run(rays_o, rays_d, **kwargs)
        return self.run_cuda(rays_o, rays_d, rays_gt, **kwargs)  # never staged with cuda_ray (ref: :576)
