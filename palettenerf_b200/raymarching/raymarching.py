"""raymarching — drop-in for the reference package of the same name (raymarching/raymarching.py:19-473).

Same twelve callables with the same positional signatures, dtypes and in-place conventions; the kernels behind
`_backend` are the sm_100a ones in csrc/raymarch.cu and csrc/composite.cu. Differences a caller can observe:
  * march_rays_train assigns sample offsets by a deterministic scan in ray order (the reference's order is an
    atomicAdd race) — `rays[i] == (i, offset_i, count_i)`;
  * no torch.cuda.empty_cache() after marching (raymarching.py:231) — it only defeats the caching allocator.
"""
import weakref

import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .backend import _backend, OCC_FLOATS

__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train", "composite_rays_flex_train", "march_rays", "composite_rays", "composite_rays_flex",
           "spread_ray_to_sample"]

_fwd32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def _cuda(t):
    return t if t.is_cuda else t.cuda()


class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _backend.near_far_from_aabb(rays_o, rays_d, _cuda(aabb).contiguous(), N, min_near, nears, fars)
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _backend.sph_from_ray(rays_o, rays_d, radius, N, coords)
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords)
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        _backend.morton3D(coords.int().contiguous(), N, indices)
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices)
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        _backend.morton3D_invert(indices.int().contiguous(), N, coords)
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _cuda(grid).contiguous()
        C, H3 = grid.shape[0], grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
            _backend.packbits(grid, N, thresh, bitfield)
            return bitfield
        # in-place convention of the reference (packbits(grid, thresh, bitfield), keep using `bitfield`): the kernel writes
        # through a raw pointer, so tell autograd / the version counter — occupied_bounds() caches per bitfield version
        _backend.packbits(grid, N, thresh, bitfield)
        ctx.mark_dirty(bitfield)
        return bitfield


packbits = _packbits.apply


_OCC_CACHE = {}          # data_ptr -> (weakref to the bitfield tensor, (version, C, H, bound), occ_aabb)
T_LIST_MAX_BYTES = 256 << 20


def occupied_bounds(density_bitfield, C, H, bound):
    """[6] world-space bounds of the occupied cells (pnerf_occupied_bounds), cached per bitfield TENSOR OBJECT until it is
    written again: torch bumps `_version` on every in-place write, and packbits() into an existing buffer marks it dirty
    itself (its kernel writes through a raw pointer). The entry holds a weak reference: a different tensor that happens to
    reuse the address of a freed one never hits."""
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the kernel is recorded: every replay recomputes the bounds from the live bitfield
        # (a cached tensor would go stale when the density grid is refreshed between replays)
        occ = torch.empty(OCC_FLOATS, dtype=torch.float32, device=density_bitfield.device)
        _backend.occupied_bounds(density_bitfield, C, H, bound, occ)
        return occ
    key = density_bitfield.data_ptr()
    tag = (density_bitfield._version, C, H, float(bound))
    hit = _OCC_CACHE.get(key)
    if hit is not None and hit[0]() is density_bitfield and hit[1] == tag:
        return hit[2]
    occ = torch.empty(OCC_FLOATS, dtype=torch.float32, device=density_bitfield.device)
    _backend.occupied_bounds(density_bitfield, C, H, bound, occ)
    if len(_OCC_CACHE) > 64:
        _OCC_CACHE.clear()
    _OCC_CACHE[key] = (weakref.ref(density_bitfield), tag, occ)
    return occ


class _march_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024, static=False):
        """`static=True` (not in the reference): return the full-capacity buffers without reading the sample count back
        (no D2H sync, shapes independent of the data -> CUDA-graph capturable) plus a 5th output `valid` (int32 [1], device):
        the number of leading rows that hold samples. Slot order is the deterministic ray order, so rows [0, valid) are all
        written; when the total count (step_counter[0]) exceeds the capacity M, `valid` is the offset of the first ray that
        did not fit (that ray and every later one write nothing, raymarching.cu:418-419) and rows >= valid are
        UNINITIALISED — consumers must not read them."""
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        density_bitfield = _cuda(density_bitfield).contiguous()
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        dev, dt = rays_o.device, rays_o.dtype
        if static:   # data-independent capacity: reusable scratch (see arena.py), rows beyond the count stay uninitialised
            from ..arena import ARENA
            xyzs = ARENA.get("march_xyzs", (M, 3), dt, dev).detach()
            dirs = ARENA.get("march_dirs", (M, 3), dt, dev).detach()
            deltas = ARENA.get("march_deltas", (M, 2), dt, dev).detach()
        else:
            xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
            dirs = torch.zeros(M, 3, dtype=dt, device=dev)
            deltas = torch.zeros(M, 2, dtype=dt, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        valid = torch.empty(1, dtype=torch.int32, device=dev) if static else None
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=dt, device=dev) if perturb else torch.zeros(N, dtype=dt, device=dev)
        # (a backend with only the reference's surface, e.g. the reference's own extension swapped in by the perf test,
        # takes the two-walk entry point)
        if hasattr(_backend, "march_rays_train_ws") and N * max_steps * 4 <= T_LIST_MAX_BYTES and (H * H * H) % 32 == 0:
            # one walk of the grid: the counting pass records the sample parameters (scratch) and stops at the occupied bounds
            from ..arena import ARENA
            t_list = ARENA.get("march_t_list", (N * max_steps,), torch.float32, dev)
            occ = occupied_bounds(density_bitfield, C, H, bound)
            _backend.march_rays_train_ws(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                         xyzs, dirs, deltas, rays, step_counter, noises, t_list, occ, valid)
        else:
            if static:
                raise RuntimeError("march_rays_train(static=True) needs the one-walk entry point (march_rays_train_ws)")
            _backend.march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, N, C, H, M, nears, fars,
                                      xyzs, dirs, deltas, rays, step_counter, noises)
        if static:
            ctx.mark_non_differentiable(valid)
            return xyzs, dirs, deltas, rays, valid
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()  # D2H sync, as in the reference (raymarching.py:224)
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
        return xyzs, dirs, deltas, rays


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, deltas = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        _backend.composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh]
        return weights_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is ignored, as in the reference (raymarching.py:275)
        grad_weights_sum, grad_image = grad_weights_sum.contiguous(), grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        grad_sigmas, grad_rgbs = torch.zeros_like(sigmas), torch.zeros_like(rgbs)
        _backend.composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum,
                                               image, M, N, T_thresh, grad_sigmas, grad_rgbs)
        return grad_sigmas, grad_rgbs, None, None, None


composite_rays_train = _composite_rays_train.apply


class _composite_rays_flex_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, input, deltas, rays, T_thresh=1e-4):
        sigmas, input, deltas = sigmas.contiguous(), input.contiguous(), deltas.contiguous()
        M, N, n_channel = sigmas.shape[0], rays.shape[0], input.shape[-1]
        output = torch.empty(N, n_channel, dtype=sigmas.dtype, device=sigmas.device)
        _backend.composite_rays_flex_train_forward(sigmas, input, deltas, rays, M, N, n_channel, T_thresh, output)
        ctx.save_for_backward(sigmas, input, deltas, rays, output)
        ctx.dims = [M, N, n_channel, T_thresh]
        return output

    @staticmethod
    @_bwd
    def backward(ctx, grad_output):
        grad_output = grad_output.contiguous()
        sigmas, input, deltas, rays, output = ctx.saved_tensors
        M, N, n_channel, T_thresh = ctx.dims
        grad_input = torch.zeros_like(input)
        _backend.composite_rays_flex_train_backward(grad_output, sigmas, input, deltas, rays, output, M, N, n_channel,
                                                    T_thresh, grad_input)
        return None, grad_input, None, None, None


composite_rays_flex_train = _composite_rays_flex_train.apply


class _march_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        rays_o = _cuda(rays_o).contiguous().view(-1, 3)
        rays_d = _cuda(rays_d).contiguous().view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        dev, dt = rays_o.device, rays_o.dtype
        xyzs = torch.zeros(M, 3, dtype=dt, device=dev)
        dirs = torch.zeros(M, 3, dtype=dt, device=dev)
        deltas = torch.zeros(M, 2, dtype=dt, device=dev)
        noises = torch.rand(n_alive, dtype=dt, device=dev) if perturb else torch.zeros(n_alive, dtype=dt, device=dev)
        _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                            density_bitfield, near, far, xyzs, dirs, deltas, noises)
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        _backend.composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas.contiguous(), rgbs.contiguous(),
                                deltas, weights_sum, depth, image)
        return tuple()


composite_rays = _composite_rays.apply


class _composite_rays_flex(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, n_channel, rays_alive, rays_t, sigmas, input, deltas, weights_sum, output,
                T_thresh=1e-2):
        _backend.composite_rays_flex(n_alive, n_step, n_channel, T_thresh, rays_alive, rays_t, sigmas.contiguous(),
                                     input.contiguous(), deltas, weights_sum, output)
        return tuple()


composite_rays_flex = _composite_rays_flex.apply


class _spread_ray_to_sample(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, input, rays, output):
        input = input.contiguous()
        if not output.is_contiguous():
            raise RuntimeError("spread_ray_to_sample writes in place: output must be contiguous")
        N, M, n_channel = input.shape[0], output.shape[0], input.shape[-1]
        _backend.spread_ray_to_sample(input, rays, M, N, n_channel, output)
        return tuple()


spread_ray_to_sample = _spread_ray_to_sample.apply
