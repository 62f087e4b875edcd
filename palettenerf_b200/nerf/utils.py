"""nerf.utils — the ray generator of the reference's trainer module (nerf/utils.py:52-151), on one CUDA kernel.

Only `get_rays` (immediately upstream of the marcher) lives here; the Trainer classes of the
reference module are control plane and out of scope.

`get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, random_size=0)` keeps the reference's
signature, sampling modes, RNG call order (the same torch.randint / torch.multinomial / torch.rand calls in the same
order, so a seeded run draws the same pixels) and result keys (`rays_o`, `rays_d`, `inds`, `inds_coarse`). What changes
is the tensor program after the indices are drawn: no [B, H*W] meshgrids, no gather / stack / norm / matmul chain —
one launch of pnerf_get_rays writes rays_o / rays_d [B, N, 3] (rays_o materialised: the marcher needs it contiguous).
Extra keyword `aabb=` (+ `min_near=`): also return `nears`, `fars` = near_far_from_aabb of these rays, computed in the
same kernel.
"""
import torch

from .. import _lib as L
from .._lib import call, ptr, require_cuda, stream

__all__ = ["get_rays", "custom_meshgrid"]


def custom_meshgrid(*args):
    """ref nerf/utils.py:34-40"""
    return torch.meshgrid(*args, indexing="ij")


def _draw_indices(B, H, W, N, error_map, patch_size, random_size, device):
    """pixel-index sampling of the reference, call for call (nerf/utils.py:76-124)"""
    extra = {}
    if patch_size > 1:
        num_patch = N // (patch_size ** 2)
        inds_x = torch.randint(0, H - patch_size, size=[num_patch], device=device)
        inds_y = torch.randint(0, W - patch_size, size=[num_patch], device=device)
        inds = torch.stack([inds_x, inds_y], dim=-1)
        pi, pj = custom_meshgrid(torch.arange(patch_size, device=device), torch.arange(patch_size, device=device))
        offsets = torch.stack([pi.reshape(-1), pj.reshape(-1)], dim=-1)
        inds = (inds.unsqueeze(1) + offsets.unsqueeze(0)).view(-1, 2)
        inds = inds[:, 0] * W + inds[:, 1]
        # note: N is not updated when patch_size**2 does not divide it; the reference's expand([B, N]) raises then too
        inds = inds.expand([B, N])
    elif random_size > 0:
        assert N % 2 == 0
        num_patch = N // 2
        inds_x = torch.randint(0, H, size=[num_patch], device=device)
        inds_y = torch.randint(0, W, size=[num_patch], device=device)
        inds = torch.stack([inds_x, inds_y], dim=-1)
        off_x = torch.randint(-random_size, random_size, size=[num_patch], device=device)
        off_y = torch.randint(-random_size, random_size, size=[num_patch], device=device)
        diff = torch.stack([(inds_x + off_x).clamp(0, H - 1), (inds_y + off_y).clamp(0, W - 1)], dim=-1)
        inds = torch.cat([inds, diff], dim=0)
        inds = (inds[:, 0] * W + inds[:, 1]).expand([B, N])
    elif error_map is None:
        inds = torch.randint(0, H * W, size=[N], device=device).expand([B, N])
    else:
        inds_coarse = torch.multinomial(error_map.to(device), N, replacement=False)   # [B, N] in [0, 128*128)
        inds_x, inds_y = inds_coarse // 128, inds_coarse % 128
        sx, sy = H / 128, W / 128
        inds_x = (inds_x * sx + torch.rand(B, N, device=device) * sx).long().clamp(max=H - 1)
        inds_y = (inds_y * sy + torch.rand(B, N, device=device) * sy).long().clamp(max=W - 1)
        inds = inds_x * W + inds_y
        extra["inds_coarse"] = inds_coarse
    return inds, extra


@torch.no_grad()
def rays_from_indices(poses, intrinsics, H, W, inds=None, aabb=None, min_near=0.2):
    """the kernel call: poses [B,4,4], inds int64 [B,N] (any batch stride, e.g. an expanded [N]) or None = all pixels.
    Returns rays_o, rays_d [B,N,3] (+ nears, fars [B,N] when aabb is given)."""
    require_cuda(poses, inds, aabb)
    poses = poses.detach().to(torch.float32).contiguous()
    B = poses.shape[0]
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    if inds is None:
        N, bstride = H * W, 0
    else:
        if inds.dtype != torch.int64:
            inds = inds.long()
        if inds.dim() == 1:
            inds = inds.unsqueeze(0).expand(B, -1)
        if inds.stride(-1) != 1:
            inds = inds.contiguous()
        N, bstride = inds.shape[-1], (inds.stride(0) if inds.shape[0] > 1 else 0)
    dev = poses.device
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    nears = fars = None
    if aabb is not None:
        aabb = aabb.detach().to(torch.float32).contiguous()
        nears = torch.empty(B, N, dtype=torch.float32, device=dev)
        fars = torch.empty(B, N, dtype=torch.float32, device=dev)
    call("pnerf_get_rays", ptr(poses), fx, fy, cx, cy, H, W, ptr(inds), bstride, N, B, ptr(rays_o), ptr(rays_d), ptr(aabb),
         float(min_near), ptr(nears), ptr(fars), stream())
    return rays_o, rays_d, nears, fars


def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, random_size=0, aabb=None, min_near=0.2):
    """ref nerf/utils.py:52-151 (same arguments, same result dict)"""
    device = poses.device
    B = poses.shape[0]
    results = {}
    if N > 0:
        N = min(N, H * W)
        inds, extra = _draw_indices(B, H, W, N, error_map, patch_size, random_size, device)
        results.update(extra)
        results["inds"] = inds
        rays_o, rays_d, nears, fars = rays_from_indices(poses, intrinsics, H, W, inds, aabb, min_near)
    else:
        results["inds"] = torch.arange(H * W, device=device).expand([B, H * W])
        rays_o, rays_d, nears, fars = rays_from_indices(poses, intrinsics, H, W, None, aabb, min_near)
    results["rays_o"] = rays_o
    results["rays_d"] = rays_d
    if nears is not None:
        results["nears"], results["fars"] = nears, fars
    return results
