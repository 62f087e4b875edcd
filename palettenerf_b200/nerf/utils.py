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

__all__ = ["get_rays", "collate", "custom_meshgrid"]


def custom_meshgrid(*args):
    """ref nerf/utils.py:34-40"""
    return torch.meshgrid(*args, indexing="ij")


# Pixel-index sampling. The four modes of the reference (nerf/utils.py:76-124) as separate samplers; each makes the
# reference's torch RNG calls in the reference's order (a seeded run draws the same pixels), everything else is this
# module's own arrangement. Every sampler returns flat pixel ids `row * W + col` of shape [B, N].
def _rand(hi, n, device, lo=0):
    return torch.randint(lo, hi, size=[n], device=device)


def _sample_patches(B, H, W, N, p, device):
    """N // p^2 random p x p patches (top-left corners drawn first along rows, then along columns); ref :79-95"""
    k = N // (p * p)
    rows, cols = _rand(H - p, k, device), _rand(W - p, k, device)
    dr, dc = custom_meshgrid(torch.arange(p, device=device), torch.arange(p, device=device))
    pix = (rows[:, None] + dr.reshape(1, -1)) * W + (cols[:, None] + dc.reshape(1, -1))      # [k, p^2]
    return pix.reshape(-1).expand([B, N])      # as in the reference, N must be a multiple of p^2


def _sample_pairs(B, H, W, N, r, device):
    """N/2 random pixels followed by one neighbour each within +-r (clamped to the image); ref :96-110"""
    assert N % 2 == 0
    k = N // 2
    rows, cols = _rand(H, k, device), _rand(W, k, device)
    d_rows, d_cols = _rand(r, k, device, lo=-r), _rand(r, k, device, lo=-r)
    first = rows * W + cols
    second = (rows + d_rows).clamp(0, H - 1) * W + (cols + d_cols).clamp(0, W - 1)
    return torch.cat([first, second]).expand([B, N])


def _sample_uniform(B, H, W, N, device):
    """N pixels with replacement; ref :111-113"""
    return _rand(H * W, N, device).expand([B, N])


def _sample_error_map(B, H, W, N, error_map, device):
    """N cells of the 128 x 128 error map without replacement, then a uniform pixel inside each cell; ref :114-126"""
    coarse = torch.multinomial(error_map.to(device), N, replacement=False)       # [B, N]
    cell_h, cell_w = H / 128, W / 128
    rows = ((coarse // 128) * cell_h + torch.rand(B, N, device=device) * cell_h).long().clamp(max=H - 1)
    cols = ((coarse % 128) * cell_w + torch.rand(B, N, device=device) * cell_w).long().clamp(max=W - 1)
    return rows * W + cols, coarse


def _draw_indices(B, H, W, N, error_map, patch_size, random_size, device):
    """mode dispatch in the reference's precedence: patches, then neighbour pairs, then uniform / error-map sampling"""
    extra = {}
    if patch_size > 1:
        inds = _sample_patches(B, H, W, N, patch_size, device)
    elif random_size > 0:
        inds = _sample_pairs(B, H, W, N, random_size, device)
    elif error_map is None:
        inds = _sample_uniform(B, H, W, N, device)
    else:
        inds, extra["inds_coarse"] = _sample_error_map(B, H, W, N, error_map, device)
    return inds, extra


@torch.no_grad()
def rays_from_indices(poses, intrinsics, H, W, inds=None, aabb=None, min_near=0.2, images=None, feat_images=None):
    """the kernel call: poses [B,4,4], inds int64 [B,N] (any batch stride, e.g. an expanded [N]) or None = all pixels.
    Returns rays_o, rays_d [B,N,3], nears, fars [B,N] (None without aabb) and — when `images` [B,H,W,C] / `feat_images`
    [B,H,W,Cf] are given — their rows at the same pixels ([B,N,C] / [B,N,Cf]), gathered by the same launch."""
    require_cuda(poses, inds, aabb, images, feat_images)
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
    if images is None and feat_images is None:
        call("pnerf_get_rays", ptr(poses), fx, fy, cx, cy, H, W, ptr(inds), bstride, N, B, ptr(rays_o), ptr(rays_d), ptr(aabb),
             float(min_near), ptr(nears), ptr(fars), stream())
        return rays_o, rays_d, nears, fars

    def prep(t):
        if t is None:
            return None, 0, None
        if t.shape[0] != B or t.shape[1] * t.shape[2] != H * W:
            raise RuntimeError("collate: images must be [B, H, W, C] of the view the rays are generated for")
        t = t.detach().to(torch.float32).contiguous()
        return t, t.shape[-1], torch.empty(B, N, t.shape[-1], dtype=torch.float32, device=dev)
    img, c_img, out_img = prep(images)
    feat, c_feat, out_feat = prep(feat_images)
    call("pnerf_get_rays_collate", ptr(poses), fx, fy, cx, cy, H, W, ptr(inds), bstride, N, B, ptr(rays_o), ptr(rays_d), ptr(aabb),
         float(min_near), ptr(nears), ptr(fars), ptr(img), c_img, ptr(out_img), ptr(feat), c_feat, ptr(out_feat), stream())
    return rays_o, rays_d, nears, fars, out_img, out_feat


def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1, random_size=0, aabb=None, min_near=0.2,
             images=None, feat_images=None):
    """ref nerf/utils.py:52-151 (same arguments, same result dict).
    images / feat_images ([B,H,W,C] ground-truth colours / semantic features, not in the reference's signature): the result
    also carries `images` / `feat_images` [B,N,C] gathered at the sampled pixels by the same launch — the two torch.gather
    calls of the data loader's collate (palette/provider.py:387-399)."""
    device = poses.device
    B = poses.shape[0]
    results = {}
    if N > 0:
        N = min(N, H * W)
        inds, extra = _draw_indices(B, H, W, N, error_map, patch_size, random_size, device)
        results.update(extra)
        results["inds"] = inds
    else:
        inds = None
        results["inds"] = torch.arange(H * W, device=device).expand([B, H * W])
    out = rays_from_indices(poses, intrinsics, H, W, inds, aabb, min_near, images, feat_images)
    results["rays_o"], results["rays_d"] = out[0], out[1]
    if out[2] is not None:
        results["nears"], results["fars"] = out[2], out[3]
    if images is not None:
        results["images"] = out[4]
    if feat_images is not None:
        results["feat_images"] = out[5]
    return results


def collate(poses, intrinsics, H, W, num_rays, images=None, feat_images=None, error_map=None, patch_size=1, random_size=0,
            training=True):
    """the ray / pixel part of NeRFDataset.collate (ref: palette/provider.py:377-403) for one batch of poses already on the
    device: result keys H, W, rays_o, rays_d, inds [, images, feat_images, inds_coarse]. In training the ground-truth pixels
    are gathered with the rays (one launch); in evaluation (num_rays = -1) the full images are passed through."""
    rays = get_rays(poses, intrinsics, H, W, num_rays, error_map, patch_size, random_size,
                    images=images if training else None, feat_images=feat_images if training else None)
    res = {"H": H, "W": W, "rays_o": rays["rays_o"], "rays_d": rays["rays_d"], "inds": rays["inds"]}
    for k, full in (("images", images), ("feat_images", feat_images)):
        if full is not None:
            res[k] = rays[k] if training else full
    if "inds_coarse" in rays:
        res["inds_coarse"] = rays["inds_coarse"]
    return res
