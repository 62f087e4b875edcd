"""Deterministic synthetic "Blender-lego-shaped" scene: occupancy grid, cameras and rays.

Used by bench.py, the tests and the oracle's CPU baseline (no dataset access on the GPU box). Pure torch/numpy on
the CPU; callers move the tensors to the device.
  * occupancy: cells whose cascade-k centre lies in the box |x|<0.55,|y|<0.30,|z|<0.40 or the sphere r=0.25 at
    (0.3,0.2,0) carry density 20, others 0; stored in Morton order like nerf/renderer.py:490-507 does;
  * camera: radius 4.031*0.8, camera_angle_x = 0.6911 (Blender lego), look-at origin, elevation 30 deg, y up;
    rays follow get_rays (nerf/utils.py:52-151): pixel centre +0.5, normalised directions, d @ R^T.
"""
import math

import numpy as np
import torch

LEGO = dict(bound=2.0, cascade=2, grid_size=128, min_near=0.2, dt_gamma=0.0, max_steps=1024, T_thresh=1e-4,
            density_thresh=10.0, radius=4.031 * 0.8, camera_angle_x=0.6911, num_basis=4, clip_dim=16)


def _spread3(v):
    v = v.astype(np.uint64)
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3d_np(x, y, z):
    return (_spread3(x) | (_spread3(y) << 1) | (_spread3(z) << 2)).astype(np.int64)


def inside_solid(xyz, scale=1.0, ground=False):
    x, y, z = xyz[..., 0] / scale, xyz[..., 1] / scale, xyz[..., 2] / scale
    box = (np.abs(x) < 0.55) & (np.abs(y) < 0.30) & (np.abs(z) < 0.40)
    sph = ((x - 0.3) ** 2 + (y - 0.2) ** 2 + z ** 2) < 0.25 ** 2
    m = box | sph
    if ground:  # mip-360-shaped variant (config 5): slab that reaches into cascade 1
        m = m | ((xyz[..., 1] > -0.6) & (xyz[..., 1] < -0.5) & (np.abs(xyz[..., 0]) < 1.8) & (np.abs(xyz[..., 2]) < 1.8))
    return m


def density_grid(bound=2.0, cascade=2, H=128, scale=1.0, ground=False, value=20.0):
    """[cascade, H^3] fp32 density grid in Morton order"""
    g = np.arange(H, dtype=np.int64)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    coords = np.stack([X.ravel(), Y.ravel(), Z.ravel()], -1)
    idx = morton3d_np(coords[:, 0], coords[:, 1], coords[:, 2])
    unit = 2.0 * coords.astype(np.float32) / (H - 1) - 1.0
    grid = np.zeros((cascade, H ** 3), np.float32)
    for k in range(cascade):
        bk = min(2.0 ** k, bound)
        xyz = unit * (bk - bk / H)
        grid[k, idx] = np.where(inside_solid(xyz, scale, ground), value, 0.0).astype(np.float32)
    return torch.from_numpy(grid)


def packbits_np(grid, thresh):
    """bit i of byte n = grid_flat[8n+i] > thresh (raymarching.cu:282-291) — numpy helper for CPU-side setup"""
    bits = (grid.reshape(-1, 8).numpy() > thresh).astype(np.uint8)
    return torch.from_numpy((bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8))


def lookat_pose(radius, azimuth_deg, elevation_deg=30.0):
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    pos = np.array([radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az)])
    fwd = -pos / np.linalg.norm(pos)
    up = np.array([0.0, 1.0, 0.0])
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, pos
    return torch.from_numpy(pose)


def camera_rays(H, W, azimuth_deg=35.0, radius=LEGO["radius"], camera_angle_x=LEGO["camera_angle_x"], inds=None):
    """rays_o, rays_d [H*W, 3] fp32 for a full view (or the given flat pixel indices)"""
    pose = lookat_pose(radius, azimuth_deg)
    fx = fy = 0.5 * W / math.tan(0.5 * camera_angle_x)
    cx, cy = W / 2, H / 2
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    i, j = i.reshape(-1) + 0.5, j.reshape(-1) + 0.5
    if inds is not None:
        i, j = i[inds], j[inds]
    d = torch.stack(((i - cx) / fx, (j - cy) / fy, torch.ones_like(i)), -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)
    rays_d = d @ pose[:3, :3].T
    rays_o = pose[:3, 3].expand_as(rays_d).contiguous()
    return rays_o.contiguous(), rays_d.contiguous()


def training_rays(n_rays, H=800, W=800, seed=0, n_views=8):
    """n_rays random pixels spread over n_views cameras (ray-batch of a training step)"""
    g = torch.Generator().manual_seed(seed)
    per = n_rays // n_views
    os_, ds_ = [], []
    for v in range(n_views):
        n = per if v < n_views - 1 else n_rays - per * (n_views - 1)
        inds = torch.randint(0, H * W, (n,), generator=g)
        o, d = camera_rays(H, W, azimuth_deg=360.0 * v / n_views + 10.0 * seed, inds=inds)
        os_.append(o)
        ds_.append(d)
    return torch.cat(os_).contiguous(), torch.cat(ds_).contiguous()


# ---------------------------------------------------------------------------------------------------------------
# model construction for the benchmark / tests
# ---------------------------------------------------------------------------------------------------------------
def make_opt(num_basis=4, clip_dim=16, pred_clip=False, **kw):
    """the subset of main_palette.py's argparse namespace that the model reads"""
    import types
    opt = types.SimpleNamespace(num_basis=num_basis, clip_dim=clip_dim, pred_clip=pred_clip, test=True,
                                use_initialization_from_rgbxy=False, color_space="srgb", smooth_sigma_xyz=0.005,
                                smooth_sigma_color=0.2, smooth_sigma_clip=0.0)
    for k, v in kw.items():
        setattr(opt, k, v)
    return opt


PALETTE_RGB = [[0.85, 0.70, 0.15], [0.20, 0.20, 0.25], [0.75, 0.75, 0.78], [0.60, 0.10, 0.10]]


def build_palette_model(device="cuda", seed=0, pred_clip=False, table_scale=None, sigma_bias=None, ground=False,
                        scene_scale=1.0):
    """random-init PaletteNetwork (torch.manual_seed(seed)) on the synthetic lego-shaped occupancy grid.
    table_scale: None keeps the reference init U(-1e-4,1e-4); a float re-draws the tables from U(-s, s) (parity runs)."""
    from .palette.network import PaletteNetwork
    torch.manual_seed(seed)
    model = PaletteNetwork(make_opt(pred_clip=pred_clip), bound=LEGO["bound"], cuda_ray=True, min_near=LEGO["min_near"],
                           density_thresh=LEGO["density_thresh"])
    with torch.no_grad():
        model.basis_color.copy_(torch.tensor(PALETTE_RGB))
        if table_scale is not None:
            for enc in (model.encoder, model.encoder_palette, model.encoder_clip):
                enc.embeddings.uniform_(-table_scale, table_scale)
        grid = density_grid(LEGO["bound"], LEGO["cascade"], LEGO["grid_size"], scale=scene_scale, ground=ground)
        model.density_grid.copy_(grid)
        model.mean_density = grid.clamp(min=0).mean().item()
        model.density_bitfield.copy_(packbits_np(grid, min(model.mean_density, LEGO["density_thresh"])))
    return model.to(device)


def build_nerf_model(device="cuda", seed=0, table_scale=None, ground=False, scene_scale=1.0):
    """random-init stage-1 NeRFNetwork (torch.manual_seed(seed)) on the same synthetic occupancy grid"""
    from .nerf.network import NeRFNetwork
    torch.manual_seed(seed)
    model = NeRFNetwork(bound=LEGO["bound"], cuda_ray=True, min_near=LEGO["min_near"], density_thresh=LEGO["density_thresh"])
    with torch.no_grad():
        if table_scale is not None:
            model.encoder.embeddings.uniform_(-table_scale, table_scale)
        grid = density_grid(LEGO["bound"], LEGO["cascade"], LEGO["grid_size"], scale=scene_scale, ground=ground)
        model.density_grid.copy_(grid)
        model.mean_density = grid.clamp(min=0).mean().item()
        model.density_bitfield.copy_(packbits_np(grid, min(model.mean_density, LEGO["density_thresh"])))
    return model.to(device)
