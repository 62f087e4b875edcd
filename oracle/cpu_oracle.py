"""numpy / ctypes front-end of the CPU oracle (see oracle/__init__.py for scope and parity status)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
sys.path.insert(0, os.path.join(_HERE, "..", "tools"))
from gen_sh import sh_eval  # noqa: E402  (shared SH derivation; independent of the CUDA code path)


def build_c_oracle():
    src = os.path.join(_HERE, "raymarch_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def clib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c_oracle())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


U, F = ctypes.c_uint32, ctypes.c_float


# ---------------------------------------------------------------- raymarching (C) ----------------------------
def near_far_from_aabb(rays_o, rays_d, aabb, min_near):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    clib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), U(N), F(min_near), _p(nears), _p(fars))
    return nears, fars


def morton3D(coords):
    coords = _i32(coords)
    out = np.empty(coords.shape[0], np.int32)
    clib().oracle_morton3D(_p(coords), U(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    out = np.empty((indices.shape[0], 3), np.int32)
    clib().oracle_morton3D_invert(_p(indices), U(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    clib().oracle_packbits(_p(grid), U(N), F(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bitfield, bound, dt_gamma, max_steps, C, H, M, nears, fars, noises, counter=None):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    xyzs, dirs = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    deltas, rays = np.zeros((M, 2), np.float32), np.zeros((N, 3), np.int32)
    counter = np.zeros(2, np.int32) if counter is None else _i32(counter)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    clib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), F(bound), F(dt_gamma), U(max_steps), U(N), U(C),
                                   U(H), U(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays), _p(counter),
                                   _p(noises))
    return xyzs, dirs, deltas, rays, counter


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, dt_gamma,
               max_steps, M=None):
    M = n_alive * n_step if M is None else M
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    clib().oracle_march_rays(U(n_alive), U(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), F(bound),
                             F(dt_gamma), U(max_steps), U(C), U(H), _p(bitfield), _p(nears), _p(fars), _p(xyzs), _p(dirs),
                             _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh, M=None):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M = sigmas.shape[0] if M is None else M
    N = rays.shape[0]
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    clib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), U(M), U(N), F(T_thresh), _p(ws),
                                               _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    grad_ws, grad_image, ws, image = _f32(grad_ws), _f32(grad_image), _f32(ws), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    clib().oracle_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                                _p(ws), _p(image), U(M), U(N), F(T_thresh), _p(gs), _p(gr))
    return gs, gr


def composite_rays_flex_train_forward(sigmas, inp, deltas, rays, T_thresh):
    sigmas, inp, deltas, rays = _f32(sigmas), _f32(inp), _f32(deltas), _i32(rays)
    M, N, nc = sigmas.shape[0], rays.shape[0], inp.shape[1]
    out = np.zeros((N, nc), np.float32)
    clib().oracle_composite_rays_flex_train_forward(_p(sigmas), _p(inp), _p(deltas), _p(rays), U(M), U(N), U(nc),
                                                    F(T_thresh), _p(out))
    return out


def composite_rays_flex_train_backward(grad_out, sigmas, deltas, rays, nc, T_thresh):
    grad_out, sigmas, deltas, rays = _f32(grad_out), _f32(sigmas), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    gi = np.zeros((M, nc), np.float32)
    clib().oracle_composite_rays_flex_train_backward(_p(grad_out), _p(sigmas), _p(deltas), _p(rays), U(M), U(N), U(nc),
                                                     F(T_thresh), _p(gi))
    return gi


def spread_ray_to_sample(inp, rays, M):
    inp, rays = _f32(inp), _i32(rays)
    N, nc = inp.shape
    out = np.zeros((M, nc), np.float32)
    clib().oracle_spread_ray_to_sample(_p(inp), _p(rays), U(M), U(N), U(nc), _p(out))
    return out


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh):
    """in place on copies; returns the updated (rays_alive, rays_t, weights_sum, depth, image)"""
    rays_alive, rays_t = _i32(rays_alive).copy(), _f32(rays_t).copy()
    ws, depth, image = _f32(weights_sum).copy(), _f32(depth).copy(), _f32(image).copy()
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    clib().oracle_composite_rays(U(n_alive), U(n_step), F(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs),
                                 _p(deltas), _p(ws), _p(depth), _p(image))
    return rays_alive, rays_t, ws, depth, image


def composite_rays_flex(n_alive, n_step, rays_alive, sigmas, inp, deltas, weights_sum, output, T_thresh):
    rays_alive, sigmas, inp, deltas = _i32(rays_alive), _f32(sigmas), _f32(inp), _f32(deltas)
    ws, out = _f32(weights_sum), _f32(output).copy()
    nc = inp.shape[1]
    clib().oracle_composite_rays_flex(U(n_alive), U(n_step), U(nc), F(T_thresh), _p(rays_alive), _p(sigmas), _p(inp),
                                      _p(deltas), _p(ws), _p(out))
    return out


def rgb_to_hsv(x):
    x = _f32(x).reshape(-1, 3)
    out = np.empty_like(x)
    clib().oracle_rgb_to_hsv(U(x.shape[0]), _p(x), _p(out))
    return out


def hsv_to_rgb(x):
    x = _f32(x).reshape(-1, 3)
    out = np.empty_like(x)
    clib().oracle_hsv_to_rgb(U(x.shape[0]), _p(x), _p(out))
    return out


def compute_rgb_histogram(colors, weights, bpc):
    colors, weights = _f32(colors).reshape(-1), _f32(weights)
    nb = 1 << (3 * bpc)
    bw, bc = np.zeros(nb, np.float64), np.zeros((nb, 3), np.float32)
    clib().oracle_rgb_histogram(_p(colors), _p(weights), ctypes.c_uint64(weights.shape[0]), ctypes.c_int(bpc), _p(bw),
                                _p(bc))
    return bw, bc


# ---------------------------------------------------------------- hash grid (numpy) --------------------------
_PRIMES = np.array([1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737], dtype=np.uint64)


def grid_offsets(input_dim=3, num_levels=16, base_resolution=16, log2_hashmap_size=19, per_level_scale=2.0,
                 align_corners=False):
    """table layout of gridencoder/grid.py:110-121"""
    offsets, offset, max_params = [], 0, 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def _level_setup(level, offsets, S, H, D, gridtype, align_corners, exp2_levels=None):
    """gridencoder.cu:97-99 + the index loop of :54-72.
    exp2_levels: optional fp32 array of exp2f(level * S) as the CUDA math library rounds it (<= 2 ulp, not correctly
    rounded): a 1-ulp difference in the scale of a 4096-cell level moves a sample by 2.4e-4 cells, which is visible
    against O(1) tables, so GPU tests pass torch.exp2 evaluated on the device (an independent route to the same
    libdevice function the reference kernel calls)."""
    size = int(offsets[level + 1]) - int(offsets[level])
    if exp2_levels is not None:
        e = np.float32(exp2_levels[level])
    else:
        e = np.exp2(np.float32(np.float32(level) * np.float32(S))).astype(np.float32)
    scale = np.float32(np.float64(e) * float(H) - 1.0)  # fmaf(exp2f(l*S), H, -1)
    resolution = int(np.ceil(scale)) + 1
    strides, stride = [], 1
    for d in range(D):
        if stride <= size:
            strides.append(stride)
            stride = (stride * (resolution if align_corners else resolution + 1)) & 0xFFFFFFFF
        else:
            strides.append(0)
    use_hash = (gridtype == 0) and (stride > size)
    return size, scale, strides, use_hash


def _corner_indices(pg, c, size, strides, use_hash, D):
    loc = [(pg[:, d] + ((c >> d) & 1)).astype(np.uint64) & 0xFFFFFFFF for d in range(D)]
    if use_hash:
        idx = np.zeros(pg.shape[0], np.uint64)
        for d in range(D):
            idx ^= (loc[d] * _PRIMES[d]) & 0xFFFFFFFF
    else:
        idx = np.zeros(pg.shape[0], np.uint64)
        for d in range(D):
            idx = (idx + loc[d] * np.uint64(strides[d])) & 0xFFFFFFFF
    return (idx % np.uint64(size)).astype(np.int64)


def _positions(inputs, scale, align_corners):
    pos = (inputs.astype(np.float64) * np.float64(scale) + (0.0 if align_corners else 0.5)).astype(np.float32)
    pg = np.floor(pos).astype(np.int64)
    frac = (pos - pg.astype(np.float32)).astype(np.float32)
    return pg, frac


def grid_encode_forward(inputs, embeddings, offsets, S, H, gridtype=0, align_corners=False, with_dy_dx=False,
                        exp2_levels=None):
    """gridencoder.cu:75-223. inputs [B,D] in [0,1]; embeddings [N,C]; returns float64 [B, L*C]
    (level-major within a row, i.e. what grid.py:52 hands to the user) and optionally dy_dx [B, L, D, C]."""
    inputs = _f32(inputs)
    emb = np.asarray(embeddings, dtype=np.float64)
    B, D = inputs.shape
    C, L = emb.shape[1], len(offsets) - 1
    out = np.zeros((B, L, C), np.float64)
    dy_dx = np.zeros((B, L, D, C), np.float64) if with_dy_dx else None
    oob = ((inputs < 0) | (inputs > 1)).any(axis=1)
    for level in range(L):
        size, scale, strides, use_hash = _level_setup(level, offsets, S, H, D, gridtype, align_corners, exp2_levels)
        pg, frac = _positions(inputs, scale, align_corners)
        table = emb[int(offsets[level]):int(offsets[level + 1])]
        for c in range(1 << D):
            w = np.ones(B, np.float32)
            for d in range(D):
                w = w * (frac[:, d] if (c >> d) & 1 else (np.float32(1) - frac[:, d]))
            idx = _corner_indices(pg, c, size, strides, use_hash, D)
            out[:, level, :] += w.astype(np.float64)[:, None] * table[idx]
        if with_dy_dx:
            for gd in range(D):
                others = [d for d in range(D) if d != gd]
                for c in range(1 << (D - 1)):
                    w = np.full(B, scale, np.float32)
                    corner = 0
                    for nd, d in enumerate(others):
                        bit = (c >> nd) & 1
                        w = w * (frac[:, d] if bit else (np.float32(1) - frac[:, d]))
                        corner |= bit << d
                    il = _corner_indices(pg, corner, size, strides, use_hash, D)
                    ir = _corner_indices(pg, corner | (1 << gd), size, strides, use_hash, D)
                    dy_dx[:, level, gd, :] += w.astype(np.float64)[:, None] * (table[ir] - table[il])
    out[oob] = 0
    if with_dy_dx:
        dy_dx[oob] = 0
        return out.reshape(B, L * C), dy_dx
    return out.reshape(B, L * C)


def grid_encode_backward(grad, inputs, n_entries, offsets, S, H, gridtype=0, align_corners=False, exp2_levels=None):
    """gridencoder.cu:226-313. grad [B, L*C] -> grad_embeddings float64 [n_entries, C] (exact scatter-add)."""
    inputs = _f32(inputs)
    B, D = inputs.shape
    L = len(offsets) - 1
    grad = np.asarray(grad, dtype=np.float64).reshape(B, L, -1)
    C = grad.shape[2]
    g = np.zeros((n_entries, C), np.float64)
    ok = ~((inputs < 0) | (inputs > 1)).any(axis=1)
    for level in range(L):
        size, scale, strides, use_hash = _level_setup(level, offsets, S, H, D, gridtype, align_corners, exp2_levels)
        pg, frac = _positions(inputs, scale, align_corners)
        for c in range(1 << D):
            w = np.ones(B, np.float32)
            for d in range(D):
                w = w * (frac[:, d] if (c >> d) & 1 else (np.float32(1) - frac[:, d]))
            idx = _corner_indices(pg, c, size, strides, use_hash, D) + int(offsets[level])
            np.add.at(g, idx[ok], w.astype(np.float64)[ok, None] * grad[ok, level, :])
    return g


# ---------------------------------------------------------------- SH / freq (numpy) ---------------------------
def sh_encode(inputs, degree, with_grad=False):
    """shencoder.cu:43-356 polynomials (derived independently in tools/gen_sh.py); float64"""
    return sh_eval(np.asarray(inputs, np.float64), degree, with_grad)


def freq_encode(inputs, degree):
    """freqencoder.cu:40-57: [x, sin(2^f x), sin(2^f x + pi/2)] blocks"""
    x = np.asarray(inputs, np.float64)
    out = [x]
    for f in range(degree):
        out.append(np.sin(x * 2.0 ** f))
        out.append(np.sin(x * 2.0 ** f + np.float32(np.pi / 2)))
    return np.concatenate(out, axis=-1)


# ---------------------------------------------------------------- ray generation (numpy) ----------------------
def get_rays(poses, intrinsics, H, W, inds=None):
    """get_rays, nerf/utils.py:52-151, the arithmetic after the pixel indices are drawn (:69-71, :132-149), float32:
    i = col + 0.5, j = row + 0.5; directions = ((i-cx)/fx, (j-cy)/fy, 1) / norm; rays_d = directions @ R^T;
    rays_o = poses[:, :3, 3] broadcast. inds: int [B, N] / [N] (row * W + col) or None = all pixels in order.
    Returns rays_o, rays_d [B, N, 3] float32."""
    poses = np.asarray(poses, np.float32).reshape(-1, 4, 4)
    B = poses.shape[0]
    fx, fy, cx, cy = (np.float32(v) for v in intrinsics)
    if inds is None:
        inds = np.arange(H * W, dtype=np.int64)
    inds = np.broadcast_to(np.asarray(inds, np.int64).reshape(-1, np.asarray(inds).shape[-1]), (B, np.asarray(inds).shape[-1]))
    i = (inds % W).astype(np.float32) + np.float32(0.5)
    j = (inds // W).astype(np.float32) + np.float32(0.5)
    xs = (i - cx) / fx
    ys = (j - cy) / fy
    zs = np.ones_like(xs)
    nrm = np.sqrt(xs * xs + ys * ys + zs * zs).astype(np.float32)
    d = np.stack([xs / nrm, ys / nrm, zs / nrm], axis=-1).astype(np.float32)
    R = poses[:, :3, :3]
    rays_d = np.zeros((B, d.shape[1], 3), np.float32)
    for k in range(3):   # accumulate in index order, float32
        acc = d[..., 0] * R[:, None, k, 0]
        acc = (acc + d[..., 1] * R[:, None, k, 1]).astype(np.float32)
        acc = (acc + d[..., 2] * R[:, None, k, 2]).astype(np.float32)
        rays_d[..., k] = acc
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).astype(np.float32).copy()
    return rays_o, rays_d
