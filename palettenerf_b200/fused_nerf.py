"""Host side of the tensor-core kernels for the stage-1 model and the density-grid refresh (csrc/field_tc.cu with
model_kind = 1, csrc/density_tc.cu).

  * `render(model, ...)`            NeRFRenderer.run_cuda's inference loop (nerf/renderer.py:329-386 of the reference) as ONE
                                    persistent kernel: the warp-per-ray tensor-core renderer with the stage-1 field (one hash
                                    grid, sigma net, colour net on SH(4) ++ geo);
  * `density(model, xyzs)`          {NeRF,Palette}Network.density in eval mode (sigma only);
  * `update_density_grid(model..)`  NeRFRenderer.update_extra_state (nerf/renderer.py:467-561): cell selection + jitter +
                                    density + scatter in one kernel, EMA-max + mean + packbits in two more, the threshold
                                    never leaves the device. Data parallel: ranks evaluate disjoint tiles with a shared seed
                                    and merge the temporary grid with one all-reduce(max).
The weight image has the palette field's layout (csrc/field_tc.cuh::TcLayer); layers the model does not have stay zero.
"""
import ctypes
from ctypes import c_float, c_uint32, c_uint64, c_void_p

import numpy as np
import torch

from . import _lib as L
from . import fused
from ._lib import ptr, stream

P, U, F = c_void_p, c_uint32, c_float
L.register("pnerf_density_tc", [P, U, P, P, P])
L.register("pnerf_density_occupied_list", [P, U, U, P, P, P, P])
L.LAUNCHES["pnerf_density_occupied_list"] = 3
L.lib.pnerf_density_occupied_chunks.argtypes = [U]
L.lib.pnerf_density_occupied_chunks.restype = c_uint32
L.register("pnerf_density_grid_sweep", [P, U, U, F, F, U, U, P, P, c_uint64, U, U, P, P, P])
L.register("pnerf_density_grid_finalize", [P, P, U, U, F, F, P, P, P, P])
L.LAUNCHES["pnerf_density_grid_finalize"] = 2
L.lib.pnerf_density_finalize_partials.argtypes = [c_uint64]
L.lib.pnerf_density_finalize_partials.restype = c_uint32

# (layer, parameter, n_pad, k_pad, row range, source columns -> destination columns) of the stage-1 / density sub-network
_LAYERS = ["s0", "s1", "d0", "d1", "d2", "v0", "v1", "v2", "b0", "b1", "h"]
_SHAPE = {"s0": (64, 32), "s1": (16, 64), "d0": (64, 64), "d1": (64, 64), "d2": (16, 64), "v0": (64, 32), "v1": (64, 64),
          "v2": (16, 64), "b0": (64, 48), "b1": (32, 64), "h": (16, 16)}      # = csrc/field_tc.cuh::tc_n / tc_k


def _index(model, with_color):
    """(names, int64 index into cat(weights..., [0])) that builds the tcgen05 weight image for this model"""
    names = ["sigma_net.0.weight", "sigma_net.1.weight"] + (["color_net.0.weight", "color_net.1.weight", "color_net.2.weight"]
                                                            if with_color else [])
    sd = dict(model.named_parameters())
    base, off = {}, 0
    for n in names:
        base[n] = off
        off += sd[n].numel()
    zero = off

    def idx(n):
        return (base[n] + torch.arange(sd[n].numel())).reshape(tuple(sd[n].shape))
    mats = {k: torch.full(_SHAPE[k], zero, dtype=torch.int64) for k in _LAYERS}
    mats["s0"] = idx("sigma_net.0.weight")
    mats["s1"][:16] = idx("sigma_net.1.weight")
    if with_color:
        c0 = idx("color_net.0.weight")
        mats["v0"][:, 0:16] = c0[:, 0:16]
        mats["v0"][:, 17:32] = c0[:, 16:31]                # column 16 would see the sigma logit: stays zero
        mats["v1"] = idx("color_net.1.weight")
        mats["v2"][0:3] = idx("color_net.2.weight")
    parts = []
    for k in _LAYERS:
        W = mats[k]
        n, kk = W.shape
        parts.append(W.reshape(n, kk // 8, 8).permute(1, 0, 2).reshape(-1))
    return names, torch.cat(parts)


def supported(model):
    """the tensor-core kernels cover the reference's default stage-1 architecture (hash grid 16 x 2, 64-wide MLPs, SH(4))"""
    try:
        enc = model.encoder
        ok = (enc.num_levels == 16 and enc.level_dim == 2 and enc.input_dim == 3 and enc.gridtype == "hash"
              and not enc.align_corners and model.hidden_dim == 64 and model.geo_feat_dim == 15 and model.num_layers == 2
              and enc.embeddings.is_cuda and model.bg_radius <= 0 and getattr(enc, "log2_hashmap_size", 19) <= 24)
        return bool(ok)
    except AttributeError:
        return False


def supported_color(model):
    return (supported(model) and getattr(model, "num_layers_color", 0) == 3 and getattr(model.encoder_dir, "degree", 0) == 4
            and tuple(model.color_net[0].weight.shape) == (64, 31))


class _Field:
    """fp16 table + weight image of the sigma (+ colour) sub-network, re-packed on every call (see fused.FieldCache for why
    a version key is not enough); buffers keep their addresses"""

    def __init__(self, model, with_color):
        self.model, self.with_color = model, with_color
        self.buf = None

    def get(self):
        m = self.model
        dev = m.encoder.embeddings.device
        if self.buf is None or self.buf["dev"] != dev:
            names, index = _index(m, self.with_color)
            n = m.encoder.embeddings.shape[0]
            self.buf = dict(dev=dev, names=names, index=index.to(dev), table=torch.empty(n, 2, dtype=torch.float16, device=dev),
                            wimg=torch.empty(index.numel(), dtype=torch.float16, device=dev),
                            zero=torch.zeros(1, dtype=torch.float32, device=dev), bias=torch.zeros(16, dtype=torch.float32, device=dev),
                            pal=torch.zeros(12, dtype=torch.float32, device=dev), offsets=m.encoder.offsets.contiguous())
            assert 2 * index.numel() == L.lib.pnerf_palette_tc_weight_bytes(0)
        b = self.buf
        with torch.no_grad():
            b["table"].copy_(m.encoder.embeddings.detach())
            sd = dict(m.named_parameters())
            flat = torch.cat([sd[k].detach().reshape(-1).float() for k in b["names"]] + [b["zero"]])
            b["wimg"].copy_(flat[b["index"]])
        f = fused.PaletteField()
        f.table_sigma = ptr(b["table"])
        f.offsets, f.wpack_tc, f.head_bias, f.palette = ptr(b["offsets"]), ptr(b["wimg"]), ptr(b["bias"]), ptr(b["pal"])
        f.L, f.H = m.encoder.num_levels, m.encoder.base_resolution
        f.pred_clip, f.clip_dim = 0, 0
        f.S = float(np.float32(np.log2(m.encoder.per_level_scale)))
        f.bound, f.density_scale = float(m.bound), float(m.density_scale)
        f.offsets_weight = f.view_dep_weight = 1.0
        f.model_kind = 1
        return f


def _field(model, with_color):
    key = "_fused_nerf_color" if with_color else "_fused_nerf_sigma"
    c = getattr(model, key, None)
    if c is None:
        c = _Field(model, with_color)
        object.__setattr__(model, key, c)
    return c


@torch.no_grad()
def density(model, xyzs):
    """sigma [M] (fp32, not scaled by density_scale) of points [M,3] on the tensor-core field"""
    L.require_cuda(xyzs)
    f = _field(model, False).get()
    xyzs = xyzs.contiguous().float().view(-1, 3)
    sigma = torch.empty(xyzs.shape[0], dtype=torch.float32, device=xyzs.device)
    L.call("pnerf_density_tc", ptr(xyzs), xyzs.shape[0], ctypes.addressof(f), ptr(sigma), stream())
    return sigma


@torch.no_grad()
def render(model, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh):
    """persistent tensor-core renderer for the stage-1 model -> dict(weights_sum [N], depth [N], image [N,3]) accumulators"""
    f = _field(model, True).get()
    N, dev = rays_o.shape[0], rays_o.device
    flat = torch.zeros(5 * N + 68, dtype=torch.float32, device=dev)
    acc = {"weights_sum": flat[0:N], "depth": flat[N:2 * N], "image": flat[2 * N:5 * N].view(N, 3)}
    queue = flat[5 * N:5 * N + 68].view(torch.int32)
    noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else None
    from .raymarching.raymarching import occupied_bounds
    occ = occupied_bounds(model.density_bitfield, model.cascade, model.grid_size, model.bound) \
        if (model.grid_size ** 3) % 32 == 0 else None
    cand = torch.empty(N, dtype=torch.int32, device=dev)
    runs = torch.empty(N * int(L.lib.pnerf_palette_render_tc_runs_bytes()), dtype=torch.uint8, device=dev)
    L.call("pnerf_palette_render_tc", ptr(rays_o), ptr(rays_d), ptr(nears), ptr(fars), ptr(noises), ptr(model.density_bitfield), N,
           model.cascade, model.grid_size, max_steps, float(dt_gamma), float(T_thresh), ctypes.addressof(f), ptr(acc["weights_sum"]),
           ptr(acc["depth"]), ptr(acc["image"]), None, None, None, None, None, None, ptr(queue), ptr(cand), ptr(runs),
           ptr(fused._t_scratch(dev, max_steps, L.lib.pnerf_palette_render_tc_warps())), ptr(occ), None, None,
           1 if fused.reproducible_render(model) else 0, stream())
    acc["_queue"] = queue
    return acc


@torch.no_grad()
def update_density_grid(model, decay=0.95, jitter=None, seed=None):
    """density_grid / density_bitfield / mean density refresh on the device (see the module docstring). Returns the [2]
    device tensor (mean density, threshold). `jitter` (tests): explicit U[0,1) numbers [C * points_per_cascade, 3]."""
    from . import distributed as D
    dev = model.density_grid.device
    C, H = model.cascade, model.grid_size
    H3 = H ** 3
    f = _field(model, False).get()
    st = getattr(model, "_density_scratch", None)
    if st is None or st["tmp"].device != dev:
        st = dict(tmp=torch.full((C, H3), -1.0, dtype=torch.float32, device=dev),
                  partials=torch.empty(int(L.lib.pnerf_density_finalize_partials(C * H3)), dtype=torch.float32, device=dev),
                  stats=torch.zeros(2, dtype=torch.float32, device=dev), occ_list=None,
                  occ_count=torch.zeros(C, dtype=torch.int32, device=dev))
        object.__setattr__(model, "_density_scratch", st)
    ws, rank = D.world()
    if seed is None:
        seed = D.shared_seed() if ws > 1 else int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    partial = model.iter_density >= 16
    n_random = H3 // 4
    if partial:
        if st["occ_list"] is None:
            st["occ_list"] = torch.empty(C, H3, dtype=torch.int32, device=dev)
            st["occ_chunks"] = torch.empty(C * int(L.lib.pnerf_density_occupied_chunks(H)), dtype=torch.int32, device=dev)
        L.call("pnerf_density_occupied_list", ptr(model.density_grid), C, H, ptr(st["occ_list"]), ptr(st["occ_count"]),
               ptr(st["occ_chunks"]), stream())
    L.call("pnerf_density_grid_sweep", ptr(st["tmp"]), C, H, float(model.bound), float(model.density_scale), int(partial), n_random,
           ptr(st["occ_list"]), ptr(st["occ_count"]), int(seed), rank, ws, ptr(jitter), ctypes.addressof(f), stream())
    if ws > 1:
        D.merge_density(st["tmp"])                 # ranks evaluated disjoint tiles: element-wise max = the full temporary grid
    L.call("pnerf_density_grid_finalize", ptr(model.density_grid), ptr(st["tmp"]), C, H, float(decay), float(model.density_thresh),
           ptr(st["partials"]), ptr(model.density_bitfield), ptr(st["stats"]), stream())
    torch.autograd.graph.increment_version(model.density_grid)
    torch.autograd.graph.increment_version(model.density_bitfield)
    return st["stats"]
