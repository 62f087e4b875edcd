"""Host side of the fused field / fused renderer (csrc/fused.cu): weight packing, fp16 table cache, C-ABI calls.

`pack_weights` turns the model's nn.Linear weights into the shared-memory image the kernels expect: per layer
[n-tile][k-step][lane] -> the two 32-bit registers of the mma.m16n8k16 B fragment. The reference's concatenations are
folded into the packing:
    diff_net.0   [64,15]  -> K=16 : column 0 multiplies the sigma logit -> zero, columns 1..15 = geo features
    color_net.0  [64,31]  -> K=32 : 0..15 SH, 16 -> zero (logit), 17..31 geo
    basis_net.0  [64,35]  -> K=48 : 0..31 palette grid, 32..34 diffuse, rest zero
    heads                 -> K=16, N=24 : rows 0..12 offsets_radiance_net.weight, rows 13..16 omega_net.0.weight
Nothing here computes on the CPU at render time; FieldCache re-packs tables and weights on the device (see its docstring
for the freshness policy).
"""
import ctypes
from ctypes import c_float, c_uint32, c_void_p

import numpy as np
import torch

from . import _lib as L
from ._lib import ptr, stream

NB = 4          # bases supported by the fused kernels
CLIP_MAX = 16


class PaletteField(ctypes.Structure):
    """mirror of `struct pnerf_palette_field` (csrc/fused.cu)"""
    _fields_ = [("table_sigma", c_void_p), ("table_palette", c_void_p), ("table_clip", c_void_p), ("offsets", c_void_p),
                ("wpack", c_void_p), ("head_bias", c_void_p), ("palette", c_void_p),
                ("L", c_uint32), ("H", c_uint32), ("pred_clip", c_uint32), ("clip_dim", c_uint32),
                ("S", c_float), ("bound", c_float), ("density_scale", c_float), ("offsets_weight", c_float),
                ("view_dep_weight", c_float), ("table_sigma_palette", c_void_p), ("wpack_tc", c_void_p), ("model_kind", c_uint32)]


class PaletteEdit(ctypes.Structure):
    """mirror of `struct pnerf_palette_edit` (include/pnerf_b200.h): GUI-time edit evaluated inside the persistent renderer"""
    _fields_ = [("mode", c_uint32), ("weight_mode", c_uint32), ("delta_hsv", c_void_p), ("mean_xyz", c_void_p),
                ("mean_clip", c_void_p), ("std_xyz", c_float), ("std_clip", c_float), ("dI", c_void_p), ("dP", c_void_p),
                ("ddelta", c_void_p)]


P, U, F = c_void_p, c_uint32, c_float
L.register("pnerf_palette_field_forward", [P, P, U, P, P, P, P, P, P, P, P])
L.register("pnerf_palette_field_forward_tc", [P, P, U, P, P, P, P, P, P, P, P])
L.lib.pnerf_palette_tc_weight_bytes.argtypes = [U]
L.lib.pnerf_palette_tc_weight_bytes.restype = c_uint32
L.register("pnerf_palette_render_fused", [P, P, P, P, P, P, U, U, U, U, F, F, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P])
L.LAUNCHES["pnerf_palette_render_fused"] = 5  # candidates + pre-pass + 2 ordering kernels + persistent kernel
L.register("pnerf_palette_render_rays", [P, P, P, P, P, P, U, U, U, U, F, F, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P])
L.LAUNCHES["pnerf_palette_render_rays"] = 2   # candidate filter + the persistent warp-per-ray kernel
L.lib.pnerf_palette_render_rays_warps.restype = c_uint32
L.register("pnerf_palette_render_tc", [P, P, P, P, P, P, U, U, U, U, F, F, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, U, P])
L.LAUNCHES["pnerf_palette_render_tc"] = 3     # candidate filter + thread-per-ray pre-pass (runs) + the persistent kernel
L.lib.pnerf_palette_render_tc_warps.restype = c_uint32
L.lib.pnerf_palette_render_tc_runs_bytes.restype = c_uint32
L.register("pnerf_field_cache_tables", [P, P, P, U, P, P, P])
L.register("pnerf_field_cache_gather", [P, U, U, U, P, P, P])
L.register("pnerf_field_cache_merge", [P, P, P, P, P, P, P])


def _frag(W, n_pad, k_pad):
    """W [N,K] (torch, any float dtype) -> fp16 B-fragment image [n_pad/8][k_pad/16][32][4]"""
    Wp = torch.zeros(n_pad, k_pad, dtype=torch.float32, device=W.device)
    Wp[: W.shape[0], : W.shape[1]] = W.float()
    nt, ks = n_pad // 8, k_pad // 16
    lane = torch.arange(32, device=W.device)
    n = lane // 4
    k = (lane % 4) * 2
    out = torch.empty(nt, ks, 32, 4, dtype=torch.float32, device=W.device)
    for a in range(nt):
        for b in range(ks):
            rows = a * 8 + n
            base = b * 16 + k
            out[a, b, :, 0] = Wp[rows, base]
            out[a, b, :, 1] = Wp[rows, base + 1]
            out[a, b, :, 2] = Wp[rows, base + 8]
            out[a, b, :, 3] = Wp[rows, base + 9]
    return out.to(torch.float16).reshape(-1)


def pack_weights(model):
    """-> (wpack fp16 [units*4], head_bias fp32 [16]) on the model's device; layer order must match enum Layer"""
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    dev = sd["sigma_net.0.weight"].device
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731

    d0 = z(64, 16); d0[:, 1:16] = sd["diff_net.0.weight"].float()
    v0 = z(64, 32); v0[:, 0:16] = sd["color_net.0.weight"][:, 0:16].float(); v0[:, 17:32] = sd["color_net.0.weight"][:, 16:31].float()
    b0 = z(64, 48); b0[:, 0:35] = sd["basis_net.0.weight"].float()
    heads = z(24, 16)
    heads[0:13, 0:15] = sd["offsets_radiance_net.weight"].float()
    heads[13:17, 0:15] = sd["omega_net.0.weight"].float()
    layers = [
        _frag(sd["sigma_net.0.weight"], 64, 32), _frag(sd["sigma_net.1.weight"], 16, 64),
        _frag(d0, 64, 16), _frag(sd["diff_net.1.weight"], 64, 64), _frag(sd["diff_net.2.weight"], 8, 64),
        _frag(v0, 64, 32), _frag(sd["color_net.1.weight"], 64, 64), _frag(sd["color_net.2.weight"], 8, 64),
        _frag(b0, 64, 48), _frag(sd["basis_net.1.weight"], 16, 64), _frag(heads, 24, 16),
    ]
    if model.opt.pred_clip:
        layers += [_frag(sd["clip_net.0.weight"], 64, 32), _frag(sd["clip_net.1.weight"], 16, 64)]
    bias = z(16)
    bias[0:13] = sd["offsets_radiance_net.bias"].float()
    return torch.cat(layers).contiguous(), bias.contiguous()


def supported(model):
    """the fused kernels cover the reference's default architecture; anything else uses the unfused path"""
    try:
        enc = model.encoder
        return (model.num_basis == NB and model.opt.clip_dim <= CLIP_MAX and model.num_layers == 2 and model.hidden_dim == 64
                and model.geo_feat_dim == 15 and model.num_layers_color == 3 and enc.num_levels == 16 and enc.level_dim == 2
                and enc.input_dim == 3 and enc.gridtype == "hash" and not enc.align_corners
                and getattr(model.encoder_dir, "degree", 0) == 4 and model.bg_radius <= 0
                and getattr(enc, "log2_hashmap_size", 19) <= 24        # raw-word hashing of the fast gathers (grid_common.cuh)
                and model.encoder.embeddings.is_cuda)
    except AttributeError:
        return False


class FieldCache:
    """device-resident fp16 tables + packed weights of one model for the inference kernels.

    Freshness: the default (`model.fused_weights_static` unset / False) re-packs on EVERY call — three strided
    fp32->fp16 table conversions into persistent buffers (pnerf_field_cache_tables: 152 MB of HBM traffic, one launch) and
    one gather of the ~45 k MLP weight-image halfs, head bias and palette (pnerf_field_cache_gather). A version key alone is not enough: `param.data.copy_()` — what torch_ema's store / copy_to / restore do
    around every evaluate() of the reference trainer (nerf/utils.py:934-944), and what checkpoint loading or GUI edits
    may do — does not bump `Parameter._version`, so a cache keyed on it would render EMA weights after restore().
    `model.fused_weights_static = True` opts into the version key (interactive viewers that never write through .data).
    All buffers keep their addresses, so the C struct is built once."""

    def __init__(self, model):
        self.model = model
        self.key = None
        self.buf = None
        self.field = None

    def _version_key(self):
        m = self.model
        ps = list(m.parameters())
        return tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps) + (m.density_scale, m.offsets_weight,
                                                                                 m.view_dep_weight, id(m.basis_color))

    def _allocate(self):
        from . import fused_train
        m = self.model
        st = fused_train._state(m)
        dev = m.encoder.embeddings.device
        n = m.encoder.embeddings.shape[0]
        n_fwd, n_tc = int(st["n_fwd"]), int(st["tc_index"].numel())
        n_fwd_pad = (n_fwd + 7) // 8 * 8                      # keeps the tcgen05 image 16-byte aligned behind the mma image
        w16 = torch.zeros(n_fwd_pad + n_tc, dtype=torch.float16, device=dev)
        f32 = torch.zeros(16 + m.num_basis * 3, dtype=torch.float32, device=dev)
        self.buf = dict(
            pair=torch.empty(n, 2, 2, dtype=torch.float16, device=dev),
            clip=torch.empty(n, 2, dtype=torch.float16, device=dev) if m.opt.pred_clip else None,
            w16=w16, f32=f32, n_fwd_pad=n_fwd_pad,
            wpack=w16[:n_fwd], wpack_tc=w16[n_fwd_pad:], tc_index=st["tc_index"],
            bias=f32[:16], palette=f32[16:].view(m.num_basis, 3),
            offsets=m.encoder.offsets.contiguous(), index=st["index"][:st["n_fwd"]], names=st["names"], zero=st["zero"],
            device=dev)
        self._addr_key, self._addr = None, None
        b = self.buf
        f = PaletteField()
        # the kernels read both grids through the interleaved table; the separate-table pointers are kept non-NULL for the
        # ABI's argument check only (every table GridEncoder can build takes the interleaved fast path)
        f.table_sigma = f.table_palette = f.table_sigma_palette = ptr(b["pair"])
        f.table_clip = ptr(b["clip"])
        f.offsets, f.wpack, f.head_bias, f.palette = ptr(b["offsets"]), ptr(b["wpack"]), ptr(b["bias"]), ptr(b["palette"])
        f.wpack_tc = ptr(b["wpack_tc"])
        assert 2 * b["wpack_tc"].numel() == L.lib.pnerf_palette_tc_weight_bytes(int(bool(m.opt.pred_clip)))
        f.L, f.H = m.encoder.num_levels, m.encoder.base_resolution
        f.pred_clip, f.clip_dim = int(bool(m.opt.pred_clip)), m.opt.clip_dim
        f.S = float(np.float32(np.log2(m.encoder.per_level_scale)))
        self.field = f

    def get(self):
        m = self.model
        if self.buf is None or self.buf["device"] != m.encoder.embeddings.device:
            self._allocate()
            self.key = None
        f, b = self.field, self.buf
        f.bound, f.density_scale = float(m.bound), float(m.density_scale)
        f.offsets_weight, f.view_dep_weight = float(m.offsets_weight), float(m.view_dep_weight)
        if getattr(m, "fused_weights_static", False):
            key = self._version_key()
            if key == self.key:
                return f
            self.key = key
        src = self._addresses()
        if src is not None:
            # two launches (csrc/field_cache.cu): the tables as one HBM stream, every small tensor through an address table
            emb_c = m.encoder_clip.embeddings if b["clip"] is not None else None
            L.call("pnerf_field_cache_tables", ptr(m.encoder.embeddings), ptr(m.encoder_palette.embeddings), ptr(emb_c),
                   m.encoder.embeddings.shape[0], ptr(b["pair"]), ptr(b["clip"]), stream())
            L.call("pnerf_field_cache_gather", ptr(src), b["w16"].numel(), b["f32"].numel(), 16, ptr(b["w16"]), ptr(b["f32"]),
                   stream())
            # the two product layers of the tcgen05 image (field_tc.cuh layer table), after the gather has zeroed them
            L.call("pnerf_field_cache_merge", ptr(m.sigma_net[1].weight), ptr(m.diff_net[0].weight), ptr(m.basis_net[1].weight),
                   ptr(m.offsets_radiance_net.weight), ptr(m.omega_net[0].weight), ptr(b["wpack_tc"]), stream())
            return f
        with torch.no_grad():       # parameters that are not plain contiguous fp32 tensors: the same refresh as torch ops
            b["pair"][:, 0, :].copy_(m.encoder.embeddings.detach())
            b["pair"][:, 1, :].copy_(m.encoder_palette.embeddings.detach())
            if b["clip"] is not None:
                b["clip"].copy_(m.encoder_clip.embeddings.detach())
            sd = dict(m.named_parameters())
            flat = torch.cat([sd[n].detach().reshape(-1).float() for n in b["names"]] + [b["zero"]])
            b["wpack"].copy_(flat[b["index"]])
            b["wpack_tc"].copy_(flat[b["tc_index"]])
            from . import fused_train
            fused_train.tc_merge_layers(m, b["wpack_tc"])
            b["bias"][0:13].copy_(m.offsets_radiance_net.bias.detach())
            b["palette"].copy_(m.basis_color.detach().float().clamp(0, 1))
        return f

    def _addresses(self):
        """int64 device tensor: for every element of (w16 | f32) the ADDRESS of the fp32 parameter element it is made from
        (0 = the value 0). Rebuilt only when a parameter's storage moves. None when a source is not contiguous fp32."""
        m, b = self.model, self.buf
        sd = dict(m.named_parameters())
        tens = [sd[n] for n in b["names"]]
        bias, pal = m.offsets_radiance_net.bias, m.basis_color
        tabs = [m.encoder.embeddings, m.encoder_palette.embeddings] + ([m.encoder_clip.embeddings] if b["clip"] is not None else [])
        if any(t.dtype != torch.float32 or not t.is_contiguous() for t in tens + [bias, pal] + tabs):
            return None
        key = tuple(t.data_ptr() for t in tens) + (bias.data_ptr(), pal.data_ptr())
        if key != self._addr_key:
            sizes = [t.numel() for t in tens]
            flat = np.zeros(sum(sizes) + 1, dtype=np.int64)       # flat weight vector (fused_train._state) -> address; last = 0.0
            off = 0
            for t, k in zip(tens, sizes):
                flat[off:off + k] = t.data_ptr() + 4 * np.arange(k, dtype=np.int64)
                off += k
            a16 = np.zeros(b["w16"].numel(), dtype=np.int64)
            n_fwd = b["wpack"].numel()
            a16[:n_fwd] = flat[b["index"].cpu().numpy()]
            a16[b["n_fwd_pad"]:] = flat[b["tc_index"].cpu().numpy()]
            a32 = np.zeros(b["f32"].numel(), dtype=np.int64)
            a32[:13] = bias.data_ptr() + 4 * np.arange(13, dtype=np.int64)
            a32[16:] = pal.data_ptr() + 4 * np.arange(pal.numel(), dtype=np.int64)
            self._addr = torch.from_numpy(np.concatenate([a16, a32])).to(b["device"])
            self._addr_key = key
        return self._addr

    def invalidate(self):
        self.key = None


def _cache(model):
    c = getattr(model, "_fused_cache", None)
    if c is None:
        c = FieldCache(model)
        object.__setattr__(model, "_fused_cache", c)
    return c


FIELD_KERNEL = __import__("os").environ.get("PNERF_FIELD_KERNEL", "mma")    # "mma" (mma.sync chain) | "tc" (tcgen05 / TMEM)


@torch.no_grad()
def field_forward(model, xyzs, dirs, kernel=None):
    """fused PaletteNetwork.forward (eval): -> (sigma [M], clip [M,cd], omega [M,4], offsets_radiance [M,13],
    view_dep [M,3], diffuse [M,3]), fp32. kernel: "mma" = mma.sync register chain (csrc/fused.cu), "tc" = tcgen05 with TMEM
    accumulators (csrc/field_tc.cu)"""
    L.require_cuda(xyzs, dirs)
    f = _cache(model).get()
    xyzs, dirs = xyzs.contiguous().float(), dirs.contiguous().float()
    M, dev = xyzs.shape[0], xyzs.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
    sigma, omega, off_rad, view_dep, diffuse = e(M), e(M, NB), e(M, 13), e(M, 3), e(M, 3)
    cd = model.opt.clip_dim
    clip = e(M, cd) if model.opt.pred_clip else torch.zeros(M, cd, dtype=torch.float32, device=dev)
    entry = "pnerf_palette_field_forward_tc" if (kernel or FIELD_KERNEL) == "tc" else "pnerf_palette_field_forward"
    L.call(entry, ptr(xyzs), ptr(dirs), M, ctypes.addressof(f), ptr(sigma),
           ptr(clip) if model.opt.pred_clip else None, ptr(omega), ptr(off_rad), ptr(view_dep), ptr(diffuse), stream())
    return sigma, clip, omega, off_rad, view_dep, diffuse


_T_SCRATCH = {}


def _t_scratch(dev, max_steps, warps=None):
    """per-warp sample lists of the warp-per-ray renderers: [resident warps, max_steps] fp32, allocated once per device"""
    warps = int(L.lib.pnerf_palette_render_rays_warps()) if warps is None else int(warps)
    key = (str(dev), int(max_steps), warps)
    t = _T_SCRATCH.get(key)
    if t is None:
        t = _T_SCRATCH[key] = torch.empty(warps * int(max_steps), dtype=torch.float32, device=dev)
    return t


RENDER_KERNEL = __import__("os").environ.get("PNERF_RENDER_KERNEL", "tc")    # "rays" (round 2) | "lanes" (round 1, A/B)


def edit_struct(model, dev):
    """-> (PaletteEdit | None, tensors to keep alive) for model.edit (RegionEdit) / model.stylizer (Stylizer); the kernel reads
    the parameters through device pointers, so nothing is copied to the host"""
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
    if getattr(model, "stylizer", None) is not None:
        st = model.stylizer
        keep = (f32(st.dI), f32(st.dP).reshape(-1), f32(st.ddelta).reshape(-1))
        e = PaletteEdit()
        e.mode = 2
        e.dI, e.dP, e.ddelta = ptr(keep[0]), ptr(keep[1]), ptr(keep[2])
        return e, keep
    ed = getattr(model, "edit", None)
    if ed is None:
        return None, ()
    dh = f32(ed.delta_hsv).reshape(-1)
    mx = None if ed.mean_xyz is None else f32(ed.mean_xyz).reshape(-1)
    mc = None if ed.mean_clip is None else f32(ed.mean_clip).reshape(-1)
    e = PaletteEdit()
    e.mode, e.weight_mode = 1, int(bool(ed.weight_mode))
    e.delta_hsv, e.mean_xyz, e.mean_clip = ptr(dh), ptr(mx), ptr(mc)
    e.std_xyz, e.std_clip = float(ed.std_xyz), float(ed.std_clip)
    return e, (dh, mx, mc)


def accumulator_layout(N, nb, cd, gui_mode):
    """name -> (offset, shape) of the per-ray output maps inside ONE flat fp32 buffer, and its length in floats"""
    shapes = {"weights_sum": (N,), "depth": (N,), "image": (N, 3), "clip_feat": (N, cd)}
    if not gui_mode:
        shapes.update(direct_rgb=(N, 3), view_dep_rgb=(N, 3), basis_acc=(N, nb), basis_rgb=(N, 3 * nb),
                      unscaled_basis_rgb=(N, 3 * nb))
    lay, off = {}, 0
    for k, shp in shapes.items():
        lay[k] = (off, shp)
        off += int(np.prod(shp))
    return lay, off


def reproducible_render(model, out=None):
    """PNERF_RENDER_REPRODUCIBLE (include/pnerf_b200.h): bit-identical maps from run to run and from shard to shard, at ~4 %
    of the render time. On for tile-sharded views (their contract is equality with the single-GPU image), for models with
    `fused_reproducible = True`, and with PNERF_RENDER_REPRODUCIBLE=1 in the environment."""
    env = __import__("os").environ.get("PNERF_RENDER_REPRODUCIBLE")
    if env is not None:
        return env not in ("0", "")
    return out is not None or bool(getattr(model, "fused_reproducible", False))


@torch.no_grad()
def render(model, rays_o, rays_d, nears, fars, perturb, dt_gamma, max_steps, T_thresh, gui_mode, kernel=None, out=None,
           out_index=None):
    """persistent fused renderer -> dict of accumulators like PaletteRenderer._infer_loop.
    out / out_index (tile-sharded views, kernel "tc" only): `out` = dict of pre-zeroed output maps with as many rows as the
    WHOLE view (they may alias a peer GPU's memory); ray n of this call writes row out_index[n].
    kernel: "tc" = warp-per-ray kernel with the field on tcgen05 / TMEM (csrc/field_tc.cu), "rays" = warp-per-ray kernel on
    mma.sync (csrc/render_rays.cu), "lanes" = round 1's lane-per-ray kernel"""
    f = _cache(model).get()
    N, dev = rays_o.shape[0], rays_o.device
    nb, cd = model.num_basis, model.opt.clip_dim
    kernel = kernel or RENDER_KERNEL
    edit, edit_keep = edit_struct(model, dev)
    if edit is not None:
        kernel = "tc"                    # the edit epilogue exists in the tensor-core renderer only
        if edit.mode == 2:
            gui_mode = True              # the Stylizer produces no debug maps (palette/renderer.py:474-475, 496-506)
    if kernel == "tc" and out is None and edit is None and model.opt.pred_clip and \
            "PNERF_RENDER_KERNEL" not in __import__("os").environ:
        # the semantic branch gathers a third table; until that gather is interleaved with the other two the lane-per-ray
        # kernel renders such models faster (config 5: 17.8 vs 19.6 ms per 1297x840 view)
        kernel = "lanes"
    if out is not None:
        if kernel != "tc" or out_index is None:
            raise RuntimeError("sharded output needs the tensor-core renderer (kernel='tc') and an out_index map")
        acc = dict(out)
        queue = torch.zeros(68, dtype=torch.int32, device=dev)
    else:
        # every accumulator (and the queue counters) is a view of ONE zero-filled buffer: one fill launch per view
        lay, total = accumulator_layout(N, nb, cd, gui_mode)
        flat = torch.zeros(total + 68, dtype=torch.float32, device=dev)
        acc = {k: flat[o:o + int(np.prod(shp))].view(*shp) for k, (o, shp) in lay.items()}
        queue = flat[total:total + 68].view(torch.int32)         # counters (+ round 1's 32-bucket histogram and cursors)
    noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else None
    aux = (lambda k: ptr(acc[k])) if not gui_mode else (lambda k: None)
    from .raymarching.raymarching import occupied_bounds
    occ = occupied_bounds(model.density_bitfield, model.cascade, model.grid_size, model.bound) \
        if (model.grid_size ** 3) % 32 == 0 else None
    common = (ptr(rays_o), ptr(rays_d), ptr(nears), ptr(fars), ptr(noises), ptr(model.density_bitfield), N, model.cascade,
              model.grid_size, max_steps, float(dt_gamma), float(T_thresh), ctypes.addressof(f), ptr(acc["weights_sum"]),
              ptr(acc["depth"]), ptr(acc["image"]), aux("direct_rgb"), aux("view_dep_rgb"), aux("basis_acc"), aux("basis_rgb"),
              aux("unscaled_basis_rgb"), ptr(acc["clip_feat"]) if model.opt.pred_clip else None, ptr(queue))
    if kernel == "tc":
        cand = torch.empty(N, dtype=torch.int32, device=dev)
        runs = torch.empty(N * int(L.lib.pnerf_palette_render_tc_runs_bytes()), dtype=torch.uint8, device=dev)
        L.call("pnerf_palette_render_tc", *common, ptr(cand), ptr(runs),
               ptr(_t_scratch(dev, max_steps, L.lib.pnerf_palette_render_tc_warps())), ptr(occ), ptr(out_index),
               None if edit is None else ctypes.addressof(edit), 1 if reproducible_render(model, out) else 0, stream())
        del edit_keep
    elif kernel == "rays":
        cand = torch.empty(N, dtype=torch.int32, device=dev)
        L.call("pnerf_palette_render_rays", *common, ptr(cand), ptr(_t_scratch(dev, max_steps)), ptr(occ), stream())
    else:
        hit_list = torch.empty(2 * N, dtype=torch.int32, device=dev)   # ordered hit list + samples per ray
        t_first, t_last = torch.empty(N, dtype=torch.float32, device=dev), torch.empty(N, dtype=torch.float32, device=dev)
        L.call("pnerf_palette_render_fused", *common, ptr(hit_list), ptr(t_first), ptr(t_last), ptr(occ), stream())
    acc["_queue"] = queue   # [ray cursor, samples shaded, rays with samples, tiles evaluated, ...]; read lazily (no sync here)
    acc["_kernel"] = kernel
    return acc
