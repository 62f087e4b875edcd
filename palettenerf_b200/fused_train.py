"""Host side of the fused TRAINING field (csrc/fused_train.cu): weight packing by one gather, the autograd Function that
routes gradients to the reference's parameters, and the mapping from the packed weight-gradient buffer to nn.Linear shapes.

Replaces the field evaluation of the training branch (palette/renderer.py:322-359 calling palette/network.py:156-280) and
its autograd graph. The reference's detach() placements are part of the contract (see fused_train.cu header).

Packing: the fp16 mma B-fragment images of all forward layers and all transposed (backward) layers are ONE
`flat[index]` gather from the concatenated fp32 parameters; `index` is built once per model (padding -> a trailing 0.0,
column permutations of the concatenated inputs, head bias -> column 15 of the head layer).
"""
import ctypes
from ctypes import c_float, c_uint32, c_uint64, c_void_p

import numpy as np
import torch
from torch.autograd import Function

from . import _lib as L
from ._lib import ptr, stream
from . import fused as _fused
from .arena import ARENA

NB = 4


class PaletteTrain(ctypes.Structure):
    """mirror of `struct pnerf_palette_train` (include/pnerf_b200.h)"""
    _fields_ = [("table_sigma", c_void_p), ("table_palette", c_void_p), ("table_clip", c_void_p), ("offsets", c_void_p),
                ("wfwd", c_void_p), ("wbwd", c_void_p), ("palette", c_void_p), ("m_dev", c_void_p),
                ("L", c_uint32), ("H", c_uint32), ("pred_clip", c_uint32), ("clip_dim", c_uint32),
                ("S", c_float), ("bound", c_float), ("density_scale", c_float), ("table_sigma_palette", c_void_p)]


P, U = c_void_p, c_uint32
L.register("pnerf_palette_train_forward", [P, P, U, P, P, P, P, P, P])
L.register("pnerf_palette_train_backward", [U, P, P, P, P, P, P, P, P, P, P])
L.register("pnerf_palette_train_wgrad", [U, U, P, P, P, P, P])
F_ = c_float
L.register("pnerf_palette_composite_train_forward", [P, P, P, P, P, U, U, U, F_, P, P, P, P, P])
L.register("pnerf_palette_composite_train_backward", [P, P, P, P, P, U, U, U, F_, P, P, P])
L.register("pnerf_palette_smooth_forward", [P, P, P, P, U, P, U, U, U, U, F_, F_, F_, F_, P, P])
L.register("pnerf_palette_smooth_backward", [P, P, P, P, P, U, P, U, U, U, U, P])
L.register("pnerf_grid_encode_backward_counted", [P, P, P, P, U, U, F_, U, U, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, F_, P])
for _n in ("pnerf_palette_train_xbuf_bytes", "pnerf_palette_train_ybuf_bytes"):
    getattr(L.lib, _n).argtypes = [U, U]
    getattr(L.lib, _n).restype = c_uint64
for _n in ("pnerf_palette_train_dw_floats", "pnerf_palette_train_wfwd_units", "pnerf_palette_train_wbwd_units"):
    getattr(L.lib, _n).argtypes = [U]
    getattr(L.lib, _n).restype = c_uint32

# parameter order of the flat vector (and of the Function's weight arguments)
WEIGHT_NAMES = ["sigma_net.0.weight", "sigma_net.1.weight", "diff_net.0.weight", "diff_net.1.weight", "diff_net.2.weight",
                "color_net.0.weight", "color_net.1.weight", "color_net.2.weight", "basis_net.0.weight", "basis_net.1.weight",
                "offsets_radiance_net.weight", "offsets_radiance_net.bias", "omega_net.0.weight"]
CLIP_NAMES = ["clip_net.0.weight", "clip_net.1.weight"]

# packed weight-gradient buffer: (name, N_pad, K_pad) in kernel order (fused_train.cu::DwLayer)
DW_LAYERS = [("D0", 64, 16), ("D1", 64, 64), ("D2", 16, 64), ("V0", 64, 32), ("V1", 64, 64), ("V2", 16, 64), ("B0", 64, 48),
             ("B1", 16, 64), ("H", 32, 16), ("C0", 64, 32), ("C1", 16, 64)]


def _frag_idx(Wi):
    """index matrix [n_pad, k_pad] (int64) -> B-fragment order [n_pad/8][k_pad/16][32 lanes][4]; same map as fused._frag"""
    n_pad, k_pad = Wi.shape
    nt, ks = n_pad // 8, k_pad // 16
    lane = torch.arange(32)
    n, k = lane // 4, (lane % 4) * 2
    out = torch.empty(nt, ks, 32, 4, dtype=torch.int64)
    for a in range(nt):
        for b in range(ks):
            rows, base = a * 8 + n, b * 16 + k
            out[a, b, :, 0] = Wi[rows, base]
            out[a, b, :, 1] = Wi[rows, base + 1]
            out[a, b, :, 2] = Wi[rows, base + 8]
            out[a, b, :, 3] = Wi[rows, base + 9]
    return out.reshape(-1)


def padded_layers(shapes, pred_clip, clip_dim):
    """-> ({layer: int64 index matrix [n_pad, k_pad] into the flat weight vector}, zero_slot): the reference's concatenations
    folded into zero-padded weight matrices (shared by the mma.sync fragment packing and the tcgen05 operand image)"""
    names = WEIGHT_NAMES + (CLIP_NAMES if pred_clip else [])
    base, off = {}, 0
    for nme in names:
        base[nme] = off
        off += int(np.prod(shapes[nme]))
    zero = off

    def idx(nme):
        return (base[nme] + torch.arange(int(np.prod(shapes[nme])))).reshape(tuple(shapes[nme]))

    def pad(n_pad, k_pad):
        return torch.full((n_pad, k_pad), zero, dtype=torch.int64)
    m = {}
    m["s0"], m["s1"] = idx("sigma_net.0.weight"), pad(16, 64)
    m["s1"][:16] = idx("sigma_net.1.weight")
    m["d0"] = pad(64, 16); m["d0"][:, 1:16] = idx("diff_net.0.weight")
    m["d1"] = idx("diff_net.1.weight")
    m["d2"] = pad(16, 64); m["d2"][0:3] = idx("diff_net.2.weight")
    c0 = idx("color_net.0.weight")
    m["v0"] = pad(64, 32); m["v0"][:, 0:16] = c0[:, 0:16]; m["v0"][:, 17:32] = c0[:, 16:31]
    m["v1"] = idx("color_net.1.weight")
    m["v2"] = pad(16, 64); m["v2"][0:3] = idx("color_net.2.weight")
    m["b0"] = pad(64, 48); m["b0"][:, 0:35] = idx("basis_net.0.weight")
    m["b1"] = pad(16, 64); m["b1"][0:15] = idx("basis_net.1.weight")
    m["h"] = pad(32, 16)
    m["h"][0:13, 0:15] = idx("offsets_radiance_net.weight")
    m["h"][13:17, 0:15] = idx("omega_net.0.weight")
    if pred_clip:
        m["q0"] = idx("clip_net.0.weight")
        m["q1"] = pad(16, 64); m["q1"][0:clip_dim] = idx("clip_net.1.weight")
    return m, zero


def tc_pack_index(shapes, pred_clip, clip_dim):
    """index (int64) that gathers the flat weight vector into the tcgen05 B-operand image of csrc/field_tc.cuh: per layer
    [k-chunk][n][8 halfs] (K-major, no swizzle), layers in TcLayer order. The layers d0 (64 x 64), b1 (32 x 64) are PRODUCTS of
    reference layers (field_tc.cuh layer table): the gather leaves them zero and `tc_merge_layers` / pnerf_field_cache_merge
    fills them; h (16 x 16) is unused padding."""
    m, zero = padded_layers(shapes, pred_clip, clip_dim)
    for k, shp in TC_MERGED_SHAPES.items():
        m[k] = torch.full(shp, zero, dtype=torch.int64)
    order = ["s0", "s1", "d0", "d1", "d2", "v0", "v1", "v2", "b0", "b1", "h"] + (["q0", "q1"] if pred_clip else [])
    parts = []
    for k in order:
        W = m[k]
        n, kk = W.shape
        parts.append(W.reshape(n, kk // 8, 8).permute(1, 0, 2).reshape(-1))
    return torch.cat(parts)


TC_MERGED_SHAPES = {"d0": (64, 64), "b1": (32, 64), "h": (16, 16)}
_TC_ORDER_SHAPES = [("s0", 64, 32), ("s1", 16, 64), ("d0", 64, 64), ("d1", 64, 64), ("d2", 16, 64), ("v0", 64, 32), ("v1", 64, 64),
                    ("v2", 16, 64), ("b0", 64, 48), ("b1", 32, 64), ("h", 16, 16)]


def tc_layer_offset(name):
    """offset (in halfs) of a layer in the tcgen05 weight image"""
    off = 0
    for k, n, kk in _TC_ORDER_SHAPES:
        if k == name:
            return off
        off += n * kk
    raise KeyError(name)


def tc_merge_layers(model, wimage):
    """torch restatement of pnerf_field_cache_merge (used when a parameter is not a plain contiguous fp32 tensor): writes the
    product layers d0 = diff_net.0 x sigma_net.1[1:16] and b1 = [offsets_radiance_net ; omega_net.0] x basis_net.1 into the image"""
    sd = dict(model.named_parameters())
    f = lambda n: sd[n].detach().float()  # noqa: E731
    d0 = f("diff_net.0.weight") @ f("sigma_net.1.weight")[1:16]
    b1 = torch.zeros(32, 64, dtype=torch.float32, device=d0.device)
    b1[0:13] = f("offsets_radiance_net.weight") @ f("basis_net.1.weight")
    b1[13:17] = f("omega_net.0.weight") @ f("basis_net.1.weight")
    for name, W in (("d0", d0), ("b1", b1)):
        n, kk = W.shape
        off = tc_layer_offset(name)
        wimage[off:off + n * kk].copy_(W.reshape(n, kk // 8, 8).permute(1, 0, 2).reshape(-1))


def build_pack_index(shapes, pred_clip, clip_dim):
    """shapes: {name: torch.Size}. -> (index int64 [fwd halfs + bwd halfs], n_fwd_halfs, zero_slot)"""
    names = WEIGHT_NAMES + (CLIP_NAMES if pred_clip else [])
    base, off = {}, 0
    for nme in names:
        base[nme] = off
        off += int(np.prod(shapes[nme]))
    zero = off                                  # index of the trailing 0.0

    def idx(nme):
        return (base[nme] + torch.arange(int(np.prod(shapes[nme])))).reshape(tuple(shapes[nme]))

    def pad(n_pad, k_pad):
        return torch.full((n_pad, k_pad), zero, dtype=torch.int64)

    s0, s1 = idx("sigma_net.0.weight"), pad(16, 64)
    s1[:16] = idx("sigma_net.1.weight")
    d0 = pad(64, 16); d0[:, 1:16] = idx("diff_net.0.weight")
    d1 = idx("diff_net.1.weight")
    d2 = pad(16, 64); d2[0:3] = idx("diff_net.2.weight")
    c0 = idx("color_net.0.weight")
    v0 = pad(64, 32); v0[:, 0:16] = c0[:, 0:16]; v0[:, 17:32] = c0[:, 16:31]
    v1 = idx("color_net.1.weight")
    v2 = pad(16, 64); v2[0:3] = idx("color_net.2.weight")
    b0 = pad(64, 48); b0[:, 0:35] = idx("basis_net.0.weight")
    b1 = pad(16, 64); b1[0:15] = idx("basis_net.1.weight")
    h = pad(32, 16)
    h[0:13, 0:15] = idx("offsets_radiance_net.weight")
    h[0:13, 15] = idx("offsets_radiance_net.bias")       # the head input carries a constant 1 in column 15
    h[13:17, 0:15] = idx("omega_net.0.weight")
    h_nobias = h.clone(); h_nobias[:, 15] = zero
    fwd = [_frag_idx(s0), _frag_idx(s1), _frag_idx(d0), _frag_idx(d1), _frag_idx(d2[:8]), _frag_idx(v0), _frag_idx(v1),
           _frag_idx(v2[:8]), _frag_idx(b0), _frag_idx(b1), _frag_idx(h[:24])]
    bwd = [_frag_idx(h_nobias.t().contiguous()), _frag_idx(b1.t().contiguous()), _frag_idx(b0[:, 0:32].t().contiguous()),
           _frag_idx(d2.t().contiguous()), _frag_idx(d1.t().contiguous()), _frag_idx(v2.t().contiguous()),
           _frag_idx(v1.t().contiguous())]
    if pred_clip:
        q0 = idx("clip_net.0.weight")
        q1 = pad(16, 64); q1[0:clip_dim] = idx("clip_net.1.weight")
        fwd += [_frag_idx(q0), _frag_idx(q1)]
        bwd += [_frag_idx(q1.t().contiguous()), _frag_idx(q0.t().contiguous())]
    fwd, bwd = torch.cat(fwd), torch.cat(bwd)
    return torch.cat([fwd, bwd]), fwd.numel(), zero


def dw_views(dwbuf, pred_clip, clip_dim):
    """packed fp32 weight-gradient buffer -> {parameter name: gradient in the parameter's own shape}"""
    v, off = {}, 0
    for nme, n, k in DW_LAYERS:
        if nme.startswith("C") and not pred_clip:
            break
        v[nme] = dwbuf[off:off + n * k].view(n, k)
        off += n * k
    g = {
        "diff_net.0.weight": v["D0"][:, 1:16], "diff_net.1.weight": v["D1"], "diff_net.2.weight": v["D2"][0:3],
        "color_net.0.weight": torch.cat([v["V0"][:, 0:16], v["V0"][:, 17:32]], dim=1), "color_net.1.weight": v["V1"],
        "color_net.2.weight": v["V2"][0:3], "basis_net.0.weight": v["B0"][:, 0:35], "basis_net.1.weight": v["B1"][0:15],
        "offsets_radiance_net.weight": v["H"][0:13, 0:15], "offsets_radiance_net.bias": v["H"][0:13, 15],
        "omega_net.0.weight": v["H"][13:17, 0:15],
    }
    if pred_clip:
        g["clip_net.0.weight"] = v["C0"]
        g["clip_net.1.weight"] = v["C1"][0:clip_dim]
    return g


class _Tables:
    """fp16 copies of the hash tables, each refreshed only when ITS parameter changed (the sigma grid is frozen in the
    palette stage, so it is converted once; the reference casts every table on every forward, gridencoder/grid.py:38-39)"""

    def __init__(self):
        self.c = {}

    def get(self, name, p, always=False):
        """always=True: convert unconditionally (tables that are trained change every step; a CUDA-graph capture must
        contain the conversion no matter what the cache looked like at capture time)"""
        key = (p._version, p.data_ptr())
        hit = self.c.get(name)
        if always and hit is not None and hit[1].shape == p.shape:
            # a trained table: FusedAdam keeps this fp16 copy current in its own pass (pnerf_adam_tensor.mirror, contiguous
            # half2 entries: offset 0, 4 bytes per entry) and records the parameter version it did that for
            mir = getattr(p, "_pnerf_half_mirror", None)
            if mir is not None and mir[0] is hit[1] and getattr(p, "_pnerf_mirror_version", None) == p._version:
                return hit[1]
            hit[1].copy_(p.detach())                  # same buffer every step (stable address for the mirror / a graph)
            p._pnerf_half_mirror, p._pnerf_mirror_version = (hit[1], 0, 4), None
            self.c[name] = (key, hit[1])
            return hit[1]
        if always or hit is None or hit[0] != key:
            hit = (key, p.detach().to(torch.float16).contiguous())
            self.c[name] = hit
        return hit[1]

    def pair(self, sigma_p, palette_p, refresh_palette):
        """density + palette tables interleaved per entry, fp16 [n, 2, 2] (one 8-byte gather per lattice corner serves
        both grids). The buffer is persistent: the frozen density half is written when that parameter changes, the
        trained palette half on every call (one strided fp32 -> fp16 copy, the conversion the step needs anyway)."""
        key_s = (sigma_p._version, sigma_p.data_ptr())
        key_p = (palette_p._version, palette_p.data_ptr())
        hit = self.c.get("pair")
        if hit is None or hit[2].shape[0] != sigma_p.shape[0] or hit[2].device != sigma_p.device:
            hit = [None, None, torch.empty(sigma_p.shape[0], 2, 2, dtype=torch.float16, device=sigma_p.device)]
            self.c["pair"] = hit
        if hit[0] != key_s:
            hit[2][:, 0, :].copy_(sigma_p.detach())
            hit[0] = key_s
        # FusedAdam writes fp16(p) into this buffer in its own pass (pnerf_adam_tensor.mirror) and records the parameter
        # version it did that for: then there is nothing to refresh
        mir = getattr(palette_p, "_pnerf_half_mirror", None)
        if mir is None or mir[0] is not hit[2]:
            palette_p._pnerf_half_mirror = (hit[2], 4, 8)      # (buffer, byte offset of the palette slot, bytes per entry)
            palette_p._pnerf_mirror_version = None
        if getattr(palette_p, "_pnerf_mirror_version", None) == palette_p._version and hit[1] is not None:
            hit[1] = key_p
            return hit[2]
        if refresh_palette or hit[1] != key_p:
            hit[2][:, 1, :].copy_(palette_p.detach())
            hit[1] = key_p
        return hit[2]


def _state(model):
    st = getattr(model, "_fused_train_state", None)
    dev = model.encoder.embeddings.device
    if st is None or st["device"] != dev:
        sd = dict(model.named_parameters())
        pred_clip, cd = bool(model.opt.pred_clip), int(model.opt.clip_dim)
        names = WEIGHT_NAMES + (CLIP_NAMES if pred_clip else [])
        shapes = {n: sd[n].shape for n in names}
        index, n_fwd, _ = build_pack_index(shapes, pred_clip, cd)
        st = dict(device=dev, index=index.to(dev), n_fwd=n_fwd, names=names, tables=_Tables(), pred_clip=pred_clip, cd=cd,
                  zero=torch.zeros(1, dtype=torch.float32, device=dev), tc_index=tc_pack_index(shapes, pred_clip, cd).to(dev))
        assert n_fwd == 4 * L.lib.pnerf_palette_train_wfwd_units(int(pred_clip))
        assert index.numel() - n_fwd == 4 * L.lib.pnerf_palette_train_wbwd_units(int(pred_clip))
        object.__setattr__(model, "_fused_train_state", st)
    return st


def supported(model):
    return _fused.supported(model)


class _TrainField(Function):
    """(model, count | None, xyzs, dirs, palette, embeddings_palette, embeddings_clip | None, *weights) -> sigma, rgb, flex
    count: optional int32 device tensor holding the number of valid rows of xyzs/dirs (static-capacity mode)"""

    @staticmethod
    def forward(ctx, model, count, xyzs, dirs, palette, emb_palette, emb_clip, *weights):
        st = _state(model)
        if any(ctx.needs_input_grad):
            st["pending"] = st.get("pending", 0) + 1        # forwards whose backward has not run yet (see backward: DP slots)
        L.require_cuda(xyzs, dirs, emb_palette)
        dev = xyzs.device
        xyzs, dirs = xyzs.detach().contiguous().float(), dirs.detach().contiguous().float()
        M, pc, cd = xyzs.shape[0], st["pred_clip"], st["cd"]
        nflex = 13 + cd + NB
        flat = torch.cat([w.detach().reshape(-1).float() for w in weights] + [st["zero"]])
        blob = flat[st["index"]].to(torch.float16)
        tabs = st["tables"]
        sig = model.encoder.embeddings
        if sig.shape == emb_palette.shape and sig.shape[1] == 2:
            t_pair = tabs.pair(sig, emb_palette, refresh_palette=emb_palette.requires_grad)
            t_sigma = t_pal = None
        else:
            t_pair = None
            t_sigma = tabs.get("sigma", sig)
            t_pal = tabs.get("palette", emb_palette, always=emb_palette.requires_grad)
        t_clip = tabs.get("clip", emb_clip, always=emb_clip.requires_grad) if pc else None
        pal = palette.detach().float().contiguous()
        offsets = model.encoder.offsets
        f = PaletteTrain()
        f.table_sigma, f.table_palette, f.table_clip = ptr(t_sigma), ptr(t_pal), ptr(t_clip)
        f.table_sigma_palette = ptr(t_pair)
        f.offsets, f.palette = ptr(offsets), ptr(pal)
        f.wfwd, f.wbwd = blob.data_ptr(), blob.data_ptr() + 2 * st["n_fwd"]
        f.m_dev = ptr(count)
        f.L, f.H, f.pred_clip, f.clip_dim = model.encoder.num_levels, model.encoder.base_resolution, int(pc), cd
        f.S = float(np.float32(np.log2(model.encoder.per_level_scale)))
        f.bound, f.density_scale = float(model.bound), float(model.density_scale)
        # static-capacity mode: the (large, data-independent) scratch comes from the arena; .detach() = a fresh alias
        # without autograd history, which also marks the buffer as in use for as long as the alias lives
        static = count is not None
        new = (lambda name, *shape, dtype=torch.float32: ARENA.get(name, shape, dtype, dev).detach()) if static else \
              (lambda name, *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device=dev))
        xbuf = new("xbuf", int(L.lib.pnerf_palette_train_xbuf_bytes(M, int(pc))), dtype=torch.uint8)
        sigma, rgb, flex = new("sigma", M), new("rgb", M, 3), new("flex", M, nflex)
        L.call("pnerf_palette_train_forward", ptr(xyzs), ptr(dirs), M, ctypes.addressof(f), ptr(xbuf), ptr(sigma), ptr(rgb),
               ptr(flex), stream())
        # `flex` is an OUTPUT: holding the returned object itself would tie ctx and its output into a reference cycle that only
        # the cyclic GC breaks — until then the arena sees the step's buffers (GBs at full capacity) as busy and allocates new
        # ones every eager step. A detached alias shares the storage without the cycle.
        ctx.keep = (f, blob, (t_sigma, t_pair), t_pal, t_clip, pal, offsets, xbuf, flex.detach(), xyzs, count)
        ctx.model, ctx.M, ctx.n_weights = model, M, len(weights)
        ctx.need_palette = palette.requires_grad
        ctx.mark_non_differentiable(sigma)
        return sigma, rgb, flex

    @staticmethod
    def backward(ctx, g_sigma, g_rgb, g_flex):
        model, M = ctx.model, ctx.M
        st = _state(model)
        pc, cd = st["pred_clip"], st["cd"]
        f, blob, t_sigma, t_pal, t_clip, pal, offsets, xbuf, flex, xyzs, count = ctx.keep
        dev = xyzs.device
        nflex = 13 + cd + NB
        g_rgb = torch.zeros(M, 3, device=dev) if g_rgb is None else g_rgb.contiguous().float()
        g_flex = torch.zeros(M, nflex, device=dev) if g_flex is None else g_flex.contiguous().float()
        static = count is not None
        new = (lambda name, *shape, dtype=torch.float32: ARENA.get(name, shape, dtype, dev).detach()) if static else \
              (lambda name, *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device=dev))
        ybuf = new("ybuf", int(L.lib.pnerf_palette_train_ybuf_bytes(M, int(pc))), dtype=torch.uint8)
        d_enc = new("d_enc", M, 32)
        d_enc_clip = new("d_enc_clip", M, 32) if pc else None
        d_pal = torch.zeros(NB, 3, dtype=torch.float32, device=dev) if ctx.need_palette else None
        dw = torch.zeros(int(L.lib.pnerf_palette_train_dw_floats(int(pc))), dtype=torch.float32, device=dev)
        L.call("pnerf_palette_train_backward", M, ctypes.addressof(f), ptr(xbuf), ptr(ybuf), ptr(g_rgb), ptr(g_flex), ptr(flex),
               ptr(d_enc), ptr(d_enc_clip), ptr(d_pal), stream())
        # hash-grid scatter of the feature gradients (run-length kernel, fp32 accumulation, [B, L*C] layout)
        from .gridencoder.backend import _backend as GB
        enc = model.encoder_palette
        x01 = ((xyzs + model.bound) / (2 * model.bound)).contiguous() if count is None else None
        S_ = float(np.float32(np.log2(enc.per_level_scale)))
        # data parallel (distributed.GradBucket attached to the model): the table gradients are scattered STRAIGHT INTO the
        # bucket (no pack copy) and their all-reduce starts on a side stream before the weight-gradient kernel below runs
        # Only when this is the step's ONLY field evaluation and nothing has been accumulated into the tables' .grad yet
        # (a second evaluation — the smooth loss — or gradient accumulation would add to a buffer already in flight).
        bucket = getattr(model, "_grad_bucket", None)
        pending = st.get("pending", 1)
        st["pending"] = max(0, pending - 1)
        if st.get("group_left", 0) > 0:                     # a later backward of a group of several evaluations
            st["group_left"] -= 1
            bucket = None
        elif pending > 1:                                   # the first backward of such a group
            st["group_left"] = pending - 1
            bucket = None
        if model.encoder_palette.embeddings.grad is not None or (pc and model.encoder_clip.embeddings.grad is not None):
            bucket = None

        def scatter(emb, d):
            g = bucket.slot(emb) if bucket is not None else None
            if g is None:
                g = L.zeros_like_fast(emb, torch.float32)        # (a memset node: 21 -> 8 us for the 50 MB of a table)
            else:
                L.zero_(g)
            if count is None:
                GB.grid_encode_backward_blc(d, x01, g, offsets, g, M, 3, 2, enc.num_levels, S_, enc.base_resolution, None,
                                            None, 0, False)
            else:   # number of rows taken from device memory
                L.call("pnerf_grid_encode_backward_counted", ptr(d), ptr(xyzs), ptr(offsets), ptr(g), M, enc.num_levels, S_,
                       enc.base_resolution, 0, 0, L.F32, L.LAYOUT_BLC, ptr(count), float(model.bound), stream())
            return g
        emb_pal, emb_clip = model.encoder_palette.embeddings, model.encoder_clip.embeddings
        g_pal_tab = scatter(emb_pal, d_enc) if M > 0 else torch.zeros_like(emb_pal)
        g_clip_tab = (scatter(emb_clip, d_enc_clip) if M > 0 else torch.zeros_like(emb_clip)) if pc else None
        if bucket is not None and M > 0:
            bucket.early([emb_pal] + ([emb_clip] if pc else []))
        # basis_net: the reference leaves it out of get_params (palette/network.py:283-308), so nothing ever steps it and its
        # weight gradients are dead work; they are computed only when the model asks (`fused_basis_net_grad = True`),
        # otherwise those parameters receive NO gradient (None) on this path
        basis = bool(getattr(model, "fused_basis_net_grad", False))
        L.call("pnerf_palette_train_wgrad", M, int(pc) | (2 if basis else 0), ptr(xbuf), ptr(ybuf), ptr(dw), ptr(count), stream())
        gw = dw_views(dw, pc, cd)
        grads = [None if (not basis and n.startswith("basis_net")) else gw.get(n) for n in st["names"]]   # sigma_net.* -> None
        return (None, None, None, None, d_pal, g_pal_tab, g_clip_tab, *grads)


def field(model, xyzs, dirs, palette, count=None):
    """fused training field: -> (sigmas [M] (density_scale applied, no grad), rgbs [M,3], channels [M, 13+clip+Nb]).
    count: optional int32 device tensor (1 element) = number of valid rows; rows beyond it are neither read nor written."""
    st = _state(model)
    if "weights" not in st:
        sd = dict(model.named_parameters())
        st["weights"] = [sd[n] for n in st["names"]]
    emb_clip = model.encoder_clip.embeddings if st["pred_clip"] else None
    return _TrainField.apply(model, count, xyzs, dirs, palette, model.encoder_palette.embeddings, emb_clip, *st["weights"])


class _SmoothGate(Function):
    """smooth-loss channel (ref: palette/renderer.py:360-381) from the field's channel rows and the rows of the same field at
    jittered positions (csrc/loss.cu: pnerf_palette_smooth_forward / _backward): column 3 of `channels` is written IN PLACE
    (the field's backward never reads that column), the gate is a constant of the gradient, rows beyond `count` are not
    touched. Replaces ~30 elementwise launches over the full static capacity (3.7 ms per step) by two kernels."""

    @staticmethod
    def forward(ctx, channels, ch_j, xyzs, xyzs_j, count, dims):
        nb, cd, pc, bound, s_xyz, s_color, s_clip = dims
        assert channels.is_contiguous() and ch_j.is_contiguous() and channels.dtype == torch.float32 == ch_j.dtype
        M, nflex = channels.shape
        xyzs, xyzs_j = xyzs.contiguous().float(), xyzs_j.contiguous().float()
        gate = ARENA.get("smooth_gate", (M,), torch.float32, channels.device).detach()
        L.call("pnerf_palette_smooth_forward", ptr(channels), ptr(ch_j), ptr(xyzs), ptr(xyzs_j), M, ptr(count), nflex, cd, nb,
               int(pc), float(bound), float(s_xyz), float(s_color), float(s_clip), ptr(gate), stream())
        ctx.mark_dirty(channels)
        ctx.save_for_backward(channels, ch_j, gate)
        ctx.count, ctx.dims = count, (M, nflex, nb, cd, int(pc))
        return channels

    @staticmethod
    def backward(ctx, g):
        channels, ch_j, gate = ctx.saved_tensors
        M, nflex, nb, cd, pc = ctx.dims
        g = g.contiguous().float()
        g_j = ARENA.get("g_ch_j", (M, nflex), torch.float32, g.device).detach()
        L.call("pnerf_palette_smooth_backward", ptr(g), ptr(g_j), ptr(channels), ptr(ch_j), ptr(gate), M, ptr(ctx.count), nflex,
               cd, nb, pc, stream())
        return g, g_j, None, None, None, None


def smooth_gate(model, channels, ch_j, xyzs, xyzs_j, count):
    o = model.opt
    dims = (model.num_basis, o.clip_dim, bool(o.pred_clip), model.bound, o.smooth_sigma_xyz, o.smooth_sigma_color,
            o.smooth_sigma_clip if o.pred_clip else 0.0)
    return _SmoothGate.apply(channels, ch_j, xyzs, xyzs_j, count, dims)


class _CompositeTrain(Function):
    """one-pass compositor of the palette training step (csrc/composite_palette.cu): (sigmas, rgbs, channels, deltas, rays)
    -> weights_sum [N], depth [N], image [N,3], maps [N, nflex]. Gradients flow to rgbs and channels only (sigma is a
    constant of this stage); the backward kernel writes every sample row, so no gradient buffer is memset."""

    @staticmethod
    def forward(ctx, sigmas, rgbs, channels, deltas, rays, T_thresh):
        sigmas, rgbs, channels, deltas = (t.contiguous().float() for t in (sigmas, rgbs, channels, deltas))
        M, N, nf = sigmas.shape[0], rays.shape[0], channels.shape[1]
        dev = sigmas.device
        ws, depth = torch.empty(N, device=dev), torch.empty(N, device=dev)
        image, maps = torch.empty(N, 3, device=dev), torch.empty(N, nf, device=dev)
        L.call("pnerf_palette_composite_train_forward", ptr(sigmas), ptr(rgbs), ptr(channels), ptr(deltas), ptr(rays), M, N, nf,
               float(T_thresh), ptr(ws), ptr(depth), ptr(image), ptr(maps), stream())
        ctx.save_for_backward(sigmas, deltas, rays)
        ctx.dims = (M, N, nf, float(T_thresh))
        ctx.mark_non_differentiable(depth)
        return ws, depth, image, maps

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image, g_maps):
        sigmas, deltas, rays = ctx.saved_tensors
        M, N, nf, T_thresh = ctx.dims
        dev = sigmas.device
        g_image = torch.zeros(N, 3, device=dev) if g_image is None else g_image.contiguous().float()
        g_maps = torch.zeros(N, nf, device=dev) if g_maps is None else g_maps.contiguous().float()
        g_rgbs = ARENA.get("g_rgbs", (M, 3), torch.float32, dev).detach()
        g_ch = ARENA.get("g_ch", (M, nf), torch.float32, dev).detach()
        L.call("pnerf_palette_composite_train_backward", ptr(g_image), ptr(g_maps), ptr(sigmas), ptr(deltas), ptr(rays), M, N, nf,
               T_thresh, ptr(g_rgbs), ptr(g_ch), stream())
        return None, g_rgbs, g_ch, None, None, None


def composite(sigmas, rgbs, channels, deltas, rays, T_thresh=1e-4):
    return _CompositeTrain.apply(sigmas, rgbs, channels, deltas, rays, T_thresh)
