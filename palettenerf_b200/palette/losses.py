"""palette.losses — the per-ray losses of PaletteTrainer.train_step (ref: palette/utils.py:486-567) as ONE kernel.

    loss, terms, per_ray = palette_loss(outputs, gt_rgb, lambda_sparsity=..., lambda_offsets=..., lambda_view_dep=..., ...)

`outputs` is the dict PaletteRenderer.render returns in training mode. The reference sums ~10 small tensor expressions
(MSE on image and direct_rgb, feature MSE, three regulariser means, smooth / blending-weight / palette terms): ~16
launch-bound kernels forward and ~25 backward on [N]-sized maps. Here `pnerf_palette_loss` computes the value, the
lambda-weighted terms (the reference's `loss_dict`) and the gradients of every input in one launch; backward is one
`pnerf_scale_buffers` launch (gradients x the upstream gradient, i.e. the GradScaler's scale).

The regulariser / feature / blending-weight entries of `outputs` are column views of ONE [N, C] tensor (the renderer's
channel composite): the loss reads them in place and hands autograd a single [N, C] gradient for that tensor instead of
one zero-filled tensor per slice. Entries that are not such views (another renderer, hand-made dicts) are gathered into a
temporary [N, C] tensor first — same results, a few more launches.

Not covered (caller adds them in torch if used): the patch-smooth term (:495-518) and the error-map update (:575-596;
`per_ray` is the quantity it needs).
"""
import ctypes
from ctypes import c_float, c_int32, c_uint32, c_void_p

import torch
from torch.amp import custom_bwd, custom_fwd
from torch.autograd import Function

from .. import _lib as L
from .._lib import ptr, stream

TERMS = ("total", "rgb", "direct", "clip_feat", "sparsity", "offsets", "view_dep", "smooth", "weight", "palette")


class _Args(ctypes.Structure):
    _fields_ = [("image", c_void_p), ("direct_rgb", c_void_p), ("gt_rgb", c_void_p), ("maps", c_void_p), ("stride", c_uint32),
                ("col_sparsity", c_int32), ("col_offsets", c_int32), ("col_view_dep", c_int32), ("col_smooth", c_int32),
                ("col_clip", c_int32), ("col_basis", c_int32), ("clip_dim", c_uint32), ("num_basis", c_uint32),
                ("gt_clip", c_void_p), ("gt_weights", c_void_p), ("basis_color", c_void_p), ("basis_color_origin", c_void_p),
                ("lambda_sparsity", c_float), ("lambda_offsets", c_float), ("lambda_view_dep", c_float),
                ("lambda_smooth", c_float), ("lambda_weight", c_float), ("lambda_palette", c_float), ("N", c_uint32),
                ("terms", c_void_p), ("per_ray", c_void_p), ("g_image", c_void_p), ("g_direct", c_void_p),
                ("g_maps", c_void_p), ("g_basis_color", c_void_p), ("partials", c_void_p)]


L.register("pnerf_palette_loss", [c_void_p, c_void_p])
L.lib.pnerf_palette_loss_partials.argtypes = [c_uint32]
L.lib.pnerf_palette_loss_partials.restype = c_uint32
L.LAUNCHES["pnerf_palette_loss"] = 2
L.register("pnerf_scale_buffers", [c_void_p, c_uint32, c_void_p, c_uint32, c_void_p, c_uint32, c_void_p, c_uint32, c_void_p,
                                   c_void_p])


class _PaletteLoss(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, image, direct_rgb, maps, basis_color, gt_rgb, gt_clip, gt_weights, basis_color_origin, cols, lambdas):
        L.require_cuda(image, direct_rgb, maps, basis_color, gt_rgb, gt_clip, gt_weights, basis_color_origin)
        ctx.shapes = (image.shape, direct_rgb.shape, maps.shape, None if basis_color is None else basis_color.shape)
        image, direct_rgb, gt_rgb = (t.reshape(-1, 3).contiguous() for t in (image, direct_rgb, gt_rgb))
        maps = maps.contiguous()
        N, stride = maps.shape
        if image.shape[0] != N or direct_rgb.shape[0] != N or gt_rgb.shape[0] != N:
            raise RuntimeError("palette_loss: image / direct_rgb / gt_rgb / maps disagree on the number of rays")
        dev = image.device
        a = _Args()
        a.image, a.direct_rgb, a.gt_rgb, a.maps, a.stride, a.N = ptr(image), ptr(direct_rgb), ptr(gt_rgb), ptr(maps), stride, N
        a.col_sparsity, a.col_offsets, a.col_view_dep, a.col_smooth = (cols[k] for k in ("sparsity", "offsets", "view_dep", "smooth"))
        a.col_clip, a.clip_dim = cols["clip"], cols["clip_dim"]
        a.col_basis, a.num_basis = cols["basis"], cols["num_basis"]
        keep = [image, direct_rgb, gt_rgb, maps]
        if a.col_clip >= 0:
            gt_clip = gt_clip.reshape(N, a.clip_dim).contiguous().float()
            a.gt_clip = ptr(gt_clip)
            keep.append(gt_clip)
        if a.col_basis >= 0:
            gt_weights = gt_weights.reshape(N, a.num_basis).contiguous().float()
            a.gt_weights = ptr(gt_weights)
            keep.append(gt_weights)
        g_basis = None
        if basis_color is not None:
            basis_color = basis_color.contiguous()
            basis_color_origin = basis_color_origin.contiguous().float()
            if cols["num_basis"] == 0:
                a.num_basis = basis_color.shape[0]
            if basis_color.shape != (a.num_basis, 3) or basis_color_origin.shape != basis_color.shape:
                raise RuntimeError("palette_loss: basis_color / basis_color_origin must be [num_basis, 3]")
            g_basis = torch.empty_like(basis_color)
            a.basis_color, a.basis_color_origin, a.g_basis_color = ptr(basis_color), ptr(basis_color_origin), ptr(g_basis)
            keep += [basis_color, basis_color_origin]
        (a.lambda_sparsity, a.lambda_offsets, a.lambda_view_dep, a.lambda_smooth, a.lambda_weight,
         a.lambda_palette) = (float(v) for v in lambdas)
        terms = torch.empty(len(TERMS), dtype=torch.float32, device=dev)
        per_ray = torch.empty(N, dtype=torch.float32, device=dev)
        g_image, g_direct, g_maps = torch.empty_like(image), torch.empty_like(direct_rgb), torch.empty_like(maps)
        partials = torch.empty(L.lib.pnerf_palette_loss_partials(N), dtype=torch.float32, device=dev)
        a.partials = ptr(partials)
        a.terms, a.per_ray, a.g_image, a.g_direct, a.g_maps = ptr(terms), ptr(per_ray), ptr(g_image), ptr(g_direct), ptr(g_maps)
        L.call("pnerf_palette_loss", ctypes.addressof(a), stream())
        del keep
        ctx.grads = (g_image, g_direct, g_maps, g_basis)
        ctx.mark_non_differentiable(terms, per_ray)
        return terms[0], terms, per_ray

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g_total, _g_terms, _g_per_ray):
        g_image, g_direct, g_maps, g_basis = ctx.grads
        ctx.grads = None
        s = g_total.reshape(1).to(torch.float32).contiguous()
        L.call("pnerf_scale_buffers", ptr(g_image), g_image.numel(), ptr(g_direct), g_direct.numel(), ptr(g_maps),
               g_maps.numel(), ptr(g_basis), 0 if g_basis is None else g_basis.numel(), ptr(s), stream())
        si, sd, sm, sb = ctx.shapes
        return (g_image.view(si), g_direct.view(sd), g_maps.view(sm), None if g_basis is None else g_basis.view(sb),
                None, None, None, None, None, None)


def _column_view(t, width):
    """(base [N, C] contiguous, first column) if element (n, j) of `t` is element (n, col + j) of such a base, else None"""
    base = t._base
    if base is None or base.dim() != 2 or not base.is_contiguous() or base.dtype != torch.float32 or t.numel() == 0:
        return None
    N, C = base.shape
    if t.numel() != N * width:
        return None
    try:
        v = t.view(N, width)
    except RuntimeError:
        return None
    if (N > 1 and v.stride(0) != C) or (width > 1 and v.stride(1) != 1):
        return None
    col = v.storage_offset() - base.storage_offset()
    if col < 0 or col + width > C:
        return None
    return base, col


def palette_loss(outputs, gt_rgb, lambda_sparsity=0.0, lambda_offsets=0.0, lambda_view_dep=0.0, lambda_smooth=0.0,
                 gt_clip_feat=None, gt_weights=None, lambda_weight=0.0, basis_color=None, basis_color_origin=None,
                 lambda_palette=0.0):
    """-> (loss, terms, per_ray). loss: 0-dim tensor with autograd history; terms: [10] detached tensor in the order of
    `TERMS` (lambda-weighted, = the reference's loss_dict values; `dict(zip(TERMS, terms))` after one D2H copy);
    per_ray: [N] detached mean_c (image - gt)^2. Terms: gt_clip_feat given -> feature MSE on outputs['clip_feat'];
    gt_weights given -> blending-weight supervision on outputs['basis_acc']; basis_color given -> palette term."""
    image, direct = outputs["image"], outputs["direct_rgb"]
    want = [("sparsity", "omega_sparsity", 1, True), ("offsets", "offsets_norm", 1, True), ("view_dep", "view_dep_norm", 1, True),
            ("smooth", "smooth_norm", 1, lambda_smooth != 0.0 and "smooth_norm" in outputs),
            ("clip", "clip_feat", None, gt_clip_feat is not None), ("basis", "basis_acc", None, gt_weights is not None)]
    used = [(name, outputs[key], (outputs[key].shape[-1] if w is None else w)) for name, key, w, on in want if on]
    cols = {"sparsity": -1, "offsets": -1, "view_dep": -1, "smooth": -1, "clip": -1, "basis": -1, "clip_dim": 0, "num_basis": 0}
    views = [(_column_view(t, w)) for _, t, w in used]
    if views and all(v is not None for v in views) and all(v[0] is views[0][0] for v in views):
        maps = views[0][0]
        for (name, _, w), (_, col) in zip(used, views):
            cols[name] = col
            if name == "clip":
                cols["clip_dim"] = w
            if name == "basis":
                cols["num_basis"] = w
    else:   # generic inputs: gather the columns (autograd scatters the gradient back)
        N = image.reshape(-1, 3).shape[0]
        parts, c = [], 0
        for name, t, w in used:
            parts.append(t.reshape(N, w).float())
            cols[name] = c
            if name == "clip":
                cols["clip_dim"] = w
            if name == "basis":
                cols["num_basis"] = w
            c += w
        maps = torch.cat(parts, dim=1)
    lambdas = (lambda_sparsity, lambda_offsets, lambda_view_dep, lambda_smooth, lambda_weight, lambda_palette)
    return _PaletteLoss.apply(image, direct, maps, basis_color, gt_rgb, gt_clip_feat, gt_weights, basis_color_origin, cols,
                              lambdas)
