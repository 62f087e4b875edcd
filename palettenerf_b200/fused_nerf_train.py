"""Host side of the fused TRAINING kernels of the stage-1 model (csrc/nerf_train.cu).

Replaces, inside NeRFRenderer.run_cuda's training branch (nerf/renderer.py:282-330 of the reference):
  * `sigmas, rgbs = self(xyzs, dirs)` (NeRFNetwork.forward, nerf/network.py:78-124) and its autograd graph by one forward
    kernel, one data-gradient kernel, the split-K weight-gradient kernel and the run-length hash-grid backward;
  * spread_ray_to_sample + the per-sample squared error + the two composite_rays_train calls (:301-327) by one compositor
    forward and one backward (`composite`).
Geometry is trained in this stage: gradients reach `sigma_net` and the density hash grid (unlike the palette stage).
Weight packing follows fused_train.py: the fp16 mma B-fragment images of the forward and the transposed layers are ONE
`flat[index]` gather from the concatenated fp32 parameters.
"""
import ctypes
from ctypes import c_float, c_uint32, c_uint64, c_void_p

import numpy as np
import torch
from torch.autograd import Function

from . import _lib as L
from . import fused_nerf
from ._lib import ptr, stream
from .arena import ARENA
from .fused_train import _Tables, _frag_idx


class NerfTrain(ctypes.Structure):
    """mirror of `struct pnerf_nerf_train` (include/pnerf_b200.h)"""
    _fields_ = [("table", c_void_p), ("offsets", c_void_p), ("wfwd", c_void_p), ("wbwd", c_void_p), ("m_dev", c_void_p),
                ("L", c_uint32), ("H", c_uint32), ("S", c_float), ("bound", c_float), ("density_scale", c_float)]


P, U, F_ = c_void_p, c_uint32, c_float
L.register("pnerf_nerf_train_forward", [P, P, U, P, P, P, P, P])
L.register("pnerf_nerf_train_backward", [U, P, P, P, P, P, P, P, P, P])
L.register("pnerf_nerf_train_wgrad", [U, P, P, P, P, P])
L.register("pnerf_nerf_composite_train_forward", [P, P, P, P, P, U, U, F_, P, P, P, P, P])
L.register("pnerf_nerf_composite_train_backward", [P, P, P, P, P, P, P, P, P, P, P, U, U, F_, P, P, P])
if "pnerf_grid_encode_backward_counted" not in L._SIGS:
    L.register("pnerf_grid_encode_backward_counted", [P, P, P, P, U, U, F_, U, U, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, F_, P])
for _n in ("pnerf_nerf_train_xbuf_bytes", "pnerf_nerf_train_ybuf_bytes"):
    getattr(L.lib, _n).argtypes = [U]
    getattr(L.lib, _n).restype = c_uint64
for _n in ("pnerf_nerf_train_dw_floats", "pnerf_nerf_train_wfwd_units", "pnerf_nerf_train_wbwd_units"):
    getattr(L.lib, _n).argtypes = []
    getattr(L.lib, _n).restype = c_uint32

WEIGHT_NAMES = ["sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight"]
# packed weight-gradient buffer: (block, N_pad, K_pad) in kernel order (nerf_train.cu)
DW_LAYERS = [("S0", 64, 32), ("S1", 16, 64), ("V0", 64, 32), ("V1", 64, 64), ("V2", 16, 64)]


def build_pack_index(shapes):
    """shapes: {name: torch.Size} -> (index int64 [fwd halfs + bwd halfs], n_fwd_halfs)"""
    base, off = {}, 0
    for nme in WEIGHT_NAMES:
        base[nme] = off
        off += int(np.prod(shapes[nme]))
    zero = off                                  # index of the trailing 0.0

    def idx(nme):
        return (base[nme] + torch.arange(int(np.prod(shapes[nme])))).reshape(tuple(shapes[nme]))

    def pad(n_pad, k_pad):
        return torch.full((n_pad, k_pad), zero, dtype=torch.int64)

    s0, s1 = idx("sigma_net.0.weight"), pad(16, 64)
    s1[:16] = idx("sigma_net.1.weight")
    c0 = idx("color_net.0.weight")
    v0 = pad(64, 32); v0[:, 0:16] = c0[:, 0:16]; v0[:, 17:32] = c0[:, 16:31]      # column 16 faces the sigma logit: zero
    v1 = idx("color_net.1.weight")
    v2 = pad(16, 64); v2[0:3] = idx("color_net.2.weight")
    fwd = [_frag_idx(s0), _frag_idx(s1), _frag_idx(v0), _frag_idx(v1), _frag_idx(v2[:8])]
    bwd = [_frag_idx(v2.t().contiguous()), _frag_idx(v1.t().contiguous()), _frag_idx(v0[:, 16:32].t().contiguous()),
           _frag_idx(s1.t().contiguous()), _frag_idx(s0.t().contiguous())]
    fwd, bwd = torch.cat(fwd), torch.cat(bwd)
    return torch.cat([fwd, bwd]), fwd.numel()


def dw_views(dwbuf):
    """packed fp32 weight-gradient buffer -> {parameter name: gradient in the parameter's own shape}"""
    v, off = {}, 0
    for nme, n, k in DW_LAYERS:
        v[nme] = dwbuf[off:off + n * k].view(n, k)
        off += n * k
    return {"sigma_net.0.weight": v["S0"], "sigma_net.1.weight": v["S1"],
            "color_net.0.weight": torch.cat([v["V0"][:, 0:16], v["V0"][:, 17:32]], dim=1), "color_net.1.weight": v["V1"],
            "color_net.2.weight": v["V2"][0:3]}


def supported(model):
    return fused_nerf.supported_color(model)


def _state(model):
    st = getattr(model, "_fused_nerf_train_state", None)
    dev = model.encoder.embeddings.device
    if st is None or st["device"] != dev:
        sd = dict(model.named_parameters())
        index, n_fwd = build_pack_index({n: sd[n].shape for n in WEIGHT_NAMES})
        st = dict(device=dev, index=index.to(dev), n_fwd=n_fwd, tables=_Tables(), weights=[sd[n] for n in WEIGHT_NAMES],
                  zero=torch.zeros(1, dtype=torch.float32, device=dev))
        assert n_fwd == 4 * L.lib.pnerf_nerf_train_wfwd_units()
        assert index.numel() - n_fwd == 4 * L.lib.pnerf_nerf_train_wbwd_units()
        object.__setattr__(model, "_fused_nerf_train_state", st)
    return st


class _NerfTrainField(Function):
    """(model, count | None, xyzs, dirs, embeddings, *weights) -> sigma [M] (density_scale applied), rgb [M,3]
    count: optional int32 device tensor = number of valid rows of xyzs / dirs (static-capacity mode)"""

    @staticmethod
    def forward(ctx, model, count, xyzs, dirs, emb, *weights):
        st = _state(model)
        L.require_cuda(xyzs, dirs, emb)
        dev = xyzs.device
        xyzs, dirs = xyzs.detach().contiguous().float(), dirs.detach().contiguous().float()
        M = xyzs.shape[0]
        flat = torch.cat([w.detach().reshape(-1).float() for w in weights] + [st["zero"]])
        blob = flat[st["index"]].to(torch.float16)
        table = st["tables"].get("sigma", emb, always=emb.requires_grad)
        enc = model.encoder
        offsets = enc.offsets
        f = NerfTrain()
        f.table, f.offsets = ptr(table), ptr(offsets)
        f.wfwd, f.wbwd = blob.data_ptr(), blob.data_ptr() + 2 * st["n_fwd"]
        f.m_dev = ptr(count)
        f.L, f.H = enc.num_levels, enc.base_resolution
        f.S = float(np.float32(np.log2(enc.per_level_scale)))
        f.bound, f.density_scale = float(model.bound), float(model.density_scale)
        static = count is not None
        new = (lambda name, *shape, dtype=torch.float32: ARENA.get(name, shape, dtype, dev).detach()) if static else \
              (lambda name, *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device=dev))
        xbuf = new("nerf_xbuf", int(L.lib.pnerf_nerf_train_xbuf_bytes(M)), dtype=torch.uint8)
        sigma, rgb = new("nerf_sigma", M), new("nerf_rgb", M, 3)
        L.call("pnerf_nerf_train_forward", ptr(xyzs), ptr(dirs), M, ctypes.addressof(f), ptr(xbuf), ptr(sigma), ptr(rgb), stream())
        # (aliases of the outputs: holding the returned objects themselves would tie ctx and its outputs into a cycle)
        ctx.keep = (f, blob, table, offsets, xbuf, sigma.detach(), rgb.detach(), xyzs, count)
        ctx.model, ctx.M = model, M
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        model, M = ctx.model, ctx.M
        f, blob, table, offsets, xbuf, sigma, rgb, xyzs, count = ctx.keep
        dev = xyzs.device
        g_sigma = torch.zeros(M, device=dev) if g_sigma is None else g_sigma.contiguous().float()
        g_rgb = torch.zeros(M, 3, device=dev) if g_rgb is None else g_rgb.contiguous().float()
        static = count is not None
        new = (lambda name, *shape, dtype=torch.float32: ARENA.get(name, shape, dtype, dev).detach()) if static else \
              (lambda name, *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype, device=dev))
        ybuf = new("nerf_ybuf", int(L.lib.pnerf_nerf_train_ybuf_bytes(M)), dtype=torch.uint8)
        d_enc = new("nerf_d_enc", M, 32)
        dw = torch.zeros(int(L.lib.pnerf_nerf_train_dw_floats()), dtype=torch.float32, device=dev)
        L.call("pnerf_nerf_train_backward", M, ctypes.addressof(f), ptr(xbuf), ptr(ybuf), ptr(g_sigma), ptr(g_rgb), ptr(sigma),
               ptr(rgb), ptr(d_enc), stream())
        enc = model.encoder
        g_tab = None
        # data parallel (distributed.GradBucket attached to the model): the density table's gradient — 99.9 % of the stage-1
        # bucket — is scattered STRAIGHT INTO the bucket (no pack copy) and its all-reduce starts on a side stream before the
        # weight-gradient kernel below runs; not when gradients are being accumulated into an existing .grad
        bucket = getattr(model, "_grad_bucket", None)
        if enc.embeddings.grad is not None or M == 0:
            bucket = None
        early = False
        if ctx.needs_input_grad[4]:
            g_tab = bucket.slot(enc.embeddings) if bucket is not None else None
            if g_tab is None:
                g_tab = L.zeros_like_fast(enc.embeddings, torch.float32)     # (a memset node instead of torch's fill kernel)
            else:
                L.zero_(g_tab)
                early = True
            if M > 0:
                S_ = float(np.float32(np.log2(enc.per_level_scale)))
                if count is None:
                    from .gridencoder.backend import _backend as GB
                    x01 = ((xyzs + model.bound) / (2 * model.bound)).contiguous()
                    GB.grid_encode_backward_blc(d_enc, x01, g_tab, offsets, g_tab, M, 3, 2, enc.num_levels, S_, enc.base_resolution,
                                                None, None, 0, False)
                else:   # number of rows taken from device memory
                    L.call("pnerf_grid_encode_backward_counted", ptr(d_enc), ptr(xyzs), ptr(offsets), ptr(g_tab), M, enc.num_levels,
                           S_, enc.base_resolution, 0, 0, L.F32, L.LAYOUT_BLC, ptr(count), float(model.bound), stream())
        if early:
            bucket.early([enc.embeddings])
        L.call("pnerf_nerf_train_wgrad", M, ptr(xbuf), ptr(ybuf), ptr(dw), ptr(count), stream())
        gw = dw_views(dw)
        return (None, None, None, None, g_tab, *[gw[n] for n in WEIGHT_NAMES])


def field(model, xyzs, dirs, count=None):
    """fused stage-1 training field -> (sigmas [M] with density_scale applied, rgbs [M,3]); rows >= count are not touched"""
    st = _state(model)
    return _NerfTrainField.apply(model, count, xyzs, dirs, model.encoder.embeddings, *st["weights"])


class _NerfComposite(Function):
    """(sigmas, rgbs, deltas, rays, gt | None) -> weights_sum [N], depth [N], image [N,3], err_map [N]"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, deltas, rays, gt, T_thresh, static):
        sigmas, rgbs, deltas = (t.contiguous().float() for t in (sigmas, rgbs, deltas))
        gt = None if gt is None else gt.detach().contiguous().float()
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        ws, depth, err = (torch.empty(N, device=dev) for _ in range(3))
        image = torch.empty(N, 3, device=dev)
        L.call("pnerf_nerf_composite_train_forward", ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), ptr(gt), M, N, float(T_thresh),
               ptr(ws), ptr(depth), ptr(image), ptr(err), stream())
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, image, err)
        ctx.gt, ctx.dims, ctx.static = gt, (M, N, float(T_thresh)), bool(static)
        ctx.mark_non_differentiable(depth)
        return ws, depth, image, err

    @staticmethod
    def backward(ctx, g_ws, g_depth, g_image, g_err):
        sigmas, rgbs, deltas, rays, ws, image, err = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        dev = sigmas.device
        g_image = torch.zeros(N, 3, device=dev) if g_image is None else g_image.contiguous().float()
        g_ws = None if g_ws is None else g_ws.contiguous().float()
        g_err = None if g_err is None else g_err.contiguous().float()
        if ctx.static:
            # every row a ray owns is written by the kernel; rows behind the valid count are never read by the field's backward
            g_sig = ARENA.get("nerf_g_sigmas", (M,), torch.float32, dev).detach()
            g_rgb = ARENA.get("nerf_g_rgbs", (M, 3), torch.float32, dev).detach()
        else:   # exact-count buffers are padded to the march's alignment: the padding rows belong to no ray
            g_sig, g_rgb = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
        L.call("pnerf_nerf_composite_train_backward", ptr(g_ws), ptr(g_image), ptr(g_err), ptr(sigmas), ptr(rgbs), ptr(deltas),
               ptr(rays), ptr(ctx.gt), ptr(ws), ptr(image), ptr(err), M, N, T_thresh, ptr(g_sig), ptr(g_rgb), stream())
        return g_sig, g_rgb, None, None, None, None, None


def composite(sigmas, rgbs, deltas, rays, gt=None, T_thresh=1e-4, static=False):
    """static: the sample buffers have a data-independent capacity and their consumers read only the rows rays own"""
    return _NerfComposite.apply(sigmas, rgbs, deltas, rays, gt, T_thresh, static)
