"""optim — `FusedAdam`: torch.optim.Adam's update for every tensor of a parameter group in one launch of
`pnerf_adam_step` (csrc/optim.cu), for the optimizer the reference's trainers build
(`optim.Adam(model.get_params(lr), betas=(0.9, 0.99), eps=1e-15)`, stepped through `GradScaler`; ref palette/utils.py:719-724).

Drop-in for `torch.optim.Adam(..., fused=True, capturable=True)` on fp32 CUDA parameters:
  * same hyper-parameters, same state keys (`step` device scalar, `exp_avg`, `exp_avg_sq`), so `state_dict()` /
    `load_state_dict()` interchange with torch's Adam and with the reference's checkpoints (nerf/utils.py:1153-1160);
  * `_step_supports_amp_scaling`: `GradScaler.step(opt)` hands it `grad_scale` / `found_inf`; the gradients are
    unscaled on the fly and a step with non-finite gradients is skipped on the device — no host synchronisation, CUDA-graph
    capturable; `lr` may be a device tensor (learning-rate schedules inside a captured graph);
  * no CPU path: parameters must be CUDA tensors.
Not supported (raise): amsgrad, maximize, sparse gradients, non-fp32 parameters.
"""
import ctypes
from ctypes import c_float, c_uint32, c_uint64, c_void_p

import torch

from . import _lib as L
from ._lib import ptr, stream


class _AdamTensor(ctypes.Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("step", c_void_p), ("n", c_uint64),
                ("mirror", c_void_p), ("mirror_stride", c_uint64)]


MAX_TENSORS = 32
L.register("pnerf_adam_step", [c_void_p, c_uint32, c_float, c_void_p, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                               c_void_p])
L.LAUNCHES["pnerf_adam_step"] = 2
L.register("pnerf_found_inf", [c_void_p, c_uint32, c_void_p, c_void_p])


class FusedAdam(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise RuntimeError("FusedAdam: amsgrad / maximize are not implemented")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False))

    def _state_of(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        elif not torch.is_tensor(st["step"]) or not st["step"].is_cuda:   # a state dict written by a non-capturable Adam
            st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=p.device)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        if grad_scale is not None:
            grad_scale = grad_scale.to(torch.float32).reshape(1)
        if found_inf is not None:
            found_inf = found_inf.to(torch.float32).reshape(1)
        # parameter groups with identical hyper-parameters share launches (the reference builds ~10 groups that differ
        # in nothing: palette/network.py:283-308)
        merged = {}
        for group in self.param_groups:
            lr = group["lr"]
            key = (id(lr) if torch.is_tensor(lr) else float(lr), tuple(group["betas"]), float(group["eps"]),
                   float(group["weight_decay"]))
            todo = merged.setdefault(key, (group, []))[1]
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam needs CUDA parameters (there is no CPU fallback)")
                if g.is_sparse or p.dtype != torch.float32 or g.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: dense fp32 parameters and gradients only")
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters must be contiguous")
                st = self._state_of(p)
                todo.append((p, g.contiguous(), st["exp_avg"], st["exp_avg_sq"], st["step"]))
        for group, todo in merged.values():
            lr = group["lr"]
            lr_dev = lr.to(torch.float32).reshape(1) if torch.is_tensor(lr) else None
            if lr_dev is not None and not lr_dev.is_cuda:
                lr, lr_dev = float(lr), None
            b1, b2 = group["betas"]
            for i in range(0, len(todo), MAX_TENSORS):
                chunk = todo[i:i + MAX_TENSORS]
                arr = (_AdamTensor * len(chunk))()
                mirrored = []
                for k, (p, g, m, v, s) in enumerate(chunk):
                    arr[k].p, arr[k].g, arr[k].m, arr[k].v, arr[k].step, arr[k].n = ptr(p), ptr(g), ptr(m), ptr(v), ptr(s), p.numel()
                    # fp16 mirror registered on the parameter (fused_train: the interleaved table the forward kernel reads)
                    mir = getattr(p, "_pnerf_half_mirror", None)
                    if mir is not None and mir[0].device == p.device and p.numel() % 2 == 0:
                        arr[k].mirror, arr[k].mirror_stride = mir[0].data_ptr() + mir[1], mir[2]
                        mirrored.append(p)
                L.call("pnerf_adam_step", ctypes.addressof(arr), len(chunk), 0.0 if lr_dev is not None else float(lr),
                       ptr(lr_dev), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), ptr(grad_scale),
                       ptr(found_inf), stream())
                # the kernel writes through raw pointers: tell autograd / version-keyed caches (fused.FieldCache, the fp16
                # table copies) that the tensors changed, as an in-place torch op would
                torch.autograd.graph.increment_version([t for tup in chunk for t in (tup[0], tup[2], tup[3], tup[4])])
                for p in mirrored:        # the mirror holds fp16(p) for exactly this version of p
                    p._pnerf_mirror_version = p._version
        return loss


class GradScaler(torch.amp.GradScaler):
    """torch.amp.GradScaler whose non-finite check in front of a `FusedAdam` step is ONE streaming pass of `pnerf_found_inf`
    over the gradients instead of torch's multi-tensor check-and-unscale kernel with a unit scale (22 us -> ~10 us for the
    50 MB of a hash table's gradient). Everything else — scale growth / backoff, `unscale_()`, other optimizers — is torch's."""

    def __init__(self, device="cuda", **kw):
        super().__init__(device, **kw)

    def _check_inf_per_device(self, optimizer):
        if not isinstance(optimizer, FusedAdam):
            return super()._check_inf_per_device(optimizer)
        _scale, _ = self._check_scale_growth_tracker("_check_inf_per_device")
        found_inf = torch.full((), 0.0, dtype=torch.float32, device=_scale.device)
        grads = [p.grad for group in optimizer.param_groups for p in group["params"] if p.grad is not None]
        if any(g.is_sparse or g.dtype != torch.float32 or not g.is_cuda or g.device != _scale.device for g in grads):
            return super()._check_inf_per_device(optimizer)
        grads = [g.contiguous() for g in grads]
        for i in range(0, len(grads), MAX_TENSORS):
            chunk = grads[i:i + MAX_TENSORS]
            arr = (_AdamTensor * len(chunk))()
            for k, g in enumerate(chunk):
                arr[k].g, arr[k].n = ptr(g), g.numel()
            L.call("pnerf_found_inf", ctypes.addressof(arr), len(chunk), ptr(found_inf), stream())
        state = self._per_optimizer_states[id(optimizer)]
        state["found_inf_per_device"] = {_scale.device: found_inf}
        return state["found_inf_per_device"]
