"""`_backend` for gridencoder: the reference's pybind11 surface (gridencoder/src/bindings.cpp, gridencoder.h:12-13)
with the same names / argument order / [L,B,C] layouts, bound to the C ABI. `grid_encode_forward_blc` /
`grid_encode_backward_blc` are the native-layout variants ([B, L*C] rows written/read directly, no permute copy)."""
import torch

from .. import _lib as L
from .._lib import ptr, stream, call, require_cuda, dtype_id


def _check(inputs, embeddings, offsets, *others):
    require_cuda(inputs, embeddings, offsets, *others)
    for name, t in (("inputs", inputs), ("embeddings", embeddings), ("offsets", offsets)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous tensor")
    if inputs.dtype != torch.float32:
        raise RuntimeError("inputs must be a float32 tensor")
    if offsets.dtype != torch.int32:
        raise RuntimeError("offsets must be an int tensor")


class _Backend:
    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L_, S, H, dy_dx, gridtype, align_corners,
                            _layout=L.LAYOUT_LBC):
        _check(inputs, embeddings, offsets, outputs, dy_dx)
        if outputs.dtype != embeddings.dtype or not outputs.is_contiguous():
            raise RuntimeError("outputs must be contiguous and of the embeddings' dtype")
        call("pnerf_grid_encode_forward", ptr(inputs), ptr(embeddings), ptr(offsets), ptr(outputs), B, D, C, L_, float(S),
             H, ptr(dy_dx), gridtype, int(bool(align_corners)), dtype_id(embeddings.dtype), _layout, stream())

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L_, S, H, dy_dx, grad_inputs,
                             gridtype, align_corners, _layout=L.LAYOUT_LBC):
        _check(inputs, embeddings, offsets, grad, grad_embeddings, dy_dx, grad_inputs)
        if grad.dtype != grad_embeddings.dtype or not grad.is_contiguous() or not grad_embeddings.is_contiguous():
            raise RuntimeError("grad / grad_embeddings must be contiguous and of one dtype")
        if grad.dtype == torch.float16 and D == 3 and C == 2 and grad_embeddings.numel() % 4 == 0 and dy_dx is None:
            # fp16 table: reduce into an fp32 workspace (red.global.add.v2.f32) and fold it into grad_embeddings once
            ws = torch.empty(grad_embeddings.numel(), dtype=torch.float32, device=grad.device)
            call("pnerf_grid_encode_backward_ws", ptr(grad), ptr(inputs), ptr(offsets), ptr(grad_embeddings), ptr(ws),
                 grad_embeddings.shape[0], B, L_, float(S), H, gridtype, int(bool(align_corners)), _layout, stream())
            return
        call("pnerf_grid_encode_backward", ptr(grad), ptr(inputs), ptr(embeddings), ptr(offsets), ptr(grad_embeddings), B,
             D, C, L_, float(S), H, ptr(dy_dx), ptr(grad_inputs), gridtype, int(bool(align_corners)),
             dtype_id(grad.dtype), _layout, stream())

    @staticmethod
    def grid_encode_forward_blc(*a):
        _Backend.grid_encode_forward(*a, _layout=L.LAYOUT_BLC)

    @staticmethod
    def grid_encode_backward_blc(*a):
        _Backend.grid_encode_backward(*a, _layout=L.LAYOUT_BLC)


_backend = _Backend()
