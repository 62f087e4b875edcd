"""`_backend` for shencoder (reference surface: shencoder/src/shencoder.h:9-10)."""
import torch

from .._lib import ptr, stream, call, require_cuda


def _f32c(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_contiguous():
            raise RuntimeError("tensor must be contiguous")
        if t.dtype != torch.float32:
            raise RuntimeError("tensor must be float32")


class _Backend:
    @staticmethod
    def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
        require_cuda(inputs, outputs, dy_dx)
        _f32c(inputs, outputs, dy_dx)
        call("pnerf_sh_encode_forward", ptr(inputs), ptr(outputs), B, D, C, ptr(dy_dx), stream())

    @staticmethod
    def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
        require_cuda(grad, inputs, dy_dx, grad_inputs)
        _f32c(grad, inputs, dy_dx, grad_inputs)
        call("pnerf_sh_encode_backward", ptr(grad), ptr(inputs), B, D, C, ptr(dy_dx), ptr(grad_inputs), stream())


_backend = _Backend()
