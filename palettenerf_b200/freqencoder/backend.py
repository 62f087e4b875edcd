"""`_backend` for freqencoder (reference surface: freqencoder/src/freqencoder.h:7-10)."""
import torch

from .._lib import ptr, stream, call, require_cuda


class _Backend:
    @staticmethod
    def freq_encode_forward(inputs, B, D, deg, C, outputs):
        require_cuda(inputs, outputs)
        if inputs.dtype != torch.float32 or outputs.dtype != torch.float32:
            raise RuntimeError("freq_encode_forward: float32 tensors required")
        call("pnerf_freq_encode_forward", ptr(inputs), B, D, deg, C, ptr(outputs), stream())

    @staticmethod
    def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
        require_cuda(grad, outputs, grad_inputs)
        call("pnerf_freq_encode_backward", ptr(grad), ptr(outputs), B, D, deg, C, ptr(grad_inputs), stream())


_backend = _Backend()
