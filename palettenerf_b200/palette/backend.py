"""palette backend ops — drop-in for the reference's `_palette_func` module (palette/src/bindings.cpp:96-104) and the
autograd wrappers in palette/utils.py:256-295: rgb_to_hsv, hsv_to_rgb (CUDA) and compute_RGB_histogram (host)."""
import ctypes

import numpy as np
import torch
from torch.autograd import Function
from torch.amp import custom_fwd

from .. import _lib as L
from .._lib import ptr, stream, call, require_cuda


class _Backend:
    @staticmethod
    def rgb_to_hsv(n, input, output):
        require_cuda(input, output)
        call("pnerf_rgb_to_hsv", n, ptr(input), ptr(output), stream())

    @staticmethod
    def hsv_to_rgb(n, input, output):
        require_cuda(input, output)
        call("pnerf_hsv_to_rgb", n, ptr(input), ptr(output), stream())

    @staticmethod
    def compute_RGB_histogram(colors_rgb, weights, bits_per_channel):
        colors = np.ascontiguousarray(colors_rgb, dtype=np.float32).reshape(-1)
        weights = np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
        if colors.shape[0] != 3 * weights.shape[0]:
            raise RuntimeError("compute_RGB_histogram: colors must hold 3 floats per weight")
        nb = 1 << (3 * bits_per_channel)
        bw, bc = np.empty(nb, np.float64), np.empty((nb, 3), np.float32)
        L.check(L.lib.pnerf_compute_rgb_histogram(colors.ctypes.data_as(ctypes.c_void_p),
                                                  weights.ctypes.data_as(ctypes.c_void_p), weights.shape[0],
                                                  int(bits_per_channel), bw.ctypes.data_as(ctypes.c_void_p),
                                                  bc.ctypes.data_as(ctypes.c_void_p)), "compute_RGB_histogram")
        return bw, bc


_backend = _Backend()
compute_RGB_histogram = _backend.compute_RGB_histogram


class _rgb_to_hsv(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, input):
        if not input.is_cuda:
            input = input.cuda()
        prefix = input.shape[:-1]
        input = input.contiguous().view(-1, 3)
        n = input.shape[0]
        output = torch.empty(n, 3, device=input.device, dtype=input.dtype)
        _backend.rgb_to_hsv(n, input, output)
        return output.reshape(*prefix, 3)


class _hsv_to_rgb(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, input):
        if not input.is_cuda:
            input = input.cuda()
        prefix = input.shape[:-1]
        input = input.contiguous().view(-1, 3)
        n = input.shape[0]
        output = torch.empty(n, 3, device=input.device, dtype=input.dtype)
        _backend.hsv_to_rgb(n, input, output)
        return output.reshape(*prefix, 3)


rgb_to_hsv = _rgb_to_hsv.apply
hsv_to_rgb = _hsv_to_rgb.apply
