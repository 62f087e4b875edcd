"""ctypes binding of libpnerf_b200.so (the C ABI declared in include/pnerf_b200.h).

There is NO CPU fallback: if the library is missing it is built once with nvcc (palettenerf_b200/build.py);
if that fails, importing raises. Every wrapper passes raw device pointers + the current CUDA stream and turns a
non-zero status into RuntimeError (the reference surfaces TORCH_CHECK failures the same way).
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_uint32, c_uint64, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("PNERF_LIB", os.path.join(_PKG, "libpnerf_b200.so"))   # PNERF_LIB: variant build (experiments)

F16, F32, F64 = 0, 1, 2
LAYOUT_LBC, LAYOUT_BLC = 0, 1
_DTYPE_ID = {torch.float16: F16, torch.float32: F32, torch.float64: F64}


def _load():
    if not os.path.exists(_LIB_PATH):
        from . import build as _build
        _build.build()
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing and could not be built; pnerf_b200 has no CPU fallback")
    return ctypes.CDLL(_LIB_PATH)


lib = _load()

P, U, F, I = c_void_p, c_uint32, c_float, c_int
_SIGS = {
    "pnerf_near_far_from_aabb": [P, P, P, U, F, P, P, P],
    "pnerf_sph_from_ray": [P, P, F, U, P, P],
    "pnerf_morton3D": [P, U, P, P],
    "pnerf_morton3D_invert": [P, U, P, P],
    "pnerf_packbits": [P, U, F, P, P],
    "pnerf_march_rays_train": [P, P, P, F, F, U, U, U, U, U, P, P, P, P, P, P, P, P, P],
    "pnerf_march_rays_train_ws": [P, P, P, F, F, U, U, U, U, U, P, P, P, P, P, P, P, P, P, P, P, P],
    "pnerf_occupied_bounds": [P, U, U, F, P, P],
    "pnerf_mark_untrained_grid": [P, U, F, F, F, F, U, U, F, F, I, P, P, P],
    "pnerf_composite_rays_train_forward": [P, P, P, P, U, U, F, P, P, P, P],
    "pnerf_composite_rays_train_backward": [P, P, P, P, P, P, P, P, U, U, F, P, P, P],
    "pnerf_composite_rays_flex_train_forward": [P, P, P, P, U, U, U, F, P, P],
    "pnerf_composite_rays_flex_train_backward": [P, P, P, P, P, P, U, U, U, F, P, P],
    "pnerf_spread_ray_to_sample": [P, P, U, U, U, P, P],
    "pnerf_march_rays": [U, U, P, P, P, P, F, F, U, U, U, P, P, P, P, P, P, P, P],
    "pnerf_composite_rays": [U, U, F, P, P, P, P, P, P, P, P, P],
    "pnerf_composite_rays_flex": [U, U, U, F, P, P, P, P, P, P, P, P],
    "pnerf_grid_encode_forward": [P, P, P, P, U, U, U, U, F, U, P, U, I, I, I, P],
    "pnerf_grid_encode_backward": [P, P, P, P, P, U, U, U, U, F, U, P, P, U, I, I, I, P],
    "pnerf_grid_encode_backward_ws": [P, P, P, P, P, c_uint64, U, U, F, U, U, I, I, P],
    "pnerf_get_rays": [P, F, F, F, F, U, U, P, c_uint64, U, U, P, P, P, F, P, P, P],
    "pnerf_get_rays_collate": [P, F, F, F, F, U, U, P, c_uint64, U, U, P, P, P, F, P, P, P, U, P, P, U, P, P],
    "pnerf_peer_allreduce": [P, U, U, c_uint64, F, P],
    "pnerf_peer_allreduce_mc": [c_uint64, U, U, c_uint64, F, P],
    "pnerf_render_tail_forward": [U, P, P, P, P, P, P, U, P, U, P, P, P, P],
    "pnerf_render_tail_backward": [U, P, P, P, U, U, U, P, P, P],
    "pnerf_sh_encode_forward": [P, P, U, U, U, P, P],
    "pnerf_sh_encode_backward": [P, P, U, U, U, P, P, P],
    "pnerf_freq_encode_forward": [P, U, U, U, U, P, P],
    "pnerf_freq_encode_backward": [P, P, U, U, U, U, P, P],
    "pnerf_rgb_to_hsv": [U, P, P, P],
    "pnerf_hsv_to_rgb": [U, P, P, P],
    "pnerf_compute_rgb_histogram": [P, P, c_uint64, I, P, P],
}
EXPORTS = sorted(list(_SIGS) + ["pnerf_status_string", "pnerf_last_cuda_error", "pnerf_abi_version", "pnerf_build_arch",
                  "pnerf_palette_field_forward", "pnerf_palette_render_fused"])

for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = c_int
lib.pnerf_status_string.argtypes = [c_int]
lib.pnerf_status_string.restype = c_char_p
lib.pnerf_last_cuda_error.restype = c_char_p
lib.pnerf_abi_version.restype = c_int
lib.pnerf_build_arch.restype = c_char_p
lib.pnerf_zero_fill.argtypes = [c_void_p, ctypes.c_uint64, c_void_p]
lib.pnerf_zero_fill.restype = c_int


def zeros_like_fast(t, dtype=None):
    """torch.zeros_like(t, dtype) for large CUDA buffers: torch.empty + pnerf_zero_fill (a memset node at the HBM write rate
    instead of torch's elementwise fill kernel: 21 -> 8 us for a 50 MB table gradient)"""
    out = torch.empty_like(t, dtype=dtype or t.dtype, memory_format=torch.contiguous_format)
    zero_(out)
    return out


def zero_(t):
    """t.zero_() for a contiguous CUDA tensor through pnerf_zero_fill"""
    if not t.is_cuda or not t.is_contiguous():
        return t.zero_()
    check(lib.pnerf_zero_fill(t.data_ptr(), t.numel() * t.element_size(), stream()), "pnerf_zero_fill")
    torch.autograd.graph.increment_version(t)
    return t


def register(name, argtypes):
    """declare a further entry point (used by the fused-kernel modules)"""
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = c_int
    _SIGS[name] = argtypes
    return fn


def ptr(t):
    """device (or host) address of a tensor / None -> NULL"""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(status, what):
    if status != 0:
        msg = lib.pnerf_status_string(status).decode()
        if status == -3:
            msg += ": " + lib.pnerf_last_cuda_error().decode()
        raise RuntimeError(f"pnerf_b200.{what} failed: {msg}")


# kernels launched per C-ABI call (for bench.py's gpu_launches claim and per-kernel CUDA-event timing)
LAUNCHES = {"pnerf_march_rays_train": 3, "pnerf_march_rays_train_ws": 3, "pnerf_grid_encode_backward_ws": 2,
            "pnerf_occupied_bounds": 2}
launch_count = 0
_profile = None  # when set: dict name -> [list of (start_event, end_event), units]


def profile_start():
    """record a CUDA-event pair around every C-ABI call (on torch's current stream, where the kernels run)"""
    global _profile
    _profile = {}


def profile_stop():
    """-> {name: (total_ms, n_calls)}; synchronises"""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in (prof or {}).items()}


def call(name, *args):
    global launch_count
    launch_count += LAUNCHES.get(name, 1)
    if _profile is None:
        check(getattr(lib, name)(*args), name)
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(getattr(lib, name)(*args), name)
    b.record()
    _profile.setdefault(name, []).append((a, b))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pnerf_b200 kernels need CUDA tensors (there is no CPU fallback)")


def dtype_id(dt):
    try:
        return _DTYPE_ID[dt]
    except KeyError:
        raise RuntimeError(f"pnerf_b200: unsupported dtype {dt}")
