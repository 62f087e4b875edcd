"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/pnerf_b200.h
declares; the Python packages expose the reference's operator names. No compute calls (no GPU here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "pnerf_b200.h")).read()
    return sorted(set(re.findall(r"PNERF_API\s+[\w\s\*]+?\b(pnerf_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from palettenerf_b200 import _lib
    names = _declared()
    assert len(names) >= 27
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/pnerf_b200.h but not exported"
    assert _lib.lib.pnerf_abi_version() >= 1
    assert _lib.lib.pnerf_build_arch() == b"sm_100a"


def test_status_strings():
    from palettenerf_b200 import _lib
    assert _lib.lib.pnerf_status_string(0) == b"ok"
    assert b"unsupported" in _lib.lib.pnerf_status_string(-2)


def test_null_pointer_is_rejected_without_touching_the_gpu():
    from palettenerf_b200 import _lib
    assert _lib.lib.pnerf_morton3D(None, 16, None, None) == -1
    assert _lib.lib.pnerf_packbits(None, 16, 0.5, None, None) == -1


def test_rgb_histogram_host_function_matches_oracle():
    import numpy as np
    import oracle
    from palettenerf_b200.palette.backend import compute_RGB_histogram
    r = np.random.RandomState(0)
    colors = r.rand(5000, 3).astype(np.float32)
    colors[:10] = 1.5  # clamped to 0.999
    colors[10:20] = -0.5
    w = r.rand(5000).astype(np.float32)
    for bpc in (1, 3, 5):
        bw, bc = compute_RGB_histogram(colors.reshape(-1), w, bpc)
        ow, oc = oracle.compute_rgb_histogram(colors, w, bpc)
        assert bw.dtype == np.float64 and bc.dtype == np.float32 and bc.shape == (1 << (3 * bpc), 3)
        assert np.array_equal(bw, ow) and np.array_equal(bc, oc)
        assert abs(bw.sum() - w.astype(np.float64).sum()) < 1e-6


def test_python_operator_surface():
    import palettenerf_b200.raymarching as rm
    for n in ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
              "composite_rays_train", "composite_rays_flex_train", "march_rays", "composite_rays", "composite_rays_flex",
              "spread_ray_to_sample"]:
        assert callable(getattr(rm, n))
    from palettenerf_b200.raymarching.backend import _backend as b
    for n in ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
              "composite_rays_train_forward", "composite_rays_train_backward", "composite_rays_flex_train_forward",
              "composite_rays_flex_train_backward", "march_rays", "composite_rays", "composite_rays_flex",
              "spread_ray_to_sample"]:
        assert callable(getattr(b, n))
    from palettenerf_b200.gridencoder import GridEncoder
    from palettenerf_b200.shencoder import SHEncoder
    from palettenerf_b200.freqencoder import FreqEncoder
    from palettenerf_b200.encoding import get_encoder
    enc, dim = get_encoder("hashgrid", desired_resolution=4096)
    assert dim == 32 and enc.embeddings.shape == (6328848, 2) and enc.offsets.shape == (17,)
    assert SHEncoder(degree=4).output_dim == 16 and FreqEncoder(3, 6).output_dim == 39


def test_product_fails_loudly_without_cuda_tensors():
    import pytest
    import torch
    from palettenerf_b200.raymarching.backend import _backend as b
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b.morton3D(torch.zeros(4, 3, dtype=torch.int32), 4, torch.zeros(4, dtype=torch.int32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "palettenerf_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
