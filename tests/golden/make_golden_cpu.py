#!/usr/bin/env python
"""Golden vectors from the reference's own PYTHON, generated in the build container (where /root/reference exists).

TEST INFRASTRUCTURE. Run:  python tests/golden/make_golden_cpu.py   ->  tests/golden/ref_python.npz

The reference ships no stored known-answer vectors for the hot path, and its CUDA kernels cannot run in
the GPU-less build container. What CAN run here is imported from where it lies under /root/reference (nothing is copied
into this repository) and its outputs are committed as fixtures:

  freq        encoding.FreqEncoder                  (encoding.py:5-43, the pure-torch frequency encoder)
  sh          testing/test_shencoder.py::SHEncoder_torch (degree 1..5, :8-89) — the class statement is exec'd from the
              reference file (the module itself runs CUDA code at import time)
  offsets     gridencoder.GridEncoder.__init__      (gridencoder/grid.py:91-129, the level-offset table)
  trunc_exp   activation.trunc_exp                  (activation.py:4-17, forward + backward)
  get_rays    nerf/utils.py::get_rays               (:52-151) with custom_meshgrid (:34-40) — the two function statements
              are exec'd from the reference file (the module imports tensorboardX / trimesh / mcubes / lpips, absent
              here); every sampling mode: full image, random pixels, patches, random_size pairs, error-map sampling
  histogram   _palette_func.compute_RGB_histogram   (palette/src/bindings.cpp:52-91, host code) through the reference
              extension compiled by oracle/build_ref.py into oracle/_ref/

The CUDA-kernel outputs of the reference are pinned by tests/golden/make_golden_gpu.py on the GPU box.
"""
import ast
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PNERF_REFERENCE_ROOT", "/root/reference")


def _load_ref_ext(name, alias):
    """oracle/_ref/_ref_<name>.so -> sys.modules[alias] so the reference's `import _<name> as _backend` resolves"""
    path = os.path.join(ROOT, "oracle", "_ref", f"_ref_{name}.so")
    spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[alias] = mod
    return mod


def _class_from_source(path, cls):
    """exec one top-level class statement of a reference file (in place, nothing copied)"""
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
    ns = {"torch": torch, "nn": torch.nn, "np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[cls]


def _functions_from_source(path, names, ns):
    """exec top-level function statements of a reference file (in place, nothing copied)"""
    tree = ast.parse(open(path).read())
    nodes = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    exec(compile(ast.Module(body=nodes, type_ignores=[]), path, "exec"), ns)
    return [ns[n] for n in names]


def _golden_get_rays(out):
    """the reference's get_rays run on the CPU with seeded torch RNG; inputs and outputs become fixtures"""
    from packaging import version as pver
    ns = {"torch": torch, "pver": pver, "np": np}
    (get_rays,) = _functions_from_source(os.path.join(REF, "nerf", "utils.py"), ["get_rays"],
                                         dict(ns, **{"custom_meshgrid": _functions_from_source(
                                             os.path.join(REF, "nerf", "utils.py"), ["custom_meshgrid"], ns)[0]}))
    g = torch.Generator().manual_seed(7)
    H, W = 37, 53
    # two cameras: rotation from a QR factorisation (a proper orthonormal frame), translation O(3)
    poses = torch.eye(4).repeat(2, 1, 1)
    for b in range(2):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        poses[b, :3, :3] = q
        poses[b, :3, 3] = 3.0 * torch.randn(3, generator=g)
    intr = np.array([61.7, 59.3, W / 2 + 0.25, H / 2 - 0.5], np.float32)
    out["rays_poses"], out["rays_intrinsics"], out["rays_HW"] = poses.numpy(), intr, np.array([H, W])
    err = torch.rand(2, 128 * 128, generator=g) + 1e-3
    cases = {"full": dict(N=-1), "rand": dict(N=64), "patch": dict(N=64, patch_size=4), "pair": dict(N=64, random_size=3),
             "err": dict(N=64, error_map=err)}
    for name, kw in cases.items():
        torch.manual_seed(11)
        r = get_rays(poses, intr, H, W, **kw)
        out[f"rays_{name}_inds"] = r["inds"].contiguous().numpy()
        out[f"rays_{name}_o"] = r["rays_o"].contiguous().numpy()
        out[f"rays_{name}_d"] = r["rays_d"].contiguous().numpy()
        if "inds_coarse" in r:
            out[f"rays_{name}_inds_coarse"] = r["inds_coarse"].numpy()
    out["rays_error_map"] = err.numpy()


def main():
    sys.path.insert(0, REF)
    out = {}
    rng = np.random.default_rng(20260101)

    # ---- frequency encoder (pure torch) ----
    import encoding as ref_encoding
    x = rng.uniform(-1, 1, size=(48, 3)).astype(np.float32)
    enc = ref_encoding.FreqEncoder(input_dim=3, max_freq_log2=5, N_freqs=6, log_sampling=True)
    out["freq_in"] = x
    out["freq_out_deg6"] = enc(torch.from_numpy(x)).numpy()

    # ---- spherical harmonics, degree 1..5 ----
    SHT = _class_from_source(os.path.join(REF, "testing", "test_shencoder.py"), "SHEncoder_torch")
    d = rng.normal(size=(64, 3))
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    out["sh_in"] = d
    for deg in range(1, 6):
        out[f"sh_out_deg{deg}"] = SHT(degree=deg)(torch.from_numpy(d).double()).numpy()

    # ---- hash-grid level offsets ----
    _load_ref_ext("gridencoder", "_gridencoder")
    from gridencoder.grid import GridEncoder as RefGrid
    cfgs = [dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=4096),
            dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19, desired_resolution=2048),
            dict(input_dim=2, num_levels=8, level_dim=4, base_resolution=8, log2_hashmap_size=14, per_level_scale=1.7,
                 gridtype="tiled"),
            dict(input_dim=3, num_levels=6, level_dim=2, base_resolution=4, log2_hashmap_size=10, per_level_scale=1.5,
                 align_corners=True)]
    for i, c in enumerate(cfgs):
        # the full default table is 12.7 M entries; only its offsets and scale are fixtures
        g = RefGrid(**c)
        out[f"offsets_{i}"] = g.offsets.numpy().astype(np.int32)
        out[f"offsets_{i}_scale"] = np.float64(g.per_level_scale)
        out[f"offsets_{i}_cfg"] = np.array(repr(sorted(c.items())))

    # ---- trunc_exp ----
    import activation as ref_act
    v = torch.linspace(-20, 20, 81, dtype=torch.float32, requires_grad=True)
    y = ref_act.trunc_exp(v)
    y.backward(torch.ones_like(y))
    out["trunc_exp_in"] = v.detach().numpy()
    out["trunc_exp_out"] = y.detach().numpy()
    out["trunc_exp_grad"] = v.grad.numpy()

    # ---- RGB histogram (host function of the palette extension) ----
    pf = _load_ref_ext("palette_func", "_palette_func")
    colors = rng.uniform(0, 1, size=(5000, 3)).astype(np.float32)
    colors[:16] = np.array([0, 1, 1], np.float32)   # exact edges: the 1.0 channel falls in the clamped top bin
    weights = rng.uniform(0, 2, size=(5000,)).astype(np.float32)
    out["hist_colors"], out["hist_weights"] = colors, weights
    for bpc in (3, 5):
        bw, bc = pf.compute_RGB_histogram(colors.reshape(-1).copy(), weights, bpc)   # numpy in / numpy out
        out[f"hist_bin_weights_b{bpc}"] = np.asarray(bw)
        out[f"hist_bin_centers_b{bpc}"] = np.asarray(bc)

    # ---- ray generation ----
    _golden_get_rays(out)

    path = os.path.join(HERE, "ref_python.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
