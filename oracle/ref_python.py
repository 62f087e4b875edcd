"""Import the UNMODIFIED reference Python stack (palette.network / palette.renderer / palette.utils ...) on top of the
reference's own CUDA kernels (oracle/_ref/_ref_*.so). TEST INFRASTRUCTURE: used by tests/golden/make_golden_palette.py
and tests/test_ref_python_gpu.py as the checker; never imported by the product.

Where the modules come from: /root/reference when it exists (build container), else ``oracle/_ref/py`` (staged by
oracle/stage_ref_py.py, travels to the GPU box). Third-party imports the image lacks resolve to ``compat/``.
The reference wrappers do ``import _raymarching as _backend`` (raymarching/raymarching.py:9-12 and the like); the
compiled reference extensions are registered under exactly those names, so the reference's JIT build is never run.
"""
import importlib
import importlib.util
import os
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_EXT = {"_raymarching": "raymarching", "_gridencoder": "gridencoder", "_shencoder": "shencoder",
        "_freqencoder": "freqencoder", "_palette_func": "palette_func"}
_loaded = None


def source_root():
    ref = os.environ.get("PNERF_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(ref, "palette")):
        return ref
    staged = os.path.join(HERE, "_ref", "py")
    return staged if os.path.isdir(os.path.join(staged, "palette")) else None


def available():
    if source_root() is None:
        return False
    return all(os.path.exists(os.path.join(HERE, "_ref", f"_ref_{n}.so")) for n in _EXT.values())


def load():
    """-> namespace with .network (palette.network), .renderer, .utils, .nerf_network, .nerf_renderer, .raymarching"""
    global _loaded
    if _loaded is not None:
        return _loaded
    src = source_root()
    if src is None:
        raise RuntimeError("reference Python sources not available (neither /root/reference nor oracle/_ref/py)")
    import torch  # noqa: F401  (the extensions link against libtorch)
    for alias, name in _EXT.items():
        if alias in sys.modules:
            continue
        path = os.path.join(HERE, "_ref", f"_ref_{name}.so")
        spec = importlib.util.spec_from_file_location(f"_ref_{name}", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[alias] = mod
    # the palette-extraction helper compiles a .pyx at import time (pyximport); it is not on the path -> empty module
    stub = types.ModuleType("palette.rgbsg.fastLayerDecomposition.GteDistPointTriangle")
    sys.modules.setdefault("palette.rgbsg.fastLayerDecomposition.GteDistPointTriangle", stub)
    for p in (os.path.join(ROOT, "compat"), src):
        if p not in sys.path:
            sys.path.insert(0, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns = types.SimpleNamespace(
            root=src,
            raymarching=importlib.import_module("raymarching"),
            network=importlib.import_module("palette.network"),
            renderer=importlib.import_module("palette.renderer"),
            utils=importlib.import_module("palette.utils"),
            nerf_network=importlib.import_module("nerf.network"),
            nerf_renderer=importlib.import_module("nerf.renderer"),
        )
    _loaded = ns
    return ns
