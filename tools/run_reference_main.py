#!/usr/bin/env python
"""Run the reference's UNMODIFIED main_nerf.py / main_palette.py on this repository's drop-in packages (verdict row b2).

    python tools/run_reference_main.py main_nerf.py <dataset> --workspace <ws> -O --iters 40 ...
    python tools/run_reference_main.py main_palette.py <dataset> <nerf workspace> -O --iters 40 ...

What is swapped (sys.modules entries installed before the main is executed; nothing in the reference tree is edited):
    raymarching, gridencoder, shencoder, freqencoder   -> palettenerf_b200.{raymarching,gridencoder,shencoder,freqencoder}
    nerf.renderer, palette.renderer                     -> palettenerf_b200.{nerf,palette}.renderer   (NeRFRenderer / PaletteRenderer)
    _palette_func                                       -> palettenerf_b200.palette.backend           (rgb<->hsv, histogram)
Everything else — the mains, Trainer / PaletteTrainer, NeRFDataset, NeRFNetwork / PaletteNetwork, encoding.get_encoder — is the
reference's own code (from /root/reference, or oracle/_ref/py on the GPU box); third-party packages the image lacks resolve to
compat/. The reference's CUDA extensions are not loaded.
"""
import collections
import json
import os
import runpy
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install():
    ref = os.environ.get("PNERF_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "palette")):
        ref = os.path.join(ROOT, "oracle", "_ref", "py")
    if not os.path.isdir(os.path.join(ref, "palette")):
        raise SystemExit("reference Python sources not found (neither /root/reference nor oracle/_ref/py)")
    for p in (ref, os.path.join(ROOT, "compat"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import palettenerf_b200.raymarching as rm
    import palettenerf_b200.gridencoder as ge
    import palettenerf_b200.shencoder as sh
    import palettenerf_b200.freqencoder as fe
    import palettenerf_b200.nerf.renderer as nr
    import palettenerf_b200.palette.renderer as pr
    from palettenerf_b200.palette import backend as pb
    sys.modules.update({"raymarching": rm, "gridencoder": ge, "shencoder": sh, "freqencoder": fe, "nerf.renderer": nr,
                        "palette.renderer": pr})
    shim = types.ModuleType("_palette_func")
    for k in ("rgb_to_hsv", "hsv_to_rgb", "compute_RGB_histogram"):
        shim.__dict__[k] = getattr(pb._backend, k)
    sys.modules["_palette_func"] = shim
    stub = types.ModuleType("palette.rgbsg.fastLayerDecomposition.GteDistPointTriangle")
    sys.modules.setdefault("palette.rgbsg.fastLayerDecomposition.GteDistPointTriangle", stub)
    _tally_schedules(nr.NeRFRenderer, pr.PaletteRenderer)
    return ref


SCHEDULES = collections.Counter()


def _tally_schedules(nerf_cls, palette_cls):
    """Count which schedule every renderer call took (the renderers record it on the instance): the log line at the end is
    what tests/test_reference_mains_gpu.py asserts on."""
    def wrap(cls, name, attrs):
        inner = getattr(cls, name)

        def call(self, *a, **k):
            for attr in attrs:
                self.__dict__.pop(attr, None)
            out = inner(self, *a, **k)
            for attr in attrs:
                if attr in self.__dict__:
                    mode = "train" if self.training else "eval"
                    SCHEDULES[f"{cls.__name__}.{name}[{mode}]:{attr.replace('_last_', '')}={self.__dict__[attr]}"] += 1
            return out
        setattr(cls, name, call)
    wrap(nerf_cls, "update_extra_state", ("_last_update_schedule",))
    wrap(nerf_cls, "run_cuda", ("_last_schedule", "_last_train_schedule"))
    wrap(palette_cls, "run_cuda", ("_last_schedule", "_last_train_schedule"))


def main():
    ref = install()
    script = sys.argv[1]
    sys.argv = [os.path.join(ref, script)] + sys.argv[2:]
    runpy.run_path(sys.argv[0], run_name="__main__")
    import palettenerf_b200._lib as L
    print(f"[run_reference_main] {script} finished on the drop-in packages; C-ABI kernel launches: {L.launch_count}")
    print("[run_reference_main] schedules " + json.dumps(dict(sorted(SCHEDULES.items()))))


if __name__ == "__main__":
    main()
