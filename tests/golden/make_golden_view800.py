#!/usr/bin/env python
"""Golden vectors for ONE full 800 x 800 view (BASELINE config 3's size) from the reference's own PaletteRenderer.run_cuda on
the reference's own CUDA kernels, tables U(-0.5, 0.5) (`noclip` model of palette_cases.py), fp32 and autocast, density
scales 1 and 40 (rays that run their full length / rays that terminate). TEST INFRASTRUCTURE; runs on the GPU box:
    gpurun -- 'python tests/golden/make_golden_view800.py'  ->  gpurun_out/ref_view800.npz  (committed as
                                                               tests/golden/ref_view800.npz)
Stored per (density scale, precision, output key): the rows of every STRIDE-th ray (4 886 of 640 000) and the fp64 mean over
all rays — the first pins individual rays, the second catches anything systematic between them."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import palette_cases as PC  # noqa: E402
from make_golden_palette import autocast, ref_palette_model  # noqa: E402
from oracle import ref_python  # noqa: E402

SIDE, STRIDE = 800, 131


def view_rays():
    from palettenerf_b200 import synthetic as S
    return S.camera_rays(SIDE, SIDE, azimuth_deg=35.0)


def main():
    dev = torch.device("cuda:0")
    R = ref_python.load()
    ours = PC.build_model("noclip", "cpu")
    m = ref_palette_model(R, ours, False)
    m.eval()
    o, d = view_rays()
    o, d = o.to(dev)[None], d.to(dev)[None]
    out = {"stride": np.int64(STRIDE), "side": np.int64(SIDE)}
    for ds in PC.DENSITY_SCALES:
        m.density_scale = ds
        for prec in ("fp32", "f16"):
            with torch.no_grad(), autocast(prec):
                res = m.render(o, d, staged=True, max_ray_batch=4096, bg_color=1, perturb=False, gui_mode=False, **PC.RENDER_KW)
            for k, t in res.items():
                t = t.detach().float().reshape(SIDE * SIDE, -1)
                out[f"ds{int(ds)}_{prec}_{k}_rows"] = t[::STRIDE].cpu().numpy()
                out[f"ds{int(ds)}_{prec}_{k}_mean"] = t.double().mean(dim=0).cpu().numpy()
    dst = os.environ.get("PNERF_GOLDEN_OUT", os.path.join(ROOT, "gpurun_out", "ref_view800.npz"))
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print(f"[make_golden_view800] {len(out)} arrays, {os.path.getsize(dst) / 1e6:.2f} MB -> {dst}")


if __name__ == "__main__":
    main()
