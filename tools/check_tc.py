"""GPU check + A/B timing of the tcgen05 field kernel against the mma.sync one (same inputs, same weights).
usage (GPU box): python tools/check_tc.py [clip]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from palettenerf_b200 import fused, synthetic as S  # noqa: E402
import palettenerf_b200.raymarching as rm  # noqa: E402

dev = torch.device("cuda:0")
clip = len(sys.argv) > 1 and sys.argv[1] == "clip"
m = S.build_palette_model(dev, seed=1, pred_clip=clip, table_scale=0.5)
m.eval()
side = int(os.environ.get("SIDE", "800"))
o, d = S.camera_rays(side, side)
o, d = o.to(dev), d.to(dev)
nears, fars = rm.near_far_from_aabb(o, d, m.aabb_infer, m.min_near)
counter = torch.zeros(2, dtype=torch.int32, device=dev)
xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter, -1,
                                               False, -1, True, 0.0, 1024)
M = xyzs.shape[0]
print("samples", M)
names = ["sigma", "clip", "omega", "offsets_radiance", "view_dep", "diffuse"]
a = fused.field_forward(m, xyzs, dirs, kernel="mma")
torch.cuda.synchronize()
b = fused.field_forward(m, xyzs, dirs, kernel="tc")
torch.cuda.synchronize()
sub = slice(0, 200000)
with torch.no_grad():
    ref = m(xyzs[sub], dirs[sub])
for n, x, y, r in zip(names, a, b, ref):
    r = r.float().reshape(x[sub].shape)
    print(f"{n:18s} tc-vs-mma {(x - y).abs().max().item():.3e}   mma-vs-fp32 {(x[sub] - r).abs().max().item():.3e}   "
          f"tc-vs-fp32 {(y[sub] - r).abs().max().item():.3e}   finite {bool(torch.isfinite(y).all())}")
for kern in ("mma", "tc", "mma", "tc"):
    for _ in range(2):
        fused.field_forward(m, xyzs, dirs, kernel=kern)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(5):
        fused.field_forward(m, xyzs, dirs, kernel=kern)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 5
    print(f"{kern}: {ms:.3f} ms per call, {M / ms / 1e6:.2f} G samples/s")
