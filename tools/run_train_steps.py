"""Runs a few eager fused palette training steps (BASELINE config 4) — the command ncu wraps for the training kernels."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from palettenerf_b200 import synthetic as S  # noqa: E402
from palettenerf_b200.graphs import make_palette_train_step  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = S.build_palette_model(dev, seed=0, pred_clip="--clip" in sys.argv)
model.train()
opt = torch.optim.Adam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True)
scaler = torch.amp.GradScaler("cuda")
o, d = S.training_rays(4096, seed=0)
o, d = o.to(dev)[None].contiguous(), d.to(dev)[None].contiguous()
gt = torch.rand(1, 4096, 3, device=dev)
step = make_palette_train_step(model, opt, scaler, o, d, lambda out: ((out["image"] - gt) ** 2).mean()
                               + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean()
                               + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean())
for i in range(n):
    if i == n - 1:
        torch.cuda.nvtx.range_push("laststep")   # ncu --nvtx --nvtx-include "laststep/": the launches of ONE warm step
    step()
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("samples:", int(model.step_counter[(model.local_step - 1) % 16, 0].item()), "schedule:", model._last_train_schedule)
