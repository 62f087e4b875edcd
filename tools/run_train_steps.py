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
from palettenerf_b200.optim import FusedAdam  # noqa: E402
from palettenerf_b200.palette.losses import palette_loss  # noqa: E402
opt = FusedAdam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
scaler = torch.amp.GradScaler("cuda")
o, d = S.training_rays(4096, seed=0)
o, d = o.to(dev)[None].contiguous(), d.to(dev)[None].contiguous()
gt = torch.rand(1, 4096, 3, device=dev)
step = make_palette_train_step(model, opt, scaler, o, d,
                               lambda out: palette_loss(out, gt, lambda_sparsity=2e-4, lambda_offsets=0.03, lambda_view_dep=0.1)[0])
for i in range(n):
    if i == n - 1:
        torch.cuda.nvtx.range_push("laststep")   # ncu --nvtx --nvtx-include "laststep/": the launches of ONE warm step
    step()
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("samples:", int(model.step_counter[(model.local_step - 1) % 16, 0].item()), "schedule:", model._last_train_schedule)
