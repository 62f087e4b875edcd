"""Runs a few eager fused stage-1 (NeRF) training steps — the command ncu wraps for csrc/nerf_train.cu's kernels."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from palettenerf_b200 import synthetic as S  # noqa: E402
from palettenerf_b200.optim import FusedAdam  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = S.build_nerf_model(dev, seed=0)
model.train()
opt = FusedAdam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
scaler = torch.amp.GradScaler("cuda")
o, d = S.training_rays(4096, seed=0)
o, d = o.to(dev)[None].contiguous(), d.to(dev)[None].contiguous()
gt = torch.rand(1, 4096, 3, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render(o, d, rays_gt=gt, staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
        loss = (((out["image"] - gt) ** 2).mean(-1) + 0.05 * out["rgb_norm"]).mean()
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()


for i in range(n):
    if i == n - 1:
        torch.cuda.nvtx.range_push("laststep")   # ncu --nvtx --nvtx-include "laststep/": the launches of ONE warm step
    step()
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("samples:", int(model.step_counter[(model.local_step - 1) % 16, 0].item()), "schedule:", model._last_train_schedule)
