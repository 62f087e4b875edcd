"""Kernel-time breakdown of one palette training step (BASELINE config 4) with torch.profiler (CUPTI)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from palettenerf_b200 import synthetic as S

dev = torch.device("cuda:0")
pred_clip = "--clip" in sys.argv
model = S.build_palette_model(dev, seed=0, pred_clip=pred_clip)
model.train()
opt = torch.optim.Adam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
scaler = torch.amp.GradScaler("cuda")
o, d = S.training_rays(4096, seed=0)
o, d = o.to(dev), d.to(dev)
gt = torch.rand(1, 4096, 3, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render(o[None], d[None], staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
        loss = ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean() \
            + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean()
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
