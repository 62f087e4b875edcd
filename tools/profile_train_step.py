"""Kernel-time breakdown of one palette training step (BASELINE config 4): section wall times (with syncs), the
bench-style event timing, and a torch.profiler (CUPTI) table. Usage: python tools/profile_train_step.py [--clip] [--torch]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from palettenerf_b200 import synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
pred_clip = "--clip" in sys.argv
fused = None if "--torch" not in sys.argv else False
model = S.build_palette_model(dev, seed=0, pred_clip=pred_clip)
model.train()
opt = torch.optim.Adam(model.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True)
scaler = torch.amp.GradScaler("cuda")
o, d = S.training_rays(4096, seed=0)
o, d = o.to(dev), d.to(dev)
gt = torch.rand(1, 4096, 3, device=dev)
sec = {}


def tick(name, t0):
    torch.cuda.synchronize()
    sec[name] = sec.get(name, 0.0) + time.perf_counter() - t0
    return time.perf_counter()


def step(timed=False):
    t = time.perf_counter()
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = model.render(o[None], d[None], staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0,
                           max_steps=1024, fused=fused)
        if timed:
            t = tick("render_fwd", t)
        loss = ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean() \
            + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean()
    if timed:
        t = tick("loss", t)
    scaler.scale(loss).backward()
    if timed:
        t = tick("backward", t)
    scaler.step(opt)
    scaler.update()
    if timed:
        t = tick("optimizer", t)


if "--graph" not in sys.argv:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    for _ in range(5):
        step(timed=True)
    print("schedule:", getattr(model, "_last_train_schedule", "?"), " section wall ms/step (synchronised):",
          {k: round(v / 5 * 1e3, 3) for k, v in sec.items()})
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(); step(); b.record()
        cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        ts.append((a.elapsed_time(b), cpu * 1e3))
    print("event ms/step:", [round(x, 3) for x, _ in ts], " cpu issue ms/step:", [round(c, 3) for _, c in ts])
run = step
if "--graph" in sys.argv:
    from palettenerf_b200.graphs import GraphedStep
    g = GraphedStep(step, warmup=2)
    run = g.replay
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("cuda-graph replay ms/step:", [round(x, 3) for x in ts])
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        run()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
