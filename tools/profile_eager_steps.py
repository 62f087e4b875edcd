"""cProfile of eager training steps of both stages (torch.optim.Adam as the reference's trainers build it): where the host time of an
eager step goes. usage (under gpurun): python tools/profile_eager_steps.py"""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd())
import torch
from palettenerf_b200 import synthetic as S
dev = torch.device("cuda:0")
def make(kind):
    if kind == "nerf":
        m = S.build_nerf_model(dev, seed=0)
    else:
        m = S.build_palette_model(dev, seed=0, pred_clip=False)
    m.train()
    opt = torch.optim.Adam(m.get_params(1e-2), betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    o, d = S.training_rays(4096, seed=0)
    o, d = o.to(dev)[None].contiguous(), d.to(dev)[None].contiguous()
    gt = torch.rand(1, 4096, 3, device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            if kind == "nerf":
                out = m.render(o, d, rays_gt=gt, staged=False, bg_color=1, perturb=True, force_all_rays=False, dt_gamma=0.0, max_steps=1024)
                loss = (((out["image"] - gt) ** 2).mean(-1) + 0.05 * out["rgb_norm"]).mean()
            else:
                out = m.render(o, d, staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0, max_steps=1024)
                loss = ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() + 2e-4 * out["omega_sparsity"].mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
    return m, step
for kind in ("nerf", "palette"):
    m, step = make(kind)
    for _ in range(20): step()
    if kind == "nerf":
        m.mean_count = int(m.step_counter[:16, 0].sum().item() / 16)
        for _ in range(5): step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(100): step()
    torch.cuda.synchronize()
    print(kind, "eager ms/step", (time.perf_counter() - t0) * 10)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(100): step()
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
    print(s.getvalue()[:6000])
