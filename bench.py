#!/usr/bin/env python
"""bench.py — headline benchmark of the PaletteNeRF hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's PyTorch-only CPU path (restated)

Metric (BASELINE.json): rays/sec of the palette-mode render. Workload at every N: BASELINE config 3 — one 800x800
view (640 000 rays) of the lego-shaped synthetic scene per GPU, palette model with 4 palettes and random-init
weights (seed 0), cuda_ray, fp16 autocast, all six auxiliary maps (gui_mode=False). One "step" = one view per rank;
ranks render different azimuths (weak scaling, no data-path collective: rays are independent).
Also reported in the same JSON line: the palette training step (config 4: 4096 rays/GPU, fwd+bwd+Adam, one NCCL
all-reduce of the gradients when N>1) and the hash-grid microbenchmark (config 2).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VIEW = 800
N_RAYS = VIEW * VIEW
TRAIN_RAYS = 4096
GRID_POINTS = 1 << 22
GRID_BYTES_PER_POINT = {"f16": 588, "f32": 1164}  # 12 B xyz + 16 levels * 8 corners * F*s B + 32*s B out


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# -------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's PyTorch-only CPU path (restated in oracle/cpu_render.py)
# -------------------------------------------------------------------------------------------------------------------
def _scene_module():
    """palettenerf_b200/synthetic.py (pure torch / numpy scene + camera helpers) loaded BY PATH: the reference arm must not
    import the product package, whose field modules map libpnerf_b200.so into the process"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_pnerf_scene", os.path.join(ROOT, "palettenerf_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import cpu_render
    S = _scene_module()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_rays = 1024
    params, _ = cpu_render.random_palette_params(seed=0, pred_clip=False)     # no product code in this process
    assert not any("libpnerf_b200" in l for l in open("/proc/self/maps")), "reference arm mapped the product library"
    g = torch.Generator().manual_seed(0)
    times = []
    for it in range(args.warmup + args.steps):
        inds = torch.randint(0, N_RAYS, (sample_rays,), generator=g)
        o, d = S.camera_rays(VIEW, VIEW, inds=inds)
        t0 = time.perf_counter()
        cpu_render.render_sampler(params, o, d, num_steps=512, pred_clip=False)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = sample_rays / (ms / 1e3)
    sample = (f"{sample_rays} random pixels of the 800x800 view per step, 512 uniform samples per ray "
              "(NeRFRenderer.run sampler, upsample_steps=0), torch fp32 CPU")
    line = {"impl": "reference", "metric": "rays/sec palette-mode render", "value": value, "unit": "rays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "palette-mode inference render 800x800 lego-shaped, 4 palettes (BASELINE config 3), "
                                   "SAMPLED: each step renders 1024 random pixels of the view with the reference's PyTorch-only "
                                   "CPU path (oracle port); the product library is not loaded in this process"},
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------------------------
# main arm
# -------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the train-step / hash-grid / cpu-baseline sections")
    ap.add_argument("--sections", default="hashgrid,train,mip360,strong,nerf,cpu",
                    help="extra sections to run (comma list of hashgrid,train,mip360,strong,nerf,cpu)")
    ap.add_argument("--torch-loss", action="store_true", help="training sections: the loss as torch tensor expressions "
                    "instead of palette_loss (A/B)")
    ap.add_argument("--torch-adam", action="store_true", help="training sections: torch.optim.Adam(fused, capturable) "
                    "instead of palettenerf_b200.optim.FusedAdam (A/B)")
    ap.add_argument("--nccl-allreduce", action="store_true", help="N > 1 training: dist.all_reduce (NCCL) for the gradient "
                    "bucket instead of pnerf_peer_allreduce over symmetric memory (A/B)")
    ap.add_argument("--no-graph", action="store_true", help="run the training step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--gui-mode", action="store_true", help="skip the five debug maps (reference gui_mode=True)")
    ap.add_argument("--fused", type=int, default=-1, help="-1 auto, 0 compatibility loop, 1 fused schedule")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from palettenerf_b200 import _lib as L
    from palettenerf_b200 import synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")   # the gradient all-reduce is captured in a CUDA graph
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    hbm_peak, tf_peak, peak_kind = _peaks()
    model = S.build_palette_model(dev, seed=0, pred_clip=False)
    model.eval()
    fused = None if args.fused < 0 else bool(args.fused)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # one view per rank, different azimuths
    o_host, d_host = S.camera_rays(VIEW, VIEW, azimuth_deg=35.0 + 45.0 * rank)
    o_pin, d_pin = o_host.pin_memory(), d_host.pin_memory()
    o_dev, d_dev = o_host.to(dev), d_host.to(dev)
    img_host = torch.empty(N_RAYS, 3).pin_memory()

    def render(o, d):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.render(o[None], d[None], staged=True, bg_color=1, perturb=False, gui_mode=args.gui_mode,
                                fused=fused, dt_gamma=S.LEGO["dt_gamma"], max_steps=S.LEGO["max_steps"], T_thresh=1e-4)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        if profile:
            L.profile_start()
        l0 = L.launch_count
        torch.cuda.nvtx.range_push("timed")     # ncu --nvtx --nvtx-include "timed/" lists exactly the timed region
        for _ in range(steps):
            flush.fill_(1)  # L2 flush between timed iterations (not timed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.nvtx.range_pop()
        barrier()
        prof = L.profile_stop() if profile else None
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        return max_over_ranks(ms), L.launch_count - l0, prof

    # ---- headline: device-resident rays --------------------------------------------------------------------------
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    import ctypes
    from palettenerf_b200 import fused as _fused
    rk = _fused.RENDER_KERNEL                     # "tc" (tcgen05 warp-per-ray), "rays" (mma.sync warp-per-ray), "lanes" (round 1)
    hook_on, hook_ms, kname = {"tc": ("pnerf_render_tc_timing", "pnerf_render_tc_last_ms", "k_render_rays_tc"),
                               "rays": ("pnerf_render_rays_timing", "pnerf_render_rays_last_ms", "k_render_rays"),
                               "lanes": ("pnerf_render_kernel_timing", "pnerf_render_kernel_last_ms", "k_render_fused")}[rk]
    getattr(L.lib, hook_ms).restype = ctypes.c_float
    getattr(L.lib, hook_on)(1)                    # event pair around the persistent kernel itself (roofline.kernel_ms)
    ms_step, launches, prof = timed(lambda: render(o_dev, d_dev), args.steps, args.warmup, profile=True)
    k_ms_last = float(getattr(L.lib, hook_ms)())  # the last timed launch (all launches are the same view)
    getattr(L.lib, hook_on)(0)
    clock_info = clocks.stop() if rank == 0 else None
    value = world * N_RAYS / (ms_step / 1e3)

    # ---- e2e: pinned host rays -> H2D -> render through the public API -> D2H image --------------------------------
    def e2e_step():
        o = o_pin.to(dev, non_blocking=True)
        d = d_pin.to(dev, non_blocking=True)
        out = render(o, d)
        img_host.copy_(out["image"].view(-1, 3), non_blocking=True)
    ms_e2e, _, _ = timed(e2e_step, args.steps, args.warmup)
    e2e = {"value": world * N_RAYS / (ms_e2e / 1e3), "unit": "rays/s", "h2d_bytes_per_step": 2 * N_RAYS * 12,
           "d2h_bytes_per_step": N_RAYS * 12, "ms_per_step": ms_e2e}

    # ---- e2e from the camera pose: 64 B pose H2D -> pnerf_get_rays -> render -> D2H image ------------
    import math
    from palettenerf_b200.nerf.utils import get_rays
    pose_pin = S.lookat_pose(S.LEGO["radius"], 35.0 + 45.0 * rank)[None].contiguous().pin_memory()
    focal = 0.5 * VIEW / math.tan(0.5 * S.LEGO["camera_angle_x"])
    intr = [focal, focal, VIEW / 2, VIEW / 2]

    def e2e_pose_step():
        r = get_rays(pose_pin.to(dev, non_blocking=True), intr, VIEW, VIEW, N=-1)
        out = render(r["rays_o"][0], r["rays_d"][0])
        img_host.copy_(out["image"].view(-1, 3), non_blocking=True)
    ms_pose, _, _ = timed(e2e_pose_step, args.steps, args.warmup)
    e2e["from_pose"] = {"value": world * N_RAYS / (ms_pose / 1e3), "ms_per_step": ms_pose, "h2d_bytes_per_step": 64,
                        "d2h_bytes_per_step": N_RAYS * 12,
                        "note": "rays generated on the device by pnerf_get_rays from the pinned-host camera pose"}

    # ---- roofline of the dominant kernel of the step (CUDA events around each C-ABI launch, timed region) ---------
    # pnerf_palette_render_fused = candidates + pre-pass + 2 ordering kernels + the persistent k_render_fused (92 % of the call,
    # profiles/r01_launches_bench_render_*.csv). It is the fused march -> hash-grid gather -> MLP -> blend -> composite
    # kernel; SURVEY §8(d) puts the MLP on the tensor roofline (36 094 FLOP per sample without the semantic branch).
    total_kernel_ms = sum(v[0] for v in prof.values()) or 1.0
    top = max(prof.items(), key=lambda kv: kv[1][0])
    shares = {k.replace("pnerf_", ""): round(v[0] / total_kernel_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
    q = getattr(model, "_last_queue", None)          # [hit-list cursor, samples shaded, rays with samples, tiles] of the last view
    samples_per_step = int(q[1].item()) if q is not None else None
    tile_fill = (float(q[1].item()) / (32.0 * max(1, int(q[3].item())))) if q is not None else None
    call_ms = top[1][0] / max(1, top[1][1])
    is_render = top[0].startswith("pnerf_palette_render")
    kernel_ms = k_ms_last if (k_ms_last > 0 and is_render) else call_ms
    FLOP_PER_SAMPLE = 36094                          # palette field without clip, forward
    GATHER_B_PER_SAMPLE = 2 * 16 * 8 * 4             # two fp16 F=2 tables, 16 levels, 8 corners
    roofline = {"bound": "tensor", "achieved": None, "peak": tf_peak, "unit": "TFLOP/s", "frac": None,
                "traffic": _ncu_traffic(kname), "traffic_source": "profiles/traffic.json (ncu --set full capture of this "
                "kernel, committed; not measured in this run)",
                "kernel": top[0] + f" ({kname})", "kernel_ms": kernel_ms, "call_ms": call_ms,
                "peak_kind": peak_kind + " (bf16 dense, burst; fp16 assumed equal)",
                "algorithmic": f"{FLOP_PER_SAMPLE} FLOP/sample x {samples_per_step} samples per launch",
                "kernel_time_share_of_own_kernels": shares}
    if samples_per_step and is_render:
        ach = FLOP_PER_SAMPLE * samples_per_step / (kernel_ms / 1e3) / 1e12
        roofline.update(achieved=ach, frac=ach / tf_peak,
                        note=_ncu_note(kname),
                        l2_gather={"achieved_gbs": GATHER_B_PER_SAMPLE * samples_per_step / (kernel_ms / 1e3) / 1e9,
                                   "algorithmic": f"{GATHER_B_PER_SAMPLE} B gathered per sample (2 tables x 16 levels x 8 corners x 4 B)",
                                   "hbm_peak_gbs_for_scale": hbm_peak})

    extras = {}
    sections = set() if args.no_extras else set(args.sections.split(","))
    if not args.no_extras:
        # config 3's second variant: the same view without / with the five debug maps (reference gui_mode)
        other = not args.gui_mode

        def render_other():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                return model.render(o_dev[None], d_dev[None], staged=True, bg_color=1, perturb=False, gui_mode=other, fused=fused,
                                    dt_gamma=S.LEGO["dt_gamma"], max_steps=S.LEGO["max_steps"], T_thresh=1e-4)
        ms_other, _, _ = timed(render_other, max(3, min(args.steps, 10)), 3)
        extras["render_gui_mode_variant"] = {"gui_mode": other, "ms_per_step": ms_other, "rays_per_s": world * N_RAYS / (ms_other / 1e3),
                                             "note": "gui_mode=True skips direct / view-dependent / basis maps (palette/renderer.py:436-443)"}
    if "hashgrid" in sections:
        extras.update(bench_hashgrid(torch, dev, L, hbm_peak, flush))
    if "train" in sections:
        try:
            extras.update(bench_train(torch, dist, dev, world, rank, S, L, barrier, max_over_ranks, flush,
                                      use_graph=not args.no_graph, torch_loss=args.torch_loss, torch_adam=args.torch_adam,
                                      nccl_allreduce=args.nccl_allreduce))
        except Exception as e:  # noqa: BLE001  (the headline line must still be printed)
            extras["train"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if "strong" in sections and world > 1:
        try:
            extras.update(bench_strong(torch, dev, rank, world, S, model, barrier, max_over_ranks, flush, args))
        except Exception as e:  # noqa: BLE001
            extras["render_strong"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if "nerf" in sections and rank == 0:
        try:
            extras.update(bench_nerf(torch, dev, S, flush))
        except Exception as e:  # noqa: BLE001
            extras["nerf_stage"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    if "mip360" in sections:
        try:
            extras.update(bench_mip360(torch, dev, rank, S, L, barrier, max_over_ranks, flush, world))
        except Exception as e:  # noqa: BLE001
            extras["mip360"] = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    # secondary roofline: the stand-alone hash-grid kernel of BASELINE config 2 (HBM-bound I/O, L2-resident table)
    if "hashgrid" in extras:
        hg = extras["hashgrid"]
        roofline["hashgrid_microbench"] = {
            "bound": "hbm", "kernel": "k_grid_fwd_coop_h (fp16, 2^22 points)", "achieved": hg["fwd_f16_hbm_gbs"], "peak": hbm_peak,
            "unit": "GB/s", "frac": hg["fwd_f16_hbm_gbs"] / hbm_peak,
            "traffic": _ncu_traffic("k_grid_fwd_coop_h"),
            "algorithmic": "compulsory HBM bytes: 76 B/point (12 B coordinates + 64 B fp16 output) x 2^22 points; the 24 MiB table is "
                           "L2-resident and read once",
            "l2_gather": {"achieved_gbs": hg["fwd_f16_gather_gbs"], "peak_gbs": hg["l2_read_peak_gbs_measured"],
                          "frac": hg["fwd_f16_gather_gbs"] / hg["l2_read_peak_gbs_measured"],
                          "algorithmic": "512 B gathered per point (16 levels x 8 corners x 4 B) vs a measured L2 read peak"},
            "ncu": _ncu_entry("k_grid_fwd_coop_h"),
            "note": _ncu_note("k_grid_fwd_coop_h")}

    cpu_baseline = None
    if rank == 0 and world == 1 and "cpu" in sections:
        cpu_baseline = bench_cpu_baseline(S)

    if rank == 0:
        line = {"metric": "rays/sec palette-mode render", "value": value, "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "palette-mode inference render 800x800 lego-shaped, cuda_ray, 4 palettes "
                                       "(BASELINE config 3), one view per GPU",
                           "rays_per_step_per_gpu": N_RAYS, "gui_mode": bool(args.gui_mode),
                           "schedule": getattr(model, "_last_schedule", "loop"), "render_kernel": rk,
                           "samples_per_step": samples_per_step, "tile_fill": tile_fill, "l2": "flushed between timed iterations (256 MB write)"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clock_info, "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _ncu_note(kernel):
    e = _ncu_entry(kernel) or {}
    return e.get("note", "see profiles/README.md for the ncu capture of this kernel")


def _ncu_entry(kernel):
    """the committed ncu figures of `kernel` (profiles/traffic.json) or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        e = json.load(open(p)).get(kernel)
        return {k: v for k, v in e.items() if k != "source"} if e else None
    except Exception:
        return None


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json, written from the .ncu-rep by tools/ncu_summary.py), or None"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def _measure_l2_read_peak(torch, dev):
    """bandwidth of L2-resident buffers (the better of a 48 MiB reduction and a 24 -> 24 MiB copy, 20 back-to-back launches):
    the denominator for the gather traffic of the hash grid, whose table lives in L2"""
    best = 0.0
    x = torch.ones(12 << 20, dtype=torch.float32, device=dev)
    y = torch.empty(6 << 20, dtype=torch.float32, device=dev)
    for fn, nbytes in ((lambda: x.sum(), x.numel() * 4), (lambda: y.copy_(x[:6 << 20]), 2 * y.numel() * 4)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = max(best, 20 * nbytes / (a.elapsed_time(b) / 1e3) / 1e9)
    return best


def bench_hashgrid(torch, dev, L, hbm_peak, flush):
    """BASELINE config 2 as SURVEY 8d specifies it: 2^22 points U[0,1)^3, 16 levels, F=2, log2T=19, fwd + bwd, fp16 and fp32
    tables, per_level_scale 1.447269 (desired resolution 4096) and 1.381913 (2048), cold (L2 flushed before every launch)
    and warm (back to back). Two units, never mixed: compulsory HBM bytes per point (12 B coordinates + 32*s B output; the table is
    L2-resident) against the measured HBM copy peak, and gathered bytes per point (16 levels x 8 corners x F*s B) against a
    measured L2 read peak."""
    from palettenerf_b200.gridencoder import GridEncoder
    from palettenerf_b200.gridencoder.backend import _backend as GB
    import numpy as np
    res = {}
    l2_peak = _measure_l2_read_peak(torch, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(GRID_POINTS, 3, device=dev, generator=g)
    for res_name, desired in (("", 4096), ("_res2048", 2048)):
        enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, desired_resolution=desired).to(dev)
        S_ = float(np.log2(enc.per_level_scale))
        for name, dt in (("f16", torch.float16), ("f32", torch.float32)):
            if res_name and name == "f32":
                continue
            emb = enc.embeddings.detach().to(dt)
            out = torch.empty(GRID_POINTS, 32, device=dev, dtype=dt)
            grad = torch.randn(GRID_POINTS, 32, device=dev, generator=g).to(dt)
            gemb = torch.zeros_like(emb)

            def fwd():
                GB.grid_encode_forward_blc(x, emb, enc.offsets, out, GRID_POINTS, 3, 2, 16, S_, 16, None, 0, False)

            def bwd():
                GB.grid_encode_backward_blc(grad, x, emb, enc.offsets, gemb, GRID_POINTS, 3, 2, 16, S_, 16, None, None, 0, False)
            for fn, tag in ((fwd, "fwd"), (bwd, "bwd")):
                if res_name and tag == "bwd":
                    continue
                for _ in range(5):
                    fn()
                for temp in ("cold", "warm"):
                    ts = []
                    for _ in range(20):
                        if temp == "cold":
                            flush.fill_(1)
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record(); fn(); b.record()
                        torch.cuda.synchronize()
                        ts.append(a.elapsed_time(b))
                    ms = sum(ts) / len(ts)
                    s = 2 if name == "f16" else 4
                    key = f"{tag}_{name}{res_name}" + ("" if temp == "cold" else "_warm")
                    res[f"{key}_ms"] = ms
                    res[f"{key}_hbm_gbs"] = (12 + 32 * s) * GRID_POINTS / (ms / 1e3) / 1e9
                    res[f"{key}_gather_gbs"] = 16 * 8 * 2 * s * GRID_POINTS / (ms / 1e3) / 1e9
                    if temp == "cold":     # (kept for continuity with round 1: algorithmic 588 / 1164 B per point)
                        res[f"{key}_eff_gbs"] = GRID_BYTES_PER_POINT[name] * GRID_POINTS / (ms / 1e3) / 1e9
    res["points"] = GRID_POINTS
    res["l2_read_peak_gbs_measured"] = l2_peak
    res["fwd_f16_frac_of_hbm_peak"] = res["fwd_f16_hbm_gbs"] / hbm_peak
    res["fwd_f16_gather_frac_of_l2_peak"] = res["fwd_f16_gather_gbs"] / l2_peak
    res["note"] = ("hbm_gbs = compulsory 12 + 32*s B per point vs the measured HBM copy peak; gather_gbs = 16 levels x 8 corners x F*s B "
                   "per point vs l2_read_peak_gbs_measured (48 MiB buffer re-read by a reduction kernel); cold = L2 flushed before every "
                   "launch (256 MB write), warm = back to back; _res2048 = per_level_scale 1.381913")
    return {"hashgrid": res}


def _allreduce_name(bucket, nccl_allreduce):
    if nccl_allreduce or bucket is None or bucket._pm is None:
        return "nccl all_reduce"
    if bucket._pm.mc_ptr:
        return "pnerf_peer_allreduce_mc (in-switch reduction: multimem.ld_reduce / multimem.st on the multicast mapping)"
    return "pnerf_peer_allreduce (two-shot over NVLink peer memory)"


def _allreduce_check(torch, dist, dev, rank, world):
    """one-shot parity check of the peer all-reduce kernel against NCCL on the same data (max abs error over 4 M floats)"""
    from palettenerf_b200.distributed import PeerMemory
    try:
        n = 1 << 22
        pm = PeerMemory(n, dev)
        x = torch.randn(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1234 + rank))
        pm.buf[:n].copy_(x)
        pm.all_reduce_(average=True)
        ref = x.clone()
        dist.all_reduce(ref)
        ref.div_(world)
        err = (pm.buf[:n] - ref).abs().max()
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        return float(err.item())
    except Exception as e:  # noqa: BLE001
        return f"check failed: {type(e).__name__}: {str(e)[:120]}"


def _adam(torch, params, torch_adam):
    if torch_adam:
        return torch.optim.Adam(params, betas=(0.9, 0.99), eps=1e-15, fused=True, capturable=True), \
            "torch Adam(0.9,0.99,1e-15,fused,capturable)+GradScaler"
    from palettenerf_b200.optim import FusedAdam
    return FusedAdam(params, betas=(0.9, 0.99), eps=1e-15), \
        "FusedAdam(0.9,0.99,1e-15: pnerf_adam_step)+GradScaler(non-finite check: pnerf_found_inf)"


def _scaler(torch, torch_adam):
    """torch's GradScaler; with FusedAdam its non-finite check in front of the step is one pass of pnerf_found_inf"""
    if torch_adam:
        return torch.amp.GradScaler("cuda")
    from palettenerf_b200.optim import GradScaler
    return GradScaler("cuda")


def _kernel_breakdown(torch, step, barrier, reps=3):
    """device time per kernel name of one step (torch.profiler / CUPTI over `reps` replays, warm L2, outside every timed
    region): says where a step's time goes — in particular what data parallelism adds (pack copy, all-reduce, barriers)"""
    try:
        from torch.profiler import ProfilerActivity, profile
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                step()
            torch.cuda.synchronize()
        barrier()
        rows = {}
        for e in prof.events():
            if getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type):
                k = e.name[:64]
                rows[k] = rows.get(k, 0.0) + float(e.device_time if hasattr(e, "device_time") else e.cuda_time) / reps
        top = sorted(rows.items(), key=lambda kv: -kv[1])[:24]
        return {"total_us": round(sum(rows.values()), 1), "kernels": {k: round(v, 1) for k, v in top}}
    except Exception as e:   # noqa: BLE001  (a diagnostic: never fails the bench)
        return {"error": f"{type(e).__name__}: {e}"}


def bench_train(torch, dist, dev, world, rank, S, L, barrier, max_over_ranks, flush, use_graph=True, torch_loss=False,
                torch_adam=False, nccl_allreduce=False):
    """BASELINE config 4: palette-stage training step, 4096 rays per GPU, fwd + bwd + Adam under fp16 autocast with
    GradScaler; ray-batch data parallel with ONE all-reduce over a flat gradient bucket when N > 1.
    The step (static-capacity march, fused field fwd/bwd/wgrad, one-pass compositor, loss, all-reduce, GradScaler, fused
    Adam) has data-independent shapes and no host synchronisation, so it is captured ONCE in a CUDA graph and replayed."""
    from palettenerf_b200.distributed import GradBucket
    from palettenerf_b200.graphs import GraphedStep, make_palette_train_step
    model = S.build_palette_model(dev, seed=0, pred_clip=False)
    model.train()
    opt, opt_name = _adam(torch, model.get_params(1e-2), torch_adam)
    params = [p for grp in opt.param_groups for p in grp["params"] if p.requires_grad]
    scaler = _scaler(torch, torch_adam)
    bucket = GradBucket(params, peer=not nccl_allreduce) if world > 1 else None
    o, d = S.training_rays(TRAIN_RAYS, seed=rank)
    o, d = o.to(dev)[None].contiguous(), d.to(dev)[None].contiguous()
    gt = torch.rand(1, TRAIN_RAYS, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))

    from palettenerf_b200.palette.losses import palette_loss

    def loss_fn(out):
        # PaletteTrainer.train_step's loss (palette/utils.py:486-567: rgb + direct rgb + the three regularisers with the
        # reference's default lambdas) in one kernel; the palette term is off (lambda_palette = 0 until basis colours unfreeze)
        if torch_loss:
            return ((out["image"] - gt) ** 2).mean() + ((out["direct_rgb"] - gt) ** 2).mean() \
                + 2e-4 * out["omega_sparsity"].mean() + 0.03 * out["offsets_norm"].mean() + 0.1 * out["view_dep_norm"].mean()
        return palette_loss(out, gt, lambda_sparsity=2e-4, lambda_offsets=0.03, lambda_view_dep=0.1)[0]

    step_fn = make_palette_train_step(model, opt, scaler, o, d, loss_fn, bucket=bucket)
    mode = "eager"
    step = step_fn
    l0 = L.launch_count
    if use_graph:
        g = GraphedStep(step_fn, warmup=3)      # 3 eager steps on the capture stream, then one captured step
        step, mode = g.replay, "cuda_graph"
        own_launches = (L.launch_count - l0) // 4
    else:
        step_fn()
        own_launches = L.launch_count - l0      # C-ABI kernels per step (torch's own kernels are not counted)
    for _ in range(3):
        step()
    barrier()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    barrier()
    m = int(model.step_counter[(model.local_step - 1) % 16, 0].item())
    ms = max_over_ranks(sum(ts) / len(ts))
    breakdown = _kernel_breakdown(torch, step, barrier)
    # config 4 WITH the smooth loss (ref: palette/renderer.py:360-381, on from epoch 30 of the reference schedule): a second
    # fused field evaluation on jittered points + the gate arithmetic; single GPU only (same all-reduce as above otherwise)
    smooth = None
    if world == 1 and use_graph and not torch_loss:
        m2 = S.build_palette_model(dev, seed=0, pred_clip=False)
        m2.train()
        m2.require_smooth_loss = True
        opt2, _ = _adam(torch, m2.get_params(1e-2), torch_adam)
        sc2 = _scaler(torch, torch_adam)

        def loss2(out):
            return palette_loss(out, gt, lambda_sparsity=2e-4, lambda_offsets=0.03, lambda_view_dep=0.1, lambda_smooth=4e-3)[0]
        g2 = GraphedStep(make_palette_train_step(m2, opt2, sc2, o, d, loss2), warmup=3)
        for _ in range(3):
            g2.replay()
        t2 = []
        for _ in range(10):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g2.replay(); b.record()
            torch.cuda.synchronize()
            t2.append(a.elapsed_time(b))
        smooth = {"ms_per_step": sum(t2) / len(t2), "rays_per_s": TRAIN_RAYS / (sum(t2) / len(t2) / 1e3),
                  "schedule": getattr(m2, "_last_train_schedule", "torch"), "lambda_smooth": 4e-3,
                  "kernel_us_per_step": _kernel_breakdown(torch, g2.replay, barrier)}
        del g2, m2, opt2
    # the optimizer alone (HBM stream: 16 B read + 12 B written per element with a gradient), same events / L2 flush
    adam = None
    if not torch_adam:
        n_el = sum(p.numel() for p in params if p.grad is not None)
        tt = []
        for _ in range(10):
            flush.fill_(1)
            torch.cuda.synchronize()
            L.profile_start()                    # CUDA events around the C-ABI call itself (not around Python's step())
            opt.step()
            pr = L.profile_stop()
            tt.append(pr["pnerf_adam_step"][0] / max(1, pr["pnerf_adam_step"][1]))
        ams = sum(tt) / len(tt)
        adam = {"ms": ams, "elements": n_el, "gbs": 28.0 * n_el / (ams / 1e3) / 1e9,
                "algorithmic": "28 B/element (p, g, m, v read; p, m, v written), all tensors of the step in one launch"}
    # the stage-1 (NeRF) step data parallel: same bucket machinery, the density table's gradient all-reduced early
    stage1 = None
    if world > 1 and use_graph and not torch_adam:
        try:
            from palettenerf_b200.graphs import make_nerf_train_step
            m1 = S.build_nerf_model(dev, seed=0)
            m1.train()
            opt1, _ = _adam(torch, m1.get_params(1e-2), False)
            p1 = [p for grp in opt1.param_groups for p in grp["params"] if p.requires_grad]
            b1 = GradBucket(p1, peer=not nccl_allreduce)
            g1 = GraphedStep(make_nerf_train_step(m1, opt1, _scaler(torch, False), o, d, gt, bucket=b1), warmup=3)
            for _ in range(3):
                g1.replay()
            barrier()
            t1 = []
            for _ in range(10):
                flush.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g1.replay(); b.record()
                torch.cuda.synchronize()
                t1.append(a.elapsed_time(b))
            barrier()
            ms1 = max_over_ranks(sum(t1) / len(t1))
            stage1 = {"ms_per_step": ms1, "rays_per_s": world * TRAIN_RAYS / (ms1 / 1e3), "schedule": m1._last_train_schedule,
                      "allreduce_bucket_bytes": None if b1.flat is None else 4 * b1.flat.numel(), "early_allreduces": b1.early_count,
                      "workload": "stage-1 (NeRF) training step, 4096 rays per GPU, MSE + rgb_norm loss, one CUDA graph per step"}
            del g1, m1, opt1, b1
        except Exception as e:   # noqa: BLE001  (a secondary record)
            stage1 = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
    return {"train": {"rays_per_s": world * TRAIN_RAYS / (ms / 1e3), "ms_per_step": ms, "rays_per_gpu": TRAIN_RAYS, "stage1_dp": stage1,
                      "samples_per_step_rank0": m, "schedule": getattr(model, "_last_train_schedule", "torch"),
                      "launch": mode, "own_kernel_launches_per_step": own_launches,
                      "optimizer": opt_name, "adam_alone": adam,
                      "allreduce": None if world == 1 else _allreduce_name(bucket, nccl_allreduce),
                      "allreduce_bucket_bytes": None if bucket is None or bucket.flat is None else 4 * bucket.flat.numel(),
                      "allreduce_max_abs_err": None if world == 1 else _allreduce_check(torch, dist, dev, rank, world),
                      "kernel_us_per_step_rank0": breakdown, "with_smooth_loss": smooth,
                      "loss": "torch expressions" if torch_loss else "palette_loss (fused: 2 launches fwd + 1 bwd)",
                      "workload": "palette-stage training step (BASELINE config 4), force_all_rays, no smooth loss"}}


def bench_nerf(torch, dev, S, flush):
    """stage-1 model (NeRFNetwork: one hash grid, sigma + colour nets): the 800x800 render on the persistent tensor-core
    renderer vs this repository's host-loop schedule (the reference's run_cuda loop on the new per-op kernels), and the
    density-grid refresh (update_extra_state, full sweep and partial) as kernels vs the torch-op schedule. Rank 0 only."""
    import torch.distributed as dist
    import palettenerf_b200.distributed as D
    model = S.build_nerf_model(dev, seed=0)
    model.eval()
    o, d = S.camera_rays(VIEW, VIEW, azimuth_deg=35.0)
    o, d = o.to(dev)[None], d.to(dev)[None]

    def timeit(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sum(ts) / len(ts)

    def render(fused):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.render(o, d, staged=True, bg_color=1, perturb=False, fused=fused, dt_gamma=0.0, max_steps=1024)
    res = {"render_fused_ms": timeit(lambda: render(None)), "render_schedule": model._last_schedule,
           "render_loop_ms": timeit(lambda: render(False), reps=2, warm=1), "rays": N_RAYS}
    res["render_rays_per_s"] = N_RAYS / (res["render_fused_ms"] / 1e3)
    real_world = D.world
    D.world = lambda: (1, 0)          # the refresh below is timed on this rank alone (no tile sharding / all-reduce)
    try:
        for name, it0 in (("full", 0), ("partial", 16)):
            def upd(fused):
                model.iter_density = it0
                with torch.autocast("cuda", dtype=torch.float16):
                    model.update_extra_state(fused=fused)
            res[f"density_update_{name}_fused_ms"] = timeit(lambda: upd(None), reps=3, warm=1)
            res[f"density_update_{name}_schedule"] = model._last_update_schedule
            res[f"density_update_{name}_torch_ms"] = timeit(lambda: upd(False), reps=2, warm=1)
    finally:
        D.world = real_world
    # stage-1 TRAINING step (ref: Trainer.train_step, nerf/utils.py:485-560 with the MSE criterion + rays_gt error channel):
    # 4096 rays, fwd + bwd + Adam under autocast + GradScaler. Three schedules on the same model and rays: the fused step
    # (csrc/nerf_train.cu: one field forward kernel, hand-written backward, one-pass compositor) eager and replayed from a
    # CUDA graph, and the reference's schedule on this repository's per-op kernels with the MLPs in torch.
    try:
        from palettenerf_b200.graphs import GraphedStep
        from palettenerf_b200.optim import FusedAdam
        to, td = S.training_rays(TRAIN_RAYS, seed=0)
        to, td = to.to(dev)[None].contiguous(), td.to(dev)[None].contiguous()
        gt = torch.rand(1, TRAIN_RAYS, 3, device=dev)

        def make(fused):
            tm = S.build_nerf_model(dev, seed=0)
            tm.train()
            opt = FusedAdam(tm.get_params(1e-2), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
            scaler = _scaler(torch, False)

            def train_step():
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.float16):
                    out = tm.render(to, td, rays_gt=gt, staged=False, bg_color=1, perturb=True, force_all_rays=True, dt_gamma=0.0,
                                    max_steps=1024, fused=fused)
                    loss = (((out["image"] - gt) ** 2).mean(-1) + 0.05 * out["rgb_norm"]).mean()   # lambda_sparse default, main_nerf.py:67
                scaler.scale(loss).backward()
                scaler.step(opt)
                scaler.update()
                return loss.detach()
            return tm, train_step
        tm, step = make(None)
        gs = GraphedStep(step, warmup=3)              # the capture stream must be the first to step this model
        res["train_step_ms"] = timeit(gs.replay, reps=20, warm=5)
        res["train_schedule"] = tm._last_train_schedule + " (csrc/nerf_train.cu), one CUDA graph per step"
        res["train_rays_per_s"] = TRAIN_RAYS / (res["train_step_ms"] / 1e3)
        res["train_samples_per_step"] = int(tm.step_counter[(tm.local_step - 1) % 16, 0].item())
        tm2, step2 = make(None)
        res["train_step_fused_eager_ms"] = timeit(step2, reps=10, warm=5)
        tm3, step3 = make(False)
        res["train_step_per_op_ms"] = timeit(step3, reps=10, warm=5)
        res["train_per_op_schedule"] = tm3._last_train_schedule + ": per-op kernels + torch MLPs, eager (round-2 record: 2.66 ms)"
        from palettenerf_b200 import _lib as L_
        L_.profile_start()
        step2()
        prof = L_.profile_stop()
        res["train_kernel_us_fused_eager"] = {k.replace("pnerf_", ""): round(1e3 * v[0] / max(1, v[1]), 1) for k, v in
                                              sorted(prof.items(), key=lambda kv: -kv[1][0])}
    except Exception as e:   # noqa: BLE001  (a secondary record: never fails the bench)
        res["train_step_error"] = f"{type(e).__name__}: {e}"
    res["note"] = ("render: csrc/field_tc.cu (model_kind 1) vs the host loop of nerf/renderer.py:329-386 on the per-op kernels; "
                   "density update: csrc/density_tc.cu (3-4 launches, threshold on the device) vs the torch-op schedule of "
                   "nerf/renderer.py:467-561 (the only host read left in the fused path is mean_count)")
    return {"nerf_stage": res}


def bench_strong(torch, dev, rank, world, S, model, barrier, max_over_ranks, flush, args):
    """STRONG scaling of the render (north_star: "rendering is split by ray/image tile across the 8 GPUs"): ONE 800x800 view
    (and one 1297x840 mip-360-shaped view with the semantic branch) dealt out in interleaved 256-ray tiles over the N ranks;
    every rank's persistent kernel stores its finished rays straight into rank 0's image in symmetric (peer-mapped)
    memory, two cross-GPU barriers bracket it. rays/s = rays of the ONE view / max-over-ranks device time."""
    from palettenerf_b200.distributed import ShardedView
    res = {}
    cases = [("800x800", model, VIEW, VIEW, S.LEGO["dt_gamma"])]
    m5 = S.build_palette_model(dev, seed=0, pred_clip=True, ground=True, scene_scale=1.5)
    m5.min_near = 0.05
    m5.eval()
    cases.append(("1297x840_clip", m5, 840, 1297, 1.0 / 128))
    for name, mdl, Hh, Ww, dtg in cases:
        o, d = S.camera_rays(Hh, Ww, azimuth_deg=35.0)
        o, d = o.to(dev), d.to(dev)
        view = ShardedView(mdl, Hh * Ww, gui_mode=args.gui_mode, tile=256)
        for _ in range(3):
            view.render(o, d, bg_color=1, dt_gamma=dtg)
        barrier()
        ts = []
        for _ in range(max(args.steps, 5)):
            flush.fill_(1)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); view.render(o, d, bg_color=1, dt_gamma=dtg); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = max_over_ranks(sum(ts) / len(ts))
        res[name] = {"ms_per_view": ms, "rays_per_s": Hh * Ww / (ms / 1e3), "rays": Hh * Ww,
                     "samples_rank0": int(view.queue[1].item())}
    res["scaling"] = "strong"
    res["how"] = ("one view, interleaved 256-ray tiles over the ranks, retire() stores into rank 0's maps over NVLink peer memory "
                  "(ShardedView); timed region = zero-fill of the maps on rank 0 + barrier + per-rank near/far, candidate filter, "
                  "pre-pass, persistent kernel + barrier + per-ray epilogue on rank 0")
    return {"render_strong": res}


def bench_mip360(torch, dev, rank, S, L, barrier, max_over_ranks, flush, world):
    """BASELINE config 5: Mip-NeRF-360-shaped unbounded scene, 1297x840 view (1 089 480 rays), dt_gamma = 1/128,
    min_near = 0.05, semantic-feature branch on (--pred_clip, clip_dim 16: third hash grid + clip net): one render
    and one training step (4096 rays, rgb + feature loss) per GPU."""
    from palettenerf_b200.graphs import GraphedStep, make_palette_train_step
    Hm, Wm, DTG = 840, 1297, 1.0 / 128
    res = {"workload": "mip-360-shaped scene (solid x1.5 + ground slab reaching cascade 1), 1297x840, dt_gamma 1/128, "
                       "min_near 0.05, pred_clip clip_dim 16 (BASELINE config 5)", "rays": Hm * Wm}
    model = S.build_palette_model(dev, seed=0, pred_clip=True, ground=True, scene_scale=1.5)
    model.min_near = 0.05
    model.eval()
    o, d = S.camera_rays(Hm, Wm, azimuth_deg=35.0 + 45.0 * rank)
    o, d = o.to(dev), d.to(dev)

    def render():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.render(o[None], d[None], staged=True, bg_color=1, perturb=False, gui_mode=False, dt_gamma=DTG,
                                max_steps=1024, T_thresh=1e-4)
    for _ in range(3):
        render()
    barrier()
    ts = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); render(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = max_over_ranks(sum(ts) / len(ts))
    q = getattr(model, "_last_queue", None)
    res.update(render_ms=ms, render_rays_per_s=world * Hm * Wm / (ms / 1e3), schedule=getattr(model, "_last_schedule", "loop"),
               samples_per_view=int(q[1].item()) if q is not None else None)

    # training step with the semantic-feature loss (palette/utils.py:486-567: rgb + direct rgb + regularisers + feature MSE)
    tm = S.build_palette_model(dev, seed=0, pred_clip=True, ground=True, scene_scale=1.5)
    tm.min_near = 0.05
    tm.train()
    opt, _ = _adam(torch, tm.get_params(1e-2), False)
    scaler = _scaler(torch, False)
    to, td = S.training_rays(TRAIN_RAYS, H=Hm, W=Wm, seed=rank)
    to, td = to.to(dev)[None].contiguous(), td.to(dev)[None].contiguous()
    g = torch.Generator(device=dev).manual_seed(rank)
    gt = torch.rand(1, TRAIN_RAYS, 3, device=dev, generator=g)
    feat = torch.randn(1, TRAIN_RAYS, 16, device=dev, generator=g)

    from palettenerf_b200.palette.losses import palette_loss

    def loss_fn(out):
        return palette_loss(out, gt, lambda_sparsity=2e-4, lambda_offsets=0.03, lambda_view_dep=0.1, gt_clip_feat=feat)[0]
    bucket5 = None
    if world > 1:
        from palettenerf_b200.distributed import GradBucket
        bucket5 = GradBucket([p for grp in opt.param_groups for p in grp["params"] if p.requires_grad], peer=True)
    step_fn = make_palette_train_step(tm, opt, scaler, to, td, loss_fn, render_kwargs=dict(dt_gamma=DTG), bucket=bucket5)
    gs = GraphedStep(step_fn, warmup=3)
    for _ in range(3):
        gs.replay()
    barrier()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gs.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = max_over_ranks(sum(ts) / len(ts))
    res.update(train_ms=ms, train_rays_per_s=world * TRAIN_RAYS / (ms / 1e3),
               train_samples_rank0=int(tm.step_counter[(tm.local_step - 1) % 16, 0].item()),
               train_schedule=getattr(tm, "_last_train_schedule", "torch"),
               train_launch="cuda_graph" + ("" if world == 1 else ", data parallel: one all-reduce of the gradient bucket"),
               train_allreduce=None if bucket5 is None else _allreduce_name(bucket5, False),
               train_bucket_bytes=None if bucket5 is None or bucket5.flat is None else 4 * bucket5.flat.numel())
    return {"mip360": res}


def bench_cpu_baseline(S):
    """BASELINE config 1 in full: ONE 200x200 view (40 000 rays x 512 samples = 20.5 M samples) of the same scene through the
    oracle's restatement of the reference's PyTorch-only CPU path, on this box's host cores: one small
    warm-up, one timed repetition (about 20 s; a reported baseline, not the optimisation target)"""
    import torch
    from oracle import cpu_render
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    params, _ = cpu_render.random_palette_params(seed=0, pred_clip=False)
    side = 200
    o, d = S.camera_rays(side, side)
    cpu_render.render_sampler(params, o[:512], d[:512], num_steps=512)  # warm-up
    t0 = time.perf_counter()
    cpu_render.render_sampler(params, o, d, num_steps=512)
    dt = time.perf_counter() - t0
    return {"value": side * side / dt, "unit": "rays/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"the whole {side}x{side} view of BASELINE config 1 (40 000 rays x 512 uniform samples, NeRFRenderer.run "
                      "sampler in 4096-ray chunks, palette field + blend), torch fp32, one repetition after a 512-ray warm-up"}


if __name__ == "__main__":
    main()
