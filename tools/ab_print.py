"""one-line summary of a bench.py JSON line (A/B runs of library variants): python tools/ab_print.py <file> <label>"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
out = [sys.argv[2], "ms/view", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4)]
if "nerf_stage" in d:
    n = d["nerf_stage"]
    out += ["nerf render", round(n["render_fused_ms"], 4), "nerf train", round(n.get("train_step_ms", 0), 4), "density", round(n["density_update_full_fused_ms"], 4)]
if "mip360" in d:
    out += ["mip360", round(d["mip360"]["render_ms"], 4)]
if "hashgrid" in d:
    out += ["grid fwd16", round(d["hashgrid"]["fwd_f16_ms"], 4)]
if "train" in d:
    out += ["train", round(d["train"]["ms_per_step"], 4)]
print(*out)
