"""The driver's bench contract, checked where it can run without a GPU: `bench.py --impl reference` (the reference's
PyTorch-only CPU path, restated in oracle/cpu_render.py) must print ONE JSON line with the agreed keys, and the native
arm must refuse to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout,
                          env=e, cwd=ROOT)


def test_reference_arm_json_line():
    p = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "3"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rays/sec palette-mode render" and j["unit"] == "rays/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 3
    assert j["value"] > 0 and abs(j["e2e"]["value"] - j["value"]) < 1e-9
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "1024" in cb["sample"]
    assert "workload" in j["config"] and j["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_native_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run(["--steps", "1"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
