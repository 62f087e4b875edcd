"""Row b2: the reference's UNMODIFIED main_nerf.py and main_palette.py run end to end on the drop-in packages
(tools/run_reference_main.py installs them under the reference's import names; compat/ stands in for the third-party
packages this image lacks). Stage 1 trains a NeRF for 8 epochs on a 12-view synthetic Blender-format scene, evaluates, tests
and exports; stage 2 loads that checkpoint into the palette model and trains / evaluates it. Asserted: both mains exit 0,
the epoch-mean loss falls, and the renderer calls took the fused schedules (density refresh, eval render, stage-1 and
palette train steps). The reference sources come from /root/reference here and from oracle/_ref/py (staged by build()) on the GPU box."""
import json
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

RUNNER = os.path.join(ROOT, "tools", "run_reference_main.py")
COMMON = ["-O", "--bound", "2", "--scale", "0.8", "--dt_gamma", "0", "--num_rays", "1024"]


def _have_reference():
    return os.path.isdir("/root/reference/palette") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "py", "palette"))


def _run(args, cwd):
    env = dict(os.environ)
    for k in ("PNERF_RENDER_KERNEL", "PNERF_FIELD_KERNEL"):
        env.pop(k, None)
    p = subprocess.run([sys.executable, RUNNER] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900, env=env)
    log = p.stdout
    assert p.returncode == 0, log[-6000:]
    # one entry per epoch: the running mean tqdm prints with the 100 % bar, just before the trainer's "Finished Epoch"
    epoch_means = [float(m) for m in re.findall(r"loss=[0-9.]+ \(([0-9.]+)\), lr=[0-9.]+: : 100%[^\n]*\n==> Finished Epoch", log)]
    sched = json.loads(re.search(r"\[run_reference_main\] schedules (\{.*\})", log).group(1))
    launches = int(re.search(r"C-ABI kernel launches: (\d+)", log).group(1))
    return epoch_means, sched, launches, log


@pytest.mark.gpu
def test_reference_mains_run_unchanged_on_the_drop_in_packages(cuda, tmp_path):
    if not _have_reference():
        pytest.skip("reference Python sources not staged (oracle/_ref/py)")
    cwd = str(tmp_path)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_synthetic_dataset.py"), "ds", "64", "12"], cwd=cwd,
                   check=True, capture_output=True, timeout=300)

    means, sched, launches, log = _run(["main_nerf.py", "ds", "--workspace", "synth", "--iters", "96"] + COMMON, cwd)
    assert len(means) == 8 and means[-1] < 0.5 * means[0], means
    assert launches > 500
    assert sched.get("NeRFRenderer.update_extra_state[train]:update_schedule=fused", 0) >= 5, sched
    assert not any("update_schedule=torch" in k for k in sched), sched
    assert sched.get("NeRFRenderer.run_cuda[eval]:schedule=fused", 0) >= 6, sched      # 3 eval + 3 test views
    assert not any("schedule=loop" in k for k in sched), sched
    assert sched.get("NeRFRenderer.run_cuda[train]:train_schedule=fused", 0) == 96, sched      # csrc/nerf_train.cu
    assert not any("train_schedule=torch" in k for k in sched), sched
    ckpts = os.listdir(os.path.join(cwd, "results", "synth", "version_1", "checkpoints"))
    assert any(c.endswith(".pth") for c in ckpts), ckpts

    means, sched, launches, log = _run(["main_palette.py", "ds", "results/synth", "--iters", "48", "--datatype", "blender"]
                                       + COMMON, cwd)
    assert len(means) == 4 and means[-1] < 0.75 * means[0], means
    assert launches > 200
    assert sched.get("PaletteRenderer.run_cuda[train]:train_schedule=fused", 0) == 48, sched
    assert not any("train_schedule=torch" in k for k in sched), sched
    assert sched.get("PaletteRenderer.run_cuda[eval]:schedule=fused", 0) >= 3, sched
    assert not any("schedule=loop" in k for k in sched), sched
