#!/usr/bin/env python
"""Stage the reference's PYTHON modules where the GPU box can import them (TEST INFRASTRUCTURE).

/root/reference does not exist on the GPU box; its compiled kernels travel as oracle/_ref/_ref_*.so
(oracle/build_ref.py). This script does the same for the Python stack above them: it copies the unmodified
.py / .pyx files the volumetric-rendering path imports into ``oracle/_ref/py/`` (git-ignored like the .so files,
NOT gpurun-ignored, so it travels with the snapshot). Nothing is copied into the tracked tree.
``oracle/ref_python.py`` imports from there (or straight from /root/reference when that exists).

Usage: python oracle/stage_ref_py.py      (a no-op when /root/reference is absent)
"""
import os
import shutil
import sys

REF = os.environ.get("PNERF_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "py")

TOP_FILES = ["activation.py", "encoding.py", "loss.py", "main_palette.py", "main_nerf.py"]
PACKAGES = ["raymarching", "gridencoder", "shencoder", "freqencoder", "nerf", "palette"]
KEEP_EXT = (".py", ".pyx")
SKIP_DIRS = {"src", "__pycache__", "build"}


def main():
    if not os.path.isdir(REF):
        print(f"[stage_ref_py] {REF} not present; nothing staged")
        return False
    n = 0
    os.makedirs(OUT, exist_ok=True)
    for f in TOP_FILES:
        shutil.copy2(os.path.join(REF, f), os.path.join(OUT, f))
        n += 1
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(REF, pkg)):
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
            rel = os.path.relpath(root, REF)
            for f in files:
                if f.endswith(KEEP_EXT) and f != "setup.py":
                    os.makedirs(os.path.join(OUT, rel), exist_ok=True)
                    shutil.copy2(os.path.join(root, f), os.path.join(OUT, rel, f))
                    n += 1
    print(f"[stage_ref_py] staged {n} files under {OUT}")
    return True


if __name__ == "__main__":
    sys.exit(0 if main() is not None else 1)
