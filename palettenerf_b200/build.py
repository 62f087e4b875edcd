"""Build libpnerf_b200.so (in-tree) with nvcc for sm_100a.

`python -m palettenerf_b200.build` compiles every csrc/*.cu into an object (in parallel, only when stale) and
links them into palettenerf_b200/libpnerf_b200.so — a plain C-ABI CUDA library with no torch/pybind dependency.
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.environ.get("PNERF_OBJ_DIR", os.path.join(PKG, "build"))
LIB = os.environ.get("PNERF_LIB_OUT", os.path.join(PKG, "libpnerf_b200.so"))   # variant builds for A/B experiments
EXTRA = os.environ.get("PNERF_EXTRA_NVCC_FLAGS", "").split()
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-fvisibility=hidden",
         "-Xptxas", "-v"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(PKG, "..", "include", "pnerf_b200.h"))
    return hs


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if not _stale(obj, [src] + _headers()):
        return obj, ""
    cmd = [NVCC, "-c", src, "-o", obj] + ARCH + FLAGS + EXTRA
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(_compile, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
