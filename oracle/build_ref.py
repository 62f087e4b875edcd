"""Build recipe for the GPU-side reference checker (TEST INFRASTRUCTURE, not product).

Compiles the five reference CUDA extensions *where they lie* under /root/reference
(raymarching/src, gridencoder/src, shencoder/src, freqencoder/src, palette/src) with
direct nvcc/g++ invocations (NOT the reference's setup.py / JIT build system) into
``oracle/_ref/_ref_<name>.so`` under renamed module names, so the reference kernels and
the new kernels can be loaded side by side in one process on the GPU box.

Only edits relative to the reference's own flags (raymarching/setup.py:7-13):
``-std=c++17`` (torch 2.11 headers need it) and an explicit sm_100a gencode.
No reference source is copied into this repository; outputs go only to oracle/_ref/
(git-ignored, NOT gpurun-ignored so the .so files travel to the GPU box).

Usage:  python oracle/build_ref.py [name ...]      (names: raymarching gridencoder shencoder freqencoder palette)
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF = os.environ.get("PNERF_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    # name: (src dir under REF, module name, extra nvcc flags)
    "raymarching": ("raymarching/src", "_ref_raymarching", []),
    "gridencoder": ("gridencoder/src", "_ref_gridencoder", []),
    "shencoder": ("shencoder/src", "_ref_shencoder", []),
    "freqencoder": ("freqencoder/src", "_ref_freqencoder", ["-use_fast_math"]),
    "palette": ("palette/src", "_ref_palette_func", ["-use_fast_math"]),
}


def _flags():
    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce
    import pybind11
    inc = ce.include_paths("cuda") + [sysconfig.get_paths()["include"], pybind11.get_include()]
    libdirs = ce.library_paths("cuda")
    return inc, libdirs


def build_one(name):
    srcdir, mod, extra = EXTS[name]
    srcdir = os.path.join(REF, srcdir)
    if not os.path.isdir(srcdir):
        print(f"[build_ref] {srcdir} not present; skipping {name}")
        return False
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, mod + ".so")
    srcs = sorted(os.path.join(srcdir, f) for f in os.listdir(srcdir) if f.endswith((".cu", ".cpp")))
    newest = max(os.path.getmtime(s) for s in srcs)
    if os.path.exists(out) and os.path.getmtime(out) > newest:
        print(f"[build_ref] {mod}.so up to date")
        return True
    inc, libdirs = _flags()
    common = ["-std=c++17", "-O3", f"-DTORCH_EXTENSION_NAME={mod}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              "-D_GLIBCXX_USE_CXX11_ABI=1"]
    for i in inc:
        common += ["-isystem", i]
    objs = []
    for s in srcs:
        o = os.path.join(OUT, f"{mod}_{os.path.basename(s)}.o")
        if s.endswith(".cu"):
            cmd = ["nvcc", "-c", s, "-o", o, "-gencode", "arch=compute_100a,code=sm_100a",
                   "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_HALF2_OPERATORS__",
                   "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC"] + extra + common
        else:
            cmd = ["g++", "-c", s, "-o", o, "-fPIC"] + common
        print("[build_ref]", " ".join(cmd[:6]), "...")
        subprocess.check_call(cmd)
        objs.append(o)
    link = ["g++", "-shared", "-o", out] + objs
    for d in libdirs:
        link += [f"-L{d}", f"-Wl,-rpath,{d}"]
    link += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    subprocess.check_call(link)
    for o in objs:
        os.remove(o)
    print(f"[build_ref] built {out}")
    return True


def main(names=None):
    names = names or list(EXTS)
    with ThreadPoolExecutor(max_workers=int(os.environ.get("PNERF_REF_JOBS", "3"))) as ex:
        return list(ex.map(build_one, names))


if __name__ == "__main__":
    main(sys.argv[1:] or None)
