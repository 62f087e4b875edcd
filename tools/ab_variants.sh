#!/bin/bash
# A/B of library variants built with PNERF_LIB_OUT / PNERF_EXTRA_NVCC_FLAGS (palettenerf_b200/build.py):
# usage (under gpurun): bash tools/ab_variants.sh <variant> [<variant> ...]   ("main" = the in-tree library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = main ]; then lib=palettenerf_b200/libpnerf_b200.so; else lib=palettenerf_b200/variants/lib_$v.so; fi
  PNERF_LIB=$PWD/$lib python bench.py --no-extras --steps 10 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().split("\n")[-1])
    print("$v", round(d["ms_per_step"],3), "ms", round(d["e2e"]["ms_per_step"],3), d["config"]["tile_fill"])
except Exception as e: print("$v failed", e)
P
done
