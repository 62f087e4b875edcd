#!/bin/bash
# ncu --set full captures of the stand-alone hash-grid kernels (BASELINE config 2: 2^22 points, fp16 table)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_fwd_d3c2 -s 3 -c 1 -o gpurun_out/grid_fwd -f \
  python bench.py --steps 1 --warmup 3 --sections hashgrid > gpurun_out/grid_fwd.log 2>&1; echo "fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_grid_bwd_runs -s 3 -c 1 -o gpurun_out/grid_bwd -f \
  python bench.py --steps 1 --warmup 3 --sections hashgrid > gpurun_out/grid_bwd.log 2>&1; echo "bwd rc=$?"
ls -la gpurun_out | grep grid_
