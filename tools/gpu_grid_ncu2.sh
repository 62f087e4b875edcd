#!/bin/bash
# ncu capture of the stand-alone fp16 hash-grid forward kernel at BASELINE config 2 (2^22 random points)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_grid_fwd_coop_h -s 6 -c 1 -o gpurun_out/r02_grid_fwd_coop -f \
  python bench.py --sections hashgrid --steps 1 > gpurun_out/r02_grid_ncu.log 2>&1
echo "ncu rc=$?"
