#!/bin/bash
# Row b2: the reference's UNMODIFIED main_nerf.py and main_palette.py on the drop-in packages, on a tiny synthetic
# Blender-format dataset (stage 1 for a few epochs, then the palette stage from its checkpoint). Logs -> gpurun_out/.
set -u
REPO=$(pwd)
mkdir -p gpurun_out
WORK=$(mktemp -d)
cd "$WORK"
python "$REPO/tools/make_synthetic_dataset.py" ds 64 12 > "$REPO/gpurun_out/ref_main_dataset.log" 2>&1
timeout 600 python "$REPO/tools/run_reference_main.py" main_nerf.py ds --workspace synth -O --bound 2 --scale 0.8 --dt_gamma 0 \
  --iters ${NERF_ITERS:-96} --num_rays 1024 > "$REPO/gpurun_out/ref_main_nerf.log" 2>&1
echo "main_nerf rc=$?"
grep -E "loss=|Epoch|PSNR|finished on the drop-in|Error|Traceback" "$REPO/gpurun_out/ref_main_nerf.log" | tail -12
timeout 600 python "$REPO/tools/run_reference_main.py" main_palette.py ds results/synth -O --bound 2 --scale 0.8 --dt_gamma 0 \
  --iters ${PALETTE_ITERS:-48} --num_rays 1024 --datatype blender > "$REPO/gpurun_out/ref_main_palette.log" 2>&1
echo "main_palette rc=$?"
grep -E "loss=|Epoch|PSNR|finished on the drop-in|Error|Traceback" "$REPO/gpurun_out/ref_main_palette.log" | tail -12
