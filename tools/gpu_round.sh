#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the launch list and one full ncu capture of the render kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-run}
MODE=${2:-tests}
mkdir -p gpurun_out
if [ "$MODE" = "tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; head -c 600 gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 400 --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_render -s 3 -c 1 \
  -o gpurun_out/${TAG}_render -f python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/${TAG}_ncu_render.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out | tail -12
