#!/bin/bash
# GPU-box visit: targeted parity tests + full bench (usage under gpurun: bash tools/gpu_visit2.sh <tag> [full])
TAG=${1:-v}
mkdir -p gpurun_out
if [ "$2" = "full" ]; then T="tests"; else T="tests/test_rays.py tests/test_encoders_gpu.py tests/test_golden.py tests/test_fused_gpu.py tests/test_render_gpu.py tests/test_fused_train_gpu.py"; fi
timeout 1500 python -m pytest $T -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err; head -c 400 gpurun_out/${TAG}_bench.json
