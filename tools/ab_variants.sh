python -m pytest tests/test_fused_gpu.py tests/test_render_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/ab_pytest.log
for v in main w12_lv2 w16_lv1 w16_lv2; do
  if [ $v = main ]; then lib=palettenerf_b200/libpnerf_b200.so; else lib=palettenerf_b200/variants/lib_$v.so; fi
  PNERF_LIB=$PWD/$lib python bench.py --no-extras --steps 5 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/ab_$v.json").read().strip().split("\n")[-1])
    print("$v", round(d["ms_per_step"],3), "ms", round(d["e2e"]["ms_per_step"],3), d["config"]["tile_fill"])
except Exception as e: print("$v failed", e)
P
done
cat gpurun_out/ab_pytest.log
