# round 2, GPU call 40: does the inlined result epilogue (zero-copy / AoS / compact branches, added after the 172.9 ms capture) perturb the walk loop's
# code generation?  default library vs. the epilogue as a real call (-DG2_STORE_NOINLINE, kernel argument __grid_constant__) vs. __grid_constant__ alone
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for v in default noinline gridconst noinline default; do
  lib=$V/libg2gpu_$v.so; [ $v = default ] && lib=gadget-2.0.7-ngravs_b200/libg2gpu.so
  G2GPU_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench40_$v.json 2> gpurun_out/r2_bench40_$v.err || tail -3 gpurun_out/r2_bench40_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench40_$v.json")); print("$v", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()})
except Exception as e: print("$v", "ERR", e)
PY
done
G2GPU_LIB=$PWD/$V/libg2gpu_noinline.so timeout 900 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_40.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_40.log
