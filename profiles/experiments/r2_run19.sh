# round 2, GPU call 19: one-sweep pass with a windowed look-back (4 predecessor tiles per step, both bins of a thread in one 8-byte word)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_tree_walk.py -m gpu -q -x > gpurun_out/r2_gpu_tests_19.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_19.log
G2GPU_SORT_ITEMS=8 timeout 900 python -m pytest tests/test_gpu_stage1.py -m gpu -q -x > gpurun_out/r2_gpu_tests_19b.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_19b.log
for wl in periodic256 periodic128 hernquist1m; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench19_${wl}.json 2> gpurun_out/r2_bench19_${wl}.err
done
G2GPU_SORT_ITEMS=8 timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench19_periodic256_items8.json 2> gpurun_out/r2_bench19_periodic256_items8.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench19_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "sort frac", round(d.get("roofline_sort",{}).get("frac",0),3))
    except Exception as e: print(f, "ERR", e)
PY
