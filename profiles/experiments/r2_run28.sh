# round 2, GPU call 28: guard-band tolerances as immediates, table clamp bound in a uniform register
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py tests/test_gpu_laws.py tests/test_gpu_group.py tests/test_gpu_lattice.py -m gpu -q -x > gpurun_out/r2_gpu_tests_28.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_28.log
for wl in periodic256 periodic128 hernquist1m periodic256x4; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench28_${wl}.json 2> gpurun_out/r2_bench28_${wl}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench28_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia", d["ia_per_particle"], "rewalked", d.get("rewalked_targets"))
    except Exception as e: print(f, "ERR", e)
PY
