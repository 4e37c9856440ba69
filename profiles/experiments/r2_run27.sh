# round 2, GPU call 27: full GPU suite and the four workloads after the instruction-count work on walk_kernel (constants held in registers, float clamp,
# one softening test per visit for TreePM cell terms, flush thinning) 
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_27.log 2>&1; tail -5 gpurun_out/r2_gpu_tests_27.log
for wl in periodic256 periodic128 hernquist1m periodic256x4; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench27_${wl}.json 2> gpurun_out/r2_bench27_${wl}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench27_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia", d["ia_per_particle"], "rewalked", d.get("rewalked_targets"), "pot", (d.get("potential_walk") or {}).get("ms_per_call"))
    except Exception as e: print(f, "ERR", e)
PY
