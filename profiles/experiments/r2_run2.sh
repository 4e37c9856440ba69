# round 2, GPU call 2: tightened guard bands + group path; exact vs fast kernel; rewalk counts
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_2.log 2>&1; tail -25 gpurun_out/r2_gpu_tests_2.log
for wl in periodic256 hernquist1m periodic128; do
  for ex in 1 0; do
    timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --walk-exact $ex > gpurun_out/r2_bench2_${wl}_ex${ex}.json 2> gpurun_out/r2_bench2_${wl}_ex${ex}.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench2_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia/part %.1f"%d["ia_per_particle"], "rewalked", d.get("rewalked_targets"), (d.get("e2e") or {}).get("ms_per_step"), d.get("roofline",{}).get("frac"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
