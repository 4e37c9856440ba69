# round 2, GPU call 6: count-only guard bands + cooperative FP64 re-walk of all flagged targets; whole-visit WRAP split; one-sweep sort
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_6.log 2>&1; tail -12 gpurun_out/r2_gpu_tests_6.log
for ex in 1 0; do
  timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim --walk-exact $ex > gpurun_out/r2_bench6_periodic256_ex${ex}.json 2> gpurun_out/r2_bench6_periodic256_ex${ex}.err
done
G2GPU_SORT_ONESWEEP=0 timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench6_periodic256_oldsort.json 2> gpurun_out/r2_bench6_periodic256_oldsort.err
for wl in hernquist1m periodic128; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench6_${wl}.json 2> gpurun_out/r2_bench6_${wl}.err
  G2GPU_SORT_ONESWEEP=0 timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench6_${wl}_oldsort.json 2> gpurun_out/r2_bench6_${wl}_oldsort.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench6_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia/part %.1f"%d["ia_per_particle"], "flagged", d.get("fp64_checked_comparisons"), "rewalked", d.get("rewalked_targets"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "launches", d.get("gpu_launches"), round(d.get("roofline",{}).get("frac"),4))
    except Exception as e: print(f, "ERR", e)
PY
