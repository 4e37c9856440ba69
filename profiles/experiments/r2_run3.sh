# round 2, GPU call 3: verify + cooperative FP64 re-walk; full-size parity tests; bench through the group path with e2e, e2e_shim, cpu_baseline, parity
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_3.log 2>&1; tail -40 gpurun_out/r2_gpu_tests_3.log
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench3_periodic256.json 2> gpurun_out/r2_bench3_periodic256.err; tail -3 gpurun_out/r2_bench3_periodic256.err
for wl in hernquist1m periodic128; do
  timeout 600 python bench.py --workload $wl --steps 5 --no-cpu-baseline > gpurun_out/r2_bench3_${wl}.json 2> gpurun_out/r2_bench3_${wl}.err; tail -2 gpurun_out/r2_bench3_${wl}.err
done
timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim --walk-exact 0 > gpurun_out/r2_bench3_periodic256_ex0.json 2> gpurun_out/r2_bench3_periodic256_ex0.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench3_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia/part %.1f"%d["ia_per_particle"], "checked", d.get("fp64_checked_comparisons"), "rewalked", d.get("rewalked_targets"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "shim", (d.get("e2e_shim") or {}), "parity", d.get("parity"), d.get("roofline",{}).get("frac"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
