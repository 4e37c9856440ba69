# round-1 final measurements: full GPU test suite, then every bench workload (JSON lines under gpurun_out/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_final.log 2>&1; tail -4 gpurun_out/gpu_tests_final.log
timeout 900 python bench.py --steps 5 > gpurun_out/bench_periodic256_r1e.json 2> gpurun_out/bench_periodic256_r1e.err
for wl in periodic256x4 periodic128 hernquist1m; do
  timeout 900 python bench.py --workload $wl --steps 5 > gpurun_out/bench_${wl}_r1e.json 2> gpurun_out/bench_${wl}_r1e.err
done
timeout 600 python bench.py --active-frac 0.25 --steps 5 --no-cpu-baseline > gpurun_out/bench_periodic256_active25_r1e.json 2> gpurun_out/bench_active25_r1e.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm_r1e.json 2> gpurun_out/bench_reference_arm_r1e.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_r1e.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "%.3e"%d.get("interactions_per_s",0), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("pm_long_range") or {}).get("ms_per_call"), d.get("roofline",{}).get("frac"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
