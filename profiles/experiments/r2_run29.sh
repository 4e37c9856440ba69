# round 2, GPU call 29: potential walk with the stock-wiring instantiation (tests + timings); ncu source counters of the current walk_kernel at 256^3
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_potential.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2_gpu_tests_29.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_29.log
for wl in periodic256 periodic128 hernquist1m; do
  timeout 600 python bench.py --workload $wl --steps 2 --no-cpu-baseline --no-shim > gpurun_out/r2_bench29_${wl}.json 2> gpurun_out/r2_bench29_${wl}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench29_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "walk", round(d["stages_ms"]["walk_kernel_ms"],3), "pot", (d.get("potential_walk") or {}).get("ms_per_call"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p256_b python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof29_ncu.log 2>&1
ls -la gpurun_out/r2_walk_p256_b.ncu-rep
