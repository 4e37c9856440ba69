# round 2, GPU call 33: chunk counters cleared right before the walk launch (the lattice walk uses the same counters); lattice walk with per-SM chunk blocks
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_lattice.py tests/test_gpu_potential.py tests/test_gpu_dropin.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r2_gpu_tests_33.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_33.log
for sl in 1 0; do
for wl in periodic128nopm; do
  G2GPU_WALK_SM_LOCAL=$sl timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench33_${wl}_sl${sl}.json 2> gpurun_out/r2_bench33_${wl}_sl${sl}.err
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench33_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "pot", (d.get("potential_walk") or {}).get("ms_per_call"), d.get("lattice_correction"))
    except Exception as e: print(f, "ERR", e)
PY
