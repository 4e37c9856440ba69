# round 2, GPU call 9: deferred-term ring walk (walk_defer) -- parity suites with it on (default), A/B timing against the in-place kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_edge_cases.py tests/test_gpu_laws.py tests/test_gpu_group.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2_gpu_tests_9.log 2>&1; tail -8 gpurun_out/r2_gpu_tests_9.log
for wl in periodic256 periodic128 hernquist1m; do
  for df in 1 0; do
    G2GPU_WALK_DEFER=$df timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench9_${wl}_defer${df}.json 2> gpurun_out/r2_bench9_${wl}_defer${df}.err
  done
done
G2GPU_WALK_DEFER=1 timeout 600 python bench.py --workload periodic256 --steps 3 --no-cpu-baseline --no-shim --walk-exact 0 > gpurun_out/r2_bench9_periodic256_defer1_ex0.json 2> gpurun_out/r2_bench9_periodic256_defer1_ex0.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench9_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "ia/part %.1f"%d["ia_per_particle"], "rewalked", d.get("rewalked_targets"), "parity", (d.get("parity") or {}).get("median"), (d.get("parity") or {}).get("cost_mismatch"), round(d.get("roofline",{}).get("frac"),4))
    except Exception as e: print(f, "ERR", e)
PY
