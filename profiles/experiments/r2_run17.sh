# round 2, GPU call 17: one-sweep pass with 2048-pair tiles (8 pairs per thread; 5 or 4 CTAs per SM) against 4096-pair tiles (16 per thread, 3 CTAs)
mkdir -p gpurun_out
for it in 8 84; do
  G2GPU_SORT_ITEMS=$it timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_tree_walk.py -m gpu -q -x > gpurun_out/r2_gpu_tests_17_items${it}.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_17_items${it}.log
done
for wl in periodic256 hernquist1m; do for it in 16 8 84; do
  G2GPU_SORT_ITEMS=$it timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench17_${wl}_items${it}.json 2> gpurun_out/r2_bench17_${wl}_items${it}.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench17_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "sort frac", round(d.get("roofline_sort",{}).get("frac",0),3), "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),2))
    except Exception as e: print(f, "ERR", e)
PY
