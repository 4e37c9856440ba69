# round 2, GPU call 20: look-back window 4 / 8 x tile 4096 / 2048 pairs; e2e with the direct (uncompacted) result path on one device
mkdir -p gpurun_out
G2GPU_SORT_ITEMS=8 G2GPU_SORT_WINDOW=8 timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_20.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_20.log
for it in 16 8; do for w in 4 8; do
  G2GPU_SORT_ITEMS=$it G2GPU_SORT_WINDOW=$w timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench20_p256_items${it}_w${w}.json 2> gpurun_out/r2_bench20_p256_items${it}_w${w}.err
  G2GPU_SORT_ITEMS=$it G2GPU_SORT_WINDOW=$w timeout 600 python bench.py --workload hernquist1m --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench20_h1m_items${it}_w${w}.json 2> gpurun_out/r2_bench20_h1m_items${it}_w${w}.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench20_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "sort frac", round(d.get("roofline_sort",{}).get("frac",0),3), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
