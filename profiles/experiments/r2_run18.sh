# round 2, GPU call 18: ranking by warp votes per digit bit instead of MATCH.ANY in the one-sweep pass
mkdir -p gpurun_out
G2GPU_SORT_RANK_BALLOT=1 timeout 900 python -m pytest tests/test_gpu_stage1.py tests/test_gpu_tree_walk.py -m gpu -q -x > gpurun_out/r2_gpu_tests_18.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_18.log
for rb in 0 1; do
  G2GPU_SORT_RANK_BALLOT=$rb timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench18_p256_ballot${rb}.json 2> gpurun_out/r2_bench18_p256_ballot${rb}.err
  G2GPU_SORT_RANK_BALLOT=$rb timeout 600 python bench.py --workload hernquist1m --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench18_h1m_ballot${rb}.json 2> gpurun_out/r2_bench18_h1m_ballot${rb}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench18_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "sort frac", round(d.get("roofline_sort",{}).get("frac",0),3))
    except Exception as e: print(f, "ERR", e)
PY
