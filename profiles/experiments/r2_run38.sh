# round 2, GPU call 38: zero-copy results into the reference's P[] (g2gpu_group_bind_results_aos, shim gravity_tree()): drop-in / whole-program / group tests,
# e2e_shim with and without it
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_fullrun.py tests/test_gpu_group.py tests/test_gpu_lattice.py -m gpu -q -x > gpurun_out/r2_gpu_tests_38.log 2>&1; tail -5 gpurun_out/r2_gpu_tests_38.log
for zc in 1 0; do
  G2GPU_ZERO_COPY=$zc timeout 900 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_bench38_zc${zc}.json 2> gpurun_out/r2_bench38_zc${zc}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench38_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "shim", d.get("e2e_shim"))
    except Exception as e: print(f, "ERR", e)
PY
