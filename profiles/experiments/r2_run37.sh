# round 2, GPU call 37 (2 GPUs): zero-copy results (the walk kernel stores into the caller's pinned arrays) against the staged download; group tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_37.log 2>&1; tail -5 gpurun_out/r2_gpu_tests_37.log
for zc in 1 0; do
for n in 1 2; do
  G2GPU_ZERO_COPY=$zc timeout 600 python bench.py --gpus $n --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench37_n${n}_zc${zc}.json 2> gpurun_out/r2_bench37_n${n}_zc${zc}.err
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench37_*.json")):
    try:
        d=json.load(open(f)); print(f, "n_gpus", d["n_gpus"], round(d["ms_per_step"],3), "walk", round(d["stages_ms"]["walk_kernel_ms"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"], d["e2e"]["results"][:30], "checksum", d["e2e"]["checksum"], "slices", d.get("target_slices"))
    except Exception as e: print(f, "ERR", e)
PY
