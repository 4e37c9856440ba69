# round 2, GPU call 35 (8 GPUs): strong scaling of the default workload through the C multi-GPU group with the final kernels; multi-GPU group tests
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2 1; do
  timeout 600 python bench.py --gpus $n --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_scale35_n${n}.json 2> gpurun_out/r2_scale35_n${n}.err
done
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2_gpu_tests_35.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_35.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_scale35_*.json")):
    try:
        d=json.load(open(f)); print(f, "n_gpus", d["n_gpus"], round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", round((d.get("e2e") or {}).get("ms_per_step",0),2), "slices", d.get("target_slices"))
    except Exception as e: print(f, "ERR", e)
PY
