# round 2, GPU call 7 (2 GPUs): the multi-GPU group inside libg2gpu.so -- bit-equality with one device, bench through the C path, torchrun launch
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2_gpu_tests_7.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_7.log
timeout 900 python bench.py --gpus 2 --steps 5 --no-cpu-baseline > gpurun_out/r2_bench7_n2.json 2> gpurun_out/r2_bench7_n2.err; tail -3 gpurun_out/r2_bench7_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench7_n2_torchrun.json 2> gpurun_out/r2_bench7_n2_torchrun.err; tail -3 gpurun_out/r2_bench7_n2_torchrun.err
timeout 600 python bench.py --gpus 1 --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench7_n1.json 2> gpurun_out/r2_bench7_n1.err
timeout 600 python bench.py --gpus 2 --steps 5 --no-cpu-baseline --no-shim --equal-slices > gpurun_out/r2_bench7_n2_equal.json 2> gpurun_out/r2_bench7_n2_equal.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench7_*.json")):
    try:
        d=json.load(open(f)); print(f, "n_gpus", d["n_gpus"], round(d["ms_per_step"],3), "%.3e"%d["value"], {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "shim", (d.get("e2e_shim") or {}).get("ms_per_step"), "slices", d.get("target_slices"), "allgather", d.get("allgather_bytes_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
