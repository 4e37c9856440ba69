# round 2, GPU call 39: final record on HEAD (zero-copy results): the whole GPU test suite, the default bench line with every leg, the reference arm,
# smoke, and the launch list of whole steps
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_39.log 2>&1 ) 2> gpurun_out/r2_gpu_tests_39.time; tail -4 gpurun_out/r2_gpu_tests_39.log
( time timeout 900 python bench.py --impl reference > gpurun_out/r2_bench39_reference.json 2> gpurun_out/r2_bench39_reference.err ) 2> gpurun_out/r2_bench39_reference.time
( time timeout 900 python bench.py > gpurun_out/r2_bench39_default.json 2> gpurun_out/r2_bench39_default.err ) 2> gpurun_out/r2_bench39_default.time; cat gpurun_out/r2_bench39_default.time; tail -3 gpurun_out/r2_bench39_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke39.log 2>&1; tail -2 gpurun_out/r2_smoke39.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench39_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "shim", (d.get("e2e_shim") or {}).get("ms_per_step"), "parity", d.get("parity"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches39.csv python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_launches39.log 2>&1
