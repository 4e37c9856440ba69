# round 2, GPU call 45: final record on HEAD (walk kernel 161 ms): the whole GPU test suite, the default bench line with every leg, smoke, ncu --set full of walk_kernel
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -m gpu -q -x > gpurun_out/r2_gpu_tests_45.log 2>&1 ) 2> gpurun_out/r2_gpu_tests_45.time; tail -3 gpurun_out/r2_gpu_tests_45.log
( time timeout 120 python bench.py > gpurun_out/r2_bench45_default.json 2> gpurun_out/r2_bench45_default.err ) 2> gpurun_out/r2_bench45_default.time; tail -2 gpurun_out/r2_bench45_default.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench45_default.json")); print(round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), "shim", (d.get("e2e_shim") or {}).get("ms_per_step"), "parity", d.get("parity"), "roofline", d["roofline"]["frac"], "pot", d["potential_walk"]["ms_per_call"])
except Exception as e: print("ERR", e)
PY
timeout 100 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p256_d python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof45_ncu.log 2>&1
ls -la gpurun_out/r2_walk_p256_d.ncu-rep
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke45.log 2>&1; tail -1 gpurun_out/r2_smoke45.log
