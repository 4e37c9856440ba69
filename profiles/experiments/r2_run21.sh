# round 2, GPU call 21: full GPU suite + the default bench line (all legs), timed
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_21.log 2>&1 ) 2> gpurun_out/r2_gpu_tests_21.time; tail -8 gpurun_out/r2_gpu_tests_21.log; cat gpurun_out/r2_gpu_tests_21.time
( time timeout 1200 python bench.py > gpurun_out/r2_bench21_default.json 2> gpurun_out/r2_bench21_default.err ) 2> gpurun_out/r2_bench21_default.time; cat gpurun_out/r2_bench21_default.time; tail -3 gpurun_out/r2_bench21_default.err
( time timeout 1200 python bench.py --impl reference > gpurun_out/r2_bench21_reference.json 2> gpurun_out/r2_bench21_reference.err ) 2> gpurun_out/r2_bench21_reference.time; cat gpurun_out/r2_bench21_reference.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke21.log 2>&1; tail -2 gpurun_out/r2_smoke21.log
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench21_default.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["stages_ms"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["d2h_bytes_per_step"], "shim", (d.get("e2e_shim") or {}).get("ms_per_step"))
print("roofline", d["roofline"]["frac"], d["roofline"]["traffic"], "sort", d["roofline_sort"]["frac"], d["roofline_sort"]["traffic"])
print("parity", d.get("parity")); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r=json.load(open("gpurun_out/r2_bench21_reference.json")); print("ref", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], r["config"]==d["config"])
PY
