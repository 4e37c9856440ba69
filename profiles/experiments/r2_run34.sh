# round 2, GPU call 34: record of the final kernels: ncu --set full of walk_kernel (per-SM chunk blocks, 1024-thread CTAs), launch list of whole steps,
# the other workloads, the default bench line with all legs, the reference arm, smoke
mkdir -p gpurun_out
for wl in periodic128 hernquist1m periodic256x4; do
  timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench34_${wl}.json 2> gpurun_out/r2_bench34_${wl}.err
done
( time timeout 1200 python bench.py > gpurun_out/r2_bench34_default.json 2> gpurun_out/r2_bench34_default.err ) 2> gpurun_out/r2_bench34_default.time; cat gpurun_out/r2_bench34_default.time; tail -3 gpurun_out/r2_bench34_default.err
( time timeout 1200 python bench.py --impl reference > gpurun_out/r2_bench34_reference.json 2> gpurun_out/r2_bench34_reference.err ) 2> gpurun_out/r2_bench34_reference.time; cat gpurun_out/r2_bench34_reference.time
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke34.log 2>&1; tail -2 gpurun_out/r2_smoke34.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench34_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()}, "pot", (d.get("potential_walk") or {}).get("ms_per_call"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "shim", (d.get("e2e_shim") or {}).get("ms_per_step"), "parity", d.get("parity"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches34.csv python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_launches34.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p256_c python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof34_ncu.log 2>&1
ls -la gpurun_out/r2_walk_p256_c.ncu-rep
