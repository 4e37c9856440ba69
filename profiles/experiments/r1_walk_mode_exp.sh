run() { tag=$1; shift
  for wl in hernquist1m periodic128 periodic256; do
    timeout 400 python bench.py --workload $wl --walk-mode 1 --steps 3 --no-cpu-baseline > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err
  done; }
timeout 300 python -m pytest tests/test_gpu_walk_modes.py -x -q -s > gpurun_out/modes_test.log 2>&1; tail -12 gpurun_out/modes_test.log
run mb4
make -C gadget-2.0.7-ngravs_b200 -j16 EXTRA="-DG2_FAST_BUILD -DWB_MINBLOCKS=3" > gpurun_out/make_mb3.log 2>&1
run mb3
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*_mb*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, round(d["ia_per_particle"],2), d["cell_visits_per_warp_step"], round(d["roofline"]["frac"],4), d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
