# round 2, GPU call 22: L1 carveout of the walk kernel (more of the 256 KB array as cache for cell records) and a prefetch of the sibling record
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  for wl in periodic256 hernquist1m; do
    env "$@" timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench22_${wl}_${name}.json 2> gpurun_out/r2_bench22_${wl}_${name}.err
  done
}
run base G2GPU_WALK_PREFETCH=0
run carve44 G2GPU_WALK_CARVEOUT=44
run carve30 G2GPU_WALK_CARVEOUT=30
run prefetch G2GPU_WALK_PREFETCH=1
run carve44_prefetch G2GPU_WALK_CARVEOUT=44 G2GPU_WALK_PREFETCH=1
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench22_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d.get("stages_ms",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
