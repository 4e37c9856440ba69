# sibling-record prefetch (prefetch.global.L1, no destination register) in the cursor walk (round 1)
run() { tag=$1; shift
  make -C gadget-2.0.7-ngravs_b200 -j16 EXTRA="-DG2_FAST_BUILD $*" > gpurun_out/make_$tag.log 2>&1
  for wl in hernquist1m periodic256; do
    timeout 400 python bench.py --workload $wl --steps 3 --no-cpu-baseline > gpurun_out/bench3_${wl}_$tag.json 2> gpurun_out/bench3_${wl}_$tag.err
  done; }
mkdir -p gpurun_out
run base
run prefetch -DG2_WALK_PREFETCH
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench3_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), round(d["stages_ms"]["walk_kernel_ms"],3), round(d["ia_per_particle"],3), d["e2e"]["checksum"])
    except Exception as e: print(f, "ERR", e)
PY
