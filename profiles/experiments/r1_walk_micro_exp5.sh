run() { tag=$1; shift
  make -C gadget-2.0.7-ngravs_b200 -j16 EXTRA="-DG2_FAST_BUILD $*" > gpurun_out/make_$tag.log 2>&1
  timeout 400 python bench.py --workload periodic256x4 --steps 3 --no-cpu-baseline > gpurun_out/bench4_$tag.json 2> gpurun_out/bench4_$tag.err
}
mkdir -p gpurun_out
run w9 -DWALK_MINBLOCKS_WIDE=9
run w10 -DWALK_MINBLOCKS_WIDE=10
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench4_w*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), round(d["stages_ms"]["walk_kernel_ms"],3), round(d["ia_per_particle"],3), d["e2e"]["checksum"])
    except Exception as e: print(f, "ERR", e)
PY
