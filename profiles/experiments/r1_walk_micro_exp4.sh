# occupancy of the D >= 3 cursor walk (config 4: 6 particle types on 4 species), round 1
run() { tag=$1; shift
  make -C gadget-2.0.7-ngravs_b200 -j16 EXTRA="-DG2_FAST_BUILD $*" > gpurun_out/make_$tag.log 2>&1
  grep -A3 "walk_kernelILi4ELb1ELb1ELb1ELb1EdLi32" gpurun_out/make_$tag.log | grep -E "spill|registers" | head -2
  timeout 400 python bench.py --workload periodic256x4 --steps 3 --no-cpu-baseline > gpurun_out/bench4_$tag.json 2> gpurun_out/bench4_$tag.err
}
mkdir -p gpurun_out
run w8 -Xptxas -v
run w7 -Xptxas -v -DWALK_MINBLOCKS_WIDE=7
run w6 -Xptxas -v -DWALK_MINBLOCKS_WIDE=6
run w5 -Xptxas -v -DWALK_MINBLOCKS_WIDE=5
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench4_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), round(d["stages_ms"]["walk_kernel_ms"],3), round(d["ia_per_particle"],3), d["e2e"]["checksum"])
    except Exception as e: print(f, "ERR", e)
PY
