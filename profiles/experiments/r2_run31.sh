# round 2, GPU call 31: CTA size of walk_kernel (128 / 256 / 512 / 1024 threads, same resident warps per SM: fewer copies of the short-range table and
# less shared memory per SM) x shared-memory carveout (more L1 for cell records), with the per-SM chunk blocks of call 30
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for kt in 128 256 512 1024; do
  lib=$V/libg2gpu_kt$kt.so; [ $kt = 128 ] && lib=gadget-2.0.7-ngravs_b200/libg2gpu.so
  for cv in -1 15 30 45; do
    for wl in periodic256 hernquist1m; do
      G2GPU_LIB=$PWD/$lib G2GPU_WALK_CARVEOUT=$cv timeout 600 python bench.py --workload $wl --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench31_${wl}_kt${kt}_cv${cv}.json 2> gpurun_out/r2_bench31_${wl}_kt${kt}_cv${cv}.err
    done
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_bench31_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],3), "walk", round(d["stages_ms"]["walk_kernel_ms"],3), "ia", d["ia_per_particle"])
    except Exception as e: print(f, "ERR", e)
PY
