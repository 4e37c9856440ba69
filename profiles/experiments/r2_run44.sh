# round 2, GPU call 44: m = -DG2_WALK_NOMASSBR (stock TreePM term path: no branch around the term of a species without mass -- its term is exactly zero)
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for v in base4 m m; do
  G2GPU_LIB=$PWD/$V/libg2gpu_$v.so timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench44_$v.json 2> gpurun_out/r2_bench44_$v.err || tail -3 gpurun_out/r2_bench44_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench44_$v.json")); print("$v", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()}, d["ia_per_particle"], d["rewalked_targets"], d["e2e"]["checksum"])
except Exception as e: print("$v", "ERR", e)
PY
done
