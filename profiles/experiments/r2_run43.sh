# round 2, GPU call 43: third batch: base3 = (call epilogue + ballot + relative-criterion instantiation), k = -DG2_WALK_KCONST (rcut, rcut^2, cull margin
# handed to the visit code through the lane struct instead of the constant bank), p = -DG2_POT_RELCRIT (relative-criterion instantiation of pot_kernel)
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for v in base3 k p base3 k; do
  G2GPU_LIB=$PWD/$V/libg2gpu_$v.so timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench43_$v.json 2> gpurun_out/r2_bench43_$v.err || tail -3 gpurun_out/r2_bench43_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench43_$v.json")); print("$v", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()}, d["ia_per_particle"], d["rewalked_targets"], "pot", d["potential_walk"]["ms_per_call"])
except Exception as e: print("$v", "ERR", e)
PY
done
G2GPU_LIB=$PWD/$V/libg2gpu_k.so timeout 900 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_43.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_43.log
G2GPU_LIB=$PWD/$V/libg2gpu_p.so timeout 900 python -m pytest tests/test_gpu_potential.py -m gpu -q -x > gpurun_out/r2_gpu_tests_43p.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_43p.log
