# round 2, GPU call 42: second batch of walk-loop micro-variants on top of (call epilogue + ballot): u = -DG2_WALK_RELCRIT (an instantiation of the default
# kernel for the relative criterion alone: no run-time criterion switch per decision), c = -DG2_WALK_CELLS_REG (cell array base held in a register pair),
# l = -DG2_WALK_LATESLEEP (non-opening lanes put to sleep after the vote, only when the warp descends; spills at 64 registers)
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for v in base2 u c uc l base2 u; do
  G2GPU_LIB=$PWD/$V/libg2gpu_$v.so timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench42_$v.json 2> gpurun_out/r2_bench42_$v.err || tail -3 gpurun_out/r2_bench42_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench42_$v.json")); print("$v", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()}, d["ia_per_particle"], d["rewalked_targets"], d["first_pass_barnes_hut"])
except Exception as e: print("$v", "ERR", e)
PY
done
G2GPU_LIB=$PWD/$V/libg2gpu_u.so timeout 900 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_42.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_42.log
