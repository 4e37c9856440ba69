# round 2, GPU call 41: micro-variants of the walk loop on top of the call-epilogue default (D = 2 unit only; 256^3, 5 steps each):
#   a = -DG2_WALK_NOSLEEPBR (sleeping lanes ride along as culled lanes, no branch around the visit), b = -DG2_WALK_BALLOT (openers as a ballot mask: no
#   byte-packed bool across the vote), d = -DG2_WALK_WRAPVOTE (warps near a box face take the image-free path for cells on their own side), combinations;
#   then the other workloads with the new default library, and the walk tests with the bd variant
mkdir -p gpurun_out
V=gadget-2.0.7-ngravs_b200/variants
for v in base b d bd a ab ad abd base bd; do
  G2GPU_LIB=$PWD/$V/libg2gpu_$v.so timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench41_$v.json 2> gpurun_out/r2_bench41_$v.err || tail -3 gpurun_out/r2_bench41_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench41_$v.json")); print("$v", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()}, d["ia_per_particle"], d["rewalked_targets"])
except Exception as e: print("$v", "ERR", e)
PY
done
for wl in periodic256x4 hernquist1m periodic128; do
  timeout 600 python bench.py --workload $wl --steps 5 --no-cpu-baseline --no-shim > gpurun_out/r2_bench41_wl_$wl.json 2> gpurun_out/r2_bench41_wl_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench41_wl_$wl.json")); print("$wl", round(d["ms_per_step"],3), {k:round(x,3) for k,x in d["stages_ms"].items()})
except Exception as e: print("$wl", "ERR", e)
PY
done
G2GPU_LIB=$PWD/$V/libg2gpu_bd.so timeout 900 python -m pytest tests/test_gpu_tree_walk.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2_gpu_tests_41.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_41.log
