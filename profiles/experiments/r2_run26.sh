# round 2, GPU call 26: bisect of the non-periodic regression of call 25 (library built with -DG2_SPL_HOIST=0)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tree_walk.py -m gpu -q -x -k nonperiodic > gpurun_out/r2_gpu_tests_26.log 2>&1; tail -3 gpurun_out/r2_gpu_tests_26.log
timeout 600 python bench.py --workload hernquist1m --steps 3 --no-cpu-baseline --no-shim > gpurun_out/r2_bench26_hernquist1m.json 2> gpurun_out/r2_bench26_hernquist1m.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench26_hernquist1m.json')); print(d['ms_per_step'], d['ia_per_particle'], d['stages_ms'])"
