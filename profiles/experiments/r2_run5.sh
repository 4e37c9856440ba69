# round 2, GPU call 5: ncu --set full of the walk kernel (exact and FP32-only instantiations), periodic 128^3; BAM potential / group tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_potential.py -m gpu -q > gpurun_out/r2_gpu_tests_5.log 2>&1; tail -5 gpurun_out/r2_gpu_tests_5.log
timeout 600 python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof5_plain.json 2> gpurun_out/r2_prof5_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p128_exact python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof5_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p128_fp32 python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim --walk-exact 0 > gpurun_out/r2_prof5_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
