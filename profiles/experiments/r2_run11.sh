# round 2, GPU call 11: new potential tests; launch list of one 256^3 step (all kernels, incl. walk_redo_kernel and the build); ncu of walk_redo_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_potential.py -m gpu -q -k "periodic" > gpurun_out/r2_gpu_tests_11.log 2>&1; tail -6 gpurun_out/r2_gpu_tests_11.log
timeout 600 python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof11_plain.json 2> gpurun_out/r2_prof11_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches11_p256.csv python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof11_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_redo_kernel -s 2 -c 1 -o gpurun_out/r2_redo_p128 python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof11_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
