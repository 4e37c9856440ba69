# round 2, GPU call 4: where does the exact mode's time go?  per-launch times of the walk kernels (ncu, time only), 256^3 and Hernquist
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_potential.py -m gpu -q > gpurun_out/r2_gpu_tests_4.log 2>&1; tail -15 gpurun_out/r2_gpu_tests_4.log
timeout 600 python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof4_plain.json 2> gpurun_out/r2_prof4_plain.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:walk --csv --log-file gpurun_out/r2_launches4_walk_p256.csv python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof4_ncu.log 2>&1
grep -E "walk" gpurun_out/r2_launches4_walk_p256.csv | awk -F'","' '{print $5, $NF}' | head -40
