# round 2, GPU call 10: ncu --set full of the deferred-ring walk kernel, periodic 128^3 (compare with r2_walk_p128_exact of call 5)
mkdir -p gpurun_out
timeout 600 python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof10_plain.json 2> gpurun_out/r2_prof10_plain.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p128_defer python bench.py --profile --workload periodic128 --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof10_ncu1.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
