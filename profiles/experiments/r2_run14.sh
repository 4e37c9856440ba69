# round 2, GPU call 14: ncu --set full + source counters of walk_kernel at periodic 256^3 (the bench workload)
mkdir -p gpurun_out
timeout 600 python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof14_plain.json 2> gpurun_out/r2_prof14_plain.err &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:walk_kernel -s 2 -c 1 -o gpurun_out/r2_walk_p256 python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof14_ncu.log 2>&1
ls -la gpurun_out/r2_walk_p256.ncu-rep
