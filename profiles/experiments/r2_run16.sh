# round 2, GPU call 16: ncu --set full of one os_pass_kernel launch (one-sweep radix pass) at 16.8 M pairs
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:os_pass_kernel -s 16 -c 1 -o gpurun_out/r2_ospass_p256 python bench.py --profile --steps 1 --no-cpu-baseline --no-shim > gpurun_out/r2_prof16_ncu.log 2>&1
ls -la gpurun_out/r2_ospass_p256.ncu-rep
