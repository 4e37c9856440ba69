# round 2, GPU call 36: the complete GPU suite on the final code
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_36.log 2>&1 ) 2> gpurun_out/r2_gpu_tests_36.time; tail -8 gpurun_out/r2_gpu_tests_36.log; cat gpurun_out/r2_gpu_tests_36.time
