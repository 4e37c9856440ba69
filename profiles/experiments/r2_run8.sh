# round 2, GPU call 8: PM potential (API, drop-in) tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pm.py tests/test_gpu_dropin.py -m gpu -q -k "potential or pm" > gpurun_out/r2_gpu_tests_8.log 2>&1; tail -15 gpurun_out/r2_gpu_tests_8.log
