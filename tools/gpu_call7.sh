mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/c10_tests.log
tail -25 gpurun_out/c10_tests.log
