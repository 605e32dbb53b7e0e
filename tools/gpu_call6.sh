mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/c9_tests.log
timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c9_rdm.log 2>&1
tail -5 gpurun_out/c9_tests.log; cat gpurun_out/c9_rdm.log
