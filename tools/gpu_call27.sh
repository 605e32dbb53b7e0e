mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_rdm_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/c30_tests.log
timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c30_rdm.log 2>&1
cat gpurun_out/c30_tests.log gpurun_out/c30_rdm.log
