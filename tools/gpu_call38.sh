mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_mex_mock_gpu.py tests/test_mex_mock_cpu.py -q -x 2>&1 | tail -30) > gpurun_out/c41_mex.log
cat gpurun_out/c41_mex.log
