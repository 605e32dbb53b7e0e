mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_mex_mock_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/c45_mex.log
cat gpurun_out/c45_mex.log
