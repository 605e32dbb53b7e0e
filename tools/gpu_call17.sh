mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_cfg4_gpu.py -m gpu -q -x -s 2>&1 | tail -25) > gpurun_out/c20_cfg4.log
cat gpurun_out/c20_cfg4.log
