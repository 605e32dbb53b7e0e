mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py -m gpu -q -x -k "cfg2" -s 2>&1 | tail -25) > gpurun_out/c22_cfg2.log
cat gpurun_out/c22_cfg2.log
