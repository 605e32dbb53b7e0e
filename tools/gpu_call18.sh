mkdir -p gpurun_out
(timeout 800 python -m pytest tests/test_cfg23_gpu.py -m gpu -q -x -s 2>&1 | tail -25) > gpurun_out/c21_cfg23.log
cat gpurun_out/c21_cfg23.log
