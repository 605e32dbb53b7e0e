mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_geometry_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/c13_tests.log
tail -25 gpurun_out/c13_tests.log
