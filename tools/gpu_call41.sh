mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_multipanel_gpu.py -m gpu -q -s 2>&1 | tail -30) > gpurun_out/c44_mp.log
(timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_multipanel_gpu.py 2>&1 | tail -8) > gpurun_out/c44_tests.log
cat gpurun_out/c44_mp.log; cat gpurun_out/c44_tests.log
