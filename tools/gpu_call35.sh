mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_properties_gpu.py -m gpu -q -x -s 2>&1 | tail -25) > gpurun_out/c38_props.log
cat gpurun_out/c38_props.log
