mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_sensing_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/c11_tests.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
tail -8 gpurun_out/c11_tests.log; tail -3 gpurun_out/c11_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c11_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print(d['config'])
PY
