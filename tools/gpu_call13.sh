mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_chest_gpu.py tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -25) > gpurun_out/c16_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err
cat gpurun_out/c16_tests.log; tail -n 3 gpurun_out/c16_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c16_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value']); print(d['roofline'])
PY
