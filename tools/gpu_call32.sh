mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_sensing_gpu.py tests/test_golden_gpu.py tests/test_cfg4_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/c35_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c35_bench.json 2> gpurun_out/c35_bench.err
cat gpurun_out/c35_tests.log; tail -n 1 gpurun_out/c35_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c35_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_all']['covariance'])
PY
