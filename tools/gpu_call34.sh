mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/c37_tests.log
timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c37_rdm.log 2>&1
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c37_bench.json 2> gpurun_out/c37_bench.err
cat gpurun_out/c37_tests.log gpurun_out/c37_rdm.log; tail -n 1 gpurun_out/c37_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c37_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline'])
PY
