mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/c18_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
cat gpurun_out/c18_tests.log; tail -n 2 gpurun_out/c18_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c18_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value']); print(d['roofline'])
PY
