mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/c32_tests.log
for i in 1 2; do
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c32_bench$i.json 2> gpurun_out/c32_bench$i.err
tail -n 1 gpurun_out/c32_bench$i.err; python - <<PY
import json
d=json.load(open('gpurun_out/c32_bench$i.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
PY
done
cat gpurun_out/c32_tests.log
