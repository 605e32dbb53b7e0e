mkdir -p gpurun_out
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err
tail -n 2 gpurun_out/c23_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c23_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']); print(d['roofline'])
PY
