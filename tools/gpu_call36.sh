mkdir -p gpurun_out
for i in 1 2; do
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c39_bench$i.json 2> gpurun_out/c39_bench$i.err
tail -n 1 gpurun_out/c39_bench$i.err; python - <<PY
import json
d=json.load(open('gpurun_out/c39_bench$i.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
PY
done
