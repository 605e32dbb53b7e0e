mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/c40_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c40_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/c40_ref.json 2> gpurun_out/c40_ref.err
cat gpurun_out/c40_tests.log; tail -n 1 gpurun_out/c40_smoke.log; tail -n 2 gpurun_out/c40_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c40_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['rd_map_sets_per_sec'], d['config']['rd_map_sets_per_sec_rdm_kernels'])
r=json.loads([l for l in open('gpurun_out/c40_ref.json') if l.startswith('{')][0])
print('reference arm', r['value'], r['cpu_baseline']['cores'], r['config']['note'])
PY
grep real gpurun_out/c40_ref.err
