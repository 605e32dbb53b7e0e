mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_cfg23_gpu.py tests/test_golden_gpu.py tests/test_chest_gpu.py -m gpu -q -x 2>&1 | tail -8) > gpurun_out/c29_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c29_bench.json 2> gpurun_out/c29_bench.err
cat gpurun_out/c29_tests.log; tail -n 2 gpurun_out/c29_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c29_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value']); print(d['roofline']['frac'], d['roofline_all']['pmi_sinr'])
PY
