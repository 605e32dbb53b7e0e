mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
tail -5 gpurun_out/c12_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['roofline']); print({k:(v.get('avg_launch_us'),v.get('share_of_step')) for k,v in d['roofline_all'].items()})
PY
