mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/c34_misc.csv python tools/profile_misc.py > gpurun_out/c34_misc.log 2>&1
tail -n 3 gpurun_out/c34_misc.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/c34_bench.json 2> gpurun_out/c34_bench.err; tail -n 2 gpurun_out/c34_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c34_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline'])
PY
