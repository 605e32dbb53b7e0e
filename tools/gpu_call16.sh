mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c19_bench2.json 2> gpurun_out/c19_bench2.err
echo "rc=$?"; tail -n 5 gpurun_out/c19_bench2.err; python - <<'PY'
import json
for l in open('gpurun_out/c19_bench2.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','gpu_launches','scaling')}, d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline'))
PY
