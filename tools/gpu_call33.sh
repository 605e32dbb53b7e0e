mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/c36_bench4.json 2> gpurun_out/c36_bench4.err
echo "rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/c36_bench4.err | tail -n 3; python - <<'PY'
import json
for l in open('gpurun_out/c36_bench4.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print({k:d.get(k) for k in ('value','n_gpus','ms_per_step','gpu_launches','scaling')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
PY
nproc; free -g | head -2
