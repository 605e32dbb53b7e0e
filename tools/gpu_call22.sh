mkdir -p gpurun_out
for m in 2 4; do
ISAC_PAIR_MINB=$m ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c25_bench_minb$m.json 2> gpurun_out/c25_bench_minb$m.err
echo "minb=$m"; tail -n 1 gpurun_out/c25_bench_minb$m.err
done
