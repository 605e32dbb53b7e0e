mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/c14_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c14_bench_own.json 2> gpurun_out/c14_bench_own.err
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline --sense-ctx shared > gpurun_out/c14_bench_shared.json 2> gpurun_out/c14_bench_shared.err
ISAC_RDM_PDL=0 ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c14_bench_own_pdl0.json 2> gpurun_out/c14_bench_own_pdl0.err
tail -5 gpurun_out/c14_tests.log; tail -3 gpurun_out/c14_bench_*.err
