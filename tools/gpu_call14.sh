mkdir -p gpurun_out
# 1) launch list of the bench command (cold-cache, serialised; shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/c17_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c17_ncu_bench.log 2>&1
# 2) full capture of the RDM kernels (current defaults: L2 hints + discard)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rdm_range4096_lean|rdm_doppler256" -s 4 -c 2 -o gpurun_out/c17_rdm_lean python tools/profile_rdm.py 1 0 > gpurun_out/c17_ncu.log 2>&1
# 3) warm-cache DRAM traffic of one chain of 4 map-sets
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none --csv --log-file gpurun_out/c17_warm.csv python tools/profile_rdm.py 4 0 > /dev/null 2>&1
# 4) RDM micro-bench + the sensing tests touched by the covariance change
timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c17_rdm.log 2>&1
(timeout 400 python -m pytest tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -8) > gpurun_out/c17_tests.log
ISAC_BENCH_DEBUG=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err
cat gpurun_out/c17_tests.log gpurun_out/c17_rdm.log; tail -n 2 gpurun_out/c17_bench.err; tail -n 3 gpurun_out/c17_ncu.log
