mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/c7_tests.log
timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c7_rdm.log 2>&1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rdm_range4096_lean|rdm_doppler256" -s 4 -c 2 -o gpurun_out/c7_rdm_lean python tools/profile_rdm.py 1 0 > gpurun_out/c7_ncu.log 2>&1
tail -5 gpurun_out/c7_tests.log; cat gpurun_out/c7_rdm.log gpurun_out/c7_bench.json
