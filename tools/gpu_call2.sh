mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/c4_tests.log
timeout 120 python tools/dev_rdm_bench.py > gpurun_out/c4_rdm.log 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_rdm_launches.csv python tools/profile_rdm.py 2 0 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rdm_range4096_lean|rdm_doppler256" -s 4 -c 2 -o gpurun_out/c4_rdm_lean python tools/profile_rdm.py 1 0 > gpurun_out/c4_ncu.log 2>&1
tail -5 gpurun_out/c4_tests.log; cat gpurun_out/c4_rdm.log
