mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/c6_tests.log
ISAC_RDM_PDL=1 timeout 120 python tools/dev_rdm_bench.py 0 > gpurun_out/c6_rdm_pdl1.log 2>&1
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
tail -5 gpurun_out/c6_tests.log; cat gpurun_out/c6_rdm_pdl1.log gpurun_out/c6_bench.json
