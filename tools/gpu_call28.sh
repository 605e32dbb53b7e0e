mkdir -p gpurun_out
: > gpurun_out/c31_direct.log
for v in 0 2 3 4; do
  echo "== ISAC_RDM_DIRECT=$v" >> gpurun_out/c31_direct.log
  ISAC_RDM_DIRECT=$v timeout 120 python tools/dev_rdm_bench.py 0 2>&1 | grep "3276\|rror" >> gpurun_out/c31_direct.log
done
(ISAC_RDM_DIRECT=3 timeout 300 python -m pytest tests/test_rdm_gpu.py -m gpu -q -x 2>&1 | tail -4) >> gpurun_out/c31_direct.log
cat gpurun_out/c31_direct.log
