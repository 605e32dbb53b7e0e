mkdir -p gpurun_out
: > gpurun_out/c33_early.log
for h in 0x149 0x349; do
  echo "== ISAC_RDM_HINTS=$h" >> gpurun_out/c33_early.log
  ISAC_RDM_HINTS=$h timeout 120 python tools/dev_rdm_bench.py 0 2>&1 | grep "3276\|rror" >> gpurun_out/c33_early.log
done
(ISAC_RDM_HINTS=0x349 timeout 300 python -m pytest tests/test_rdm_gpu.py -m gpu -q -x 2>&1 | tail -4) >> gpurun_out/c33_early.log
cat gpurun_out/c33_early.log
