mkdir -p gpurun_out
: > gpurun_out/c15_hints.log
for h in 0 0x40 0x41 0x49 0x59 0x100 0x149 0x159 0x4a; do
  echo "== ISAC_RDM_HINTS=$h" >> gpurun_out/c15_hints.log
  ISAC_RDM_HINTS=$h timeout 120 python tools/dev_rdm_bench.py 0 2>&1 | grep "B=4 3276\|B=16 3276\|Error\|error" >> gpurun_out/c15_hints.log
done
(ISAC_RDM_HINTS=0x159 timeout 300 python -m pytest tests/test_rdm_gpu.py -m gpu -q 2>&1 | tail -5) > gpurun_out/c15_tests_discard.log
cat gpurun_out/c15_hints.log gpurun_out/c15_tests_discard.log
