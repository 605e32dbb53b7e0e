#!/bin/bash
# One entry point for the round-2 GPU calls: tools/gpu_r2.sh <step> ; outputs under gpurun_out/r2_<step>*
set -u
mkdir -p gpurun_out
step=$1
case $step in
pmi1)
  (timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_cfg23_gpu.py tests/test_golden_gpu.py tests/test_chest_gpu.py tests/test_mex_mock_gpu.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2_pmi1_tests.log
  cat gpurun_out/r2_pmi1_tests.log
  for v in "ISAC_PMI_FUSED=0" "ISAC_PAIR_G=1" "ISAC_PAIR_G=2" "ISAC_PAIR_G=2 ISAC_PAIR_T=256" "ISAC_PAIR_G=4" "ISAC_PAIR_G=4 ISAC_PAIR_T=384"; do
    env $v timeout 300 python tools/dev_pmi_variants.py 2>&1 | tail -3
  done > gpurun_out/r2_pmi1_variants.log
  cat gpurun_out/r2_pmi1_variants.log
  ;;
pmi2)
  for v in "ISAC_PAIR_G=1" "ISAC_PAIR_G=4" "ISAC_PAIR_G=4 ISAC_PAIR_T=384"; do
    echo "== $v"
    env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmi_ -s 8 -c 4 --csv python tools/dev_pmi_variants.py 2>/dev/null | grep -E "pmi_" | awk -F'","' '{print $5, $NF}' | cut -c1-160
  done > gpurun_out/r2_pmi2_launches.log 2>&1
  cat gpurun_out/r2_pmi2_launches.log
  ISAC_PAIR_G=4 ISAC_PAIR_T=384 timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmi_pair_fused -s 4 -c 1 -o gpurun_out/r2_pmi2_fused python tools/dev_pmi_variants.py > /dev/null 2>&1
  ls -la gpurun_out/
  ;;
pmi3)
  (timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_cfg23_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_pmi3_tests.log
  cat gpurun_out/r2_pmi3_tests.log
  for v in "ISAC_PMI_FUSED=0" "ISAC_PAIR_G=1" "ISAC_PAIR_G=4" "ISAC_PAIR_G=4 ISAC_PAIR_T=384"; do
    env $v timeout 300 python tools/dev_pmi_variants.py 2>&1 | tail -2 | head -1
    env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmi_ -s 8 -c 3 --csv python tools/dev_pmi_variants.py 2>/dev/null | grep -E "pmi_" | awk -F'","' '{print $5, $NF}' | cut -c1-160
  done > gpurun_out/r2_pmi3_variants.log 2>&1
  cat gpurun_out/r2_pmi3_variants.log
  ;;
pmi4)
  (timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_cfg23_gpu.py tests/test_golden_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_pmi4_tests.log
  cat gpurun_out/r2_pmi4_tests.log
  for v in "ISAC_PAIR_G=1" "ISAC_PAIR_G=2" "ISAC_PAIR_G=4" "ISAC_PAIR_G=4 ISAC_PAIR_T=384" "ISAC_PAIR_G=4 ISAC_PAIR_T=384 ISAC_PAIR_BP=0"; do
    env $v timeout 300 python tools/dev_pmi_variants.py 2>&1 | tail -2 | head -1
    env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmi_ -s 8 -c 2 --csv python tools/dev_pmi_variants.py 2>/dev/null | grep -E "pmi_" | awk -F'","' '{print $5, $NF}' | cut -c1-160
  done > gpurun_out/r2_pmi4_variants.log 2>&1
  cat gpurun_out/r2_pmi4_variants.log
  ;;
pmiq)
  [ -n "${PMIQ_NOTEST:-}" ] || (timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_golden_gpu.py tests/test_multipanel_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_pmiq_tests.log
  cat gpurun_out/r2_pmiq_tests.log
  for v in ${PMIQ_VARIANTS:-"ISAC_PAIR_ASC=0,ISAC_PAIR_RR=0" "ISAC_PAIR_ASC=1,ISAC_PAIR_RR=0" "ISAC_PAIR_ASC=1,ISAC_PAIR_RR=1" "ISAC_PAIR_ASC=0,ISAC_PAIR_RR=1" "ISAC_PAIR_UNIT=1100" "ISAC_PAIR_UNIT=2200" "ISAC_PAIR_UNIT=1"}; do
    v=${v//,/ }
    echo "== $v"
    env $v timeout 300 python tools/dev_pmi_variants.py 2>&1 | tail -2 | head -1
    env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pmi_ -s 8 -c 2 --csv python tools/dev_pmi_variants.py 2>/dev/null | grep -E "pmi_" | awk -F'","' '{print $5, $NF}' | cut -c1-160
  done > gpurun_out/r2_pmiq_variants.log 2>&1
  cat gpurun_out/r2_pmiq_variants.log
  ;;
profpmi)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmi_pair_fused -s 4 -c 1 -o gpurun_out/r2_prof_pmi_${TAG:-x} python tools/dev_pmi_variants.py > /dev/null 2>&1
  ls -la gpurun_out/*.ncu-rep
  ;;
misc)
  (timeout 900 python -m pytest tests/test_sensing_gpu.py tests/test_golden_gpu.py tests/test_cdl.py tests/test_properties_gpu.py tests/test_mex_mock_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_misc_tests.log
  cat gpurun_out/r2_misc_tests.log
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:echo_|cov_|eig_|music_|cdl_|prg_" -c 80 --csv python tools/profile_misc.py 2>/dev/null | grep -E "echo_|cov_|eig_|music_|cdl_|prg_" | awk -F'","' '{print $5, $NF}' | cut -c1-120 > gpurun_out/r2_misc_times.log
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:cdl_|prg_|ul_" -s 10 -c 40 --csv python tools/profile_comm.py 2>/dev/null | grep -E "cdl_|prg_|ul_" | awk -F'","' '{print $5, $NF}' | cut -c1-120 >> gpurun_out/r2_misc_times.log
  sort gpurun_out/r2_misc_times.log | uniq -c | sort -k2 | head -60
  ;;
prof2)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdl_response_umma -s 9 -c 1 -o gpurun_out/r2_prof2_cdl_dl python tools/profile_comm.py > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:echo_|cov_|eig_small|prg_precode" -s 8 -c 6 -o gpurun_out/r2_prof2_misc python tools/profile_misc.py > /dev/null 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_prof2_bench_launches.csv python bench.py --steps 1 --warmup 1 --frames-per-step 1 --no-cpu-baseline --skip-host-h > gpurun_out/r2_prof2_bench_under_ncu.json 2> /dev/null
  ls -la gpurun_out | tail -6
  ;;
full)
  (timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2_full_tests.log
  cat gpurun_out/r2_full_tests.log
  (timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3)
  timeout 600 python bench.py > gpurun_out/r2_full_bench.json 2> gpurun_out/r2_full_bench.err
  tail -n 3 gpurun_out/r2_full_bench.err; cat gpurun_out/r2_full_bench.json | cut -c1-1500
  ;;
t1)
  (timeout 1500 python -m pytest tests/test_cfg3_full_gpu.py tests/test_mex_mock_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/r2_t1_tests.log
  cat gpurun_out/r2_t1_tests.log
  ;;
t2)
  (timeout 1700 python -m pytest tests/test_cfg5_gpu.py tests/test_link_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_t2_tests.log
  cat gpurun_out/r2_t2_tests.log
  ;;
b1)
  timeout 900 python bench.py > gpurun_out/r2_b1_bench.json 2> gpurun_out/r2_b1_bench.err; tail -n 5 gpurun_out/r2_b1_bench.err; cut -c1-3000 gpurun_out/r2_b1_bench.json
  timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2_b1_cfg5.json 2> gpurun_out/r2_b1_cfg5.err; tail -n 5 gpurun_out/r2_b1_cfg5.err; cut -c1-2000 gpurun_out/r2_b1_cfg5.json
  ;;
b2)
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload cfg5 --steps 5 --warmup 2 > gpurun_out/r2_b2_cfg5_n2.json 2> gpurun_out/r2_b2_cfg5_n2.err; tail -n 5 gpurun_out/r2_b2_cfg5_n2.err; cut -c1-2000 gpurun_out/r2_b2_cfg5_n2.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_b2_cfg2_n2.json 2> gpurun_out/r2_b2_cfg2_n2.err; tail -n 3 gpurun_out/r2_b2_cfg2_n2.err; cut -c1-600 gpurun_out/r2_b2_cfg2_n2.json
  ;;
rdm1)
  (timeout 600 python -m pytest tests/test_rdm_gpu.py tests/test_sensing_gpu.py tests/test_golden_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_rdm1_tests.log; cat gpurun_out/r2_rdm1_tests.log
  for v in "ISAC_CFAR_FUSED=0" "ISAC_CFAR_FUSED=1" "ISAC_RDM_DIRECT=2 ISAC_RDM_HINTS=0x349" "ISAC_RDM_DIRECT=3 ISAC_RDM_HINTS=0x349" "ISAC_RDM_DIRECT=4 ISAC_RDM_HINTS=0x349" "ISAC_RDM_DIRECT=3"; do
    echo "== $v"; env $v timeout 300 python tools/dev_rdm_bench.py 0 2>&1 | grep "variant=0"
  done > gpurun_out/r2_rdm1_variants.log 2>&1
  cat gpurun_out/r2_rdm1_variants.log
  ;;
b3)
  (timeout 900 python -m pytest tests/test_cfg5_gpu.py tests/test_cdl.py -m gpu -q 2>&1 | tail -5)
  timeout 900 python bench.py --workload cfg3 --steps 5 --warmup 2 > gpurun_out/r2_b3_cfg3.json 2> gpurun_out/r2_b3_cfg3.err; tail -n 5 gpurun_out/r2_b3_cfg3.err; cut -c1-1500 gpurun_out/r2_b3_cfg3.json
  timeout 900 python bench.py --workload cfg5 --steps 10 --warmup 3 > gpurun_out/r2_b3_cfg5.json 2> gpurun_out/r2_b3_cfg5.err; tail -n 5 gpurun_out/r2_b3_cfg5.err; cut -c1-300 gpurun_out/r2_b3_cfg5.json
  ;;
prof)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pmi_pair_fused -s 4 -c 1 -o gpurun_out/r2_prof_pmi python tools/dev_pmi_variants.py > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cdl_response_umma -s 2 -c 1 -o gpurun_out/r2_prof_cdl python tools/profile_comm.py > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:rdm_range4096_lean|rdm_doppler256_tma|cfar2d" -s 16 -c 4 -o gpurun_out/r2_prof_rdm python tools/profile_rdm.py 4 0 > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:echo_|cov_" -s 8 -c 6 -o gpurun_out/r2_prof_misc python tools/profile_misc.py > /dev/null 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_prof_bench_launches.csv python bench.py --steps 1 --warmup 1 --frames-per-step 1 --no-cpu-baseline --skip-host-h > gpurun_out/r2_prof_bench_under_ncu.json 2> /dev/null
  ls -la gpurun_out | tail -8
  ;;
scale)
  N=$2
  run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:3}" > gpurun_out/r2_scale_$2_n$N.json 2> gpurun_out/r2_scale_$2_n$N.err; tail -n 1 gpurun_out/r2_scale_$2_n$N.json | cut -c1-400; }
  run 29521 cfg5 --workload cfg5 --steps 20 --warmup 5
  run 29522 cfg3 --workload cfg3 --steps 10 --warmup 3
  if [ "${3:-}" = "cfg2" ]; then run 29523 cfg2 --steps 10 --warmup 3 --no-cpu-baseline; fi
  ;;
*) echo "unknown step $step"; exit 1;;
esac
