mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pmi_pair_kernel" -s 1 -c 1 -o gpurun_out/c28_pmi_pair python tools/profile_comm.py > gpurun_out/c28_ncu.log 2>&1
tail -n 3 gpurun_out/c28_ncu.log
