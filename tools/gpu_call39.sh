mkdir -p gpurun_out
timeout 300 python tools/dev_next_rows_bench.py > gpurun_out/c42_next_rows.log 2>&1
cat gpurun_out/c42_next_rows.log | tail -12
