mkdir -p gpurun_out
timeout 120 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/c8_warm.csv python tools/profile_rdm.py 4 0 > /dev/null 2>&1
grep -E "rdm_|cfar" gpurun_out/c8_warm.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | tail -60
