"""Turn ncu captures brought back in gpurun_out/ into the small text summaries kept under profiles/.

  python tools/summarise_ncu.py rep   gpurun_out/x.ncu-rep  profiles/x_summary.txt  "header comment"
  python tools/summarise_ncu.py list  gpurun_out/x.csv      profiles/x_launches.txt "header comment"
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEEP = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "inst_executed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__maximum_warps_per_active_cycle_pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
)


def rep(src, dst, note):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {note}\n# source: {src} (ncu --set full --clock-control none --import-source on)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[head.index('Kernel Name')]}  (launch id {r[head.index('ID')]})\n")
            for i, name in enumerate(head):
                if name in KEEP and r[i] != "":
                    f.write(f"{name} [{units[i]}] = {r[i]}\n")


def launches(src, dst, note):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.reader(lines))
    head = rows[0]
    ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
        k = re.sub(r"\(.*", "", r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {note}\n# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised launches)\n")
        f.write(f"# {n} launches, {tot / 1e3:.3f} ms of kernel time\n")
        f.write(f"{'kernel':72s} {'n':>6s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:72]:72s} {c:6d} {t:12.1f} {t / c:10.2f} {t / tot:7.3f}\n")


if __name__ == "__main__":
    {"rep": rep, "list": launches}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
