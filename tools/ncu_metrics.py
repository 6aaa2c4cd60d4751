"""Print selected metrics of every kernel in an .ncu-rep (reads `ncu -i X --page raw --csv` from stdin)."""
import csv
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_lsu.avg.pct",
        "sm__inst_executed_pipe_fp16.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled", "sm__pipe_tensor_cycles_active.avg.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "sm__cycles_elapsed.max")
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for r in rows[2:]:
    print("---", r[hdr.index("Kernel Name")][:60])
    for i, k in enumerate(hdr):
        if any(k == x or k.startswith(x) for x in WANT):
            if "warp_issue_stalled" in k and not k.endswith("per_warp_active.pct"):
                continue
            print("    %-90s %s %s" % (k, r[i], rows[1][i]))
