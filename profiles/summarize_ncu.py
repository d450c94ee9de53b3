#!/usr/bin/env python
"""Summarise an ncu report (--page raw --csv) into the handful of metrics DESIGN.md / bench.py cite.
Usage: python profiles/summarize_ncu.py gpurun_out/prof_x.ncu-rep > profiles/r1_x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    print("==", r[ki][:90], "| id", r[0])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("   %-85s %14s %s" % (k, r[i], units[i]))
