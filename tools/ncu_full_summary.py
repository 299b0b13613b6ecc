#!/usr/bin/env python
"""Key metrics of every launch in an `ncu --set full` report.  usage: python tools/ncu_full_summary.py report.ncu-rep"""
import csv
import io
import signal
import subprocess
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)
KEYS = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("--- launch")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-80s %s %s" % (k, r[i][:90], units[i]))
