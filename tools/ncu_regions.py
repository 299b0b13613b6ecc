#!/usr/bin/env python
"""Stall samples of an `ncu --set full --import-source on` capture, aggregated over consecutive SASS regions, plus the
headline raw metrics.  usage: python tools/ncu_regions.py report.ncu-rep [bucket=100] [min_samples=300]"""
import collections
import csv
import io
import signal
import subprocess
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)
rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100
MIN = int(sys.argv[3]) if len(sys.argv) > 3 else 300
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
KEYS = ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size")
for i, k in enumerate(raw[0]):
    if k in KEYS:
        print("%-70s %s %s" % (k, raw[2][i], raw[1][i]))
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hi = next(i for i, r in enumerate(src) if "Source" in r and "Instructions Executed" in r)
hdr = src[hi]
cs, ce, cw = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
cols = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = src[hi + 1:]
tot = sum(float(r[cw] or 0) for r in data)
print("total stall samples %.0f, warp instructions %.1f M" % (tot, sum(float(r[ce] or 0) for r in data) / 1e6))
print("SASS region | samples (share) | warp instr | top stall reasons | most frequent opcodes")
for b0 in range(0, len(data), B):
    seg = data[b0:b0 + B]
    t = sum(float(r[cw] or 0) for r in seg)
    if t < MIN:
        continue
    st = sorted(((h, sum(float(r[i] or 0) for r in seg)) for h, i in cols), key=lambda kv: -kv[1])[:3]
    ops = collections.Counter((r[cs].split()[1] if r[cs].strip().startswith("@") else r[cs].split()[0]) for r in seg if r[cs].strip())
    print("[%4d,%4d) %6.0f (%.3f) %7.1f M  %s  %s" % (b0, b0 + B, t, t / tot, sum(float(r[ce] or 0) for r in seg) / 1e6,
                                                     ", ".join("%s %.0f" % kv for kv in st), " ".join(k for k, _ in ops.most_common(4))))
