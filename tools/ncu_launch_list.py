#!/usr/bin/env python
"""Per-launch listing (in launch order) of an ncu CSV with gpu__time_duration.sum + dram__bytes_{read,write}.sum.
usage: python tools/ncu_launch_list.py launches.csv [first_kernel_substring]"""
import collections
import csv
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)
rows = list(csv.DictReader([l for l in open(sys.argv[1]) if not l.startswith("==")]))
start = sys.argv[2] if len(sys.argv) > 2 else None
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}
per = collections.OrderedDict()
for r in rows:
    k = int(r["ID"])
    per.setdefault(k, {"name": r["Kernel Name"], "grid": r.get("Grid Size")})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * U.get(r["Metric Unit"], 1)
on = start is None
tot = 0.0
for k in sorted(per):
    m = per[k]
    n = re.sub(r"\(.*", "", m["name"]).replace("void ", "").replace("<unnamed>::", "")[:44]
    if not on and start in n:
        on = True
    if not on:
        continue
    tot += m["gpu__time_duration.sum"]
    print("%4d %-44s %-16s %7.1f us  rd %6.0f MB  wr %6.0f MB" % (k, n, m["grid"], m["gpu__time_duration.sum"], m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6))
print("total %.1f us" % tot)
