#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list with gpu__time_duration.sum + dram__bytes_{read,write}.sum:
launches, time, share, DRAM bytes; plus the per-step totals.  usage: python tools/ncu_step_summary.py launches.csv STEPS"""
import collections
import csv
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)

path, steps = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    n = m.group(1) if m else name
    t = m.group(2) if m and m.group(2) else ""
    return n + re.sub(r"\((bool|int)\)", "", t)


per = collections.OrderedDict()
for r in rows:
    k = (r["ID"], short(r["Kernel Name"]))
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    per.setdefault(k, {})[r["Metric Name"]] = v
agg = collections.OrderedDict()
for (_, name), m in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
tot_t = sum(a[1] for a in agg.values())
tot_r = sum(a[2] for a in agg.values())
tot_w = sum(a[3] for a in agg.values())
print("# %d launches over %d step(s); per step: %.0f launches, %.2f ms (cold-cache, serialised: compare SHARES), DRAM read %.2f GB + write %.2f GB = %.2f GB"
      % (len(per), steps, len(per) / steps, tot_t / steps / 1e3, tot_r / steps / 1e9, tot_w / steps / 1e9, (tot_r + tot_w) / steps / 1e9))
print("%-58s %6s %11s %7s %10s %10s" % ("kernel", "n/step", "us/step", "share", "rd MB/step", "wr MB/step"))
for name, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %6.1f %11.1f %7.3f %10.1f %10.1f" % (name[:58], n / steps, t / steps, t / tot_t, rd / steps / 1e6, wr / steps / 1e6))
