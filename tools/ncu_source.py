"""Instruction mix / hottest SASS lines from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu -i rep --page source --csv | python tools/ncu_source.py [N]"""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
cs, ce, cw = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((float(r[ce]), float(r[cw] or 0), r[cs].strip()))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1.0
tots = sum(d[1] for d in data) or 1.0
agg = collections.Counter()
for e, s, src in data:
    t = src.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
    agg[op.split(".")[0]] += e
print("total warp instructions %.0f, stall samples %.0f" % (tot, tots))
for k, v in agg.most_common(18):
    print("  %-10s %14.0f  %.3f" % (k, v, v / tot))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
print("hottest lines by stall samples:")
for i, (e, s, src) in sorted(enumerate(data), key=lambda kv: -kv[1][1])[:n]:
    print("  [%4d] %7.0f samples %.3f | exec %12.0f | %s" % (i, s, s / tots, e, src[:90]))
