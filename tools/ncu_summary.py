"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total us, share.
usage: python tools/ncu_summary.py launches.csv [--detail]"""
import collections
import csv
import re
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # `| head` closes the pipe early


def short(name):
    m = re.search(r"(\w+)(<[^(]*>)?\(", name)
    n = m.group(1) if m else name
    t = m.group(2) if m and m.group(2) else ""
    t = re.sub(r"\(bool\)", "", t)
    return n + t


def main(path, detail=False):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        tot += us
        if detail:
            print("%-60s grid=%-18s %10.1f us" % (k[:60], r.get("Grid Size", ""), us))
    print("# %d launches, %.1f us total (cold-cache, serialised: compare shares)" % (len(rows), tot))
    print("%-62s %6s %12s %7s" % ("kernel", "n", "us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-62s %6d %12.1f %7.3f" % (k[:62], n, t, t / tot))


if __name__ == "__main__":
    main(sys.argv[1], "--detail" in sys.argv)
