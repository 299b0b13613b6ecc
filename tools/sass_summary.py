#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell path (B200_PROFILING.md): UTCHMMA/UTCQMMA
(tcgen05.mma), LDTM/STTM (TMEM loads/stores), UTMALDG/UTMASTG (TMA tensor copies), UBLKCP (bulk copies), SYNCS
(mbarrier).  usage: python tools/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mrn_b200", "libmrn_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA")
kern, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern).split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern:
        for mn in MN:
            if re.search(r"\b%s[\.\s]" % mn, line):
                counts[kern][mn] += 1
print("# SASS summary of %s (sm_100a); columns: %s" % (os.path.basename(lib), " ".join(MN)))
tc = 0
for k, c in counts.items():
    if c["UTCHMMA"] or c["UTCQMMA"]:
        tc += 1
    print("%-110s %s" % (k[:110], " ".join("%s=%d" % (mn, c[mn]) for mn in MN if c[mn])))
print("# kernels: %d total, %d with tcgen05.mma (UTCHMMA/UTCQMMA)" % (len(counts), tc))
