#!/usr/bin/env python
"""Timeline of CTA 0 of the fused mixer kernel (globaltimer stamps, ns): where one softmax stream and its MMA issuer spend
the time of a (query tile, key block) pair.  usage: SHAPE=64:1 python tools/mixer_trace.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mrn_b200 import _lib as L, ops  # noqa: E402

D, local = (int(v) for v in os.environ.get("SHAPE", "64:1").split(":"))
units, N = 1536, 32768 // D
x = torch.randn(units, N, D, device="cuda")
a16 = ops.cast_bf16(torch.randn(units, N, D, device="cuda"))
wqkv = ops.cast_bf16(torch.randn(3 * D, D, device="cuda") / D ** 0.5)
bqkv = torch.randn(3 * D, device="cuda") * 0.1
wp = ops.cast_bf16(torch.randn(D, D, device="cuda") / D ** 0.5)
bp = torch.randn(D, device="cuda") * 0.1
lib = C.CDLL(L.LIB_PATH)
for _ in range(2):
    ops.mixer_bf16(a16, wqkv, bqkv, wp, bp, x, local)
tr = torch.zeros(16, 256, dtype=torch.int64, device="cuda")
lib.mrnb_mixer_set_trace(C.c_void_p(tr.data_ptr()))
ops.mixer_bf16(a16, wqkv, bqkv, wp, bp, x, local)
torch.cuda.synchronize()
lib.mrnb_mixer_set_trace(C.c_void_p(0))
t = tr.cpu().numpy()
names = ["mma:start", "mma:S(next) issued", "mma:P avail", "mma:PV issued", "sm:start", "sm:S ready", "sm:S in regs", "sm:max done",
         "sm:PV(prev) done", "sm:exps+P stored", "sm:P published"]
t0 = t[4, 0]
print("pair | " + " | ".join("%s" % n for n in names))
for p in range(0, 40):
    if t[4, p] == 0:
        continue
    print("%4d | " % p + " | ".join("%8d" % (t[k, p] - t0) if t[k, p] else "       -" for k in range(11)))
# average per-pair durations of the softmax stream
import numpy as np
valid = [p for p in range(2, 200) if t[4, p] and t[10, p] and t[4, p + 1]]
d = lambda a, b: float(np.mean([t[b, p] - t[a, p] for p in valid]))
print("avg ns over %d pairs: wait S %.0f | ld S %.0f | max %.0f | wait PV(prev) %.0f | exps+store %.0f | fence+arrive %.0f | pair period %.0f" %
      (len(valid), d(4, 5), d(5, 6), d(6, 7), d(7, 8), d(8, 9), d(9, 10), float(np.mean([t[4, p + 1] - t[4, p] for p in valid]))))
valid_m = [p for p in range(2, 200) if t[0, p] and t[3, p]]
dm = lambda a, b: float(np.mean([t[b, p] - t[a, p] for p in valid_m]))
print("MMA issuer avg ns: wait s_empty + issue S %.0f | wait P %.0f | issue PV %.0f" % (dm(0, 1), dm(1, 2), dm(2, 3)))

# OVL (D = 128): Y-epilogue warpgroup timeline, us relative to the first softmax stamp
if t[11, 0]:
    f = lambda v: "%.2f" % ((int(v) - int(t0)) / 1e3) if v else "-"
    print("unit | y_full seen | Y drained | epilogue done ; heads drained at (4 per unit)")
    for i in range(4):
        print("%4d | %s | %s | %s ; %s" % (i, f(t[11, i]), f(t[12, i]), f(t[13, i]), " ".join(f(t[14, 4 * i + k]) for k in range(4))))
    hs = [p for p in range(0, 48, 4) if t[4, p]]
    print("softmax stream 0, first pair of each head: wait-S start / S ready:", " ".join("%s/%s" % (f(t[4, p]), f(t[5, p])) for p in hs))
