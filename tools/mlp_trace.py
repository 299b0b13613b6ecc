"""Debug timeline of the fused MLP kernel (CTA 0): python tools/mlp_trace.py [D] [M]"""
import ctypes as C
import math
import sys
import torch
sys.path.insert(0, ".")
from mrn_b200 import _lib as L, ops
D = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 128 * 4
lib = L.load()
lib.mrnb_mlp_set_trace.argtypes = [C.c_void_p]
tr = torch.zeros(10 * 64, dtype=torch.int64, device="cuda")
a = torch.randn(M, D, device="cuda").bfloat16()
w1 = (torch.randn(4 * D, D, device="cuda") / math.sqrt(D)).bfloat16()
w2 = (torch.randn(D, 4 * D, device="cuda") / math.sqrt(4 * D)).half()
b1, b2, x = torch.zeros(4 * D, device="cuda"), torch.zeros(D, device="cuda"), torch.randn(M, D, device="cuda")
for it in range(3):
    if it == 2:
        lib.mrnb_mlp_set_trace(C.c_void_p(tr.data_ptr()))
    ops.mlp_bf16(a, w1, b1, w2, b2, x)
torch.cuda.synchronize()
t = tr.cpu().view(10, 64)
base = int(t[9, 0])
names = ["mma:wait P", "mma:P ready", "mma:PV issued", "epi:wait S", "epi:S ready", "epi:GELU done", "epi:P free", "epi:P published", "epi:O ready", "epi:tile start"]
C_ = 4 * D // 128
print("D=%d chunks/tile=%d ; times in us relative to tile 0 start" % (D, C_))
for k in range(min(3 * C_, 24)):
    print("chunk %2d: " % k + "  ".join("%s=%.2f" % (names[s].split(":")[1][:9], (int(t[s, k]) - base) / 1e3) for s in (3, 4, 5, 6, 7)) +
          "  |  " + "  ".join("%s=%.2f" % (names[s][4:], (int(t[s, k]) - base) / 1e3) for s in (0, 1, 2)))
print("tile starts:", [(int(v) - base) / 1e3 for v in t[9, :4]], "O ready:", [(int(v) - base) / 1e3 for v in t[8, :4]])
