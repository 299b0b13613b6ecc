"""Timing sweep of mrnb_gate_combine (row pass) to separate per-row from per-byte cost.  GPU only."""
import sys
import torch
from mrn_b200 import ops, synth

def run(cc, B, T, soft=True, want=("E", "lpe"), iters=20):
    dev = "cuda"
    I = len(cc)
    zs = []
    for c in cc:
        ld = ops.round_up(c, 4)
        buf = torch.randn(B, T, ld, device=dev)
        zs.append(buf[:, :, :c])
    if soft:
        gate = torch.softmax(torch.randn(B, I, device=dev), -1)
    else:
        gate = torch.nn.functional.one_hot(torch.randint(0, I, (B,), device=dev), I).float()
    tgt = torch.randint(2, cc[-1], (B, 25), device=dev)
    lens = torch.randint(1, 26, (B,), device=dev, dtype=torch.int32)
    kw = dict(want_logits=False, want_E="E" in want, want_decode=False)
    args = (tgt, lens) if "lpe" in want else (None, None)
    for _ in range(3):
        ops.gate_combine(zs, gate, *args, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # rotate over 3 copies to defeat L2 (126 MB) for small cases
    e0.record()
    for _ in range(iters):
        ops.gate_combine(zs, gate, *args, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = 4.0 * B * T * sum(cc)
    print("I=%d C=%s B=%d T=%d soft=%d want=%s: %.3f ms  %.0f GB/s  %.2f us/row/SM" % (
        I, cc[-1], B, T, soft, ",".join(want), ms, nbytes / ms / 1e6, ms * 1e3 * 148 / (B * T)))

cc = synth.MLT17_CLASS_COUNTS
run(cc, 256, 64)
run(cc, 256, 64, want=("E",))
run(cc, 256, 64, want=())
run(cc, 256, 64, soft=False)
run(cc, 128, 64)
run(cc, 512, 64)
run(cc[-1:], 256, 64)
run(cc[-2:], 256, 64)
run((1024, 2048), 256, 64)
run((8192,) * 6, 128, 64)
