#!/usr/bin/env python
"""Per-kernel determinism check at the headline shapes: every tcgen05 building block is launched repeatedly on identical
inputs; any bitwise difference between reruns (or against a torch fp32 reference beyond bf16 rounding) is a race."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from mrn_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
REPS = int(os.environ.get("REPS", 6))


def report(name, outs, ref=None):
    o0 = outs[0].float()
    bad = [int((o.float() != o0).sum()) for o in outs[1:]]
    nf = [int((~torch.isfinite(o.float())).sum()) for o in outs]
    msg = "%-44s rerun-mismatch elems %s nonfinite %s" % (name, bad, nf)
    if ref is not None:
        msg += " rel-err vs torch %.3g" % float((o0 - ref).abs().max() / ref.abs().max())
    print(msg, flush=True)


for (d, N) in ((64, 512), (128, 256), (256, 128)):
    M = 6 * 256 * N
    x = torch.randn(M, d, device=dev)
    a16 = ops.cast_bf16(x)
    # plain linear: qkv (bf16 out), proj with residual (fp32 out)
    w = torch.randn(3 * d, d, device=dev) * d ** -0.5
    b = torch.randn(3 * d, device=dev)
    w16 = ops.cast_bf16(w)
    outs = [ops.linear_bf16(a16, w16, b, None, False, out_f32=False) for _ in range(REPS)]
    torch.cuda.synchronize()
    ref = a16.float() @ w16.float().t() + b
    report("qkv linear d=%d" % d, outs, ref)
    del outs, ref
    wp = torch.randn(d, d, device=dev) * d ** -0.5
    bp = torch.randn(d, device=dev)
    wp16 = ops.cast_bf16(wp)
    outs = [ops.linear_bf16(a16, wp16, bp, x, False, out_f32=True) for _ in range(REPS)]
    torch.cuda.synchronize()
    ref = a16.float() @ wp16.float().t() + bp + x
    report("proj linear+res d=%d" % d, outs, ref)
    del outs, ref
    # attention
    heads = d // 32
    H = N // 64
    qkv = (torch.randn(6 * 256, N, 3 * d, device=dev)).to(torch.bfloat16).contiguous()
    for local in ((1, 0) if d < 256 else (0,)):
        outs = [ops.svtr_attention_bf16(qkv, heads, H, 64, local) for _ in range(REPS)]
        torch.cuda.synchronize()
        report("attention d=%d local=%d" % (d, local), outs)
        del outs
    del qkv
    # fused MLP (in place on x)
    w1 = torch.randn(4 * d, d, device=dev) * d ** -0.5
    b1 = torch.randn(4 * d, device=dev) * 0.1
    w2 = torch.randn(d, 4 * d, device=dev) * (4 * d) ** -0.5
    b2 = torch.randn(d, device=dev) * 0.1
    w1_16, w2_16 = ops.cast_bf16(w1), ops.cast_f16(w2)
    for ln in ((True, False) if d <= 128 else (False,)):
        outs, lns = [], []
        g = torch.ones(d, device=dev)
        bt = torch.zeros(d, device=dev)
        for _ in range(REPS):
            xx = x.clone()
            lo = ops.mlp_bf16(a16, w1_16, b1, w2_16, b2, xx, None, 1, g if ln else None, bt if ln else None)
            outs.append(xx)
            if lo is not None:
                lns.append(lo)
        torch.cuda.synchronize()
        ref = x + torch.nn.functional.gelu(a16.float() @ w1_16.float().t() + b1) @ w2_16.float().t() + b2
        report("mlp d=%d ln=%s" % (d, ln), outs, ref)
        if lns:
            report("   ln_out", lns)
        del outs, lns, ref
    torch.cuda.empty_cache()
print("done")
