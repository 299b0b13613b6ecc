#!/usr/bin/env python
"""Fused mixer kernel vs the unfused qkv GEMM -> attention -> proj GEMM sequence at the headline unit count
(6 experts x 256 samples = 1536 units), per stage shape.  CUDA events, 3 warm-up + 10 timed launches each."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mrn_b200 import ops  # noqa: E402

torch.manual_seed(0)
units = int(os.environ.get("UNITS", 1536))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


SHAPES = [tuple(int(v) for v in t.split(":")) for t in os.environ.get("SHAPES", "64:1,128:1,128:0,256:0").split(",")]
for D, local in SHAPES:
    N, heads = 32768 // D, D // 32
    x = torch.randn(units, N, D, device="cuda")
    a16 = ops.cast_bf16(torch.randn(units, N, D, device="cuda"))
    wqkv = ops.cast_bf16(torch.randn(3 * D, D, device="cuda") / D ** 0.5)
    bqkv = torch.randn(3 * D, device="cuda") * 0.1
    wp = ops.cast_bf16(torch.randn(D, D, device="cuda") / D ** 0.5)
    bp = torch.randn(D, device="cuda") * 0.1
    g, bt = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    ln = D <= 128
    t_f = timeit(lambda: ops.mixer_bf16(a16, wqkv, bqkv, wp, bp, x, local, None, g if ln else None, bt if ln else None))
    a2 = a16.view(units * N, D)

    def unfused():
        qkv = ops.linear_bf16(a2, wqkv, bqkv, None, False, out_f32=False)
        att = ops.svtr_attention_bf16(qkv.view(units, N, 3 * D), heads, N // 64, 64, local)
        return ops.linear_bf16(att.view(units * N, D), wp, bp, x.view(units * N, D), False, out_f32=True)
    t_u = timeit(unfused) if not os.environ.get("FUSED_ONLY") else float("nan")
    flops = units * (8.0 * N * D * D + 4.0 * N * N * D)
    print("D=%3d N=%3d local=%d: fused %.3f ms (%.0f TFLOP/s dense-eq) | unfused qkv+attn+proj %.3f ms | ratio %.2f" %
          (D, N, local, t_f, flops / t_f / 1e9, t_u, t_u / t_f), flush=True)
