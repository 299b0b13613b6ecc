#!/usr/bin/env python
"""Fused MLP kernel at the headline row count (6 experts x 256 samples x N tokens = 1536 * 32768 / D rows) per stage width,
with and without the fused LayerNorm.  CUDA events, 3 warm-up + 10 timed launches each."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from mrn_b200 import ops  # noqa: E402

torch.manual_seed(0)
for D in (64, 128, 256):
    M = 1536 * 32768 // D
    a = torch.randn(M, D, device="cuda").bfloat16()
    w1 = (torch.randn(4 * D, D, device="cuda") / math.sqrt(D)).bfloat16()
    w2 = (torch.randn(D, 4 * D, device="cuda") / math.sqrt(4 * D)).half()
    b1, b2, x = torch.randn(4 * D, device="cuda") * 0.1, torch.randn(D, device="cuda") * 0.1, torch.randn(M, D, device="cuda")
    g, bt = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    for ln in ((True, False) if D <= 128 else (False,)):
        fn = lambda: ops.mlp_bf16(a, w1, b1, w2, b2, x, None, 1, g if ln else None, bt if ln else None)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("D=%3d LN=%d: %.3f ms  %.0f TFLOP/s" % (D, ln, ms, 16.0 * M * D * D / ms / 1e9), flush=True)
