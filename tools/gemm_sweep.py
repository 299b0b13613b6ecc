"""Timing sweep of the persistent tcgen05 GEMM (mrnb_linear_bf16) over the expert shapes.  GPU only."""
import torch
from mrn_b200 import ops

def run(M, N, K, f32out, res=False, iters=20):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = torch.randn(N, K, device="cuda").bfloat16()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    for _ in range(3):
        ops.linear_bf16(a, w, b, r, False, f32out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.linear_bf16(a, w, b, r, False, f32out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    byt = M * K * 2 + N * K * 2 + M * N * (4 if f32out else 2) + (M * N * 4 if res else 0)
    tiles = (M // 128) * ((N + (127 if N >= 128 and (N % 128 == 0 or N > 256) else 63)) // (128 if N >= 128 and (N % 128 == 0 or N > 256) else 64))
    print("M=%d N=%d K=%d f32=%d res=%d: %.1f us  %.0f GB/s  %.1f TF/s  %.2f us/tile/SM" % (
        M, N, K, f32out, res, ms * 1e3, byt / ms / 1e6, 2.0 * M * N * K / ms / 1e9, ms * 1e3 * 148 / tiles))

M = 786432
for N in (64, 128, 192, 256, 384):
    run(M, N, 64, False)
run(M, 64, 64, True, True)
run(M // 2, 384, 128, False)
run(M // 2, 128, 128, True, True)
run(M // 4, 768, 256, False)
run(M // 4, 256, 256, True, True)
run(98304 // 6, 5153, 256, True)
