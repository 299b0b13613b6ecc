"""Throughput of the device-side image preparation (mrn_b200.data) next to Pillow on the host cores.
usage: python tools/resize_bench.py [B]   -> one JSON line"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

from mrn_b200 import data

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
imgs = [rng.integers(0, 256, size=(int(rng.integers(24, 80)), int(rng.integers(60, 600)), 4), dtype=np.uint8) for _ in range(B)]
in_bytes = sum(a.size for a in imgs)
out_bytes = B * 4 * 32 * 256 * 4
for _ in range(3):
    out = data.resize_normalize_batch(imgs)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 10
for _ in range(n):
    out = data.resize_normalize_batch(imgs)          # includes host packing + the H2D copy (end to end)
torch.cuda.synchronize()
e2e_ms = (time.perf_counter() - t0) / n * 1e3
# kernels only: reuse one staged buffer
from mrn_b200 import _lib as L
import ctypes as C
arrs = imgs
hs = np.array([a.shape[0] for a in arrs], dtype=np.int32); ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
sizes = hs.astype(np.int64) * ws * 4
offs = np.zeros(B, dtype=np.int64); offs[1:] = np.cumsum(sizes)[:-1]
px = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs])).cuda()
d_off, d_w, d_h = torch.from_numpy(offs).cuda(), torch.from_numpy(ws).cuda(), torch.from_numpy(hs).cuda()
o = torch.empty(B, 4, 32, 256, device="cuda")
lib = L.load()
need = int(lib.mrnb_resize_workspace_bytes(B, int(hs.max()), 256))
wsb = torch.empty(need, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    L.check(lib.mrnb_resize_normalize_rgba(C.c_void_p(px.data_ptr()), C.c_void_p(d_off.data_ptr()), C.c_void_p(d_w.data_ptr()),
                                           C.c_void_p(d_h.data_ptr()), B, int(ws.max()), int(hs.max()), 32, 256,
                                           C.c_void_p(o.data_ptr()), C.c_void_p(wsb.data_ptr()), need, st), "resize")


for _ in range(3):
    run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    run()
e1.record(); torch.cuda.synchronize()
k_ms = e0.elapsed_time(e1) / 50
tmp_bytes = int((hs.astype(np.int64) * 256 * 4).sum()) * 2
cpu = None
try:
    from PIL import Image
    t0 = time.perf_counter()
    for a in imgs:
        r = np.asarray(Image.fromarray(a, "RGBA").resize((256, 32), Image.BICUBIC))
        t = torch.from_numpy(r).permute(2, 0, 1).contiguous().float().div(255).sub_(0.5).div_(0.5)
    cpu = B / (time.perf_counter() - t0)
except Exception:
    pass
print(json.dumps({"metric": "image preparation (RGBA -> 4x32x256 fp32)", "batch": B, "images_per_s_kernels": round(B / k_ms * 1e3, 1),
                  "kernel_ms": round(k_ms, 4), "images_per_s_e2e_host_pack_h2d": round(B / e2e_ms * 1e3, 1),
                  "algorithmic_bytes": in_bytes + out_bytes, "gbs_kernels": round((in_bytes + out_bytes + tmp_bytes) / k_ms / 1e6, 1),
                  "pillow_images_per_s_1_thread": None if cpu is None else round(cpu, 1)}))
