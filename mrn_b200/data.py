"""Device-side image preparation with the reference's collate interface (data/dataset.py:169-197 AlignCollate with
Aug == "None" / mode != "train", i.e. ResizeNormalize :235-246): RGBA crops of any size -> [B,4,imgH,imgW] fp32 in
[-1,1] ON THE GPU, byte-identical to the PIL + torchvision host pipeline.  The LMDB reader and image decoding stay on
the host (data/dataset.py:44-112); what moves is the per-image bicubic resize, ToTensor and normalisation, which at
B200 step times (15 ms per 256 images) would otherwise bound the loop.
"""
import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib as L


def _as_rgba_array(img) -> np.ndarray:
    """PIL.Image (any mode; converted like data/dataset.py:97 `.convert("RGBA")`) or an [H,W,4] uint8 array."""
    if isinstance(img, np.ndarray):
        a = img
    elif isinstance(img, torch.Tensor):
        a = img.cpu().numpy()
    else:                                   # PIL image: conversion / decoding is host work, as in the reference
        a = np.asarray(img.convert("RGBA"))
    if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("expected an RGBA uint8 image [H,W,4], got %s %s" % (a.dtype, a.shape))
    return np.ascontiguousarray(a)


def resize_normalize_batch(images: Sequence, img_w: int = 256, img_h: int = 32, device=None, stream=None) -> torch.Tensor:
    """[B,4,img_h,img_w] fp32 device tensor from a sequence of RGBA images.  One pinned staging buffer, one H2D copy,
    two kernels (mrnb_resize_normalize_rgba)."""
    if not torch.cuda.is_available():
        raise RuntimeError("mrn_b200.data needs a CUDA device: there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    arrs = [_as_rgba_array(im) for im in images]
    B = len(arrs)
    if B == 0:
        raise ValueError("empty batch")
    hs = np.array([a.shape[0] for a in arrs], dtype=np.int32)
    ws = np.array([a.shape[1] for a in arrs], dtype=np.int32)
    sizes = hs.astype(np.int64) * ws.astype(np.int64) * 4
    offs = np.zeros(B, dtype=np.int64)
    offs[1:] = np.cumsum(sizes)[:-1]
    total = int(sizes.sum())
    # header (offsets, widths, heights) and pixels travel in ONE pinned buffer -> ONE host-to-device copy
    head = B * 8 + B * 4 + B * 4
    head = (head + 15) // 16 * 16
    stage = torch.empty(head + total, dtype=torch.uint8).pin_memory()
    sv = stage.numpy()
    sv[:B * 8] = offs.view(np.uint8)
    sv[B * 8:B * 12] = ws.view(np.uint8)
    sv[B * 12:B * 16] = hs.view(np.uint8)
    for a, o in zip(arrs, offs):
        sv[head + o: head + o + a.size] = a.reshape(-1)
    dev = stage.to(device, non_blocking=True)
    base = dev.data_ptr()
    out = torch.empty(B, 4, img_h, img_w, device=device, dtype=torch.float32)
    max_w, max_h = int(ws.max()), int(hs.max())
    lib = L.load()
    need = int(lib.mrnb_resize_workspace_bytes(B, max_h, img_w))
    wsb = torch.empty(need, dtype=torch.uint8, device=device)
    st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream if stream is None else stream)
    rc = lib.mrnb_resize_normalize_rgba(C.c_void_p(base + head), C.c_void_p(base), C.c_void_p(base + B * 8),
                                        C.c_void_p(base + B * 12), B, max_w, max_h, img_h, img_w, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(wsb.data_ptr()), need, st)
    L.check(rc, "resize_normalize_rgba")
    return out


class AlignCollate:
    """Drop-in for data/dataset.py:169-197 (`AlignCollate(opt, mode)` as a DataLoader collate_fn or called directly on a
    list of (image, label)): returns (image_tensors ON THE DEVICE, labels).  Augmentations (opt.Aug != "None" in train
    mode: Blur / Crop / Rot / ABINet, data/dataset.py:255-330) are host-side PIL / OpenCV code outside the hot path and
    raise here."""

    def __init__(self, opt, mode="train"):
        self.opt = opt
        self.mode = mode
        if getattr(opt, "Aug", "None") != "None" and mode == "train":
            raise NotImplementedError("mrn_b200.data.AlignCollate implements the Aug='None' / test-mode path "
                                      "(ResizeNormalize); augmentations are not part of the hot path")

    def __call__(self, batch):
        images, labels = zip(*batch)
        return resize_normalize_batch(images, self.opt.imgW, self.opt.imgH), labels
