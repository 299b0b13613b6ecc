"""CPU oracle of the reference's image preparation (data/dataset.py:235-246 ResizeNormalize, called from AlignCollate
data/dataset.py:169-197): PIL `Image.resize((imgW, imgH), BICUBIC)` on an RGBA image, torchvision ToTensor, (x-0.5)/0.5.

TEST INFRASTRUCTURE (only tests/, smoke() and bench.py's cpu_baseline leg may import this).

The arithmetic lives in a third-party dependency, Pillow (the reference pins none; this container has Pillow 12.2):
  * Image.resize on mode RGBA converts to premultiplied "RGBa", resamples, converts back (PIL/Image.py resize());
    an image that already has the target size is returned unchanged;
  * libImaging/Resample.c: separable two-pass resampling, horizontal then vertical, 8-bit intermediate;
    precompute_coeffs() (double) -> normalize_coeffs_8bpc() (fixed point, PRECISION_BITS = 22) -> integer MAC,
    clip8(ss >> 22) with ss starting at 1 << 21; bicubic filter a = -0.5, support 2, antialiased when shrinking;
  * libImaging/Convert.c: rgbA2rgba (premultiply, MULDIV255) and rgba2rgbA (un-premultiply, CLIP8(255*c/a)).
Restated here in numpy integer arithmetic and pinned bit-exactly against Pillow itself (tests/test_resize_pinning.py,
fixtures from oracle/make_resize_golden.py).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bicubic_filter(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for box (0, in_size).
    Returns (bounds [out,2] int (xmin, count), kk [out, ksize] int32, ksize)."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """Resample along axis 0 of a [n, ..., C] uint8 array (8bpc integer MAC of Resample.c)."""
    bounds, kk, _ = precompute_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        ss = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        ss = ss + np.tensordot(kk[xx, :n], src[xmin:xmin + n], axes=(0, 0))
        out[xx] = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def premultiply(rgba: np.ndarray) -> np.ndarray:
    """Convert.c rgbA2rgba: c' = MULDIV255(c, a)."""
    a = rgba[..., 3:4].astype(np.int64)
    t = rgba[..., :3].astype(np.int64) * a + 128
    out = rgba.copy()
    out[..., :3] = (((t >> 8) + t) >> 8).astype(np.uint8)
    return out


def unpremultiply(rgba: np.ndarray) -> np.ndarray:
    """Convert.c rgba2rgbA: c' = c if a in (0, 255) else CLIP8(255 * c / a)."""
    a = rgba[..., 3:4].astype(np.int64)
    c = rgba[..., :3].astype(np.int64)
    safe = np.where(a == 0, 1, a)
    q = np.clip((255 * c) // safe, 0, 255)
    out = rgba.copy()
    out[..., :3] = np.where((a == 0) | (a == 255), c, q).astype(np.uint8)
    return out


def resize_rgba_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """PIL Image.resize((out_w, out_h), BICUBIC) for an RGBA uint8 array [H, W, 4] -> [out_h, out_w, 4]."""
    h, w, _ = img.shape
    if (w, h) == (out_w, out_h):
        return img.copy()
    x = premultiply(img)
    if w != out_w:                                   # horizontal pass first (Resample.c ImagingResample)
        x = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(x.transpose(1, 0, 2)), out_w).transpose(1, 0, 2))
    if h != out_h:
        x = _resample_axis0(x, out_h)
    return unpremultiply(x)


def resize_normalize(img: np.ndarray, out_w: int = 256, out_h: int = 32) -> np.ndarray:
    """ResizeNormalize.__call__ (data/dataset.py:242-246): resize -> ToTensor (uint8 / 255, CHW fp32) -> sub 0.5, div 0.5."""
    r = resize_rgba_bicubic(img, out_w, out_h)
    t = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)
    return ((t - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def align_collate(images, out_w: int = 256, out_h: int = 32) -> np.ndarray:
    """AlignCollate.__call__ (data/dataset.py:192-197) without augmentation: [B, 4, out_h, out_w] fp32."""
    return np.stack([resize_normalize(im, out_w, out_h) for im in images], 0)
