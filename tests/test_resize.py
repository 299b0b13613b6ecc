"""Image preparation (data/dataset.py:169-197,235-246): the numpy oracle is pinned bit-exactly to
Pillow (committed fixture + live Pillow when importable); the CUDA path (mrn_b200.data, through the C ABI) must
reproduce the same bytes and the same fp32 tensor."""
import os

import numpy as np
import pytest
import torch

from oracle import resize_oracle as R
from oracle.make_resize_golden import SIZES, make_images
from conftest import load_golden


def test_oracle_matches_committed_pillow_outputs_bit_exactly():
    g = load_golden("resize_rgba")
    imgs = make_images(int(g["seed"]))
    assert [tuple(s) for s in g["sizes"]] == list(SIZES)
    for k, im in enumerate(imgs):
        got = R.resize_rgba_bicubic(im, 256, 32)
        assert np.array_equal(got, g["resized"][k]), SIZES[k]


def test_oracle_matches_live_pillow_and_torch_normalisation():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    for (w, h) in [(77, 21), (256, 32), (410, 64), (3, 200), (256, 16), (1, 1), (1, 300), (300, 1), (3900, 10), (40, 1000),
                   (33, 999), (257, 33)]:
        im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        if (w + h) % 2:
            im[..., 3] = rng.integers(0, 4, size=(h, w)) * 85
        ref = np.array(Image.fromarray(im, "RGBA").resize((256, 32), Image.BICUBIC))
        assert np.array_equal(R.resize_rgba_bicubic(im, 256, 32), ref), (w, h)
        t = torch.from_numpy(ref).permute(2, 0, 1).contiguous().to(torch.float32).div(255)      # torchvision ToTensor
        t.sub_(0.5).div_(0.5)                                                                    # data/dataset.py:245
        assert np.array_equal(R.resize_normalize(im, 256, 32), t.numpy()), (w, h)


def test_tap_tables_follow_resample_c():
    # identity when sizes agree; every row of taps sums to 2^22 +- rounding; antialiased support grows with the shrink factor
    b, kk, ks = R.precompute_coeffs(32, 32)
    assert ks == 5 and all(int(kk[i].sum()) == 1 << 22 and int(kk[i].max()) == 1 << 22 for i in range(32))
    b, kk, ks = R.precompute_coeffs(1000, 256)
    assert ks == int(np.ceil(2 * 1000 / 256)) * 2 + 1 and abs(int(kk[100].sum()) - (1 << 22)) <= ks


@pytest.mark.gpu
def test_cuda_resize_normalize_is_bit_identical_to_pillow():
    from mrn_b200 import data
    g = load_golden("resize_rgba")
    imgs = make_images(int(g["seed"]))
    out = data.resize_normalize_batch(imgs, 256, 32)
    torch.cuda.synchronize()
    assert out.shape == (len(imgs), 4, 32, 256) and out.dtype == torch.float32
    got = out.cpu().numpy()
    for k, im in enumerate(imgs):
        ref_u8 = g["resized"][k]                                               # Pillow's bytes
        ref = ((ref_u8.transpose(2, 0, 1).astype(np.float32) / np.float32(255)) - np.float32(0.5)) / np.float32(0.5)
        back = np.rint((got[k] * 0.5 + 0.5) * 255).astype(np.int64).transpose(1, 2, 0)
        assert np.array_equal(back, ref_u8.astype(np.int64)), ("bytes", SIZES[k], int(np.abs(back - ref_u8).max()))
        assert np.array_equal(got[k], ref), ("fp32", SIZES[k])
    # the oracle agrees on a fresh ragged batch, and the collate interface returns device tensors + labels
    rng = np.random.default_rng(11)
    fresh = [rng.integers(0, 256, size=(int(rng.integers(8, 90)), int(rng.integers(8, 700)), 4), dtype=np.uint8) for _ in range(9)]
    out2 = data.resize_normalize_batch(fresh, 256, 32).cpu().numpy()
    assert np.array_equal(out2, R.align_collate(fresh, 256, 32))
    import argparse
    coll = data.AlignCollate(argparse.Namespace(imgW=256, imgH=32, Aug="None"), mode="test")
    t, labels = coll([(im, "x%d" % i) for i, im in enumerate(fresh)])
    assert t.is_cuda and labels == tuple("x%d" % i for i in range(9)) and np.array_equal(t.cpu().numpy(), out2)


@pytest.mark.gpu
def test_cuda_resize_rejects_unsupported_shrink_factor():
    from mrn_b200 import data
    with pytest.raises(RuntimeError, match="shrinks"):
        data.resize_normalize_batch([np.zeros((8, 9000, 4), dtype=np.uint8)], 256, 32)


@pytest.mark.gpu
def test_cuda_resize_randomised_sizes_and_extremes_match_the_oracle():
    """Forty-odd ragged sizes incl. the extremes the tap tables allow (1 x 1, one-pixel-wide / -high strips, 3900-pixel
    wide lines that shrink 15x, 1000-pixel tall crops, the identity size, constant and saturated images, every alpha
    pattern): every output byte equals the Pillow-pinned oracle."""
    from mrn_b200 import data
    rng = np.random.default_rng(2025)
    sizes = [(1, 1), (1, 300), (300, 1), (3900, 10), (40, 1000), (256, 32), (255, 32), (256, 31), (257, 33), (512, 64),
             (128, 16), (2, 2), (7, 5), (1024, 128), (33, 999)]
    sizes += [(int(rng.integers(1, 1200)), int(rng.integers(1, 200))) for _ in range(28)]
    imgs = []
    for k, (w, h) in enumerate(sizes):
        kind = k % 5
        if kind == 0:
            im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        elif kind == 1:
            im = np.full((h, w, 4), 255, dtype=np.uint8)
        elif kind == 2:
            im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8); im[..., 3] = 255
        elif kind == 3:
            im = rng.integers(0, 2, size=(h, w, 4), dtype=np.uint8) * 255
        else:
            im = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8); im[..., 3] = rng.integers(0, 4, size=(h, w)) * 85
        imgs.append(np.ascontiguousarray(im))
    got = data.resize_normalize_batch(imgs, 256, 32).cpu().numpy()
    for k, im in enumerate(imgs):
        ref = R.resize_normalize(im, 256, 32)
        assert np.array_equal(got[k], ref), (sizes[k], float(np.abs(got[k] - ref).max()))
