"""Generate tests/golden/resize_rgba.npz by running Pillow + the reference's ResizeNormalize arithmetic in the build
container (python -m oracle.make_resize_golden).  Images are deterministic pseudo-random RGBA crops of ragged sizes,
with opaque, binary and graded alpha; outputs are Pillow's own resize((256, 32), BICUBIC) bytes."""
import os

import numpy as np
from PIL import Image

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_rgba.npz")
SIZES = [(100, 32), (256, 32), (37, 19), (613, 87), (256, 40), (300, 32), (1200, 48), (5, 3), (255, 31), (2, 2), (1, 40),
         (256, 33), (480, 120), (64, 64), (31, 97), (1900, 60)]          # (width, height)


def make_images(seed=7):
    rng = np.random.default_rng(seed)
    imgs = []
    for k, (w, h) in enumerate(SIZES):
        # smooth-ish content (text crops are not white noise) plus noise
        base = rng.integers(0, 256, size=(max(1, h // 4 + 1), max(1, w // 4 + 1), 4), dtype=np.uint8)
        img = np.kron(base, np.ones((4, 4, 1), dtype=np.uint8))[:h, :w].astype(np.int64)
        img = np.clip(img + rng.integers(-20, 21, size=img.shape), 0, 255).astype(np.uint8)
        if k % 3 == 0:
            img[..., 3] = 255
        elif k % 3 == 1:
            img[..., 3] = rng.choice(np.array([0, 255, 17, 200], dtype=np.uint8), size=(h, w))
        imgs.append(np.ascontiguousarray(img))
    return imgs


if __name__ == "__main__":
    imgs = make_images()
    outs = [np.array(Image.fromarray(im, "RGBA").resize((256, 32), Image.BICUBIC)) for im in imgs]
    np.savez_compressed(OUT, resized=np.stack(outs, 0), sizes=np.array(SIZES), seed=np.int64(7),
                        pillow=np.array(Image.__version__ if hasattr(Image, "__version__") else "unknown"))
    print(OUT, os.path.getsize(OUT) // 1024, "KiB")
