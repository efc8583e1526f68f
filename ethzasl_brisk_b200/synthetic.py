"""Seeded synthetic grayscale frames for parity tests and benches.

No datasets are reachable from the build or GPU boxes, so tests and bench.py
use procedurally generated frames: a mid-grey canvas with random filled
rectangles, rotated quads and discs of uniform random intensity, a 3x3 box
blur and additive integer noise.  Frames are
non-periodic, so score ties stay local.
"""
import numpy as np


def synthetic_frame(width, height, seed, n_shapes=None, noise=6, max_size=None):
    """One u8 frame [height, width].

    Defaults give roughly 3.5 detected key points per 1000 pixels at AGAST
    threshold 60 (about 6-7 k on a 1080p frame, like the reference's sample
    image tiled to that size).
    """
    rng = np.random.default_rng(seed)
    if n_shapes is None:
        n_shapes = max(40, (width * height) // 1300)
    if max_size is None:
        max_size = 30.0
    img = np.full((height, width), 128, np.float32)
    kinds = rng.integers(0, 3, n_shapes)
    cx = rng.uniform(0, width, n_shapes)
    cy = rng.uniform(0, height, n_shapes)
    sz = rng.uniform(3, max_size, (n_shapes, 2))
    ang = rng.uniform(0, np.pi, n_shapes)
    val = rng.integers(0, 256, n_shapes)
    for i in range(n_shapes):
        if kinds[i] == 0:  # axis-aligned rectangle
            x0, x1 = int(max(cx[i] - sz[i, 0], 0)), int(min(cx[i] + sz[i, 0], width))
            y0, y1 = int(max(cy[i] - sz[i, 1], 0)), int(min(cy[i] + sz[i, 1], height))
            img[y0:y1, x0:x1] = val[i]
            continue
        r = float(np.hypot(sz[i, 0], sz[i, 1])) if kinds[i] == 1 else sz[i, 0] * 0.7
        x0, x1 = int(max(cx[i] - r, 0)), int(min(cx[i] + r + 1, width))
        y0, y1 = int(max(cy[i] - r, 0)), int(min(cy[i] + r + 1, height))
        if x1 <= x0 or y1 <= y0:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        dx, dy = xx - cx[i], yy - cy[i]
        if kinds[i] == 1:  # rotated rectangle
            c, s_ = np.cos(ang[i]), np.sin(ang[i])
            m = (np.abs(dx * c + dy * s_) <= sz[i, 0]) & (np.abs(-dx * s_ + dy * c) <= sz[i, 1])
        else:  # disc
            m = dx * dx + dy * dy <= r * r
        img[y0:y1, x0:x1][m] = val[i]
    # 3x3 box blur (edge replicated)
    p = np.pad(img, 1, mode="edge")
    img = sum(p[dy:dy + height, dx:dx + width] for dy in range(3) for dx in range(3)) / 9.0
    if noise:
        img = img + rng.integers(-noise, noise + 1, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_batch(n, width, height, seed0, **kw):
    """[n, height, width] u8, frame i seeded with seed0 + i."""
    return np.stack([synthetic_frame(width, height, seed0 + i, **kw) for i in range(n)])


def random_descriptors(n, nbytes, seed):
    """Uniform random descriptor rows [n, nbytes] u8."""
    return np.random.default_rng(seed).integers(0, 256, (n, nbytes), dtype=np.uint8)
