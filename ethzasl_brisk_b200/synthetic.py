"""Seeded synthetic grayscale frames for parity tests and benches.

No datasets are reachable from the build or GPU boxes, so tests and bench.py
use procedurally generated frames: a mid-grey canvas with random filled
rectangles, rotated quads and discs of uniform random intensity, a 3x3 box
blur and additive integer noise.  Frames are
non-periodic, so score ties stay local.
"""
import numpy as np


def _fill_poly(img, pts, val):
    """Scan-line fill of a convex polygon (pts: [k, 2] float x, y)."""
    h, w = img.shape
    y0 = max(int(np.floor(pts[:, 1].min())), 0)
    y1 = min(int(np.ceil(pts[:, 1].max())), h - 1)
    k = len(pts)
    for y in range(y0, y1 + 1):
        xs = []
        for i in range(k):
            (xa, ya), (xb, yb) = pts[i], pts[(i + 1) % k]
            if (ya <= y < yb) or (yb <= y < ya):
                xs.append(xa + (y - ya) * (xb - xa) / (yb - ya))
        if len(xs) >= 2:
            xa, xb = int(max(min(xs), 0)), int(min(max(xs), w - 1))
            if xb >= xa:
                img[y, xa:xb + 1] = val


def synthetic_frame(width, height, seed, n_shapes=None, noise=6):
    """One u8 frame [height, width]."""
    rng = np.random.default_rng(seed)
    if n_shapes is None:
        n_shapes = max(40, (width * height) // 2500)
    img = np.full((height, width), 128, np.float32)
    kinds = rng.integers(0, 3, n_shapes)
    cx = rng.uniform(0, width, n_shapes)
    cy = rng.uniform(0, height, n_shapes)
    sz = rng.uniform(6, max(12, min(width, height) / 8), (n_shapes, 2))
    ang = rng.uniform(0, np.pi, n_shapes)
    val = rng.integers(0, 256, n_shapes)
    yy, xx = None, None
    for i in range(n_shapes):
        if kinds[i] == 0:  # axis-aligned rectangle
            x0, x1 = int(max(cx[i] - sz[i, 0], 0)), int(min(cx[i] + sz[i, 0], width))
            y0, y1 = int(max(cy[i] - sz[i, 1], 0)), int(min(cy[i] + sz[i, 1], height))
            img[y0:y1, x0:x1] = val[i]
        elif kinds[i] == 1:  # rotated quad
            c, s = np.cos(ang[i]), np.sin(ang[i])
            corners = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float64) * sz[i]
            pts = np.stack([cx[i] + corners[:, 0] * c - corners[:, 1] * s,
                            cy[i] + corners[:, 0] * s + corners[:, 1] * c], 1)
            _fill_poly(img, pts, val[i])
        else:  # disc
            r = sz[i, 0] * 0.7
            x0, x1 = int(max(cx[i] - r, 0)), int(min(cx[i] + r + 1, width))
            y0, y1 = int(max(cy[i] - r, 0)), int(min(cy[i] + r + 1, height))
            if x1 > x0 and y1 > y0:
                yy, xx = np.mgrid[y0:y1, x0:x1]
                m = (xx - cx[i]) ** 2 + (yy - cy[i]) ** 2 <= r * r
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
