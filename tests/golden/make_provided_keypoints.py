#!/usr/bin/env python
"""Golden vectors for the provided-key-point modes, generated with the UNMODIFIED reference compiled into
oracle/_ref (build container only: needs /root/reference via `make -C oracle ref`):

  * BriskFeatureDetector(thresh, octaves).ComputeScale(image, keypoints)     (brisk-feature-detector.cc:87-92)
  * ScaleSpaceFeatureDetector<HarrisScoreCalculator>(0, radius, 0, maxNumKpt).detect(image, keypoints) with a
    non-empty vector ("use passed key points", scale-space-feature-detector.h:103-108)

on image 1 of the reference's own fixtures (tests/golden/brisk_verification.npz, `image0`) with seeded input lists.
Writes tests/golden/provided_keypoints.npz; tests compare the oracle (CPU) and the CUDA path (GPU) with it.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle import ref  # noqa: E402
from test_oracle_golden import _passed_points, _provided_points  # noqa: E402

COMPUTE_SCALE_CASES = [(60, 4, 2000, 21, False), (70, 3, 500, 22, True), (30, 5, 40, 5, False), (70, 0, 500, 22, True)]
PASSED_CASES = [(30.0, -1, 3000, 1, False), (3.0, 150, 700, 2, True), (0.0, 400, 3000, 1, False)]


def main():
    img = np.load(ROOT / "tests" / "golden" / "brisk_verification.npz")["image0"]
    out = {}
    for i, (thresh, octaves, n, seed, integer) in enumerate(COMPUTE_SCALE_CASES):
        k = _provided_points(img, n, seed, integer)
        res = ref.compute_scale(img, k, thresh, octaves)
        out[f"cs{i}_in"], out[f"cs{i}_out"] = k, res
        print(f"ComputeScale thresh {thresh} octaves {octaves}: {n} points in, {len(res)} key points out")
    for i, (radius, max_kpt, n, seed, ties) in enumerate(PASSED_CASES):
        k = _passed_points(img.shape, n, seed, ties)
        res = ref.harris_detect_passed(img, k, radius, max_kpt)
        out[f"hp{i}_in"], out[f"hp{i}_out"] = k, res
        print(f"passed key points radius {radius} maxNumKpt {max_kpt}: {n} points in, {len(res)} out")
    path = ROOT / "tests" / "golden" / "provided_keypoints.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())
