#!/usr/bin/env python
"""Angle / rotation-bin agreement of the CUDA descriptor path with the unmodified reference over >= 10^6 key points
(SURVEY.md H3).  Test infrastructure (runs oracle/_ref), GPU box only:

  python tests/angle_stats.py [--frames 160] > gpurun_out/angle_stats.json

The device computes the orientation with a double atan2 of its own (CUDA libm) where the reference calls glibc's; the
angle is stored as float and the rotation bin is derived from that float (brisk-descriptor-extractor.cc:732-739).  For every
key point of `--frames` distinct synthetic 1080p frames: is the float angle bit-identical, is the rotation bin identical,
is the descriptor identical.
"""
import argparse
import json
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ethzasl_brisk_b200 as bb  # noqa: E402
from oracle import ref  # noqa: E402


def theta_bin(angle):
    t = np.floor(np.float64(np.float32(1024.0) * angle.astype(np.float32)) / 360.0 + 0.5).astype(np.int64)   # int(x + 0.5), x may be negative
    t = np.where(np.float64(np.float32(1024.0) * angle.astype(np.float32)) / 360.0 + 0.5 < 0, np.ceil(np.float64(np.float32(1024.0) * angle.astype(np.float32)) / 360.0 + 0.5).astype(np.int64), t)
    t = np.where(t < 0, t + 1024, t)
    return np.where(t >= 1024, t - 1024, t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=160)
    ap.add_argument("--threads", type=int, default=16)
    a = ap.parse_args()
    ctx = bb.Context(0)
    det, ext = bb.BriskFeatureDetector(60, 4, ctx=ctx), bb.BriskDescriptorExtractor(ctx=ctx)
    total = same_angle = same_bin = same_desc = 0
    max_diff = 0.0
    batch = 16
    for b0 in range(0, a.frames, batch):
        seeds = list(range(7000 + b0, 7000 + min(b0 + batch, a.frames)))
        with ThreadPoolExecutor(a.threads) as ex:
            frames = np.stack(list(ex.map(lambda s: bb.synthetic_frame(1920, 1080, s), seeds)))
        kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=16384)

        def one(i):
            k = ref.agast_detect(frames[i], 60, 4, cap=1 << 19)
            return ref.describe(frames[i], k)
        with ThreadPoolExecutor(a.threads) as ex:
            refs = list(ex.map(one, range(len(seeds))))
        for i, (k2, d2) in enumerate(refs):
            n = int(counts[i])
            assert n == len(k2), "key point sets differ"
            a1, a2 = kps[i, :n]["angle"], k2["angle"]
            total += n
            same_angle += int(np.sum(a1.view(np.uint32) == a2.view(np.uint32)))
            same_bin += int(np.sum(theta_bin(a1) == theta_bin(a2)))
            same_desc += int(np.sum(np.all(desc[i, :n] == d2, axis=1)))
            if n:
                max_diff = max(max_diff, float(np.abs(a1.astype(np.float64) - a2.astype(np.float64)).max()))
    print(json.dumps({"frames": a.frames, "keypoints": total, "angle_bit_identical": same_angle, "angle_differs": total - same_angle,
                      "max_abs_angle_difference_deg": max_diff, "rotation_bin_identical": same_bin, "rotation_bin_differs": total - same_bin,
                      "descriptor_identical": same_desc, "descriptor_differs": total - same_desc,
                      "against": "oracle/_ref (unmodified reference), AGAST(60,4) + BRISK2 on distinct synthetic 1920x1080 frames"}))


if __name__ == "__main__":
    main()
