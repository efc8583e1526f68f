"""Small workload that launches every kernel family of libbrisk_b200.so once, for compute-sanitizer
(memcheck / racecheck / initcheck are 10-100x slower than a plain run, so the sizes are tiny):

    compute-sanitizer --tool memcheck  python tools/sanitize_workload.py
    compute-sanitizer --tool racecheck python tools/sanitize_workload.py

Prints one line per step; the sanitizer's own summary follows on stderr."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ethzasl_brisk_b200 as bb  # noqa: E402


def main():
    ctx = bb.Context(0)
    frames = np.stack([bb.synthetic_frame(320, 240, 11 + i) for i in range(3)])
    det = bb.BriskFeatureDetector(60, 3, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=4096)
    print("agast + brisk2:", counts.tolist())
    ext1 = bb.BriskDescriptorExtractor(version=1, ctx=ctx)
    k1, c1, d1 = ext1.compute_batch(frames, kps, counts)
    print("brisk1 describe:", c1.tolist(), d1.shape[-1])
    n0 = int(counts[0])
    print("compute_scale:", len(det.compute_scale(frames[0], kps[0, :n0])))
    mask = np.zeros(frames[0].shape, np.uint8); mask[:, :160] = 255
    print("masked detect:", len(det.detect(frames[0], mask=mask)))
    det0 = bb.BriskFeatureDetector(40, 0, suppressScaleNonmaxima=False, ctx=ctx)
    print("single layer, no scale suppression:", len(det0.detect(frames[1])))
    for radius in (20.0, 0.0):
        har = bb.ScaleSpaceFeatureDetector(2, radius, 0.0, maxNumKpt=300 if radius <= 0 else None, ctx=ctx)
        har.set_corner_capacity(80000)
        hk = har.detect(frames[2])
        print(f"harris scale space radius {radius}:", len(hk))
    har0 = bb.ScaleSpaceFeatureDetector(0, 10.0, ctx=ctx)
    har0.set_corner_capacity(80000)
    hk0 = har0.detect(frames[2])
    print("harris passed key points:", len(har0.detect(frames[2], keypoints=hk0)))
    sq = np.ascontiguousarray(frames[0][:240, :240])
    leg = bb.HarrisFeatureDetector(8.0, ctx=ctx)
    leg.set_corner_capacity(80000)
    print("legacy harris:", len(leg.detect(sq)))
    calc = bb.HarrisScoreCalculator(ctx=ctx); calc.SetImage(frames[0])
    print("harris score calculator maxima:", len(calc.Get2dMaxima(0)))
    img16 = (frames[0].astype(np.uint16) * 200)
    print("16-bit samplers:", ctx.halfsample16(img16).shape, ctx.twothirdsample16(img16[:, :318]).shape)
    ctx.debug_integral(frames[0])
    m = bb.BruteForceMatcher(ctx=ctx)
    for nb in (48, 64):
        q, t = bb.random_descriptors(300, nb, 1), bb.random_descriptors(9000, nb, 2)
        for variant in (0, 1, 2, 3):
            ctx.set_knn_variant(variant)
            idx, dist = m.knn(q, t, 2)
        print(f"knn {nb} bytes, four kernels:", int(dist.sum()))
    ctx.set_knn_variant(3)
    q, t = bb.random_descriptors(100, 64, 3), bb.random_descriptors(700, 64, 4)
    idx, dist = m.knn(q, t, 5)
    tm = (np.arange(100)[:, None] + np.arange(700)[None, :]) % 3 != 0
    idx, dist = m.knn(q, t, 3, mask=tm.astype(np.uint8))
    offsets, ridx, rdist = m.radius(q, t, 250.0)
    print("knn k=5, masked knn, radius:", int(dist[dist < 1e9].sum()), int(offsets[-1]))
    h = bb.Hamming(ctx=ctx)
    print("hamming pairs:", int(np.sum(h.pairs(q, t[:100]))))
    ctx.sync()
    print("workload done")


if __name__ == "__main__":
    main()
