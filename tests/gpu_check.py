#!/usr/bin/env python
"""Verbose stage-by-stage parity check of the CUDA path against the oracle
(run on a GPU box: `python tests/gpu_check.py`; test infrastructure, not collected by pytest).  Prints diffs instead of
asserting, for debugging; the asserting versions live in tests/."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ethzasl_brisk_b200 as bb  # noqa: E402
from oracle import restate  # noqa: E402


def kpeq(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in a.dtype.names)


def main():
    g = np.load(Path(__file__).resolve().parent.parent / "tests" / "golden" / "brisk_verification.npz")
    ctx = bb.Context(0, timing=True)
    imgs = [g["image0"], g["image1"], bb.synthetic_frame(752, 480, 1000), bb.synthetic_frame(500, 333, 3), bb.synthetic_frame(1920, 1080, 2000)]
    ok_all = True
    for ii, img in enumerate(imgs):
        h, w = img.shape
        # pyramid
        for octaves in (4, 1):
            mine = ctx.debug_pyramid(img, octaves)
            ref, _ = restate.pyramid(img, octaves)
            oks = [a.shape == b.shape and np.array_equal(a, b) for a, b in zip(mine, ref)]
            print(f"img{ii} {w}x{h} pyramid oct={octaves}", oks)
            ok_all &= all(oks)
            for li, (a, b) in enumerate(zip(mine, ref)):
                if a.shape == b.shape and not np.array_equal(a, b):
                    d = np.argwhere(a != b)
                    print("   layer", li, "ndiff", len(d), "first", d[:5].tolist(), "cols", sorted(set(d[:, 1].tolist()))[:20])
        ok = np.array_equal(ctx.debug_integral(img), restate.integral8(img))
        print(f"img{ii} integral", ok)
        ok_all &= ok
        # corners per layer
        det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
        c, lc = det.debug_corners(img)
        pyr, _ = restate.pyramid(img, 4)
        off = 0
        for li, layer in enumerate(pyr):
            _, rc = restate.layer_dump(layer, 60)
            mine = c[off:off + lc[li]]
            off += lc[li]
            ok = mine.shape == rc.shape and np.array_equal(mine, rc)
            print(f"img{ii} corners layer {li}: {len(mine)} vs {len(rc)} {ok}")
            ok_all &= ok
        for (t, o) in [(60, 4), (70, 3), (40, 2), (70, 0), (60, 1)]:
            det = bb.BriskFeatureDetector(t, o, ctx=ctx)
            t0 = time.time()
            a = det.detect(img)
            dt = time.time() - t0
            b = restate.agast_detect(img, t, o)
            ok = kpeq(a, b)
            print(f"img{ii} detect thr={t} oct={o}: {len(a)} vs {len(b)} {ok}  ({dt * 1e3:.1f} ms)")
            ok_all &= ok
            if not ok:
                sa = {(x["x"].tobytes(), x["y"].tobytes(), int(x["octave"])) for x in a}
                sb = {(x["x"].tobytes(), x["y"].tobytes(), int(x["octave"])) for x in b}
                print("    only mine", len(sa - sb), "only oracle", len(sb - sa))
        det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
        kp = restate.agast_detect(img, 60, 4)
        for (rot, sc, ver, ps) in [(True, True, 2, 1.0), (True, True, 1, 1.0), (False, True, 2, 1.0), (True, False, 2, 1.0), (True, True, 2, 0.5)]:
            ext = bb.BriskDescriptorExtractor(rot, sc, ver, ps, ctx=ctx)
            k1, d1 = ext.compute(img, kp)
            k2, d2 = restate.describe(img, kp, rot, sc, ver, ps)
            okk = kpeq(k1, k2)
            okd = d1.shape == d2.shape and np.array_equal(d1, d2)
            print(f"img{ii} describe rot={rot} scale={sc} v{ver} ps={ps}: {len(k1)} vs {len(k2)} kps {okk} desc {okd}")
            ok_all &= okk and okd
            if len(k1) == len(k2) and not okk:
                for f in k1.dtype.names:
                    if not np.array_equal(k1[f], k2[f]):
                        print("     field", f, "ndiff", int((k1[f] != k2[f]).sum()), "max abs", float(np.abs(k1[f] - k2[f]).max()))
            if d1.shape == d2.shape and not okd:
                bad = np.where((d1 != d2).any(1))[0]
                print("     desc rows differing", len(bad), bad[:10])
        ext = bb.BriskDescriptorExtractor(ctx=ctx)
        kps, counts, desc = bb.detect_and_compute_batch(det, ext, img)
        k2, d2 = restate.describe(img, restate.agast_detect(img, 60, 4))
        ok = kpeq(kps[0, :counts[0]], k2) and np.array_equal(desc[0, :counts[0]], d2)
        print(f"img{ii} detect+describe fused: {counts[0]} vs {len(k2)} {ok}")
        ok_all &= ok
        print("   timing", {k: round(v, 3) for k, v in ctx.last_timing()[0].items()}, ctx.last_timing()[1])
    # kNN
    m = bb.BruteForceMatcher(ctx=ctx)
    for nb in (48, 64):
        q = bb.random_descriptors(700, nb, 5)
        t = bb.random_descriptors(5000, nb, 6)
        t[100] = q[3]; t[200] = q[3]  # exact ties
        for k in (1, 2, 3):
            i1, d1 = m.knn(q, t, k)
            i2, d2 = restate.knn(q, t, k)
            ok = np.array_equal(i1, i2) and np.array_equal(d1, d2)
            print(f"knn {nb}B k={k}", ok)
            ok_all &= ok
    print("ALL OK" if ok_all else "SOME FAILED")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
