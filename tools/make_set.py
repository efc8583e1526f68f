#!/usr/bin/env python
"""Write a dataset in the reference's `.set` format from PGM images, with key points and descriptors
computed on the GPU -- what the reference's test-binary-equal.cc produces with its detector / extractor
(brisk/src/test/test-binary-equal.cc:73-89,268-299; format in ethzasl_brisk_b200/setio.py).

  python tools/make_set.py --detector ast    -o ast.set    img1.pgm img2.pgm   # BriskFeatureDetector(70), BRISK2
  python tools/make_set.py --detector harris -o harris.set img1.pgm img2.pgm   # octaves 0, radius 30, absThr 20
  python tools/make_set.py --compare a.set b.set                               # field-by-field diff of two datasets
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ethzasl_brisk_b200 as bb  # noqa: E402


def compare(a, b):
    ea, eb = bb.read_set(a), bb.read_set(b)
    ok = len(ea) == len(eb)
    print(f"{len(ea)} vs {len(eb)} entries")
    for i, (x, y) in enumerate(zip(ea, eb)):
        same_img = np.array_equal(x["image"], y["image"])
        nk = (len(x["keypoints"]), len(y["keypoints"]))
        fields = {}
        if nk[0] == nk[1]:
            for f in bb.KP_DTYPE.names:
                d = np.abs(x["keypoints"][f].astype(np.float64) - y["keypoints"][f].astype(np.float64))
                fields[f] = float(d.max()) if len(d) else 0.0
            bits = int(np.unpackbits(x["descriptors"] ^ y["descriptors"]).sum()) if x["descriptors"].shape == y["descriptors"].shape else -1
        else:
            bits = -1
        print(f"entry {i}: image equal {same_img}, key points {nk}, max |diff| per field {fields}, differing descriptor bits {bits}")
        ok = ok and same_img and nk[0] == nk[1] and bits == 0 and all(v == 0 for k, v in fields.items() if k != "angle")
    return 0 if ok else 1


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--detector", choices=["ast", "harris"], default="ast")
    ap.add_argument("--thresh", type=int, default=70)
    ap.add_argument("--octaves", type=int, default=None)
    ap.add_argument("--radius", type=float, default=30.0)
    ap.add_argument("--abs-thresh", type=float, default=20.0)
    ap.add_argument("--compare", nargs=2, metavar=("A", "B"))
    ap.add_argument("-o", "--output")
    ap.add_argument("images", nargs="*")
    args = ap.parse_args()
    if args.compare:
        return compare(*args.compare)
    if not args.output or not args.images:
        ap.error("need -o OUT.set and at least one PGM image")
    ctx = bb.Context(0)
    if args.detector == "ast":
        det = bb.BriskFeatureDetector(args.thresh, 3 if args.octaves is None else args.octaves, ctx=ctx)
    else:
        det = bb.ScaleSpaceFeatureDetector(0 if args.octaves is None else args.octaves, args.radius, args.abs_thresh, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    entries = []
    for path in args.images:
        img = bb.read_pgm(path)
        kps, desc = ext.compute(img, det.detect(img))
        entries.append(dict(path=path, image=img, keypoints=kps, descriptors=desc, blobs={"testImage": img.tobytes()}))
        print(f"{path}: {img.shape[1]}x{img.shape[0]}, {len(kps)} key points")
    bb.write_set(args.output, entries)
    return 0


if __name__ == "__main__":
    sys.exit(main())
