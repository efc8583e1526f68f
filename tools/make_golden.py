#!/usr/bin/env python
"""Extract the reference's own golden fixtures into small committed files.

Reads (build container only) the serialized datasets that the reference's
test-binary-equal.cc compares against:
  /root/reference/brisk/src/test/test_data/brisk_verification_{ast,harris}.set
(format: reference brisk/src/test/serialization.{h,cc}, bench-ds.cc:57-94;
SURVEY.md Appendix C) and writes tests/golden/brisk_verification.npz holding,
per entry, the 800x640 input image, the keypoints (x, y, size, angle, response,
octave, class_id) and the 48-byte descriptors.  The GPU box has no
/root/reference, so tests read only the npz.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np

REF = Path("/root/reference/brisk/src/test/test_data")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "brisk_verification.npz"


def parse_set(path):
    from ethzasl_brisk_b200.setio import read_set
    return [(e["path"], e["image"], e["keypoints"], e["descriptors"]) for e in read_set(path)]


def main():
    out = {}
    for kind in ("ast", "harris"):
        for i, (name, img, kps, desc) in enumerate(parse_set(REF / f"brisk_verification_{kind}.set")):
            print(kind, i, name, img.shape, len(kps), desc.shape)
            out[f"{kind}{i}_image"] = img
            out[f"{kind}{i}_kps"] = kps
            out[f"{kind}{i}_desc"] = desc
    # both datasets hold the same two images; store each once.
    for i in range(2):
        assert np.array_equal(out[f"ast{i}_image"], out[f"harris{i}_image"])
        out[f"image{i}"] = out.pop(f"ast{i}_image")
        out.pop(f"harris{i}_image")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())
