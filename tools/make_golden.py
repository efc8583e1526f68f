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
import struct
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference/brisk/src/test/test_data")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "brisk_verification.npz"


class Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, fmt):
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def string(self):
        return self.raw(self.take("I")).decode()

    def mat(self):
        rows, cols, typ, esz = self.take("iiii")
        data = self.raw(rows * cols * esz)
        return rows, cols, typ, esz, data


def parse_set(path):
    r = Reader(path.read_bytes())
    entries = []
    for _ in range(r.take("I")):
        name = r.string()
        rows, cols, typ, esz, data = r.mat()
        assert typ == 0 and esz == 1
        img = np.frombuffer(data, np.uint8).reshape(rows, cols).copy()
        nk = r.take("I")
        kps = np.zeros(nk, dtype=[("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                                  ("response", "f4"), ("octave", "i4"), ("class_id", "i4")])
        for i in range(nk):
            angle, class_id, octave, x, y, response, size = r.take("fiiffff")
            kps[i] = (x, y, size, angle, response, octave, class_id)
        drows, dcols, dtyp, desz, ddata = r.mat()
        desc = np.frombuffer(ddata, np.uint8).reshape(drows, dcols).copy()
        for _ in range(r.take("I")):
            r.string()
            r.raw(r.take("I"))
        entries.append((name, img, kps, desc))
    assert r.o == len(r.b), (r.o, len(r.b))
    return entries


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
