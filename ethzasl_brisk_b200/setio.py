"""Reader / writer of the reference's serialized test datasets (`*.set`) and of binary PGM images.

Format (little-endian; reference brisk/src/test/serialization.{h,cc}, bench-ds.cc:57-94):
  u32 n_entries; per entry
    string path                      u32 length + bytes
    Mat image                        i32 rows, cols, type, elemSize + raw bytes
    vector<KeyPoint>                 u32 n x {f32 angle, i32 class_id, i32 octave, f32 x, f32 y, f32 response, f32 size}
    Mat descriptors                  N x descriptorSize u8
    map<string, Blob>                u32 n x {string key, u32 size, bytes}
Host-side data plumbing only: nothing here computes features.
"""
import struct

import numpy as np

from .api import KP_DTYPE


class _Reader:
    def __init__(self, buf):
        self.b, self.o = buf, 0

    def take(self, fmt):
        if self.o + struct.calcsize("<" + fmt) > len(self.b):
            raise ValueError("truncated .set file")
        v = struct.unpack_from("<" + fmt, self.b, self.o)
        self.o += struct.calcsize("<" + fmt)
        return v if len(v) > 1 else v[0]

    def raw(self, n):
        if self.o + n > len(self.b):
            raise ValueError("truncated .set file")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def string(self):
        return self.raw(self.take("I")).decode()

    def mat(self):
        rows, cols, typ, esz = self.take("iiii")
        return rows, cols, typ, esz, self.raw(rows * cols * esz)


def read_set(path):
    """-> list of dicts {path, image [h,w] u8, keypoints (KP_DTYPE), descriptors [n,bytes] u8, blobs {key: bytes},
    image_type, descriptor_type} in file order."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    entries = []
    for _ in range(r.take("I")):
        name = r.string()
        rows, cols, typ, esz, data = r.mat()
        if esz != 1:
            raise ValueError("only 8-bit images are supported")
        img = np.frombuffer(data, np.uint8).reshape(rows, cols).copy()
        nk = r.take("I")
        rec = np.frombuffer(r.raw(nk * 28), np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("x", "<f4"),
                                                      ("y", "<f4"), ("response", "<f4"), ("size", "<f4")]))
        kps = np.zeros(nk, KP_DTYPE)
        for f in KP_DTYPE.names:
            kps[f] = rec[f]
        drows, dcols, dtyp, desz, ddata = r.mat()
        desc = np.frombuffer(ddata, np.uint8).reshape(drows, dcols * desz).copy()
        blobs = {}
        for _ in range(r.take("I")):
            key = r.string()
            blobs[key] = bytes(r.raw(r.take("I")))
        entries.append(dict(path=name, image=img, keypoints=kps, descriptors=desc, blobs=blobs, image_type=typ,
                            descriptor_type=dtyp))
    if r.o != len(r.b):
        raise ValueError("trailing bytes in .set file")
    return entries


def write_set(path, entries):
    """Inverse of read_set (image_type / descriptor_type default to CV_8UC1 = 0; blob order is kept)."""
    out = [struct.pack("<I", len(entries))]

    def string(s):
        b = s.encode()
        out.append(struct.pack("<I", len(b)) + b)

    def mat(a, typ):
        a = np.ascontiguousarray(a, np.uint8)
        if a.ndim != 2:
            raise ValueError("matrices must be 2-D u8")
        out.append(struct.pack("<iiii", a.shape[0], a.shape[1], typ, 1) + a.tobytes())

    for e in entries:
        string(e["path"])
        mat(e["image"], e.get("image_type", 0))
        kps = np.ascontiguousarray(e["keypoints"], KP_DTYPE)
        rec = np.zeros(len(kps), np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("x", "<f4"), ("y", "<f4"),
                                           ("response", "<f4"), ("size", "<f4")]))
        for f in KP_DTYPE.names:
            rec[f] = kps[f]
        out.append(struct.pack("<I", len(kps)) + rec.tobytes())
        desc = np.ascontiguousarray(e["descriptors"], np.uint8)
        if len(desc) != len(kps):
            raise ValueError("one descriptor row per key point")
        mat(desc.reshape(len(kps), -1) if len(kps) else desc.reshape(0, 0), e.get("descriptor_type", 0))
        blobs = e.get("blobs", {})
        out.append(struct.pack("<I", len(blobs)))
        for key, val in blobs.items():
            string(key)
            out.append(struct.pack("<I", len(val)) + bytes(val))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def read_pgm(path):
    """Binary (P5) 8-bit PGM -> [h, w] u8 (the reference's test images, brisk/src/test/image-io.cc)."""
    with open(path, "rb") as f:
        buf = f.read()
    tokens, o = [], 0
    while len(tokens) < 4:
        while buf[o:o + 1].isspace():
            o += 1
        if buf[o:o + 1] == b"#":
            o = buf.index(b"\n", o) + 1
            continue
        e = o
        while not buf[e:e + 1].isspace():
            e += 1
        tokens.append(buf[o:e])
        o = e
    if tokens[0] != b"P5" or int(tokens[3]) > 255:
        raise ValueError("not a binary 8-bit PGM")
    w, h = int(tokens[1]), int(tokens[2])
    return np.frombuffer(buf, np.uint8, w * h, o + 1).reshape(h, w).copy()


def write_pgm(path, img):
    img = np.ascontiguousarray(img, np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]) + img.tobytes())
