"""ctypes binding of oracle/_ref/libbrisk_ref.so -- the UNMODIFIED reference
sources compiled against oracle/shim (TEST INFRASTRUCTURE ONLY).

The library is built in the build container (oracle/Makefile, target `ref`,
needs /root/reference) and travels to the GPU box as a prebuilt artefact.
"""
import ctypes as C
from pathlib import Path

import numpy as np

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                     ("response", "f4"), ("octave", "i4"), ("class_id", "i4")])
assert KP_DTYPE.itemsize == 28

_LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libbrisk_ref.so"
_lib = None


def available():
    return _LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} missing: run `make -C oracle ref` in the build container")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.ref_bench_detect_describe.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _img(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img, img.shape[1], img.shape[0]


def halfsample8(img):
    img, w, h = _img(img)
    out = np.zeros((h // 2, w // 2), np.uint8)
    lib().ref_halfsample8(_p(img), w, h, _p(out))
    return out


def twothirdsample8(img):
    img, w, h = _img(img)
    out = np.zeros((2 * (h // 3), 2 * (w // 3)), np.uint8)
    lib().ref_twothirdsample8(_p(img), w, h, _p(out))
    return out


def halfsample16(img):
    img = np.ascontiguousarray(img, np.uint16)
    h, w = img.shape
    out = np.zeros((h // 2, w // 2), np.uint16)
    lib().ref_halfsample16(_p(img), w, h, _p(out))
    return out


def twothirdsample16(img):
    img = np.ascontiguousarray(img, np.uint16)
    h, w = img.shape
    out = np.zeros((2 * (h // 3), 2 * (w // 3)), np.uint16)
    lib().ref_twothirdsample16(_p(img), w, h, _p(out))
    return out


def layer_dump(img, thresh, lower=10, cap=1 << 20):
    """-> (thrmap u8 HxW, corners int32 [n,3] = x, y, score)"""
    img, w, h = _img(img)
    thr = np.zeros((h, w), np.uint8)
    c = np.zeros((cap, 3), np.int32)
    n = lib().ref_layer_dump(_p(img), w, h, int(thresh), int(lower), _p(thr), _p(c), cap)
    assert n <= cap
    return thr, c[:n].copy()


def dense_scores(img):
    img, w, h = _img(img)
    a = np.zeros((h, w), np.uint8)
    b = np.zeros((h, w), np.uint8)
    lib().ref_dense_scores(_p(img), w, h, _p(a), _p(b))
    return a, b


def pyramid(img, octaves):
    img, w, h = _img(img)
    nl = max(1, 2 * octaves)
    dims = np.zeros((nl, 2), np.int32)
    so = np.zeros((nl, 2), np.float32)
    buf = np.zeros(w * h * 3, np.uint8)
    n = lib().ref_pyramid(_p(img), w, h, int(octaves), _p(buf), _p(dims), _p(so))
    out, off = [], 0
    for i in range(n):
        cw, ch = int(dims[i, 0]), int(dims[i, 1])
        out.append(buf[off:off + cw * ch].reshape(ch, cw).copy())
        off += cw * ch
    return out, so[:n]


def agast_detect(img, thresh, octaves=3, suppress=True, mask=None, cap=1 << 18):
    img, w, h = _img(img)
    kps = np.zeros(cap, KP_DTYPE)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint8)
    n = lib().ref_agast_detect(_p(img), w, h, int(thresh), int(octaves), int(bool(suppress)), _p(mask), _p(kps), cap)
    assert n <= cap
    return kps[:n].copy()


def compute_scale(img, kps, thresh, octaves=3, suppress=True, cap=1 << 18):
    """BriskFeatureDetector(thresh, octaves, suppress).ComputeScale(img, kps) -> key points"""
    img, w, h = _img(img)
    k = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().ref_compute_scale(_p(img), w, h, int(thresh), int(octaves), int(bool(suppress)), _p(k), len(k), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def harris_detect(img, octaves, radius, abs_thr=0.0, max_kpt=-1, cap=1 << 18):
    img, w, h = _img(img)
    kps = np.zeros(cap, KP_DTYPE)
    n = lib().ref_harris_detect(_p(img), w, h, int(octaves), C.c_double(radius), C.c_double(abs_thr),
                                C.c_int64(max_kpt), _p(kps), cap)
    assert n <= cap
    return kps[:n].copy()


def harris_detect_passed(img, kps, radius, max_kpt=-1, cap=1 << 18):
    """ScaleSpaceFeatureDetector<HarrisScoreCalculator>(0, radius, 0, max_kpt).detect(img, kps) with kps non-empty"""
    img, w, h = _img(img)
    k = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().ref_harris_detect_passed(_p(img), w, h, 0, C.c_double(radius), C.c_double(0.0), C.c_int64(max_kpt), _p(k), len(k),
                                       _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def harris_legacy(img, radius, cap=1 << 18, with_scores=False):
    """brisk::HarrisFeatureDetector(radius).detect(img) (the legacy single-scale detector) -> key points [, score map]"""
    img, w, h = _img(img)
    kps = np.zeros(cap, KP_DTYPE)
    sc = np.zeros((h, w), np.int32) if with_scores else None
    n = lib().ref_harris_legacy(_p(img), w, h, C.c_double(radius), _p(sc), _p(kps), cap)
    assert n <= cap
    return (kps[:n].copy(), sc) if with_scores else kps[:n].copy()


def harris_scores(img):
    img, w, h = _img(img)
    out = np.zeros((h, w), np.int32)
    lib().ref_harris_scores(_p(img), w, h, _p(out))
    return out


def harris_maxima(img, abs_thr, cap=1 << 20):
    img, w, h = _img(img)
    out = np.zeros((cap, 3), np.int32)
    n = lib().ref_harris_maxima(_p(img), w, h, int(abs_thr), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def integral8(img):
    img, w, h = _img(img)
    out = np.zeros((h + 1, w + 1), np.int32)
    rc = lib().ref_integral8(_p(img), w, h, _p(out))
    assert rc == 0
    return out


def describe(img, kps, rot=True, scale=True, version=2, pattern_scale=1.0):
    """-> (surviving keypoints with angle written, descriptors [n, 48|64])"""
    img, w, h = _img(img)
    k = np.ascontiguousarray(kps, KP_DTYPE).copy()
    nb = C.c_int32(0)
    desc = np.zeros((max(len(k), 1), 256), np.uint8)
    flat = np.zeros(max(len(k), 1) * 256, np.uint8)
    n = lib().ref_describe(_p(img), w, h, _p(k), len(k), int(rot), int(scale), int(version),
                           C.c_float(pattern_scale), _p(flat), C.byref(nb))
    del desc
    return k[:n].copy(), flat[:n * nb.value].reshape(n, nb.value).copy()


def pattern_dump(version=2, pattern_scale=1.0):
    counts = np.zeros(4, np.int32)
    lib().ref_pattern_dump(int(version), C.c_float(pattern_scale), _p(counts), None, None, None, None, None)
    P, ns, nl, strings = (int(v) for v in counts)
    pts = np.zeros((64, 1024, P, 3), np.float32)
    scale_list = np.zeros(64, np.float32)
    size_list = np.zeros(64, np.uint32)
    sp = np.zeros((ns, 2), np.uint32)
    lp = np.zeros((nl, 4), np.int32)
    lib().ref_pattern_dump(int(version), C.c_float(pattern_scale), _p(counts), _p(pts), _p(scale_list),
                           _p(size_list), _p(sp), _p(lp))
    return dict(points=P, strings=strings, pts=pts, scale_list=scale_list, size_list=size_list,
                short_pairs=sp, long_pairs=lp)


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_hamming(_p(a), _p(b), a.size)


def knn(q, t, k, nthreads=1):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.zeros((len(q), k), np.int32)
    dist = np.zeros((len(q), k), np.int32)
    rc = lib().ref_knn(_p(q), C.c_int64(len(q)), _p(t), C.c_int64(len(t)), q.shape[1], int(k), _p(idx), _p(dist),
                       int(nthreads))
    assert rc == 0
    return idx, dist


def _matcher(q, trains, masks, k, radius, compact):
    q = np.ascontiguousarray(q, np.uint8)
    nb = q.shape[1]
    tr = [np.ascontiguousarray(t, np.uint8).reshape(-1, nb) for t in trains]
    n = len(tr)
    tp = (C.c_void_p * n)(*[t.ctypes.data if len(t) else None for t in tr])
    nts = np.array([len(t) for t in tr], np.int32)
    mp, keep = None, []
    if masks:
        for m, t in zip(masks, tr):
            keep.append(None if m is None or m.size == 0 else np.ascontiguousarray(m, np.uint8).reshape(len(q), len(t)))
        mp = (C.c_void_p * n)(*[m.ctypes.data if m is not None else None for m in keep])
    f = lib().ref_matcher
    f.restype = C.c_int64
    cap = 1 << 16
    while True:
        out = np.zeros(cap, np.int32)
        need = f(_p(q), len(q), nb, n, tp, _p(nts), mp, int(k), C.c_float(radius), int(bool(compact)), _p(out), C.c_int64(cap))
        if need <= cap:
            break
        cap = int(need)
    pos, res = 1, []
    for _ in range(int(out[0])):
        c = int(out[pos]); pos += 1
        rec = out[pos:pos + 4 * c].reshape(c, 4)
        res.append([(int(r[0]), int(r[1]), int(r[2]), float(r[3:4].view(np.float32)[0])) for r in rec])
        pos += 4 * c
    return res


def knn_match(q, trains, k, masks=None, compact=False):
    """brisk::BruteForceMatcher::knnMatch (the reference class itself, brute-force-matcher.cc:59-68,80-162) over a train
    collection: list (per query) of (queryIdx, trainIdx, imgIdx, distance)."""
    return _matcher(q, trains, masks, k, -1.0, compact)


def radius_match(q, trains, max_distance, masks=None, compact=False):
    """brisk::BruteForceMatcher::radiusMatch (brute-force-matcher.cc:70-78,164-214)."""
    return _matcher(q, trains, masks, 0, float(max_distance), compact)


def bench_detect_describe(imgs, harris=False, thresh=60, octaves=4, radius=30.0, abs_thr=20.0, nthreads=1):
    """-> (seconds, total described keypoints) for imgs [n, h, w] u8."""
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    total = C.c_int64(0)
    s = lib().ref_bench_detect_describe(_p(imgs), n, w, h, int(harris), int(thresh), int(octaves),
                                        C.c_double(radius), C.c_double(abs_thr), int(nthreads), C.byref(total))
    return float(s), int(total.value)
