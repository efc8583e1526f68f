"""ctypes binding of oracle/_build/libbrisk_oracle.so -- our CPU restatement of
the reference algorithm (TEST INFRASTRUCTURE ONLY).  Same Python surface as
oracle/ref.py so tests can swap one for the other."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from .ref import KP_DTYPE

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "libbrisk_oracle.so"
_lib = None


def build():
    subprocess.run(["make", "-C", str(_DIR), "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(_LIB_PATH))
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
    lib().orc_halfsample8(_p(img), w, h, _p(out))
    return out


def twothirdsample8(img):
    img, w, h = _img(img)
    out = np.zeros((2 * (h // 3), 2 * (w // 3)), np.uint8)
    lib().orc_twothirdsample8(_p(img), w, h, _p(out))
    return out


def halfsample16(img):
    """Halfsample16 (image-down-sampling.cc:56-139): per 2x2 block avg(avg(a, b), avg(sat(sat(c + 1) + 1), d)) with the
    rounding-up average of pavgw; the last SSE block overlaps the one before, so every output pixel follows the same rule.
    Defined for cols >= 16 (the reference writes nothing otherwise)."""
    s = np.ascontiguousarray(img, np.uint16).astype(np.int64)
    h, w = s.shape
    assert w >= 16
    hh, ww = h // 2, w // 2
    a, b = s[0:2 * hh:2, 0:2 * ww:2], s[0:2 * hh:2, 1:2 * ww:2]
    c, d = s[1:2 * hh:2, 0:2 * ww:2], s[1:2 * hh:2, 1:2 * ww:2]
    c = np.minimum(np.minimum(c + 1, 65535) + 1, 65535)
    avg = lambda x, y: (x + y + 1) >> 1
    return avg(avg(a, b), avg(c, d)).astype(np.uint16)


def twothirdsample16(img):
    """Twothirdsample16 (image-down-sampling.cc:394-548): per 3x3 block the 4:2:2:1 weighted sums of its four 2x2 corners,
    truncating division by 9, packed with SIGNED saturation (values above 32767 come out as 32767).  Needs cols >= 12."""
    s = np.ascontiguousarray(img, np.uint16).astype(np.int64)
    h, w = s.shape
    assert (w // 3) * 3 >= 12
    bh, bw = h // 3, w // 3
    p = lambda r, c: s[r:3 * bh:3, c:3 * bw:3]
    out = np.zeros((2 * bh, 2 * bw), np.int64)
    out[0::2, 0::2] = (4 * p(0, 0) + 2 * p(0, 1) + 2 * p(1, 0) + p(1, 1)) // 9
    out[0::2, 1::2] = (4 * p(0, 2) + 2 * p(0, 1) + 2 * p(1, 2) + p(1, 1)) // 9
    out[1::2, 0::2] = (4 * p(2, 0) + 2 * p(2, 1) + 2 * p(1, 0) + p(1, 1)) // 9
    out[1::2, 1::2] = (4 * p(2, 2) + 2 * p(2, 1) + 2 * p(1, 2) + p(1, 1)) // 9
    return np.minimum(out, 32767).astype(np.uint16)


def thrmap(img):
    img, w, h = _img(img)
    out = np.zeros((h, w), np.uint8)
    lib().orc_thrmap(_p(img), w, h, _p(out))
    return out


def layer_dump(img, thresh, lower=10, cap=1 << 20):
    img, w, h = _img(img)
    c = np.zeros((cap, 3), np.int32)
    n = lib().orc_layer_corners(_p(img), w, h, int(thresh), int(lower), _p(c), cap)
    assert n <= cap
    return thrmap(img), c[:n].copy()


def dense_scores(img):
    img, w, h = _img(img)
    a = np.zeros((h, w), np.uint8)
    b = np.zeros((h, w), np.uint8)
    lib().orc_dense_scores(_p(img), w, h, _p(a), _p(b))
    return a, b


def pyramid(img, octaves):
    img, w, h = _img(img)
    nl = max(1, 2 * octaves)
    dims = np.zeros((nl, 2), np.int32)
    so = np.zeros((nl, 2), np.float32)
    buf = np.zeros(w * h * 3, np.uint8)
    n = lib().orc_pyramid(_p(img), w, h, int(octaves), _p(buf), _p(dims), _p(so))
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
    n = lib().orc_agast_detect(_p(img), w, h, int(thresh), int(octaves), int(bool(suppress)), _p(mask), _p(kps), cap)
    assert n <= cap
    return kps[:n].copy()


def compute_scale(img, kps, thresh, octaves=3, suppress=True, cap=1 << 18):
    """BriskFeatureDetector(thresh, octaves, suppress).ComputeScale(img, kps) (brisk-feature-detector.cc:87-92)"""
    img, w, h = _img(img)
    k = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_compute_scale(_p(img), w, h, int(thresh), int(octaves), int(bool(suppress)), _p(k), len(k), _p(out), cap)
    assert 0 <= n <= cap
    return out[:n].copy()


def harris_detect(img, octaves, radius, abs_thr=0.0, max_kpt=-1, cap=1 << 18):
    img, w, h = _img(img)
    kps = np.zeros(cap, KP_DTYPE)
    n = lib().orc_harris_detect(_p(img), w, h, int(octaves), C.c_double(radius), C.c_double(abs_thr),
                                C.c_int64(max_kpt), _p(kps), cap)
    assert n <= cap
    return kps[:n].copy()


def harris_detect_passed(shape, kps, radius, max_kpt=-1, cap=1 << 18):
    """The "use passed key points" mode of the Harris scale-space detector on one layer; shape = (h, w)."""
    h, w = shape
    k = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_harris_detect_passed(int(w), int(h), C.c_double(radius), C.c_int64(max_kpt), _p(k), len(k), _p(out), cap)
    assert 0 <= n <= cap
    return out[:n].copy()


def harris_scores(img):
    img, w, h = _img(img)
    out = np.zeros((h, w), np.int32)
    lib().orc_harris_scores(_p(img), w, h, _p(out))
    return out


def harris_maxima(img, abs_thr, cap=1 << 20):
    img, w, h = _img(img)
    out = np.zeros((cap, 3), np.int32)
    n = lib().orc_harris_maxima(_p(img), w, h, int(abs_thr), _p(out), cap)
    assert n <= cap
    return out[:n].copy()


def integral8(img):
    img, w, h = _img(img)
    out = np.zeros((h + 1, w + 1), np.int32)
    lib().orc_integral8(_p(img), w, h, _p(out))
    return out


def describe(img, kps, rot=True, scale=True, version=2, pattern_scale=1.0):
    img, w, h = _img(img)
    k = np.ascontiguousarray(kps, KP_DTYPE).copy()
    nb = C.c_int32(0)
    flat = np.zeros(max(len(k), 1) * 256, np.uint8)
    n = lib().orc_describe(_p(img), w, h, _p(k), len(k), int(rot), int(scale), int(version),
                           C.c_float(pattern_scale), _p(flat), C.byref(nb))
    return k[:n].copy(), flat[:n * nb.value].reshape(n, nb.value).copy()


def pattern_dump(version=2, pattern_scale=1.0):
    counts = np.zeros(4, np.int32)
    lib().orc_pattern_dump(int(version), C.c_float(pattern_scale), _p(counts), None, None, None, None, None)
    P, ns, nl, strings = (int(v) for v in counts)
    pts = np.zeros((64, 1024, P, 3), np.float32)
    scale_list = np.zeros(64, np.float32)
    size_list = np.zeros(64, np.uint32)
    sp = np.zeros((ns, 2), np.uint32)
    lp = np.zeros((nl, 4), np.int32)
    lib().orc_pattern_dump(int(version), C.c_float(pattern_scale), _p(counts), _p(pts), _p(scale_list),
                           _p(size_list), _p(sp), _p(lp))
    return dict(points=P, strings=strings, pts=pts, scale_list=scale_list, size_list=size_list,
                short_pairs=sp, long_pairs=lp)


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming(_p(a), _p(b), a.size)


def knn(q, t, k):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.zeros((len(q), k), np.int32)
    dist = np.zeros((len(q), k), np.int32)
    lib().orc_knn(_p(q), C.c_int64(len(q)), _p(t), C.c_int64(len(t)), q.shape[1], int(k), _p(idx), _p(dist))
    return idx, dist


def _is_masked_out(masks, q):
    """cv::DescriptorMatcher::isMaskedOut: every mask is non-empty and has an all-zero row q."""
    if not masks:
        return False
    out = sum(1 for m in masks if m is not None and m.size and not np.any(m[q]))
    return out == len(masks)


def _possible(masks, img, q, t):
    """masks.empty() || isPossibleMatch(masks[img], q, t) (an empty mask allows everything)."""
    if not masks:
        return True
    m = masks[img]
    return m is None or m.size == 0 or m[q, t] != 0


def _sorted_matches(ms):
    """std::sort of one query's DMatch list (brute-force-matcher.cc:160,210)."""
    if len(ms) < 2:
        return ms
    t = np.array([m[1] for m in ms], np.int32)
    i = np.array([m[2] for m in ms], np.int32)
    d = np.array([m[3] for m in ms], np.float32)
    lib().orc_sort_matches(_p(t), _p(i), _p(d), len(ms))
    return [(ms[0][0], int(a), int(b), float(c)) for a, b, c in zip(t, i, d)]


def _distance_matrices(q, trains):
    out = []
    for t in trains:
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, q.shape[1])
        d = np.zeros((len(q), len(t)), np.int32)
        if len(t):
            lib().orc_hamming_matrix(_p(q), C.c_int64(len(q)), _p(t), C.c_int64(len(t)), q.shape[1], _p(d))
        out.append(d)
    return out


def knn_match(q, trains, k, masks=None, compact=False):
    """BruteForceMatcher::commonKnnMatchImpl (brute-force-matcher.cc:80-162) over a train collection with
    optional per-image masks [nq][nt_i]: list (per query) of (queryIdx, trainIdx, imgIdx, distance).
    Line by line, including what the code does once the real candidates of a query run out: masked /
    taken entries hold INT_MAX, minMaxLoc still returns a location and INT_MAX < FLT_MAX, so the list is
    padded with (trainIdx of the first INT_MAX entry, first image whose minimum it is, 2147483648.0f)."""
    q = np.ascontiguousarray(q, np.uint8)
    INT_MAX = np.iinfo(np.int32).max
    dist = _distance_matrices(q, trains)
    matches = []
    for qi in range(len(q)):
        if _is_masked_out(masks, qi):
            if not compact:
                matches.append([])
            continue
        all_d = []
        for img, d in enumerate(dist):
            row = np.full(d.shape[1], INT_MAX, np.int64)
            for t in range(d.shape[1]):
                if _possible(masks, img, qi, t):
                    row[t] = d[qi, t]
            all_d.append(row)
        cur = []
        for _ in range(k):
            best = None
            best_d = float(np.finfo(np.float32).max)
            for img, row in enumerate(all_d):
                if row.size:
                    loc = int(np.argmin(row))  # minMaxLoc: first minimum
                    if float(row[loc]) < best_d:
                        best = (qi, loc, img, float(np.float32(float(row[loc]))))
                        best_d = best[3]
            if best is None:
                break
            all_d[best[2]][best[1]] = INT_MAX
            cur.append(best)
        matches.append(_sorted_matches(cur))
    return matches


def radius_match(q, trains, max_distance, masks=None, compact=False):
    """BruteForceMatcher::commonRadiusMatchImpl (brute-force-matcher.cc:164-214)."""
    q = np.ascontiguousarray(q, np.uint8)
    dist = _distance_matrices(q, trains)
    md = np.float32(max_distance)
    matches = []
    for qi in range(len(q)):
        if _is_masked_out(masks, qi):
            if not compact:
                matches.append([])
            continue
        cur = []
        for img, d in enumerate(dist):
            for t in range(d.shape[1]):
                if _possible(masks, img, qi, t) and np.float32(d[qi, t]) < md:
                    cur.append((qi, t, img, float(d[qi, t])))
        matches.append(_sorted_matches(cur))
    return matches


# ---------------------------------------------------------------------------
# Legacy single-scale brisk::HarrisFeatureDetector (harris-feature-detector.cc:56-409, vectorized-filters.cc:54-123)
# ---------------------------------------------------------------------------

def _wrap16(a):
    return ((a.astype(np.int64) + 32768) % 65536 - 32768).astype(np.int64)


def harris_legacy_scores(img):
    """GetCovarEntries -> FilterGauss3by316S x3 -> CornerHarris (harris-feature-detector.cc:76-268, vectorized-filters.cc:54-123):
    the int32 response map.  16-bit wrap-around arithmetic as the SSE code has it; rows h-2.. and columns w-2.. stay zero,
    and so do row 0 / column 0 of the smoothed covariances.  Needs w >= 18 (narrower images never enter the SSE loops)."""
    s = np.ascontiguousarray(img, np.uint8).astype(np.int64)
    h, w = s.shape
    assert w - 2 >= 16 and h >= 3
    p = lambda dy, dx: s[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
    dx = _wrap16(24 * p(-1, -1) + 80 * p(0, -1) + 24 * p(1, -1) - 24 * p(-1, 1) - 80 * p(0, 1) - 24 * p(1, 1))
    dy = _wrap16(24 * p(-1, -1) + 80 * p(-1, 0) + 24 * p(-1, 1) - 24 * p(1, -1) - 80 * p(1, 0) - 24 * p(1, 1))
    cov = []
    for a, b in ((dx, dx), (dy, dy), (dy, dx)):
        c = np.zeros((h, w), np.int64)
        c[1:h - 1, 1:w - 1] = ((a * b) >> 16) >> 4          # pmulhw, then psraw 4
        cov.append(c)
    sm = []
    for c in cov:
        g = np.zeros((h, w), np.int64)
        q = lambda dy, dx: c[1 + dy:h - 1 + dy, 1 + dx:w - 1 + dx]
        g[1:h - 1, 1:w - 1] = _wrap16(4 * q(0, 0) + 2 * (q(-1, 0) + q(1, 0) + q(0, -1) + q(0, 1)) + q(-1, -1) + q(-1, 1) + q(1, -1) + q(1, 1))
        sm.append(g)
    a, b, c = (g[:h - 2, :w - 2] for g in sm)
    tq = _wrap16(_wrap16((a >> 1) + (b >> 1)) >> 1)
    out = np.zeros((h, w), np.int64)
    out[:h - 2, :w - 2] = a * b - c * c - tq * tq
    return ((out + 2 ** 31) % 2 ** 32 - 2 ** 31).astype(np.int32)


def harris_legacy_lut(radius):
    """SetRadius (harris-feature-detector.cc:57-72): 31 x 31 float mask centred at (radius / 2, radius / 2)."""
    r = float(radius) / 2.0
    lut = np.zeros((31, 31), np.float32)
    for x in range(31):
        for y in range(31):
            lut[y, x] = np.float32(max(1 - float((r - x) * (r - x) + (r - y) * (r - y)) / float(r * r), 0.0))
    return lut


def harris_legacy(img, radius):
    """brisk::HarrisFeatureDetector(radius).detect(img): NonmaxSuppress (:270-322; threshold 64, x reported one column to the
    right of the maximum) and EnforceUniformity (:324-393; std::sort by response, half-resolution occupancy map indexed
    with x as the ROW, the int16 cast of the response in the acceptance test).  Raises when the occupancy indices leave the
    map (the reference then reads / writes out of bounds: landscape images)."""
    sc = harris_legacy_scores(img).astype(np.int64)
    h, w = sc.shape
    flat = sc.ravel()
    pts = []
    for j in range(2, h - 2):
        row = sc[j]
        c = np.arange(0, w - 2)
        ctr = row[c]
        left = np.where(c > 0, row[np.maximum(c - 1, 0)], flat[j * w - 1])            # column -1 is the previous row's last element
        ok = (ctr >= 64) & (row[c + 1] <= ctr) & (left <= ctr)
        for dj in (1, -1):
            r2 = sc[j + dj]
            l2 = np.where(c > 0, r2[np.maximum(c - 1, 0)], flat[(j + dj) * w - 1])
            ok &= (r2[c] <= ctr) & (r2[c + 1] <= ctr) & (l2 <= ctr)
        for cc in np.flatnonzero(ok):
            pts.append((float(cc + 1), float(j), float(np.float32(ctr[cc]))))
    n = len(pts)
    kps = np.zeros(n, KP_DTYPE)
    if n == 0:
        return kps
    resp = np.array([p[2] for p in pts], np.float32)
    perm = np.arange(n, dtype=np.int32)
    lib().orc_sort_desc_by_response(_p(resp), _p(perm), n)
    H, W = h // 2 + 32, w // 2 + 32
    occ = np.zeros(H * W, np.int64)
    lut = harris_legacy_lut(radius)
    keep = []
    for i in perm:
        x, y, r = pts[i]
        cy = int(np.float32(x) / np.float32(2) + np.float32(16))
        cx = int(np.float32(y) / np.float32(2) + np.float32(16))
        if (cy + 15) * W + cx + 16 >= H * W:
            raise ValueError("occupancy index out of bounds (undefined in the reference)")
        s0 = float(occ[cy * W + cx])
        r16 = ((int(r) + 32768) % 65536) - 32768                                      # static_cast<int16_t>(float): cvttss2si, low 16 bits
        if r16 < (s0 * s0) * (s0 * s0):
            continue
        nsc = np.float32(np.sqrt(np.sqrt(np.float64(r))))
        stamp = (lut * nsc).astype(np.float32).astype(np.int64) & 0xff                  # (char)(float * float)
        for yy in range(31):
            base = (cy + yy - 15) * W + cx - 15
            occ[base:base + 31] = np.minimum(occ[base:base + 31] + stamp[yy], 255)
        keep.append(i)
    out = np.zeros(len(keep), KP_DTYPE)
    for o, i in enumerate(keep):
        out[o] = (pts[i][0], pts[i][1], 10.0, -1.0, pts[i][2], 0, -1)
    return out
