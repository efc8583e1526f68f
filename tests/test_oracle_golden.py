"""The oracle against the reference's own golden fixtures and against the
compiled reference (oracle/_ref).  CPU only."""
import numpy as np
import pytest

from conftest import kp_equal
from ethzasl_brisk_b200.synthetic import random_descriptors, synthetic_frame


@pytest.mark.parametrize("i", [0, 1])
def test_golden_ast(oracle, golden, i):
    # reference test-binary-equal.cc:323-326: BriskFeatureDetector(70) + default extractor
    img = golden[f"image{i}"]
    k = oracle.agast_detect(img, 70, 3)
    k2, d = oracle.describe(img, k)
    assert kp_equal(k2, golden[f"ast{i}_kps"])          # all 7 fields, exact
    assert np.array_equal(d, golden[f"ast{i}_desc"])    # Hamming 0 (the reference allows <= 5)


@pytest.mark.parametrize("i", [0, 1])
def test_golden_harris(oracle, golden, i):
    # reference test-binary-equal.cc:73-89,305-311: octaves 0, uniformity radius 30, abs threshold 20
    img = golden[f"image{i}"]
    k = oracle.harris_detect(img, 0, 30.0, 20.0)
    k2, d = oracle.describe(img, k, True, True)
    assert kp_equal(k2, golden[f"harris{i}_kps"])
    assert np.array_equal(d, golden[f"harris{i}_desc"])


def test_golden_match_homography(oracle, golden):
    # reference test-match.cc:50-116: BriskFeatureDetector(70, 2), best match with Hamming < 50 must
    # agree with H_1to2 within 5 px.
    H = np.array([[8.7976964e-01, 3.1245438e-01, -3.9430589e+01], [-1.8389418e-01, 9.3847198e-01, 1.5315784e+02],
                  [1.9641425e-04, -1.6015275e-05, 1.0000000e+00]])
    k1, d1 = oracle.describe(golden["image0"], oracle.agast_detect(golden["image0"], 70, 2))
    k2, d2 = oracle.describe(golden["image1"], oracle.agast_detect(golden["image1"], 70, 2))
    idx, dist = oracle.knn(d1, d2, 1)
    outliers = matches = 0
    for q in range(len(k1)):
        if dist[q, 0] < 50:
            matches += 1
            p = H @ np.array([k1["x"][q], k1["y"][q], 1.0])
            p = p[:2] / p[2]
            t = idx[q, 0]
            if np.hypot(p[0] - k2["x"][t], p[1] - k2["y"][t]) > 5:
                outliers += 1
    assert matches > 50 and outliers == 0


def test_popcount_kat(oracle):
    # reference test-popcount.cc:72-86 style known answer: bit loop vs primitive
    a = random_descriptors(1, 64, 1)[0]
    b = random_descriptors(1, 64, 2)[0]
    expect = int(np.unpackbits(a ^ b).sum())
    assert oracle.hamming(a, b) == expect
    assert oracle.hamming(a[:48], b[:48]) == int(np.unpackbits(a[:48] ^ b[:48]).sum())


@pytest.mark.parametrize("w,h", [(752, 480), (500, 320), (376, 240), (250, 160), (125, 80), (801, 601), (47, 33)])
def test_sampling_vs_ref(oracle, ref, w, h):
    img = synthetic_frame(w, h, w + h)
    assert np.array_equal(oracle.halfsample8(img), ref.halfsample8(img))
    assert np.array_equal(oracle.twothirdsample8(img), ref.twothirdsample8(img))


def test_stages_vs_ref(oracle, ref, golden):
    for img in (golden["image0"], synthetic_frame(640, 480, 7)):
        t1, c1 = ref.layer_dump(img, 60)
        t2, c2 = oracle.layer_dump(img, 60)
        assert np.array_equal(t1, t2) and np.array_equal(c1, c2)
        a1, b1 = ref.dense_scores(img)
        a2, b2 = oracle.dense_scores(img)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)
        assert np.array_equal(ref.integral8(img), oracle.integral8(img))
        assert np.array_equal(ref.harris_scores(img), oracle.harris_scores(img))
        assert np.array_equal(ref.harris_maxima(img, 20), oracle.harris_maxima(img, 20))


@pytest.mark.parametrize("thresh,octaves", [(60, 4), (70, 3), (35, 2), (70, 0), (60, 1)])
def test_agast_detect_vs_ref(oracle, ref, golden, thresh, octaves):
    for img in (golden["image1"], synthetic_frame(752, 480, 1000)):
        assert kp_equal(oracle.agast_detect(img, thresh, octaves), ref.agast_detect(img, thresh, octaves))


def _provided_points(img, n, seed, integer=False):
    """n key points anywhere on the image (also on its border), with class ids to be copied through."""
    from oracle.ref import KP_DTYPE
    rng = np.random.default_rng(seed)
    h, w = img.shape
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, w, n), rng.uniform(0, h, n)
    if integer:
        k["x"], k["y"] = np.floor(k["x"]), np.floor(k["y"])
    k["size"], k["angle"], k["class_id"] = 9.0, -1.0, np.arange(n)
    return k


@pytest.mark.parametrize("thresh,octaves", [(60, 4), (70, 3), (35, 2), (70, 0), (60, 1)])
def test_compute_scale_vs_ref(oracle, ref, golden, thresh, octaves):
    # BriskFeatureDetector::ComputeScale = GetKeypoints with provided key points (brisk-scale-space.cc:104-124):
    # detected key points fed back, arbitrary fractional points, integer points
    for img in (golden["image1"], synthetic_frame(500, 333, 3)):
        lists = [ref.agast_detect(img, thresh, octaves), _provided_points(img, 400, 11), _provided_points(img, 400, 12, True)]
        for k in lists:
            want = ref.compute_scale(img, k, thresh, octaves)
            assert len(want) > 5 and kp_equal(oracle.compute_scale(img, k, thresh, octaves), want)
    # a layer that keeps no point runs the detector (threshold map without lower bound) on that layer only
    few = _provided_points(golden["image0"], 3, 1)
    few["x"], few["y"] = [5, 6.5, 30], [5, 7.25, 9]
    want = ref.compute_scale(golden["image0"], few, thresh, octaves)
    assert kp_equal(oracle.compute_scale(golden["image0"], few, thresh, octaves), want)


def _passed_points(shape, n, seed, ties=False):
    """n key points inside the image with Harris-like responses on both sides of the 1e6 gate."""
    from oracle.ref import KP_DTYPE
    rng = np.random.default_rng(seed)
    h, w = shape
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, w - 0.01, n), rng.uniform(0, h - 0.01, n)
    r = 10.0 ** rng.uniform(5.0, 9.2, n)
    if ties:  # many equal scores: the order libstdc++'s introsort leaves them in is part of the result
        r = rng.choice(np.array([2.0e6, 3.0e6, 5.0e5, 7.5e6, 1.0e9]), n)
    k["response"], k["size"], k["angle"], k["class_id"] = r, 12.0, -1.0, np.arange(n)
    return k


@pytest.mark.parametrize("radius,max_kpt", [(30.0, -1), (8.0, -1), (3.0, 150), (0.0, 400), (-1.0, 90), (15.0, 40)])
def test_harris_passed_keypoints_vs_ref(oracle, ref, golden, radius, max_kpt):
    # detect() on a non-empty vector (scale-space-feature-detector.h:103-108): re-filtering, no detection
    img = golden["image0"]
    lists = [ref.harris_detect(img, 0, 1.0, 0.0), _passed_points(img.shape, 3000, 1), _passed_points(img.shape, 700, 2, True),
             _passed_points(img.shape, 20, 3)]
    for k in lists:
        want = ref.harris_detect_passed(img, k, radius, max_kpt)
        assert kp_equal(oracle.harris_detect_passed(img.shape, k, radius, max_kpt), want)
    low = lists[1].copy()
    low["response"] = 999999.0  # nothing passes the gate: the vector comes back untouched
    assert kp_equal(oracle.harris_detect_passed(img.shape, low, radius, max_kpt), ref.harris_detect_passed(img, low, radius, max_kpt))
    assert kp_equal(oracle.harris_detect_passed(img.shape, low, radius, max_kpt), low)


# (thresh, octaves) / (radius, maxNumKpt) of the cases in tests/golden/provided_keypoints.npz (tests/golden/make_provided_keypoints.py)
COMPUTE_SCALE_GOLDEN = [(60, 4), (70, 3), (30, 5), (70, 0)]
PASSED_GOLDEN = [(30.0, -1), (3.0, 150), (0.0, 400)]


def test_provided_keypoints_golden(oracle, golden, golden_provided):
    # the committed reference outputs (no compiled reference needed): ComputeScale and "use passed key points"
    img = golden["image0"]
    for i, (thresh, octaves) in enumerate(COMPUTE_SCALE_GOLDEN):
        assert kp_equal(oracle.compute_scale(img, golden_provided[f"cs{i}_in"], thresh, octaves), golden_provided[f"cs{i}_out"])
    for i, (radius, max_kpt) in enumerate(PASSED_GOLDEN):
        assert kp_equal(oracle.harris_detect_passed(img.shape, golden_provided[f"hp{i}_in"], radius, max_kpt), golden_provided[f"hp{i}_out"])


def test_agast_mask_vs_ref(oracle, ref, golden):
    img = golden["image0"]
    mask = np.zeros_like(img)
    mask[100:500, 200:700] = 255
    assert kp_equal(oracle.agast_detect(img, 60, 4, mask=mask), ref.agast_detect(img, 60, 4, mask=mask))


@pytest.mark.parametrize("version,ps,rot,scale", [(2, 1.0, True, True), (1, 1.0, True, True), (2, 0.5, True, True),
                                                   (2, 1.0, False, True), (2, 1.0, True, False), (1, 0.7, True, True)])
def test_describe_vs_ref(oracle, ref, golden, version, ps, rot, scale):
    img = golden["image0"]
    k = ref.agast_detect(img, 60, 4)
    a, da = ref.describe(img, k, rot, scale, version, ps)
    b, db = oracle.describe(img, k, rot, scale, version, ps)
    assert kp_equal(a, b) and np.array_equal(da, db)
    pa, pb = ref.pattern_dump(version, ps), oracle.pattern_dump(version, ps)
    for key in ("pts", "scale_list", "size_list", "short_pairs", "long_pairs"):
        assert np.array_equal(pa[key], pb[key]), key


@pytest.mark.parametrize("octaves,radius,max_kpt", [(1, 30.0, -1), (4, 30.0, -1), (2, 10.0, 300), (4, 0.0, 400), (2, -1.0, 64), (0, 0.0, 100000)])
def test_harris_detect_vs_ref(oracle, ref, golden, octaves, radius, max_kpt):
    for img in (golden["image1"], synthetic_frame(752, 480, 1001)):
        assert kp_equal(oracle.harris_detect(img, octaves, radius, 20.0, max_kpt), ref.harris_detect(img, octaves, radius, 20.0, max_kpt))


def test_knn_vs_ref_primitive(oracle, ref):
    q = random_descriptors(40, 64, 5)
    t = random_descriptors(300, 64, 6)
    t[7] = q[3]
    t[9] = q[3]  # exact tie -> lowest train index
    i1, d1 = oracle.knn(q, t, 2)
    i2, d2 = ref.knn(q, t, 2)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert i1[3, 0] == 7 and i1[3, 1] == 9 and d1[3, 0] == 0


def test_matcher_restatement_consistency(oracle):
    # knn_match / radius_match (brute-force-matcher.cc:80-214 line by line) against the plain single-image kNN
    # restatement and a numpy brute force; std::sort replay keeps lists sorted by distance
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (40, 48), dtype=np.uint8)
    t = rng.integers(0, 256, (500, 48), dtype=np.uint8)
    t[7] = q[2]; t[9] = q[2]
    idx, dist = oracle.knn(q, t, 3)
    lists = oracle.knn_match(q, [t], 3)
    assert [[m[1] for m in v] for v in lists] == idx.tolist() and [[int(m[3]) for m in v] for v in lists] == dist.tolist()
    d = np.unpackbits(q[:, None, :] ^ t[None, :, :], axis=2).sum(axis=2)
    rad = oracle.radius_match(q, [t[:250], t[250:]], 180.0)
    for qi, v in enumerate(rad):
        assert sorted((m[2] * 250 + m[1]) for m in v) == np.nonzero(d[qi] < 180)[0].tolist()
        assert all(a[3] <= b[3] for a, b in zip(v, v[1:]))
    # fewer candidates than k: the reference pads with (0, last non-empty image, 2147483648.0f)
    short = oracle.knn_match(q[:2], [t[:1], t[:0]], 3)
    assert short[0][1:] == [(0, 0, 0, 2147483648.0)] * 2


def _tie_rich(n, nbytes, seed):
    rng = np.random.default_rng(seed)
    d = np.zeros((n, nbytes), np.uint8)
    d[:, :6] = rng.integers(0, 4, (n, 6), dtype=np.uint8) * 17
    return d


@pytest.mark.parametrize("nbytes", [48, 64])
def test_matcher_vs_ref_class(oracle, ref, nbytes):
    # the restated selection rules against brisk::BruteForceMatcher ITSELF (brute-force-matcher.cc compiled
    # unmodified into oracle/_ref): collections with an empty image, masks, masked-out queries, exhausted
    # candidates (INT_MAX padding), ties across images, compactResult, radius lists with long runs of equal
    # distances (std::sort's permutation), fewer train rows than k
    rng = np.random.default_rng(11)
    q = random_descriptors(90, nbytes, 5)
    trains = [random_descriptors(300, nbytes, 6), np.zeros((0, nbytes), np.uint8), random_descriptors(170, nbytes, 7)]
    trains[0][10] = q[3]; trains[2][5] = q[3]; trains[0][200] = q[3]
    masks = [(rng.random((90, 300)) < 0.7).astype(np.uint8), None, (rng.random((90, 170)) < 0.5).astype(np.uint8)]
    masks[0][7] = 0; masks[2][7] = 0
    masks[0][9] = 0; masks[2][9] = 0; masks[2][9, 4] = 1
    for k in (1, 2, 5, 11):
        for compact in (False, True):
            assert oracle.knn_match(q, trains, k, masks, compact) == ref.knn_match(q, trains, k, masks, compact), (k, compact)
        assert oracle.knn_match(q, trains, k) == ref.knn_match(q, trains, k)
    full = [masks[0], masks[2]]
    two = [trains[0], trains[2]]
    for compact in (False, True):
        assert oracle.knn_match(q, two, 3, full, compact) == ref.knn_match(q, two, 3, full, compact)
    assert oracle.knn_match(q[:4], [trains[0][:2]], 3) == ref.knn_match(q[:4], [trains[0][:2]], 3)
    # single image: the plain kNN restatement / the multi-threaded baseline loop are the class's result too
    idx, dist = oracle.knn(q, trains[0], 3)
    lists = ref.knn_match(q, [trains[0]], 3)
    assert [[m[1] for m in v] for v in lists] == idx.tolist() and [[int(m[3]) for m in v] for v in lists] == dist.tolist()
    i2, d2 = ref.knn(q, trains[0], 3, nthreads=2)
    assert np.array_equal(i2, idx) and np.array_equal(d2, dist)
    # radiusMatch
    tq = _tie_rich(60, nbytes, 1)
    tt = [_tie_rich(400, nbytes, 2), _tie_rich(250, nbytes, 3)]
    for md in (0.0, 1.0, 7.5, 12.0, 1000.0):
        assert oracle.radius_match(tq, tt, md) == ref.radius_match(tq, tt, md), md
    rm = [(rng.random((60, 400)) < 0.6).astype(np.uint8), (rng.random((60, 250)) < 0.6).astype(np.uint8)]
    rm[0][5] = 0; rm[1][5] = 0
    for compact in (False, True):
        assert oracle.radius_match(tq, tt, 9.0, rm, compact) == ref.radius_match(tq, tt, 9.0, rm, compact)
    md = float(nbytes * 8 * 0.44)
    assert oracle.radius_match(q, [trains[0]], md) == ref.radius_match(q, [trains[0]], md)


@pytest.mark.parametrize("w,h", [(16, 8), (17, 9), (752, 480), (33, 21), (100, 50), (12, 9), (13, 10), (641, 479)])
def test_samplers_16bit_vs_ref(oracle, ref, w, h):
    # Halfsample16 / Twothirdsample16 restated in numpy against the compiled reference (no fixture exists in the reference)
    rng = np.random.default_rng(w * 7 + h)
    for hi in (65536, 4096, 300):
        img = rng.integers(0, hi, (h, w)).astype(np.uint16)
        img[0, 0] = 65535; img[1, 0] = 65535; img[-1, -1] = 65535
        if w >= 16:
            assert np.array_equal(ref.halfsample16(img), oracle.halfsample16(img))
        if (w // 3) * 3 >= 12:
            assert np.array_equal(ref.twothirdsample16(img), oracle.twothirdsample16(img))


@pytest.mark.parametrize("w,h,radius", [(480, 480, 30.0), (333, 500, 30.0), (400, 401, 20.0), (18, 40, 10.0), (200, 320, 45.0), (131, 167, 7.5)])
def test_harris_legacy_vs_ref(oracle, ref, w, h, radius):
    # the legacy single-scale HarrisFeatureDetector (harris-feature-detector.cc compiled unmodified into oracle/_ref): score
    # map of its CornerHarris stage and the detected key points (order, coordinates, responses) against the restatement
    for img in (synthetic_frame(w, h, 7), np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8)):
        k, sc = ref.harris_legacy(img, radius, with_scores=True)
        assert np.array_equal(sc, oracle.harris_legacy_scores(img))
        assert kp_equal(k, oracle.harris_legacy(img, radius))
