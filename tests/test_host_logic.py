"""CPU-side checks of the product: the device code's NMS / refinement logic
compiled for the host (tests/host_emul) against the oracle, the C-ABI library's
exported symbols, and the synthetic-data helpers.  No GPU needed."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, kp_equal
from ethzasl_brisk_b200.synthetic import synthetic_frame
from oracle.ref import KP_DTYPE


@pytest.fixture(scope="module")
def emul():
    src = ROOT / "tests" / "host_emul" / "emul_agast.cc"
    lib = ROOT / "tests" / "host_emul" / "libemul_agast.so"
    deps = [src] + list((ROOT / "ethzasl_brisk_b200" / "csrc").glob("*.cuh"))
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-msse2", "-ffp-contract=off", "-fPIC", "-shared",
                        "-Wno-unknown-pragmas", "-o", str(lib), str(src)], check=True)
    handle = C.CDLL(str(lib))

    def detect(img, thresh, octaves, cap=1 << 18):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        k = np.zeros(cap, KP_DTYPE)
        n = handle.emul_agast_detect(img.ctypes.data_as(C.c_void_p), w, h, thresh, octaves, k.ctypes.data_as(C.c_void_p), cap)
        return n if n < 0 else k[:n].copy()

    def compute_scale(img, kps, thresh, octaves, cap=1 << 18):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        kin = np.ascontiguousarray(kps, KP_DTYPE)
        k = np.zeros(cap, KP_DTYPE)
        n = handle.emul_compute_scale(img.ctypes.data_as(C.c_void_p), w, h, thresh, octaves, kin.ctypes.data_as(C.c_void_p), len(kin),
                                      k.ctypes.data_as(C.c_void_p), cap)
        return n if n < 0 else k[:n].copy()
    detect.compute_scale = compute_scale
    return detect


@pytest.mark.parametrize("thresh,octaves", [(70, 3), (60, 4), (30, 2), (70, 0), (45, 1), (20, 3)])
def test_parallel_nms_formulation_golden(emul, oracle, golden, thresh, octaves):
    for i in (0, 1):
        img = golden[f"image{i}"]
        assert kp_equal(emul(img, thresh, octaves), oracle.agast_detect(img, thresh, octaves))


@pytest.mark.parametrize("seed,w,h,thresh,octaves", [(1, 752, 480, 60, 4), (2, 640, 480, 40, 3), (3, 500, 333, 30, 4),
                                                      (5, 321, 243, 35, 2), (6, 800, 600, 60, 0)])
def test_parallel_nms_formulation_synthetic(emul, oracle, seed, w, h, thresh, octaves):
    img = synthetic_frame(w, h, seed)
    assert kp_equal(emul(img, thresh, octaves), oracle.agast_detect(img, thresh, octaves))


@pytest.mark.parametrize("thresh,octaves", [(19, 3), (15, 4), (10, 2)])
def test_parallel_nms_formulation_low_thresholds(emul, oracle, golden, thresh, octaves):
    # below 20 the closed form still holds as long as no detected corner scores <= 2 (checked on the data by
    # corner_score_check_kernel); on these images none does down to thresh 10
    for img in (golden["image0"], synthetic_frame(500, 333, 3)):
        assert kp_equal(emul(img, thresh, octaves, cap=1 << 20), oracle.agast_detect(img, thresh, octaves, cap=1 << 20))
    # ... and below 10 such corners exist: refused, not guessed
    assert emul(golden["image0"], 5, 3) == -400


def test_parallel_nms_formulation_tie_heavy(emul, oracle):
    # quantised noise: plateaus of equal scores exercise the cache-state reconstruction
    rng = np.random.default_rng(7)
    for k in range(3):
        img = (rng.integers(0, 4, (240, 320)) * 60 + rng.integers(0, 8, (240, 320))).astype(np.uint8)
        assert kp_equal(emul(img, 30 + 5 * k, 3), oracle.agast_detect(img, 30 + 5 * k, 3))


@pytest.mark.parametrize("thresh,octaves", [(60, 4), (70, 3), (35, 2), (70, 0), (60, 1), (30, 5), (12, 3)])
def test_provided_keypoints_closed_form(emul, oracle, golden, thresh, octaves):
    # ComputeScale: the three data-parallel passes of the GPU path (touch / stamp / refine, nms_logic.cuh) run on
    # the CPU against the oracle's sequential replay of the lazy score cache
    from test_oracle_golden import _provided_points
    for img in (golden["image0"], synthetic_frame(500, 333, 3), synthetic_frame(1000, 700, 5)):
        for k in (oracle.agast_detect(img, max(thresh, 40), min(octaves, 4)), _provided_points(img, 2000, 21),
                  _provided_points(img, 500, 22, True)):
            got = emul.compute_scale(img, k, thresh, octaves)
            assert not isinstance(got, int), got
            assert len(got) > 5 and kp_equal(got, oracle.compute_scale(img, k, thresh, octaves))
    # layers that keep no point are detected on (threshold map without lower bound): deep pyramids, few points
    few = _provided_points(golden["image0"], 3, 1)
    few["x"], few["y"] = [5, 6.5, 30], [5, 7.25, 9]
    for img in (golden["image0"], synthetic_frame(500, 333, 3)):
        got = emul.compute_scale(img, few, thresh, octaves)
        assert not isinstance(got, int), got
        assert kp_equal(got, oracle.compute_scale(img, few, thresh, octaves))


@pytest.fixture(scope="module")
def emul_harris():
    src = ROOT / "tests" / "host_emul" / "emul_harris.cc"
    lib = ROOT / "tests" / "host_emul" / "libemul_harris.so"
    deps = [src] + list((ROOT / "ethzasl_brisk_b200" / "csrc").glob("*.cuh"))
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-msse2", "-ffp-contract=off", "-fPIC", "-shared",
                        "-Wno-unknown-pragmas", "-o", str(lib), str(src)], check=True)
    handle = C.CDLL(str(lib))

    def detect(img, octaves, radius, abs_thr, max_kpt=-1, cap=1 << 18):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        k = np.zeros(cap, KP_DTYPE)
        n = handle.emul_harris_detect(img.ctypes.data_as(C.c_void_p), w, h, octaves, C.c_double(radius), C.c_double(abs_thr),
                                      C.c_longlong(max_kpt), k.ctypes.data_as(C.c_void_p), cap)
        return k[:n].copy()

    def passed(shape, kps, radius, max_kpt=-1, cap=1 << 18):
        kin = np.ascontiguousarray(kps, KP_DTYPE)
        k = np.zeros(cap, KP_DTYPE)
        n = handle.emul_harris_passed(int(shape[1]), int(shape[0]), C.c_double(radius), C.c_longlong(max_kpt), kin.ctypes.data_as(C.c_void_p),
                                      len(kin), k.ctypes.data_as(C.c_void_p), cap)
        return n if n < 0 else k[:n].copy()
    detect.passed = passed
    return detect


@pytest.mark.parametrize("octaves,radius,abs_thr,max_kpt", [(0, 30.0, 20.0, -1), (4, 30.0, 20.0, -1), (2, 10.0, 0.0, 300)])
def test_harris_logic_and_introsort_replay(emul_harris, oracle, golden, octaves, radius, abs_thr, max_kpt):
    # the device code's Harris arithmetic and its restatement of libstdc++'s std::sort, on tie-heavy
    # score lists, against the oracle (which calls the real std::sort)
    for img in (golden["image0"], synthetic_frame(500, 333, 3)):
        assert kp_equal(emul_harris(img, octaves, radius, abs_thr, max_kpt), oracle.harris_detect(img, octaves, radius, abs_thr, max_kpt))


@pytest.mark.parametrize("octaves,radius,abs_thr,max_kpt", [(0, 0.5, 0.0, -1), (2, 0.7, 20.0, -1), (1, 0.25, 0.0, 500)])
def test_harris_logic_small_uniformity_radius(emul_harris, oracle, ref, octaves, radius, abs_thr, max_kpt):
    # radii below 1: occupancy maps of ceil(15 / radius)^2 bytes per pixel (small images only)
    img = synthetic_frame(320, 240, 8)
    want = ref.harris_detect(img, octaves, radius, abs_thr, max_kpt)
    assert len(want) > 100 and kp_equal(oracle.harris_detect(img, octaves, radius, abs_thr, max_kpt), want)
    assert kp_equal(emul_harris(img, octaves, radius, abs_thr, max_kpt), want)


def test_set_and_pgm_io_round_trip(tmp_path, golden):
    # the reference's serialized datasets (brisk/src/test/serialization.cc, bench-ds.cc:57-94) and its PGM images
    import ethzasl_brisk_b200 as bb
    entries = [dict(path=f"img{i}.pgm", image=golden[f"image{i}"], keypoints=golden[f"ast{i}_kps"], descriptors=golden[f"ast{i}_desc"],
                    blobs={"testImage": golden[f"image{i}"].tobytes()}) for i in (0, 1)]
    entries.append(dict(path="empty", image=np.zeros((8, 8), np.uint8), keypoints=np.zeros(0, KP_DTYPE),
                        descriptors=np.zeros((0, 48), np.uint8), blobs={}))
    bb.write_set(tmp_path / "a.set", entries)
    back = bb.read_set(tmp_path / "a.set")
    assert len(back) == 3
    for e, b in zip(entries, back):
        assert b["path"] == e["path"] and np.array_equal(b["image"], e["image"]) and kp_equal(b["keypoints"], e["keypoints"])
        assert b["descriptors"].shape[0] == len(e["keypoints"]) and np.array_equal(b["descriptors"].ravel(), e["descriptors"].ravel())
        assert b["blobs"] == e["blobs"]
    bb.write_pgm(tmp_path / "a.pgm", golden["image0"])
    assert np.array_equal(bb.read_pgm(tmp_path / "a.pgm"), golden["image0"])
    with pytest.raises(ValueError):
        (tmp_path / "cut.set").write_bytes((tmp_path / "a.set").read_bytes()[:-7])
        bb.read_set(tmp_path / "cut.set")
    # the reference's own fixtures: parse -> serialize is the identity on their bytes, and they hold the committed goldens
    ref_dir = Path("/root/reference/brisk/src/test/test_data")
    if ref_dir.exists():
        for kind in ("ast", "harris"):
            src = ref_dir / f"brisk_verification_{kind}.set"
            parsed = bb.read_set(src)
            bb.write_set(tmp_path / "b.set", parsed)
            assert (tmp_path / "b.set").read_bytes() == src.read_bytes()
            for i, e in enumerate(parsed):
                assert kp_equal(e["keypoints"], golden[f"{kind}{i}_kps"]) and np.array_equal(e["descriptors"], golden[f"{kind}{i}_desc"])
        img = bb.read_pgm(ref_dir / "img1.pgm")
        assert img.shape == (640, 800)


@pytest.fixture(scope="module")
def emul_describe():
    src = ROOT / "tests" / "host_emul" / "emul_describe.cc"
    lib = ROOT / "tests" / "host_emul" / "libemul_describe.so"
    csrc = ROOT / "ethzasl_brisk_b200" / "csrc"
    deps = [src, csrc / "pattern.cc", csrc / "pattern.h", csrc / "brisk_pattern_data.inc"] + list(csrc.glob("*.cuh"))
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-msse2", "-ffp-contract=off", "-fPIC", "-shared",
                        "-Wno-unknown-pragmas", "-o", str(lib), str(src), str(csrc / "pattern.cc")], check=True)
    handle = C.CDLL(str(lib))

    def describe(img, kps, rot=True, scale=True, version=2, pattern_scale=1.0):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        k = np.ascontiguousarray(kps, KP_DTYPE).copy()
        nb = C.c_int(0)
        flat = np.zeros(max(len(k), 1) * 256, np.uint8)
        n = handle.emul_describe(img.ctypes.data_as(C.c_void_p), w, h, k.ctypes.data_as(C.c_void_p), len(k), int(rot), int(scale),
                                 int(version), C.c_float(pattern_scale), flat.ctypes.data_as(C.c_void_p), C.byref(nb))
        assert n >= 0
        return k[:n].copy(), flat[:n * nb.value].reshape(n, nb.value).copy()
    return describe


@pytest.mark.parametrize("version,ps,rot,scale", [(2, 1.0, True, True), (1, 1.0, True, True), (2, 0.5, True, True), (2, 1.0, False, True),
                                                   (2, 1.0, True, False), (1, 0.7, True, True), (2, 1.3, True, True)])
def test_descriptor_logic(emul_describe, oracle, golden, version, ps, rot, scale):
    # describe_cull_kernel + describe_kernel run serially on the CPU with the kernels' own sampler (2x2-block integral
    # layout, tabulated constants), host-built pattern tables and size breaks: key points, angles and descriptor bytes
    # must equal the oracle's -- including the angle, which is the same double atan2 here
    rng = np.random.default_rng(5)
    for img in (golden["image0"], synthetic_frame(500, 333, 3)):
        kp = oracle.agast_detect(img, 45, 4)
        given = kp.copy()
        given["angle"] = rng.uniform(-359, 719, len(kp)).astype(np.float32)  # the range in which the reference is defined
        for k in (kp, given, kp[:0]):
            k1, d1 = emul_describe(img, k, rot, scale, version, ps)
            k2, d2 = oracle.describe(img, k, rot, scale, version, ps)
            assert kp_equal(k1, k2) and np.array_equal(d1.ravel(), d2.ravel())


def test_randomized_sweep_describe(emul_describe, oracle, ref):
    rng = np.random.default_rng(20261020)
    for _ in range(40):
        w, h = int(rng.integers(120, 420)), int(rng.integers(120, 330))
        img = _sweep_image(rng, w, h)
        m = 300
        k = np.zeros(m, KP_DTYPE)
        k["x"], k["y"] = rng.uniform(0, w, m), rng.uniform(0, h, m)
        k["size"] = 10.0 ** rng.uniform(0.3, 2.2, m)  # all 64 scale indices, most large ones culled at the border
        k["angle"] = np.where(rng.integers(0, 2, m) == 1, -1.0, rng.uniform(-359, 719, m))
        k["class_id"] = np.arange(m)
        version, ps = (int(rng.choice([1, 2])), float(rng.choice([1.0, 0.6, 1.4])))
        rot, scale = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        want_k, want_d = ref.describe(img, k, rot, scale, version, ps)
        k2, d2 = oracle.describe(img, k, rot, scale, version, ps)
        assert kp_equal(k2, want_k) and np.array_equal(d2.ravel(), want_d.ravel()), (w, h, version, ps, rot, scale)
        k1, d1 = emul_describe(img, k, rot, scale, version, ps)
        assert kp_equal(k1, want_k) and np.array_equal(d1.ravel(), want_d.ravel()), (w, h, version, ps, rot, scale)


def _sweep_image(rng, w, h):
    kind = int(rng.integers(0, 3))
    if kind == 0:
        return synthetic_frame(w, h, int(rng.integers(0, 10000)))
    if kind == 1:  # quantised noise: plateaus of equal scores
        return (rng.integers(0, 4, (h, w)) * 60 + rng.integers(0, 8, (h, w))).astype(np.uint8)
    return rng.integers(0, 256, (h, w)).astype(np.uint8)


def test_randomized_sweep_agast(emul, oracle, ref):
    # odd sizes (all column regimes of the samplers, tiny top layers), thresholds and depths drawn at random: the compiled
    # reference, the oracle and the device logic compiled for the host must agree on every case
    rng = np.random.default_rng(20261017)
    for _ in range(120):
        octaves = int(rng.integers(0, 5))
        lo = max(40, 12 * 2 ** max(octaves - 1, 0) + 16)
        w, h = int(rng.integers(lo, 420)), int(rng.integers(lo, 330))
        thresh = int(rng.integers(20, 95))
        img = _sweep_image(rng, w, h)
        want = ref.agast_detect(img, thresh, octaves, cap=1 << 20)
        assert kp_equal(oracle.agast_detect(img, thresh, octaves, cap=1 << 20), want), (w, h, thresh, octaves)
        got = emul(img, thresh, octaves, cap=1 << 20)
        assert not isinstance(got, int) and kp_equal(got, want), (w, h, thresh, octaves)


def test_randomized_sweep_compute_scale(emul, oracle, ref):
    # ComputeScale on random images / depths / thresholds (also below 20: the threshold only matters on layers that fall
    # back to detection) with 1 to 400 provided points, fractional or integral
    rng = np.random.default_rng(20261019)
    for _ in range(150):
        octaves = int(rng.integers(0, 5))
        lo = max(40, 12 * 2 ** max(octaves - 1, 0) + 16)
        w, h = int(rng.integers(lo, 420)), int(rng.integers(lo, 330))
        thresh = int(rng.integers(5, 95))
        img = _sweep_image(rng, w, h)
        m = int(rng.choice([1, 3, 40, 400]))
        k = np.zeros(m, KP_DTYPE)
        # a few rows away from the bottom: there the compiled reference reads its ring pixels past the image buffer
        k["x"], k["y"], k["class_id"] = rng.uniform(0, w, m), rng.uniform(0, h - 4, m), np.arange(m)
        if rng.integers(0, 2):
            k["x"], k["y"] = np.floor(k["x"]), np.floor(k["y"])
        want = ref.compute_scale(img, k, thresh, octaves, cap=1 << 20)
        assert kp_equal(oracle.compute_scale(img, k, thresh, octaves, cap=1 << 20), want), (w, h, thresh, octaves, m)
        got = emul.compute_scale(img, k, thresh, octaves, cap=1 << 20)
        assert not isinstance(got, int) and kp_equal(got, want), (w, h, thresh, octaves, m)


def test_randomized_sweep_harris(emul_harris, oracle, ref):
    rng = np.random.default_rng(20261018)
    for _ in range(120):
        octaves = int(rng.integers(0, 5))
        lo = max(40, 12 * 2 ** max(octaves - 1, 0) + 16)
        w, h = int(rng.integers(lo, 420)), int(rng.integers(lo, 330))
        radius = float(rng.choice([30.0, 10.0, 5.0, 2.5, 1.0, 17.3, 0.0, -1.0]))
        abs_thr = float(rng.choice([0.0, 20.0, 1000.0, 1e6]))
        max_kpt = int(rng.choice([-1, 50, 300, 2000])) if radius > 0 else int(rng.choice([16, 100, 400, 3000]))  # bucketing needs a limit
        img = _sweep_image(rng, w, h)
        want = ref.harris_detect(img, octaves, radius, abs_thr, max_kpt)
        assert kp_equal(oracle.harris_detect(img, octaves, radius, abs_thr, max_kpt), want), (w, h, octaves, radius, abs_thr, max_kpt)
        assert kp_equal(emul_harris(img, octaves, radius, abs_thr, max_kpt), want), (w, h, octaves, radius, abs_thr, max_kpt)


@pytest.mark.parametrize("radius,max_kpt", [(30.0, -1), (8.0, -1), (3.0, 150), (0.0, 400), (-1.0, 90), (15.0, 40)])
def test_harris_passed_keypoints_logic(emul_harris, oracle, golden, radius, max_kpt):
    # "use passed key points": import gate / truncation, std::sort replay, thinning, unrefined emit -- as the GPU path does it
    from test_oracle_golden import _passed_points
    shape = golden["image0"].shape
    for k in (_passed_points(shape, 3000, 1), _passed_points(shape, 700, 2, True), _passed_points(shape, 20, 3), _passed_points(shape, 40000, 4)):
        got = emul_harris.passed(shape, k, radius, max_kpt)
        assert not isinstance(got, int) and kp_equal(got, oracle.harris_detect_passed(shape, k, radius, max_kpt))
    low = _passed_points(shape, 50, 5)
    low["response"] = 999999.0
    assert kp_equal(emul_harris.passed(shape, low, radius, max_kpt), low)
    bad = _passed_points(shape, 5, 6)
    bad["x"][0], bad["response"][0] = shape[1] + 5.0, 5.0e6
    assert emul_harris.passed(shape, bad, radius, max_kpt) == -4


def test_provided_keypoints_golden_logic(emul, emul_harris, golden, golden_provided):
    # the device logic against the committed outputs of the compiled reference (tests/golden/provided_keypoints.npz)
    from test_oracle_golden import COMPUTE_SCALE_GOLDEN, PASSED_GOLDEN
    img = golden["image0"]
    for i, (thresh, octaves) in enumerate(COMPUTE_SCALE_GOLDEN):
        got = emul.compute_scale(img, golden_provided[f"cs{i}_in"], thresh, octaves)
        assert not isinstance(got, int) and kp_equal(got, golden_provided[f"cs{i}_out"])
    for i, (radius, max_kpt) in enumerate(PASSED_GOLDEN):
        got = emul_harris.passed(img.shape, golden_provided[f"hp{i}_in"], radius, max_kpt)
        assert not isinstance(got, int) and kp_equal(got, golden_provided[f"hp{i}_out"])


def test_capi_exports_every_declared_symbol():
    from ethzasl_brisk_b200 import build, lib_path
    build_lib = build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    assert Path(build_lib) == lib_path()
    header = (ROOT / "include" / "brisk_b200.h").read_text()
    declared = set(re.findall(r"\b(brisk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = C.CDLL(str(lib_path()))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ethzasl_brisk_b200 as bb
    with pytest.raises(bb.BriskError):
        bb.Context(0)


def test_synthetic_frames_are_deterministic():
    a = synthetic_frame(320, 240, 42)
    b = synthetic_frame(320, 240, 42)
    c = synthetic_frame(320, 240, 43)
    assert a.dtype == np.uint8 and a.shape == (240, 320)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_matcher_collection_host_logic():
    # the host part of BruteForceMatcher (no GPU): train collection -> one matrix, masks side by side, and
    # cv::DescriptorMatcher::isMaskedOut (only when every image has a non-empty mask)
    import ethzasl_brisk_b200 as bb
    m = object.__new__(bb.BruteForceMatcher)
    m.ctx = None
    m._train = [np.full((3, 48), 1, np.uint8), np.zeros((0, 48), np.uint8), np.full((2, 48), 2, np.uint8)]
    q = np.zeros((4, 48), np.uint8)
    query, trains, train, start, mask, masked_out = m._collection(q, None, None)
    assert train.shape == (5, 48) and start.tolist() == [0, 3, 3, 5] and mask is None and not masked_out.any()
    masks = [np.array([[1, 0, 0], [0, 0, 0], [0, 0, 0], [1, 1, 1]], np.uint8), None, np.array([[0, 0], [0, 0], [0, 5], [0, 0]], np.uint8)]
    _, _, _, _, mask, masked_out = m._collection(q, None, masks)
    assert mask.tolist() == [[1, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0, 0, 0, 0, 1], [1, 1, 1, 0, 0]]
    assert not masked_out.any()  # one image has no mask: no query is "masked out"
    m._train = [m._train[0], m._train[2]]
    _, _, _, _, mask, masked_out = m._collection(q, None, [masks[0], masks[2]])
    assert masked_out.tolist() == [False, True, False, False]
    assert m.isMaskSupported()


def test_std_sort_matches_replays_libstdcxx(oracle):
    # brisk_std_sort_matches (host only): the reference's final std::sort of a match list (brute-force-matcher.cc:160,210);
    # beyond 16 entries introsort permutes equal distances -- compared with the very std::sort call of the oracle
    from ethzasl_brisk_b200.api import _ptr, load_library
    lib = load_library()
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 16, 17, 33, 100, 1000):
        d = np.sort(rng.integers(150, 150 + max(2, n // 6), n)).astype(np.float32)   # many equal distances, ascending as selected
        t = rng.permutation(n).astype(np.int32)
        i = rng.integers(0, 3, n).astype(np.int32)
        want = oracle._sorted_matches([(7, int(a), int(b), float(c)) for a, b, c in zip(t, i, d)])
        assert lib.brisk_std_sort_matches(C.c_int64(n), _ptr(t), _ptr(i), _ptr(d)) == 0
        assert [(7, int(a), int(b), float(c)) for a, b, c in zip(t, i, d)] == want, n
    assert lib.brisk_std_sort_matches(C.c_int64(-1), None, None, None) != 0


@pytest.fixture(scope="module")
def emul_packed():
    src = ROOT / "tests" / "host_emul" / "emul_packed.cc"
    lib = ROOT / "tests" / "host_emul" / "libemul_packed.so"
    deps = [src] + list((ROOT / "ethzasl_brisk_b200" / "csrc").glob("*.cuh"))
    if not lib.exists() or any(d.stat().st_mtime > lib.stat().st_mtime for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(lib), str(src)], check=True)
    return C.CDLL(str(lib))


def test_packed_fast_rows_on_the_host(emul_packed, oracle):
    # fast_packed.cuh (two pixels per register, 16-bit lanes) compiled for the host with the SIMD intrinsics restated in
    # C++: dense scores of fast916_row<1..3> at every alignment against the oracle's cornerScore
    rng = np.random.default_rng(4)
    for w, h in ((67, 41), (96, 33), (131, 29)):
        for img in (synthetic_frame(w, h, 9), rng.integers(0, 256, (h, w), dtype=np.uint8),
                    (rng.integers(0, 3, (h, w)) * 100 + rng.integers(0, 5, (h, w))).astype(np.uint8)):
            want, _ = oracle.dense_scores(img)
            pitch = (w + 15) // 16 * 16
            padded = np.zeros((h, pitch), np.uint8)
            padded[:, :w] = img
            for np_ in (1, 2, 3):
                got = np.zeros((h, w), np.uint8)
                emul_packed.emul_packed_scores(padded.ctypes.data_as(C.c_void_p), w, h, pitch, np_, got.ctypes.data_as(C.c_void_p))
                assert np.array_equal(got[3:-3, 3:-3], want[3:-3, 3:-3]), (w, h, np_)


def test_packed_segment_and_compass_tests_on_the_host(emul_packed):
    # the detector's second phase on pairs (largest arc minimum / smallest arc maximum against c +- b) equals the bit-mask
    # 9-of-16 run test, and its compass pre-test (two ADJACENT compass pixels beyond c +- b) never loses a corner
    rng = np.random.default_rng(8)

    def is_corner(ring, c, b):
        br = [int(v) > c + b for v in ring]
        dk = [int(v) < c - b for v in ring]
        return any(all(m[(s + j) % 16] for j in range(9)) for m in (br, dk) for s in range(16))

    n_corner = 0
    for trial in range(4000):
        base = int(rng.integers(0, 256))
        kind = trial % 4
        rings = []
        for _ in range(2):
            if kind == 0:
                ring = rng.integers(0, 256, 16)
            elif kind == 1:   # a bright or dark arc of random length on a flat background
                ring = np.full(16, base) + rng.integers(-3, 4, 16)
                s0, ln = int(rng.integers(0, 16)), int(rng.integers(6, 13))
                ring[[(s0 + j) % 16 for j in range(ln)]] += int(rng.choice([-1, 1])) * int(rng.integers(20, 120))
            else:
                ring = base + rng.integers(-40, 41, 16)
            rings.append(np.clip(ring, 0, 255).astype(np.uint8))
        cs = [int(np.clip(base + rng.integers(-5, 6), 0, 255)) for _ in range(2)]
        bs = [int(rng.integers(1, 140)) for _ in range(2)]
        got = emul_packed.emul_packed_segment_pair(rings[0].ctypes.data_as(C.c_void_p), cs[0], bs[0], rings[1].ctypes.data_as(C.c_void_p), cs[1], bs[1])
        want = [is_corner(rings[k], cs[k], bs[k]) for k in range(2)]
        assert [bool(got & 1), bool(got & 2)] == want, (trial, rings, cs, bs)
        n_corner += sum(want)
        comp = [np.ascontiguousarray(r[[0, 4, 8, 12]]) for r in rings]   # left, up, right, down
        flags = emul_packed.emul_packed_compass_pair(comp[0].ctypes.data_as(C.c_void_p), cs[0], bs[0], comp[1].ctypes.data_as(C.c_void_p), cs[1], bs[1])
        for k in range(2):
            v = [int(x) for x in comp[k]]
            adjacent = any((v[i] > cs[k] + bs[k] and v[(i + 1) % 4] > cs[k] + bs[k]) or (v[i] < cs[k] - bs[k] and v[(i + 1) % 4] < cs[k] - bs[k])
                           for i in range(4))
            assert bool(flags & (1 << k)) == adjacent, (trial, k)
            assert adjacent or not want[k]   # the pre-test is a necessary condition
    assert n_corner > 200


def test_fp4_operand_expansion_on_the_host(emul_packed):
    # every descriptor byte: bit i -> nibble i = 0x2 (+1.0 in E2M1) when set, 0xA (-1.0) when clear; the dot product of two
    # expanded rows is K - 2 hamming, which is what the FP4 matcher turns back into distances
    emul_packed.emul_e2m1_expand_byte.restype = C.c_uint32
    val = {0x2: 1.0, 0xA: -1.0}
    words = [emul_packed.emul_e2m1_expand_byte(C.c_uint32(b)) for b in range(256)]
    for b, w in enumerate(words):
        for i in range(8):
            assert (w >> (4 * i)) & 0xF == (0x2 if (b >> i) & 1 else 0xA), (b, i)
    rng = np.random.default_rng(1)
    for _ in range(50):
        x, y = int(rng.integers(0, 256)), int(rng.integers(0, 256))
        dot = sum(val[(words[x] >> (4 * i)) & 0xF] * val[(words[y] >> (4 * i)) & 0xF] for i in range(8))
        assert dot == 8 - 2 * bin(x ^ y).count("1")
