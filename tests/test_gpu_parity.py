"""Parity of the CUDA path (through the C ABI) against the oracle, the golden
fixtures and size-independent properties.  Needs a GPU: `pytest -m gpu`."""
import numpy as np
import pytest

import ethzasl_brisk_b200 as bb
from conftest import kp_equal

pytestmark = pytest.mark.gpu


def test_native_library_is_loaded(ctx):
    # the parity claims below are about libbrisk_b200.so, not a fallback
    maps = open("/proc/self/maps").read()
    assert "libbrisk_b200.so" in maps


@pytest.mark.parametrize("w,h", [(752, 480), (800, 640), (500, 333), (1920, 1080), (193, 97), (97, 200), (250, 160)])
def test_pyramid_bit_exact(ctx, oracle, w, h):
    # widths cover all rounding regimes of the reference's SSE loops
    img = bb.synthetic_frame(w, h, 100 + w)
    for octaves in (4, 1):
        mine = ctx.debug_pyramid(img, octaves)
        want, _ = oracle.pyramid(img, octaves)
        assert len(mine) == len(want)
        for a, b in zip(mine, want):
            assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(752, 480), (641, 479), (16, 8), (17, 9), (33, 21), (1920, 1080)])
def test_samplers_16bit_bit_exact(ctx, oracle, w, h):
    # Halfsample16 / Twothirdsample16 (image-down-sampling.cc:56-139, 394-548) incl. their saturation quirks; the oracle's
    # restatement is pinned to the compiled reference in tests/test_oracle_golden.py
    rng = np.random.default_rng(w + h)
    for hi in (65536, 4096, 300):
        img = rng.integers(0, hi, (h, w)).astype(np.uint16)
        img[0, 0] = 65535; img[1, 0] = 65535; img[-1, -1] = 65535; img[-2, -1] = 65534
        assert np.array_equal(ctx.halfsample16(img), oracle.halfsample16(img))
        assert np.array_equal(ctx.twothirdsample16(img), oracle.twothirdsample16(img))
    with pytest.raises(bb.BriskError):
        ctx.halfsample16(np.zeros((10, 15), np.uint16))
    with pytest.raises(bb.BriskError):
        ctx.twothirdsample16(np.zeros((10, 11), np.uint16))
    import torch
    t = torch.from_numpy(img.astype(np.int32)).to(torch.int16).cuda()   # device-resident, pitched views
    src = torch.zeros((h, w + 6), dtype=torch.int16, device="cuda"); src[:, :w] = t
    dst = torch.zeros((h // 2, w // 2 + 4), dtype=torch.int16, device="cuda")
    import ctypes as C
    ctx._check(ctx._lib.brisk_halfsample16(ctx._h, bb.api._ptr(src), w, h, C.c_size_t(2 * (w + 6)), bb.api._ptr(dst), C.c_size_t(2 * (w // 2 + 4))))
    assert np.array_equal(dst[:, :w // 2].cpu().numpy().view(np.uint16), oracle.halfsample16(img))


def test_pyramid_six_octaves(ctx, oracle):
    img = bb.synthetic_frame(1600, 1200, 5)
    mine = ctx.debug_pyramid(img, 6)
    want, _ = oracle.pyramid(img, 6)
    assert len(mine) == 12
    for a, b in zip(mine, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(752, 480), (641, 479), (131, 67)])
def test_integral_bit_exact(ctx, oracle, w, h):
    img = bb.synthetic_frame(w, h, 7)
    assert np.array_equal(ctx.debug_integral(img), oracle.integral8(img))
    white = np.full((h, w), 255, np.uint8)  # int32 head-room
    assert ctx.debug_integral(white)[-1, -1] == 255 * w * h


@pytest.mark.parametrize("w,h", [(752, 480), (641, 479), (131, 67), (500, 333), (97, 200)])
def test_dense_fast_scores_bit_exact(ctx, oracle, w, h):
    # the packed row evaluator of the NMS kernels (fast_packed.cuh: pairs of pixels per instruction, word loads at every
    # alignment) against the oracle's cornerScore on every pixel; also the 5-8 mask
    rng = np.random.default_rng(w)
    for img in (bb.synthetic_frame(w, h, 11), rng.integers(0, 256, (h, w), dtype=np.uint8),
                (rng.integers(0, 3, (h, w)) * 100 + rng.integers(0, 5, (h, w))).astype(np.uint8)):
        a1, b1 = ctx.debug_scores(img)
        a2, b2 = oracle.dense_scores(img)
        assert np.array_equal(a1, a2) and np.array_equal(b1, b2)


def test_raw_corners_match_detector_order(ctx, oracle, golden):
    img = golden["image0"]
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    c, lc = det.debug_corners(img)
    pyr, _ = oracle.pyramid(img, 4)
    off = 0
    for li, layer in enumerate(pyr):
        _, want = oracle.layer_dump(layer, 60)
        assert np.array_equal(c[off:off + lc[li]], want)  # x, y, score in raster order
        off += lc[li]


@pytest.mark.parametrize("i", [0, 1])
def test_golden_ast_fixture(ctx, golden, i):
    # the reference's own ValidationAST fixture (test-binary-equal.cc:315-333), bit for bit
    img = golden[f"image{i}"]
    det = bb.BriskFeatureDetector(70, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    k, d = ext.compute(img, det.detect(img))
    gk, gd = golden[f"ast{i}_kps"], golden[f"ast{i}_desc"]
    assert len(k) == len(gk)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k[f], gk[f]), f
    # angle: double atan2 on the device (<= 2 ulp) then rounded to float; tolerance 1e-4 degree and
    # identical descriptors (same rotation bin) -- north_star's stated bar
    assert np.abs(k["angle"] - gk["angle"]).max() <= 1e-4
    assert np.array_equal(d, gd)


@pytest.mark.parametrize("thresh,octaves", [(60, 4), (70, 3), (40, 2), (70, 0), (60, 1), (30, 4), (20, 3), (24, 4), (15, 3), (10, 4)])
def test_detect_bit_exact(ctx, oracle, golden, thresh, octaves):
    # thresholds below 20 are served while no detected corner scores <= 2 (none does on these images down to 10)
    det = bb.BriskFeatureDetector(thresh, octaves, ctx=ctx)
    det.set_corner_capacity(400000)
    for img in (golden["image0"], bb.synthetic_frame(752, 480, 1000), bb.synthetic_frame(500, 333, 3)):
        assert kp_equal(det.detect(img, cap=400000), oracle.agast_detect(img, thresh, octaves, cap=1 << 20))


@pytest.mark.parametrize("w,h,thresh,octaves", [(752, 480, 60, 5), (752, 480, 45, 6), (1600, 1200, 60, 6), (1920, 1080, 60, 5), (1024, 520, 30, 6)])
def test_detect_deep_pyramids(ctx, oracle, w, h, thresh, octaves):
    # 10- and 12-layer pyramids (kMaxLayers paths of the pyramid tiles, corner lists and the chain kernel)
    img = bb.synthetic_frame(w, h, 4000 + w)
    det = bb.BriskFeatureDetector(thresh, octaves, ctx=ctx)
    det.set_corner_capacity(400000)
    got = det.detect(img, cap=400000)
    assert kp_equal(got, oracle.agast_detect(img, thresh, octaves, cap=1 << 20))
    assert got["octave"].max() >= 2 * octaves - 3


def test_config4_4k_six_octaves(ctx, oracle):
    # BASELINE config 4: 3840x2160, 6 octaves (12 layers, 13-bit corner coordinates), detect + describe on two frames of
    # the bench's own generator settings, bit for bit against the oracle; plus batch invariance
    frames = np.stack([bb.synthetic_frame(3840, 2160, 3000 + i, n_shapes=4200) for i in range(2)])
    det = bb.BriskFeatureDetector(60, 6, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=65536)
    for f in range(2):
        k2, d2 = oracle.describe(frames[f], oracle.agast_detect(frames[f], 60, 6, cap=1 << 20))
        n = counts[f]
        assert n == len(k2) and n > 10000
        for fld in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(kps[f, :n][fld], k2[fld]), fld
        assert np.abs(kps[f, :n]["angle"] - k2["angle"]).max() <= 1e-4
        assert np.array_equal(desc[f, :n], d2)
    k1, c1, d1 = bb.detect_and_compute_batch(det, ext, frames[1:2], cap=65536)
    assert c1[0] == counts[1] and np.array_equal(d1[0, :c1[0]], desc[1, :counts[1]])


def test_detect_without_scale_suppression_single_layer(ctx, oracle, ref, golden):
    # suppressScaleNonmaxima = false is defined for one layer only (the reference indexes layer 0's corner list
    # with the other layers' counts, brisk-scale-space.cc:137); there it equals the compiled reference
    for img in (golden["image0"], bb.synthetic_frame(500, 333, 3)):
        got = bb.BriskFeatureDetector(55, 0, False, ctx=ctx).detect(img)
        assert kp_equal(got, oracle.agast_detect(img, 55, 0, False)) and len(got) > 50
        if ref is not None:
            assert kp_equal(got, ref.agast_detect(img, 55, 0, False))


@pytest.mark.parametrize("thresh,octaves", [(60, 4), (70, 3), (40, 2), (70, 0), (60, 1), (30, 5)])
def test_compute_scale_bit_exact(ctx, oracle, golden, thresh, octaves):
    # BriskFeatureDetector::ComputeScale (provided key points, brisk-scale-space.cc:104-124)
    from test_oracle_golden import _provided_points
    det = bb.BriskFeatureDetector(thresh, octaves, ctx=ctx)
    det.set_corner_capacity(300000)  # the top layers of a 5-octave pyramid keep no point and are detected on
    for img in (golden["image0"], bb.synthetic_frame(752, 480, 1000), bb.synthetic_frame(500, 333, 3)):
        for k in (oracle.agast_detect(img, max(thresh, 40), min(octaves, 4)), _provided_points(img, 2000, 21),
                  _provided_points(img, 500, 22, True)):
            got = det.compute_scale(img, k)
            assert len(got) > 5 and kp_equal(got, oracle.compute_scale(img, k, thresh, octaves))


@pytest.mark.parametrize("thresh,octaves", [(60, 3), (30, 5), (12, 3), (70, 0)])
def test_compute_scale_layers_without_points(ctx, oracle, golden, thresh, octaves):
    # a layer that keeps none of the points is detected on (threshold map without lower bound, brisk-layer.cc:103-105)
    # and its corners go through the scale checks like provided points; mixed with provided layers in one frame
    from test_oracle_golden import _provided_points
    det = bb.BriskFeatureDetector(thresh, octaves, ctx=ctx)
    det.set_corner_capacity(300000)
    few = _provided_points(golden["image0"], 3, 1)
    few["x"], few["y"] = [5, 6.5, 30], [5, 7.25, 9]
    for img in (golden["image0"], bb.synthetic_frame(500, 333, 3)):
        for k in (few, _provided_points(img, 40, 5)):
            assert kp_equal(det.compute_scale(img, k, cap=400000), oracle.compute_scale(img, k, thresh, octaves))
    # a batch in which only some frames have such layers
    imgs = np.stack([bb.synthetic_frame(500, 333, 60 + i) for i in range(3)])
    lists = [few, _provided_points(imgs[1], 600, 6), few[:1]]
    kin = np.zeros((3, 600), bb.KP_DTYPE)
    for i, k in enumerate(lists):
        kin[i, :len(k)] = k
    out, oc = det.compute_scale_batch(imgs, kin, [len(k) for k in lists], cap=400000)
    for i in range(3):
        assert kp_equal(out[i, :oc[i]], oracle.compute_scale(imgs[i], lists[i], thresh, octaves))


def test_compute_scale_full_size_1080p(ctx, oracle):
    # BASELINE config 3's frame size: the key points the detector found, fed back through ComputeScale
    img = bb.synthetic_frame(1920, 1080, 2000)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    kps = det.detect(img)
    got = det.compute_scale(img, kps)
    assert len(kps) > 3000 and len(got) > len(kps)
    assert kp_equal(got, oracle.compute_scale(img, kps, 60, 4))
    # properties that hold at any size: octave is the layer that accepted the point, sizes follow the layer scale
    assert got["octave"].min() >= 0 and got["octave"].max() <= 7 and np.all(np.diff(got["octave"]) >= 0)
    assert np.all(got["size"] >= 12.0 * 0.5) and np.all(got["angle"] == -1.0)


def test_provided_keypoints_golden_fixture(ctx, golden, golden_provided):
    # outputs of the compiled reference, committed as tests/golden/provided_keypoints.npz (tests/golden/make_provided_keypoints.py)
    from test_oracle_golden import COMPUTE_SCALE_GOLDEN, PASSED_GOLDEN
    img = golden["image0"]
    for i, (thresh, octaves) in enumerate(COMPUTE_SCALE_GOLDEN):
        det = bb.BriskFeatureDetector(thresh, octaves, ctx=ctx)
        det.set_corner_capacity(300000)
        assert kp_equal(det.compute_scale(img, golden_provided[f"cs{i}_in"], cap=400000), golden_provided[f"cs{i}_out"])
    for i, (radius, max_kpt) in enumerate(PASSED_GOLDEN):
        det = bb.ScaleSpaceFeatureDetector(0, radius, 0.0, None if max_kpt < 0 else max_kpt, ctx=ctx)
        assert kp_equal(det.detect(img, keypoints=golden_provided[f"hp{i}_in"]), golden_provided[f"hp{i}_out"])


def test_compute_scale_batch_and_errors(ctx, oracle):
    from test_oracle_golden import _provided_points
    det = bb.BriskFeatureDetector(60, 3, ctx=ctx)
    imgs = np.stack([bb.synthetic_frame(640, 360, 40 + i) for i in range(5)])
    lists = [_provided_points(imgs[i], 300 + 50 * i, 30 + i) for i in range(5)]
    cap_in = max(len(k) for k in lists)
    kin = np.zeros((5, cap_in), bb.KP_DTYPE)
    for i, k in enumerate(lists):
        kin[i, :len(k)] = k
    counts = np.array([len(k) for k in lists], np.int32)
    out, oc = det.compute_scale_batch(imgs, kin, counts, cap=6 * cap_in)
    for i in range(5):
        assert kp_equal(out[i, :oc[i]], oracle.compute_scale(imgs[i], lists[i], 60, 3))
    # the same batch in chunks (small workspace limit) and with device-resident frames
    import torch
    small = bb.Context(0, workspace_limit=64 << 20)
    det2 = bb.BriskFeatureDetector(60, 3, ctx=small)
    det2.set_corner_capacity(200000)  # 20 MB of per-frame scratch: two frames per chunk
    out2, oc2 = det2.compute_scale_batch(torch.from_numpy(imgs).cuda(), kin, counts, cap=6 * cap_in)
    assert np.array_equal(oc, oc2) and all(kp_equal(out[i, :oc[i]], out2[i, :oc[i]]) for i in range(5))
    # truncation is reported with the true counts
    with pytest.raises(bb.BriskError):
        det.compute_scale_batch(imgs, kin, counts, cap=10)
    # an empty list / the Harris detector: refused, not guessed
    with pytest.raises(bb.BriskError):
        det.compute_scale_batch(imgs[:1], kin[:1], [0])
    with pytest.raises(bb.BriskError):
        bb.ScaleSpaceFeatureDetector(2, 30.0, 20.0, ctx=ctx).compute_scale(imgs[0], lists[0])


def test_detect_tie_heavy(ctx, oracle):
    rng = np.random.default_rng(7)
    det = bb.BriskFeatureDetector(35, 3, ctx=ctx)
    det.set_corner_capacity(100000)
    for _ in range(3):
        img = (rng.integers(0, 4, (240, 320)) * 60 + rng.integers(0, 8, (240, 320))).astype(np.uint8)
        assert kp_equal(det.detect(img), oracle.agast_detect(img, 35, 3))


def test_detect_mask(ctx, oracle, golden):
    img = golden["image1"]
    mask = np.zeros_like(img)
    mask[50:600, 100:650] = 1
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    assert kp_equal(det.detect(img, mask), oracle.agast_detect(img, 60, 4, mask=mask))


def test_detect_empty_and_flat(ctx):
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    assert len(det.detect(np.full((480, 752), 90, np.uint8))) == 0
    kps, counts = det.detect_batch(np.zeros((0, 480, 752), np.uint8))
    assert kps.shape[0] == 0 and counts.shape[0] == 0


def test_unsupported_configurations_fail_loudly(ctx, golden):
    img = bb.synthetic_frame(320, 240, 1)
    for bad in (bb.BriskFeatureDetector(0, 4, ctx=ctx), bb.BriskFeatureDetector(256, 4, ctx=ctx), bb.BriskFeatureDetector(60, 7, ctx=ctx),
                bb.BriskFeatureDetector(60, 4, False, ctx=ctx)):
        with pytest.raises(bb.BriskError):
            bad.detect(img)
    # thresh < 20: refused only when a detected corner really scores <= 2 (the reference's cache does not keep such scores)
    low = bb.BriskFeatureDetector(5, 3, ctx=ctx)
    low.set_corner_capacity(400000)
    with pytest.raises(bb.BriskError, match="scores <= 2"):
        low.detect(golden["image0"], cap=400000)
    with pytest.raises(bb.BriskError):
        bb.BriskDescriptorExtractor(True, True, 3, ctx=ctx)


@pytest.mark.parametrize("version,ps,rot,scale", [(2, 1.0, True, True), (1, 1.0, True, True), (2, 0.5, True, True),
                                                   (2, 1.0, False, True), (2, 1.0, True, False), (1, 0.7, True, True)])
def test_describe_bit_exact_given_keypoints(ctx, oracle, golden, version, ps, rot, scale):
    ext = bb.BriskDescriptorExtractor(rot, scale, version, ps, ctx=ctx)
    assert ext.descriptorSize() == oracle.pattern_dump(version, ps)["strings"]
    for img in (golden["image0"], bb.synthetic_frame(500, 333, 3)):
        kp = oracle.agast_detect(img, 45, 4)
        k1, d1 = ext.compute(img, kp)
        k2, d2 = oracle.describe(img, kp, rot, scale, version, ps)
        assert len(k1) == len(k2)
        for f in ("x", "y", "size", "response", "octave", "class_id"):
            assert np.array_equal(k1[f], k2[f]), f
        assert np.abs(k1["angle"] - k2["angle"]).max() <= 1e-4
        assert np.array_equal(d1, d2)


def test_describe_given_angles_and_edge_cases(ctx, oracle, golden):
    img = golden["image1"]
    kp = oracle.agast_detect(img, 60, 4)
    kp["angle"] = np.linspace(-359, 719, len(kp)).astype(np.float32)  # caller-supplied angles, incl. wrap-around
    kp["angle"][::7] = -1
    kp["size"][::11] = 0.0   # log(0) path -> scale index 0
    kp["size"][::13] = 1e9   # saturates at scale index 63 -> culled by the border test
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    k1, d1 = ext.compute(img, kp)
    k2, d2 = oracle.describe(img, kp)
    assert len(k1) == len(k2) and np.array_equal(d1, d2)
    k0, d0 = ext.compute(img, kp[:0])
    assert len(k0) == 0 and d0.shape == (0, 48)


def test_pattern_tables_match_reference_construction(ctx, oracle):
    for version, ps in ((2, 1.0), (1, 1.0), (2, 0.5)):
        mine = bb.BriskDescriptorExtractor(True, True, version, ps, ctx=ctx).pattern()
        want = oracle.pattern_dump(version, ps)
        for key in ("pts", "scale_list", "size_list", "short_pairs", "long_pairs"):
            assert np.array_equal(mine[key], want[key]), (version, ps, key)


def test_fused_batch_matches_per_frame(ctx, oracle):
    frames = bb.synthetic_batch(5, 752, 480, 1000)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=8192)
    for f in range(len(frames)):
        k2, d2 = oracle.describe(frames[f], oracle.agast_detect(frames[f], 60, 4))
        n = counts[f]
        assert n == len(k2)
        for fld in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(kps[f, :n][fld], k2[fld])
        assert np.array_equal(desc[f, :n], d2)


def test_chunked_batch_equals_single_pass(ctx):
    # small workspace limit forces several chunks; results must not depend on chunking
    frames = bb.synthetic_batch(6, 640, 480, 50)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    a = bb.detect_and_compute_batch(det, ext, frames, cap=4096)
    small = bb.Context(0, workspace_limit=96 << 20)
    det2 = bb.BriskFeatureDetector(60, 4, ctx=small)
    ext2 = bb.BriskDescriptorExtractor(ctx=small)
    b = bb.detect_and_compute_batch(det2, ext2, frames, cap=4096)
    assert np.array_equal(a[1], b[1])
    for f in range(len(frames)):
        n = a[1][f]
        assert np.array_equal(a[0][f, :n], b[0][f, :n]) and np.array_equal(a[2][f, :n], b[2][f, :n])
    # many small chunks of host frames: the two-slot pipeline with its short first chunk
    frames = bb.synthetic_batch(13, 640, 480, 70)
    a = bb.detect_and_compute_batch(det, ext, frames, cap=4096)
    tiny = bb.Context(0, workspace_limit=64 << 20)
    b = bb.detect_and_compute_batch(bb.BriskFeatureDetector(60, 4, ctx=tiny), bb.BriskDescriptorExtractor(ctx=tiny), frames, cap=4096)
    assert np.array_equal(a[1], b[1]) and a[1].min() > 0
    for f in range(len(frames)):
        n = a[1][f]
        assert np.array_equal(a[0][f, :n], b[0][f, :n]) and np.array_equal(a[2][f, :n], b[2][f, :n])


def test_async_calls_chain_and_match_blocking_calls(ctx):
    # brisk_detect_describe_async: batches streamed through two sets of host buffers; every batch must equal the blocking
    # call's result, whether a call covers one chunk (slots alternate from call to call) or several
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    for n_frames, limit in ((3, None), (7, 64 << 20)):
        batches = [bb.synthetic_batch(n_frames, 640, 480, 300 + 10 * b) for b in range(5)]
        want = [bb.detect_and_compute_batch(det, ext, f, cap=4096) for f in batches]
        actx = bb.Context(0, workspace_limit=limit) if limit else bb.Context(0)
        adet, aext = bb.BriskFeatureDetector(60, 4, ctx=actx), bb.BriskDescriptorExtractor(ctx=actx)
        bufs = [(np.zeros((n_frames, 4096), bb.KP_DTYPE), np.zeros(n_frames, np.int32), np.zeros((n_frames, 4096, 48), np.uint8)) for _ in range(2)]
        got = []
        for b, frames in enumerate(batches):
            bb.detect_and_compute_batch(adet, aext, frames, cap=4096, out=bufs[b % 2], async_=True)
            if b >= 1:   # batch b - 1 is complete once the call for batch b has returned
                got.append(tuple(x.copy() for x in bufs[(b - 1) % 2]))
        actx.sync()
        got.append(tuple(x.copy() for x in bufs[(len(batches) - 1) % 2]))
        for w, g in zip(want, got):
            assert np.array_equal(w[1], g[1]) and w[1].min() > 0
            for f in range(n_frames):
                n = w[1][f]
                assert np.array_equal(w[0][f, :n], g[0][f, :n]) and np.array_equal(w[2][f, :n], g[2][f, :n])
        # a blocking call in between collects the pending chunk first; a shape change does too
        bb.detect_and_compute_batch(adet, aext, batches[0], cap=4096, out=bufs[0], async_=True)
        mid = bb.detect_and_compute_batch(adet, aext, batches[1][:2], cap=4096)
        assert np.array_equal(bufs[0][1], want[0][1]) and np.array_equal(mid[1], want[1][1][:2])
        n0 = want[0][1][n_frames - 1]
        assert np.array_equal(bufs[0][2][n_frames - 1, :n0], want[0][2][n_frames - 1, :n0])
        # device-side errors of an asynchronous call surface at sync
        small = (np.zeros((n_frames, 16), bb.KP_DTYPE), np.zeros(n_frames, np.int32), np.zeros((n_frames, 16, 48), np.uint8))   # kept alive until the sync
        bb.detect_and_compute_batch(adet, aext, batches[0], cap=16, out=small, async_=True)
        with pytest.raises(bb.BriskError) as e:
            actx.sync()
        assert e.value.code == -4
        actx.sync()   # reported once


def test_device_resident_inputs_and_outputs(ctx):
    import torch
    frames = bb.synthetic_batch(3, 752, 480, 1000)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    host = bb.detect_and_compute_batch(det, ext, frames, cap=4096)
    dev_frames = torch.from_numpy(frames).cuda()
    kps = torch.zeros((3, 4096, 7), dtype=torch.float32, device="cuda")
    counts = torch.zeros(3, dtype=torch.int32, device="cuda")
    desc = torch.zeros((3, 4096, 48), dtype=torch.uint8, device="cuda")
    bb.detect_and_compute_batch(det, ext, dev_frames, cap=4096, out=(kps, counts, desc))
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy(), host[1])
    for f in range(3):
        n = host[1][f]
        assert np.array_equal(kps[f, :n].cpu().numpy().view(np.uint8).reshape(n, 28), host[0][f, :n].view(np.uint8).reshape(n, 28))
        assert np.array_equal(desc[f, :n].cpu().numpy(), host[2][f, :n])


def test_capacity_overflow_is_reported(ctx):
    img = bb.synthetic_frame(752, 480, 1000)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    with pytest.raises(bb.BriskError) as e:
        det.detect(img, cap=10)
    assert e.value.code == -4
    # the fused call reports it too (the border cull must not hide the detector's overflow), with the true count
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    full = len(det.detect(img))
    with pytest.raises(bb.BriskError) as e:
        bb.detect_and_compute_batch(det, ext, img[None], cap=full - 1)
    assert e.value.code == -4
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, img[None], cap=full)
    assert 0 < counts[0] <= full
    import ctypes as C
    k = np.zeros((1, 16), bb.KP_DTYPE); c = np.zeros(1, np.int32); d = np.zeros((1, 16, 48), np.uint8)
    rc = ctx._lib.brisk_detect_describe(ctx._h, det._h, ext._h, bb.api._ptr(img), 1, 752, 480, C.c_size_t(752), C.c_size_t(752 * 480), None,
                                        bb.api._ptr(k), bb.api._ptr(c), 16, bb.api._ptr(d))
    assert rc == -4 and c[0] == full
    det.set_corner_capacity(64)
    with pytest.raises(bb.BriskError):
        det.detect(img)


@pytest.mark.parametrize("nbytes", [48, 64])
@pytest.mark.parametrize("k", [1, 2, 3, 8])
def test_knn_bit_exact(ctx, oracle, nbytes, k):
    q = bb.random_descriptors(700, nbytes, 5)
    t = bb.random_descriptors(5000, nbytes, 6)
    t[100] = q[3]
    t[200] = q[3]  # exact ties: lowest train index first (reference brute-force-matcher.cc:138-157)
    m = bb.BruteForceMatcher(ctx=ctx)
    i1, d1 = m.knn(q, t, k)
    i2, d2 = oracle.knn(q, t, k)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert i1[3, 0] == 100 and d1[3, 0] == 0
    from oracle import ref as r
    if r.available():  # the reference class itself on a slice of the queries
        lists = r.knn_match(q[:64], [t], k)
        assert [[mm[1] for mm in v] for v in lists] == i1[:64].tolist() and [[int(mm[3]) for mm in v] for v in lists] == d1[:64].tolist()


def test_reference_match_test_homography(ctx, golden):
    # the reference's own matching test (test-match.cc:50-116) end to end on the GPU: BriskFeatureDetector(70, 2) + BRISK2 on
    # both images, best match per query; every match with Hamming distance < 50 must agree with H_1to2 within 5 px
    H = np.array([[8.7976964e-01, 3.1245438e-01, -3.9430589e+01], [-1.8389418e-01, 9.3847198e-01, 1.5315784e+02],
                  [1.9641425e-04, -1.6015275e-05, 1.0000000e+00]])
    det = bb.BriskFeatureDetector(70, 2, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    k1, d1 = ext.compute(golden["image0"], det.detect(golden["image0"]))
    k2, d2 = ext.compute(golden["image1"], det.detect(golden["image1"]))
    idx, dist = bb.BruteForceMatcher(ctx=ctx).knn(d1, d2, 1)
    good = np.flatnonzero(dist[:, 0] < 50)
    p = (H @ np.stack([k1["x"][good], k1["y"][good], np.ones(len(good))]).astype(np.float64))
    p = p[:2] / p[2]
    t = idx[good, 0]
    err = np.hypot(p[0] - k2["x"][t], p[1] - k2["y"][t])
    assert len(good) > 50 and np.all(err <= 5)


def test_knn_edge_cases(ctx, oracle):
    m = bb.BruteForceMatcher(ctx=ctx)
    q = bb.random_descriptors(5, 64, 1)
    t = bb.random_descriptors(1, 64, 2)
    idx, dist = m.knn(q, t, 2)  # fewer train rows than k
    assert np.all(idx[:, 0] == 0) and np.all(idx[:, 1] == -1) and np.all(dist[:, 1] == -1)
    assert bb.Hamming(ctx)(q[0], t[0]) == oracle.hamming(q[0], t[0])
    # reference test-popcount.cc: one hand-checkable 128-bit pair inside a 48-byte row
    a = np.zeros(48, np.uint8)
    b = np.zeros(48, np.uint8)
    a[0], b[0], a[17], b[40] = 0xFF, 0x0F, 0x01, 0x80
    assert bb.Hamming(ctx)(a, b) == 4 + 1 + 1


def test_hamming_primitive(ctx):
    # the reference's own known-answer test (test-popcount.cc:60-113): one 128-bit pair, bit loop vs primitive
    d1 = np.zeros(16, np.uint8)
    d2 = np.zeros(16, np.uint8)
    for i, v in {0: 0x5, 3: 0x2, 6: 0x34, 8: 0x7, 10: 0x23, 13: 0x45, 15: 0x78}.items():
        d1[i] = v
    for i, v in {0: 0x22, 3: 0x78, 6: 0x12, 8: 0x32, 10: 0x1, 13: 0x23, 15: 0x75}.items():
        d2[i] = v
    want = int(np.unpackbits(d1 ^ d2).sum())
    ham = bb.Hamming(ctx)
    assert ham(d1, d2) == want and bb.Hamming.PopcntofXORed(d1, d2, 1, ctx=ctx) == want
    # any size: size // 16 whole 128-bit words are counted (hamming.h:101-113), trailing bytes ignored; batches; unaligned rows
    rng = np.random.default_rng(3)
    for nbytes in (16, 32, 48, 64, 100, 128, 15, 1000):
        a = rng.integers(0, 256, (257, nbytes), dtype=np.uint8)
        b = rng.integers(0, 256, (257, nbytes), dtype=np.uint8)
        used = (nbytes // 16) * 16
        expect = np.unpackbits(a[:, :used] ^ b[:, :used], axis=1).sum(axis=1).astype(np.int32) if used else np.zeros(257, np.int32)
        assert np.array_equal(ham.pairs(a, b), expect)
        assert ham(a[5], b[5]) == expect[5]
    import ctypes as C
    import torch
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.zeros(257, dtype=torch.int32, device="cuda")
    ctx._check(ctx._lib.brisk_hamming_distance(ctx._h, bb.api._ptr(ta), bb.api._ptr(tb), C.c_int64(257), 1000, bb.api._ptr(out)))
    assert np.array_equal(out.cpu().numpy(), expect)


def test_knn_large_properties(ctx):
    # BASELINE config 5 shape at reduced size: idempotence (train == query -> distance 0 at own
    # index), sortedness, and agreement of a split train set with the unsplit one.
    import torch
    t = torch.from_numpy(bb.random_descriptors(200000, 64, 6)).cuda()
    q = t[:20000].clone()
    m = bb.BruteForceMatcher(ctx=ctx)
    idx, dist = m.knn(q, t, 2)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    assert np.array_equal(idx[:, 0], np.arange(20000)) and np.all(dist[:, 0] == 0)
    assert np.all(dist[:, 1] >= dist[:, 0])
    # sharded: two halves -> keys -> merge == unsplit
    keys = torch.zeros((2, 20000, 2), dtype=torch.int64, device="cuda")
    m.knn_keys(q, t[:100000], 2, 0, keys[0])
    m.knn_keys(q, t[100000:], 2, 100000, keys[1])
    i2 = torch.zeros((20000, 2), dtype=torch.int32, device="cuda")
    d2 = torch.zeros((20000, 2), dtype=torch.int32, device="cuda")
    m.merge_keys(keys, 2, 20000, 2, i2, d2)
    assert np.array_equal(i2.cpu().numpy(), idx) and np.array_equal(d2.cpu().numpy(), dist)


def test_full_size_properties_1080p(ctx, oracle):
    # BASELINE config 3 frame size: one frame against the oracle, and batch invariance
    # (a frame's result must not depend on its neighbours in the batch).
    frames = bb.synthetic_batch(3, 1920, 1080, 2000)
    det = bb.BriskFeatureDetector(60, 4, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=32768)
    k2, d2 = oracle.describe(frames[0], oracle.agast_detect(frames[0], 60, 4))
    n = counts[0]
    assert n == len(k2) and np.array_equal(desc[0, :n], d2)
    for fld in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(kps[0, :n][fld], k2[fld])
    k1, c1, d1 = bb.detect_and_compute_batch(det, ext, frames[2:3], cap=32768)
    assert c1[0] == counts[2] and np.array_equal(d1[0, :c1[0]], desc[2, :counts[2]])


def test_cpp_dropin_classes(tmp_path, oracle, golden):
    # the header-only C++ classes (include/brisk/brisk.h) used like the reference's, bit for bit
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "dropin_main"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "dropin_main.cc"), "-o", str(exe),
                    f"-L{ROOT / 'ethzasl_brisk_b200'}", "-lbrisk_b200", f"-Wl,-rpath,{ROOT / 'ethzasl_brisk_b200'}"], check=True)
    img = golden["image0"]
    pgm = tmp_path / "img.pgm"
    with open(pgm, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())
    out = tmp_path / "out.bin"
    mout = tmp_path / "matches.bin"
    subprocess.run([str(exe), str(pgm), str(out), str(mout)], check=True)
    raw = out.read_bytes()
    n, nb, self_matches = np.frombuffer(raw[:12], np.int32)
    kps = np.frombuffer(raw[12:12 + n * 28], bb.KP_DTYPE)
    desc = np.frombuffer(raw[12 + n * 28:12 + n * 28 + n * nb], np.uint8).reshape(n, nb)
    off = 12 + n * 28 + n * nb
    hn = int(np.frombuffer(raw[off:off + 4], np.int32)[0])
    hk = np.frombuffer(raw[off + 4:off + 4 + hn * 28], bb.KP_DTYPE)
    hd = np.frombuffer(raw[off + 4 + hn * 28:off + 4 + hn * 76], np.uint8).reshape(hn, 48)
    off += 4 + hn * 76
    cn = int(np.frombuffer(raw[off:off + 4], np.int32)[0])
    cs = np.frombuffer(raw[off + 4:off + 4 + cn * 28], bb.KP_DTYPE)
    off += 4 + cn * 28
    pn = int(np.frombuffer(raw[off:off + 4], np.int32)[0])
    pk = np.frombuffer(raw[off + 4:off + 4 + pn * 28], bb.KP_DTYPE)
    off += 4 + pn * 28
    ham, same_again = (int(v) for v in np.frombuffer(raw[off:off + 8], np.int32))
    assert hn == len(golden["harris0_kps"]) and np.array_equal(hk["x"], golden["harris0_kps"]["x"]) and np.array_equal(hd, golden["harris0_desc"])
    gk, gd = golden["ast0_kps"], golden["ast0_desc"]
    assert n == len(gk) and nb == 48 and self_matches == n
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], gk[f]), f
    assert np.array_equal(desc, gd)
    # ComputeScale on the detected key points (their `angle` was written by compute(); it is not read)
    assert kp_equal(cs, oracle.compute_scale(img, kps, 70, 3))
    # Hamming::PopcntofXORed on the first two descriptors; detectAndCompute overrides reproduce detect() + compute()
    assert ham == int(np.unpackbits(desc[0] ^ desc[1]).sum()) and same_again == 1
    # the Harris detector handed its own key points back (non-empty vector: re-filtering instead of detection)
    assert pn > 10 and kp_equal(pk, oracle.harris_detect_passed(img.shape, hk, 30.0, -1))
    # matcher surface: masked knnMatch over a two-image collection, radiusMatch with compactResult
    nq, n0 = min(n, 150), n // 3
    trains = [desc[:n0], desc[n0:]]
    qi = np.arange(nq)[:, None]
    masks = [((qi % 7 != 3) & ((qi + np.arange(n0)[None, :]) % 3 != 0)).astype(np.uint8),
             ((qi % 7 != 3) & ((qi + 2 * np.arange(n - n0)[None, :]) % 5 != 0)).astype(np.uint8)]
    mraw = np.frombuffer(mout.read_bytes(), np.int32)

    def read_lists(pos):
        lists = int(mraw[pos]); pos += 1
        res = []
        for _ in range(lists):
            c = int(mraw[pos]); pos += 1
            rec = mraw[pos:pos + 4 * c].reshape(c, 4)
            res.append([(int(r[0]), int(r[1]), int(r[2]), float(r[3:4].view(np.float32)[0])) for r in rec])
            pos += 4 * c
        return res, pos
    knn_lists, pos = read_lists(0)
    rad_lists, pos = read_lists(pos)
    assert knn_lists == oracle.knn_match(desc[:nq], trains, 3, masks, False)
    assert rad_lists == oracle.radius_match(desc[:nq], trains, 45.0, masks, True)


def test_cpp_dropin_opencv_mode(tmp_path, oracle, golden):
    # BRISK_B200_USE_OPENCV: the classes derive from cv::Feature2D / cv::DescriptorMatcher and are driven through those
    # base classes (cv::Ptr holders, detect / compute / knnMatch / radiusMatch forwarding to the overridden virtuals, the
    # cv:: aliases).  Compiled against the stand-in OpenCV headers of oracle/shim (OpenCV is not installed here).
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "dropin_cv"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}", f"-I{ROOT / 'oracle' / 'shim'}",
                    str(ROOT / "tests" / "cpp" / "dropin_opencv_main.cc"), "-o", str(exe), f"-L{ROOT / 'ethzasl_brisk_b200'}", "-lbrisk_b200",
                    f"-Wl,-rpath,{ROOT / 'ethzasl_brisk_b200'}"], check=True)
    img = golden["image0"]
    pgm = tmp_path / "img.pgm"
    bb.write_pgm(pgm, img)
    out = tmp_path / "out.bin"
    subprocess.run([str(exe), str(pgm), str(out)], check=True)
    raw = out.read_bytes()
    n, nb, self_matches, hn, bn, nrad, nmax, s_int = (int(v) for v in np.frombuffer(raw[:32], np.int32))
    s_dbl = float(np.frombuffer(raw[32:40], np.float64)[0])
    off = 40
    kps = np.frombuffer(raw[off:off + n * 28], bb.KP_DTYPE); off += n * 28
    desc = np.frombuffer(raw[off:off + n * nb], np.uint8).reshape(n, nb); off += n * nb
    hk = np.frombuffer(raw[off:off + hn * 28], bb.KP_DTYPE); off += hn * 28
    bdesc = np.frombuffer(raw[off:off + bn * 48], np.uint8).reshape(bn, 48); off += bn * 48
    maxima = np.frombuffer(raw[off:off + nmax * 12], np.int32).reshape(nmax, 3)
    gk, gd = golden["ast0_kps"], golden["ast0_desc"]
    assert n == len(gk) and nb == 48 and self_matches == n and np.array_equal(desc, gd)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], gk[f]), f
    # (the fixture's Harris key points are the ones that survive the extractor's border cull: BriskFeature's output)
    assert bn == len(golden["harris0_kps"]) and np.array_equal(bdesc, golden["harris0_desc"])
    assert hn >= bn and kp_equal(hk, oracle.harris_detect(img, 0, 30.0, 20.0))
    assert nrad == sum(len(v) for v in oracle.radius_match(gd, [gd], 40.0))
    # HarrisScoreCalculator: SetImage / Score / Get2dMaxima against the oracle's score map and maxima
    sc = oracle.harris_scores(img)
    assert np.array_equal(maxima, oracle.harris_maxima(img, 20))
    assert s_int == int(sc[120, 100])
    ru, rv = 0.25, 0.5
    want = (1 - rv) * ((1 - ru) * float(sc[120, 100]) + ru * float(sc[120, 101])) + rv * ((1 - ru) * float(sc[121, 100]) + ru * float(sc[121, 101]))
    assert s_dbl == want


@pytest.mark.parametrize("w,h,radius", [(480, 480, 30.0), (333, 500, 30.0), (400, 401, 20.0), (18, 40, 10.0), (200, 320, 45.0), (131, 167, 7.5),
                                        (640, 640, 12.0)])
def test_harris_legacy_detector_bit_exact(ctx, oracle, w, h, radius):
    # brisk::HarrisFeatureDetector (legacy, single scale): key points in the reference's output order, bit for bit; the
    # oracle's restatement is pinned to the compiled reference in tests/test_oracle_golden.py
    det = bb.HarrisFeatureDetector(radius, ctx=ctx)
    det.set_corner_capacity(w * h // 4)
    frames = [bb.synthetic_frame(w, h, 7), np.random.default_rng(w).integers(0, 256, (h, w), dtype=np.uint8),
              (np.random.default_rng(h).integers(0, 3, (h, w)) * 90 + 20).astype(np.uint8)]   # plateaus: equal responses
    for img in frames:
        assert kp_equal(det.detect(img), oracle.harris_legacy(img, radius))
    kps, counts = det.detect_batch(np.stack(frames))
    for i, img in enumerate(frames):
        assert kp_equal(kps[i, :counts[i]], oracle.harris_legacy(img, radius))


def test_harris_legacy_detector_refuses_undefined_shapes(ctx):
    # landscape images: the reference's occupancy map is indexed with x as the row and accessed out of bounds
    with pytest.raises(bb.BriskError, match="occupancy"):
        bb.HarrisFeatureDetector(30.0, ctx=ctx).detect(bb.synthetic_frame(752, 480, 7))
    with pytest.raises(bb.BriskError):
        bb.HarrisFeatureDetector(30.0, ctx=ctx).detect(bb.synthetic_frame(17, 40, 7))


def test_harris_score_calculator(ctx, oracle, golden):
    # brisk::HarrisScoreCalculator's public methods (harris-score-calculator.h:52-90) through the Python mirror
    for img in (golden["image1"], bb.synthetic_frame(500, 333, 3), bb.synthetic_frame(131, 67, 5)):
        calc = bb.HarrisScoreCalculator(ctx=ctx)
        calc.SetImage(img)
        sc = oracle.harris_scores(img)
        assert np.array_equal(calc.scores(), sc)
        for thr in (0, 20, 5000):
            assert np.array_equal(calc.Get2dMaxima(thr), oracle.harris_maxima(img, thr))
        assert calc.Score(40, 30) == int(sc[30, 40]) and calc.Score(-1.0, 3.0) == 0.0 and calc.Score(float(img.shape[1] - 1), 3.0) == 0.0
        u, v = 40.75, 30.125
        want = (1 - 0.125) * ((1 - 0.75) * float(sc[30, 40]) + 0.75 * float(sc[30, 41])) + 0.125 * ((1 - 0.75) * float(sc[31, 40]) + 0.75 * float(sc[31, 41]))
        assert calc.Score(u, v) == want
    with pytest.raises(bb.BriskError):
        calc.Get2dMaxima(0, cap=3) if False else ctx._check(ctx._lib.brisk_harris_scores(ctx._h, None, 10, 10, 10, 0, None, None, 0, None))


def _tie_rich_descriptors(n, nbytes, seed):
    # few distinct byte values in few positions: distances fall in a narrow range, equal distances are the norm
    rng = np.random.default_rng(seed)
    d = np.zeros((n, nbytes), np.uint8)
    d[:, :6] = rng.integers(0, 4, (n, 6), dtype=np.uint8) * 17
    return d


@pytest.fixture
def matcher_oracle(oracle):
    """brisk::BruteForceMatcher itself (oracle/_ref, brute-force-matcher.cc compiled unmodified) when the prebuilt
    library travelled with the snapshot; else its restatement (pinned to it by tests/test_oracle_golden.py)."""
    from oracle import ref as r
    return r if r.available() else oracle


@pytest.mark.parametrize("nbytes", [48, 64])
def test_knn_match_collection_masks(ctx, matcher_oracle, nbytes):
    # knnMatch over a train collection (one image empty, one mask empty), ties across images, masked-out
    # queries, fewer allowed rows than k: against the compiled reference class (brute-force-matcher.cc:80-162)
    oracle = matcher_oracle
    rng = np.random.default_rng(11)
    q = bb.random_descriptors(90, nbytes, 5)
    trains = [bb.random_descriptors(300, nbytes, 6), np.zeros((0, nbytes), np.uint8), bb.random_descriptors(170, nbytes, 7)]
    trains[0][10] = q[3]; trains[2][5] = q[3]; trains[0][200] = q[3]
    masks = [(rng.random((90, 300)) < 0.7).astype(np.uint8), None, (rng.random((90, 170)) < 0.5).astype(np.uint8)]
    masks[0][7] = 0; masks[2][7] = 0                    # no allowed row at all
    masks[0][9] = 0; masks[2][9] = 0; masks[2][9, 4] = 1  # a single allowed row
    m = bb.BruteForceMatcher(ctx=ctx)
    m.add(trains)
    for k in (1, 2, 5):
        for compact in (False, True):
            assert m.knnMatch(q, None, k, masks, compact) == oracle.knn_match(q, trains, k, masks, compact), (k, compact)
        assert m.knnMatch(q, None, k) == oracle.knn_match(q, trains, k)
    # with the empty image (whose mask is empty too) query 7 is not "masked out": it gets k padding entries
    assert m.knnMatch(q, None, 2, masks)[7] == [(7, 0, 2, 2147483648.0)] * 2
    # every image has a mask: query 7 is masked out (dropped when compactResult), query 9 is padded the reference's way
    m2 = bb.BruteForceMatcher(ctx=ctx)
    m2.add([trains[0], trains[2]])
    full = [masks[0], masks[2]]
    got = m2.knnMatch(q, None, 3, full, True)
    assert got == oracle.knn_match(q, [trains[0], trains[2]], 3, full, True) and len(got) == 89
    assert got[8][1:] == [(9, 0, 1, 2147483648.0)] * 2
    assert m2.match(q, None, full) == [ms[0] for ms in oracle.knn_match(q, [trains[0], trains[2]], 1, full, True)]
    # fewer train rows than k
    assert m.knnMatch(q[:4], trains[0][:2], 3) == oracle.knn_match(q[:4], [trains[0][:2]], 3)


@pytest.mark.parametrize("nbytes", [48, 64])
def test_radius_match(ctx, matcher_oracle, nbytes):
    oracle = matcher_oracle
    # radiusMatch: long lists of equal distances (std::sort's permutation is part of the result), masks, two
    # images, compactResult, the capacity retry: reference brute-force-matcher.cc:164-214
    rng = np.random.default_rng(12)
    q = _tie_rich_descriptors(60, nbytes, 1)
    trains = [_tie_rich_descriptors(400, nbytes, 2), _tie_rich_descriptors(250, nbytes, 3)]
    m = bb.BruteForceMatcher(ctx=ctx)
    m.add(trains)
    for md in (0.0, 1.0, 7.5, 12.0, 1000.0):
        got = m.radiusMatch(q, None, md)
        assert got == oracle.radius_match(q, trains, md), md
    assert max(len(v) for v in m.radiusMatch(q, None, 12.0)) > 64
    masks = [(rng.random((60, 400)) < 0.6).astype(np.uint8), (rng.random((60, 250)) < 0.6).astype(np.uint8)]
    masks[0][5] = 0; masks[1][5] = 0
    for compact in (False, True):
        assert m.radiusMatch(q, None, 9.0, masks, compact) == oracle.radius_match(q, trains, 9.0, masks, compact)
    # random descriptors, one train image given per call
    q2, t2 = bb.random_descriptors(300, nbytes, 5), bb.random_descriptors(4000, nbytes, 6)
    md = float(nbytes * 8 * 0.44)
    assert m.radiusMatch(q2, t2, md) == oracle.radius_match(q2, [t2], md)
    # train-order output of the C ABI and its offsets
    off, idx, dist = m.radius(q2, t2, md, sort=False)
    ref = oracle.radius_match(q2, [t2], md)
    assert off[-1] == sum(len(v) for v in ref)
    for i in (0, 17, 299):
        assert sorted(idx[off[i]:off[i + 1]].tolist()) == idx[off[i]:off[i + 1]].tolist() == sorted(t for _, t, _, _ in ref[i])


@pytest.mark.parametrize("i", [0, 1])
def test_golden_harris_fixture(ctx, golden, i):
    # the reference's own ValidationHarris fixture (test-binary-equal.cc:73-89,301-313), bit for bit
    img = golden[f"image{i}"]
    feat = bb.BriskFeature(0, 30.0, 20.0, ctx=ctx)
    k, d = feat.detectAndCompute(img)
    gk, gd = golden[f"harris{i}_kps"], golden[f"harris{i}_desc"]
    assert len(k) == len(gk)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k[f], gk[f]), f
    assert np.abs(k["angle"] - gk["angle"]).max() <= 1e-4
    assert np.array_equal(d, gd)


@pytest.mark.parametrize("octaves,radius,abs_thr,max_kpt", [(0, 30.0, 20.0, None), (4, 30.0, 20.0, None), (2, 10.0, 0.0, 300),
                                                             (1, 45.0, 50.0, None), (3, 15.0, 5.0, None), (2, 2.5, 20.0, None),
                                                             (1, 1.0, 40.0, 500)])
def test_harris_detect_bit_exact(ctx, oracle, golden, octaves, radius, abs_thr, max_kpt):
    # multi-layer Harris incl. 3-D NMS and the introsort tie order (not covered by the golden fixture)
    det = bb.ScaleSpaceFeatureDetector(octaves, radius, abs_thr, max_kpt, ctx=ctx)
    det.set_corner_capacity(800 * 640)  # with abs_thr = 0 every plateau pixel is a 2-D maximum
    for img in (golden["image1"], bb.synthetic_frame(752, 480, 1000), bb.synthetic_frame(500, 333, 3)):
        want = oracle.harris_detect(img, octaves, radius, abs_thr, -1 if max_kpt is None else max_kpt)
        assert kp_equal(det.detect(img), want)


def test_harris_batch_config2(ctx, oracle):
    # BASELINE config 2 shape: 752x480 frames, Harris (4 octaves, r = 30, abs = 20) + BRISK2
    frames = bb.synthetic_batch(4, 752, 480, 1000)
    det = bb.ScaleSpaceFeatureDetector(4, 30.0, 20.0, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    kps, counts, desc = bb.detect_and_compute_batch(det, ext, frames, cap=8192)
    for f in range(len(frames)):
        k2, d2 = oracle.describe(frames[f], oracle.harris_detect(frames[f], 4, 30.0, 20.0))
        n = counts[f]
        assert n == len(k2)
        for fld in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(kps[f, :n][fld], k2[fld])
        assert np.array_equal(desc[f, :n], d2)


@pytest.mark.parametrize("octaves,radius,abs_thr,max_kpt", [(4, 0.0, 20.0, 400), (2, -1.0, 0.0, 64), (0, 0.0, 50.0, 100000), (3, 0.0, 5.0, 15)])
def test_harris_key_point_bucketing_bit_exact(ctx, oracle, golden, octaves, radius, abs_thr, max_kpt):
    # uniformityRadius <= 0: KeyPointBucketing, 4 x 4 buckets with maxNumKpt / 16 points each, per layer
    det = bb.ScaleSpaceFeatureDetector(octaves, radius, abs_thr, max_kpt, ctx=ctx)
    det.set_corner_capacity(800 * 640)
    for img in (golden["image0"], bb.synthetic_frame(752, 480, 1000), bb.synthetic_frame(322, 241, 4)):
        want = oracle.harris_detect(img, octaves, radius, abs_thr, max_kpt)
        got = det.detect(img)
        assert kp_equal(got, want)
    assert max_kpt < 16 or len(got) > 0


@pytest.mark.parametrize("radius,max_kpt", [(30.0, None), (8.0, None), (3.0, 150), (0.0, 400), (-1.0, 90), (15.0, 40)])
def test_harris_passed_keypoints(ctx, oracle, golden, radius, max_kpt):
    # detect() on a non-empty vector: the "use passed key points" mode (scale-space-feature-detector.h:103-108)
    from test_oracle_golden import _passed_points
    img = golden["image0"]
    det = bb.ScaleSpaceFeatureDetector(0, radius, 0.0, max_kpt, ctx=ctx)
    mk = -1 if max_kpt is None else max_kpt
    lists = [_passed_points(img.shape, 3000, 1), _passed_points(img.shape, 700, 2, True), _passed_points(img.shape, 20, 3),
             _passed_points(img.shape, 40000, 4)]
    if radius > 0:
        lists.append(bb.ScaleSpaceFeatureDetector(0, 1.0, 0.0, ctx=ctx).detect(img))
    for k in lists:
        assert kp_equal(det.detect(img, keypoints=k), oracle.harris_detect_passed(img.shape, k, radius, mk))
    low = lists[0].copy()
    low["response"] = 999999.0
    assert kp_equal(det.detect(img, keypoints=low), low)
    # a batch with lists of different lengths
    kin = np.zeros((3, 3000), bb.KP_DTYPE)
    for i in range(3):
        kin[i, :len(lists[i])] = lists[i]
    out, oc = det.detect_passed_batch(img.shape, kin, [len(lists[i]) for i in range(3)])
    for i in range(3):
        assert kp_equal(out[i, :oc[i]], oracle.harris_detect_passed(img.shape, lists[i], radius, mk))
    # BriskFeature::detectAndCompute(useProvidedKeypoints = true): re-filter, then describe (brisk-feature.h:80-93)
    k1, d1 = bb.BriskFeature(0, radius, 0.0, max_kpt, ctx=ctx).detectAndCompute(img, keypoints=lists[0])
    k2, d2 = oracle.describe(img, oracle.harris_detect_passed(img.shape, lists[0], radius, mk))
    assert len(k1) == len(k2) > 0 and np.array_equal(k1["x"], k2["x"]) and np.array_equal(k1["y"], k2["y"]) and np.array_equal(d1, d2)
    # refused: more than one layer, a point outside the image
    with pytest.raises(bb.BriskError):
        bb.ScaleSpaceFeatureDetector(1, 30.0, 0.0, ctx=ctx).detect(img, keypoints=lists[0])
    bad = lists[2].copy()
    bad["x"][0] = img.shape[1] + 5.0
    bad["response"][0] = 5.0e6
    with pytest.raises(bb.BriskError):
        det.detect(img, keypoints=bad)


@pytest.mark.parametrize("octaves,radius,abs_thr,max_kpt", [(0, 0.5, 0.0, None), (2, 0.7, 20.0, None), (1, 0.25, 0.0, 500)])
def test_harris_small_uniformity_radius(ctx, oracle, octaves, radius, abs_thr, max_kpt):
    # radii below 1: occupancy maps of ceil(15 / radius)^2 bytes per pixel (here up to 3600)
    img = bb.synthetic_frame(320, 240, 8)
    want = oracle.harris_detect(img, octaves, radius, abs_thr, -1 if max_kpt is None else max_kpt)
    assert len(want) > 100 and kp_equal(bb.ScaleSpaceFeatureDetector(octaves, radius, abs_thr, max_kpt, ctx=ctx).detect(img), want)


@pytest.mark.parametrize("kind", ["ast", "harris"])
def test_make_set_tool_reproduces_reference_fixtures(tmp_path, golden, kind):
    # tools/make_set.py writes what the reference's test-binary-equal.cc serialises; against the reference's own
    # fixtures (rebuilt from the committed goldens in the same format) everything but `angle` must be identical
    import subprocess
    import sys
    from conftest import ROOT
    pgms = []
    for i in (0, 1):
        pgms.append(str(tmp_path / f"img{i + 1}.pgm"))
        bb.write_pgm(pgms[-1], golden[f"image{i}"])
    bb.write_set(tmp_path / "golden.set", [dict(path=pgms[i], image=golden[f"image{i}"], keypoints=golden[f"{kind}{i}_kps"],
                                                descriptors=golden[f"{kind}{i}_desc"], blobs={"testImage": golden[f"image{i}"].tobytes()})
                                           for i in (0, 1)])
    tool = str(ROOT / "tools" / "make_set.py")
    subprocess.run([sys.executable, tool, "--detector", kind, "-o", str(tmp_path / "ours.set"), *pgms], check=True)
    r = subprocess.run([sys.executable, tool, "--compare", str(tmp_path / "ours.set"), str(tmp_path / "golden.set")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    ours = bb.read_set(tmp_path / "ours.set")
    assert all(np.abs(ours[i]["keypoints"]["angle"] - golden[f"{kind}{i}_kps"]["angle"]).max() <= 1e-4 for i in (0, 1))


def test_harris_unsupported(ctx):
    # bucketing with the default maxNumKpt = SIZE_MAX: the reference throws std::length_error (reserve)
    with pytest.raises(bb.BriskError):
        bb.ScaleSpaceFeatureDetector(4, 0.0, 20.0, ctx=ctx).detect(bb.synthetic_frame(320, 240, 1))
    # an occupancy map beyond 2 GiB per layer is refused
    with pytest.raises(bb.BriskError):
        bb.ScaleSpaceFeatureDetector(0, 0.2, 20.0, ctx=ctx).detect(bb.synthetic_frame(1920, 1080, 1))




def test_matcher_any_width_any_k(ctx, matcher_oracle):
    # rows the extractors do not produce (multiples of 16 bytes here: the reference reads rows with aligned 128-bit loads) and
    # k > 8: the general kernel, against the reference's matcher class -- with and without masks, and radiusMatch
    rng = np.random.default_rng(5)
    m = bb.BruteForceMatcher(ctx=ctx)
    for nb, nq, nt, k in ((32, 70, 300, 3), (80, 150, 900, 11), (16, 33, 64, 2), (64, 90, 400, 20), (48, 40, 13, 16), (256, 20, 50, 9), (496, 10, 40, 4)):
        q = rng.integers(0, 256, (nq, nb), dtype=np.uint8)
        t = rng.integers(0, 4, (nt, nb), dtype=np.uint8) if nb == 16 else rng.integers(0, 256, (nt, nb), dtype=np.uint8)   # (many ties)
        got = m.knnMatch(q, [t], k)
        want = matcher_oracle.knn_match(q, [t], k)
        assert got == want, (nb, k)
        mask = (rng.integers(0, 3, (nq, nt)) > 0).astype(np.uint8)
        mask[::7] = 0
        assert m.knnMatch(q, [t], k, masks=[mask]) == matcher_oracle.knn_match(q, [t], k, masks=[mask]), (nb, k, "masked")
        radius = float(np.median(m.knn(q, t, 1)[1])) + 8.0
        assert m.radiusMatch(q, [t], radius) == matcher_oracle.radius_match(q, [t], radius), (nb, "radius")


@pytest.mark.parametrize("nbytes", [48, 64])
def test_knn_tensor_core_variants_match_popc(oracle, nbytes):
    # the tcgen05 kernels (variant 3, the default: kind::mxf4 MMAs on +-1.0 E2M1 values for 64-byte rows; variant 2: kind::i8
    # MMAs on +-1 bytes; TMEM accumulators, TMA operands) and the mma.sync IMMA kernel (variant 1) must return exactly what
    # the POPC kernel returns: ties, missing rows, ragged
    # tiles, one or several train splits, queries / train rows that are no multiple of the tile sizes
    ctx2 = bb.Context(0)
    m = bb.BruteForceMatcher(ctx=ctx2)
    for nq, nt in ((700, 5000), (129, 257), (5, 1), (300, 50000), (1000, 127), (257, 128), (256, 129), (3000, 70001)):
        q = bb.random_descriptors(nq, nbytes, 5)
        t = bb.random_descriptors(nt, nbytes, 6)
        if nt > 200:
            t[100] = q[3]
            t[200] = q[3]
            t[nt - 1] = q[3]
            t[150] = q[4]; t[150, 0] ^= 1   # distance 1
        res = []
        for variant in (0, 1, 2, 3):
            ctx2.set_knn_variant(variant)
            res.append(m.knn(q, t, 2))
        for variant in (1, 2, 3):
            assert np.array_equal(res[0][0], res[variant][0]) and np.array_equal(res[0][1], res[variant][1]), (nq, nt, variant)
    i2, d2 = oracle.knn(q, t, 2)
    assert np.array_equal(res[2][0], i2) and np.array_equal(res[2][1], d2)
    # tie-rich rows (few distinct distances): the index tie rule across tiles and splits
    q = _tie_rich_descriptors(300, nbytes, 1)
    t = _tie_rich_descriptors(40000, nbytes, 2)
    i2, d2 = oracle.knn(q, t, 2)
    for variant in (2, 3):
        ctx2.set_knn_variant(variant)
        i1, d1 = m.knn(q, t, 2)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2), variant


def test_knn_tcgen05_alternative_forms():
    # the measured alternatives of the tcgen05 kernels -- queries held in tensor memory (BRISK_B200_TC5_MODE=ts), 256-row
    # train tiles (BRISK_B200_TC5_TILE_ROWS=256), and the FP4 kernel's 96-row and three-query-tile schedules -- are chosen by environment variables read once per process: run each in
    # a child process and compare with the POPC kernel
    import os
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import numpy as np, ethzasl_brisk_b200 as bb\n"
        "ctx = bb.Context(0); m = bb.BruteForceMatcher(ctx=ctx)\n"
        "for nb in (48, 64):\n"
        "    for nq, nt in ((700, 5000), (129, 257), (5, 1), (3000, 70001)):\n"
        "        q = bb.random_descriptors(nq, nb, 5); t = bb.random_descriptors(nt, nb, 6)\n"
        "        if nt > 200: t[100] = q[3]; t[200] = q[3]; t[nt - 1] = q[3]\n"
        "        ctx.set_knn_variant(0); a = m.knn(q, t, 2)\n"
        "        for v in (2, 3):\n"
        "            ctx.set_knn_variant(v); b = m.knn(q, t, 2)\n"
        "            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (nb, nq, nt, v)\n"
        "print('same')\n")
    # (the last two select other schedules of the FP4 kernel, variant 3)
    for extra in ({"BRISK_B200_TC5_MODE": "ts"}, {"BRISK_B200_TC5_TILE_ROWS": "256"}, {"BRISK_B200_TC5MX_TILE_ROWS": "96"},
                  {"BRISK_B200_TC5MX_QTILES": "3"}, {"BRISK_B200_TC5MX_EPI_WARPS": "8"}):
        env = dict(os.environ, PYTHONPATH=str(ROOT), **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "same" in r.stdout, (extra, r.stdout[-500:], r.stderr[-1500:])
