"""Python host mirror of the reference's public classes over the C ABI.

Class names, constructor arguments and semantics follow the reference headers
(brisk/include/brisk/brisk-feature-detector.h:51-84,
brisk-descriptor-extractor.h:54-202, brute-force-matcher.h:54-94,
internal/hamming.h:56-114); each method calls straight into
libbrisk_b200.so (include/brisk_b200.h).  Images are numpy u8 arrays (host) or
torch CUDA tensors (device, passed by pointer); key points are numpy structured
arrays binary-compatible with cv::KeyPoint.

There is no CPU path: if the CUDA library is missing or no GPU is usable,
construction fails.
"""
import ctypes as C
from pathlib import Path

import numpy as np

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"),
                     ("response", "f4"), ("octave", "i4"), ("class_id", "i4")])
assert KP_DTYPE.itemsize == 28

STAGES = ("h2d", "pyramid", "detect", "lists", "nms", "integral", "describe", "d2h", "knn")
_LIB_PATH = Path(__file__).resolve().parent / "libbrisk_b200.so"
_lib = None

BRISK_ERR_CAPACITY = -4


class BriskError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"brisk_b200 error {code}: {msg}")
        self.code = code


def lib_path():
    return _LIB_PATH


def load_library():
    """dlopen libbrisk_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} not found: build it with `python -m ethzasl_brisk_b200.build` "
                               "(nvcc, sm_100a); there is no CPU fallback")
        lib = C.CDLL(str(_LIB_PATH))
        lib.brisk_last_error.restype = C.c_char_p
        lib.brisk_last_error.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a):
    """(address, keep-alive) of a numpy array or torch tensor, or NULL."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):  # torch tensor (host or CUDA)
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _is_torch(a):
    return hasattr(a, "data_ptr")


class Context:
    """One CUDA device + stream + workspaces (brisk_ctx). Not thread-safe."""

    def __init__(self, device=0, stream=None, workspace_limit=None, timing=False):
        self._lib = load_library()
        self._h = C.c_void_p()
        # stream=None: the context creates a private non-blocking stream.  Any other value is the caller's stream handle
        # (e.g. torch.cuda.current_stream().cuda_stream); handle 0, the default stream, is passed as cudaStreamLegacy (0x1),
        # because a NULL handle means "create one" in brisk_ctx_create.  Inputs must be complete on that stream.
        sp = None if stream is None else C.c_void_p(int(stream) or 1)
        rc = self._lib.brisk_ctx_create(int(device), sp, C.byref(self._h))
        if rc != 0:
            raise BriskError(rc, "brisk_ctx_create failed (no usable CUDA device?)")
        self.device = int(device)
        if workspace_limit:
            self._check(self._lib.brisk_ctx_set_workspace_limit(self._h, C.c_size_t(int(workspace_limit))))
        if timing:
            self.enable_timing(True)

    def _check(self, rc, allow_capacity=False):
        if rc != 0 and not (allow_capacity and rc == BRISK_ERR_CAPACITY):
            raise BriskError(rc, (self._lib.brisk_last_error(self._h) or b"").decode())
        return rc

    def enable_timing(self, on=True):
        self._check(self._lib.brisk_ctx_enable_timing(self._h, int(bool(on))))

    def set_knn_variant(self, variant):
        """0: POPC kernel; for k == 2 and 48/64-byte rows 1: mma.sync IMMA kernel, 2: tcgen05 kind::i8 kernel, 3 (default): tcgen05
        kind::mxf4 (FP4, +-1.0 operands)."""
        self._check(self._lib.brisk_ctx_set_knn_variant(self._h, int(variant)))

    def set_pipelining(self, on=True):
        self._check(self._lib.brisk_ctx_set_pipelining(self._h, int(bool(on))))

    def last_timing(self):
        """-> (dict stage -> device ms of the last call, kernel launches of the last call)"""
        ms = (C.c_float * len(STAGES))()
        n = C.c_int64(0)
        self._check(self._lib.brisk_ctx_last_timing(self._h, ms, C.byref(n)))
        return {s: float(ms[i]) for i, s in enumerate(STAGES)}, int(n.value)

    def last_raw_corners(self):
        """Raw AGAST corners (before NMS) summed over the frames of the last detect call (timing mode)."""
        n = C.c_int64(0)
        self._check(self._lib.brisk_ctx_last_raw_corners(self._h, C.byref(n)))
        return int(n.value)

    def sync(self):
        self._check(self._lib.brisk_sync(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.brisk_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sample16(self, fn, img, dw, dh):
        a = np.ascontiguousarray(img, np.uint16)
        h, w = a.shape
        out = np.zeros((dh(h), dw(w)), np.uint16)
        self._check(fn(self._h, _ptr(a), w, h, C.c_size_t(2 * w), _ptr(out), C.c_size_t(2 * out.shape[1])))
        return out

    def halfsample16(self, img):
        """brisk::Halfsample16 on a CV_16UC1 image -> (h // 2, w // 2) u16."""
        return self._sample16(self._lib.brisk_halfsample16, img, lambda w: w // 2, lambda h: h // 2)

    def twothirdsample16(self, img):
        """brisk::Twothirdsample16 on a CV_16UC1 image -> (2 * (h // 3), 2 * (w // 3)) u16."""
        return self._sample16(self._lib.brisk_twothirdsample16, img, lambda w: 2 * (w // 3), lambda h: 2 * (h // 3))

    # --- stage dumps used by the parity tests ---
    def debug_pyramid(self, img, octaves):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        dims = np.zeros((12, 2), np.int32)
        buf = np.zeros(3 * w * h + 64, np.uint8)
        nl = C.c_int(0)
        self._check(self._lib.brisk_debug_pyramid(self._h, int(octaves), _ptr(img), w, h, C.c_size_t(w), _ptr(buf), _ptr(dims), C.byref(nl)))
        out, off = [], 0
        for i in range(nl.value):
            cw, ch = int(dims[i, 0]), int(dims[i, 1])
            out.append(buf[off:off + cw * ch].reshape(ch, cw).copy())
            off += cw * ch
        return out

    def debug_scores(self, img):
        """Dense FAST 9-16 (the NMS kernels' packed row evaluator) and AGAST 5-8 score planes of one image, threshold 1."""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        a, b = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
        self._check(self._lib.brisk_debug_scores(self._h, _ptr(img), w, h, C.c_size_t(w), _ptr(a), _ptr(b)))
        return a, b

    def debug_integral(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        out = np.zeros((h + 1, w + 1), np.int32)
        self._check(self._lib.brisk_debug_integral(self._h, _ptr(img), w, h, C.c_size_t(w), _ptr(out)))
        return out


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _frames(images):
    """-> (array-like, n, h, w, stride, frame_pitch) for [h,w] / [n,h,w] u8 input."""
    if _is_torch(images):
        t = images
        if t.dim() == 2:
            t = t.unsqueeze(0)
        assert t.dim() == 3 and t.element_size() == 1 and t.stride(2) == 1
        n, h, w = t.shape
        return t, n, h, w, t.stride(1), (t.stride(0) if n > 1 else t.stride(1) * h)
    a = np.asarray(images)
    if a.dtype != np.uint8:
        raise TypeError("images must be 8-bit grayscale (CV_8UC1)")
    if a.ndim == 2:
        a = a[None]
    a = np.ascontiguousarray(a)
    n, h, w = a.shape
    return a, n, h, w, w, w * h


def _mask_frames(masks, frames):
    """Masks share the images' geometry in the C ABI (same n, h, w, row stride and frame pitch): validated here."""
    if masks is None:
        return None
    m, n, h, w, stride, fp = _frames(masks)
    _, n2, h2, w2, stride2, fp2 = frames
    if (n, h, w) != (n2, h2, w2):
        raise ValueError(f"mask shape {(n, h, w)} differs from the image shape {(n2, h2, w2)}")
    if (stride, fp) != (stride2, fp2):
        raise ValueError("mask strides differ from the image strides: pass a contiguous mask (or one laid out like the images)")
    return m


class BriskFeatureDetector:
    """brisk::BriskFeatureDetector(thresh, octaves=3, suppressScaleNonmaxima=true)."""

    def __init__(self, thresh, octaves=3, suppressScaleNonmaxima=True, ctx=None):
        self.ctx = ctx or default_context()
        self.threshold, self.octaves = int(thresh), int(octaves)
        self._h = C.c_void_p()
        self.ctx._check(self.ctx._lib.brisk_agast_detector_create(self.ctx._h, self.threshold, self.octaves,
                                                                   int(bool(suppressScaleNonmaxima)), C.byref(self._h)))

    def set_corner_capacity(self, n):
        self.ctx._check(self.ctx._lib.brisk_detector_set_corner_capacity(self._h, int(n)))

    def detect_batch(self, images, masks=None, cap=16384, out=None):
        """images [n,h,w] u8 -> (kps [n,cap] structured, counts [n])."""
        a, n, h, w, stride, fp = _frames(images)
        m = _mask_frames(masks, (a, n, h, w, stride, fp))
        if out is None:
            kps = np.zeros((n, cap), KP_DTYPE)
            counts = np.zeros(n, np.int32)
        else:
            kps, counts = out
        self.ctx._check(self.ctx._lib.brisk_detect(self.ctx._h, self._h, _ptr(a), n, w, h, C.c_size_t(stride), C.c_size_t(fp),
                                                   _ptr(m), _ptr(kps), _ptr(counts), int(cap)))
        return kps, counts

    def detect(self, image, mask=None, cap=65536):
        """cv::Feature2D::detect for one image -> key points (structured array)."""
        kps, counts = self.detect_batch(image, None if mask is None else mask, cap)
        return kps[0, :counts[0]].copy()

    def compute_scale_batch(self, images, kps, counts, cap=65536, _allow_capacity=False):
        """BriskFeatureDetector::ComputeScale for a batch (reference brisk-feature-detector.cc:87-92):
        images [n,h,w] u8, kps [n,cap_in] structured, counts [n] -> (kps [n,cap], counts [n])."""
        a, n, h, w, stride, fp = _frames(images)
        k = np.ascontiguousarray(kps, KP_DTYPE).reshape(n, -1)
        c = np.ascontiguousarray(counts, np.int32).reshape(n)
        out = np.zeros((n, cap), KP_DTYPE)
        oc = np.zeros(n, np.int32)
        self._last_rc = self.ctx._check(self.ctx._lib.brisk_compute_scale(self.ctx._h, self._h, _ptr(a), n, w, h, C.c_size_t(stride),
                                                          C.c_size_t(fp), _ptr(k), _ptr(c), k.shape[1], _ptr(out), _ptr(oc),
                                                          int(cap)), allow_capacity=_allow_capacity)
        return out, oc

    def compute_scale(self, image, keypoints, cap=None):
        """ComputeScale(image, keypoints) for one image: the provided key points (x, y, class_id are read)
        are tested in every pyramid layer -> the key points the reference leaves in `keypoints`."""
        k = np.ascontiguousarray(keypoints, KP_DTYPE).reshape(1, -1)
        n_layers = max(1, 2 * self.octaves)
        out, oc = self.compute_scale_batch(image, k, [k.shape[1]], cap or max(1, n_layers * k.shape[1]), _allow_capacity=cap is None)
        if self._last_rc != 0:  # layers without points contributed their corners: retry with room for them
            out, oc = self.compute_scale_batch(image, k, [k.shape[1]], max(int(oc[0]), out.shape[1]))
        return out[0, :oc[0]].copy()

    def debug_corners(self, image, cap=1 << 20):
        img = np.ascontiguousarray(image, np.uint8)
        h, w = img.shape
        c = np.zeros((cap, 3), np.int32)
        lc = np.zeros(12, np.int32)
        n = self.ctx._lib.brisk_debug_corners(self.ctx._h, self._h, _ptr(img), w, h, C.c_size_t(w), _ptr(c), cap, _ptr(lc))
        if n < 0:
            self.ctx._check(n)
        return c[:n].copy(), lc

    def __del__(self):
        try:
            if self._h.value:
                self.ctx._lib.brisk_detector_destroy(self._h)
        except Exception:
            pass


class ScaleSpaceFeatureDetector(BriskFeatureDetector):
    """brisk::ScaleSpaceFeatureDetector<brisk::HarrisScoreCalculator>(octaves, uniformityRadius,
    absoluteThreshold=0, maxNumKpt=SIZE_MAX) -- reference scale-space-feature-detector.h:62-135.
    Masks are ignored, as in the reference's detectImpl."""

    def __init__(self, octaves, uniformityRadius, absoluteThreshold=0.0, maxNumKpt=None, ctx=None):
        self.ctx = ctx or default_context()
        self.octaves = int(octaves)
        self._h = C.c_void_p()
        mk = -1 if maxNumKpt is None else int(maxNumKpt)
        self.ctx._check(self.ctx._lib.brisk_harris_detector_create(self.ctx._h, self.octaves, C.c_double(uniformityRadius),
                                                                    C.c_double(absoluteThreshold), C.c_int64(mk), C.byref(self._h)))


    def detect(self, image, mask=None, cap=65536, keypoints=None):
        """detect(image, keypoints): a non-empty `keypoints` switches to the reference's "use passed key points" mode
        (scale-space-feature-detector.h:103-108)."""
        if keypoints is not None and len(keypoints):
            a = image if _is_torch(image) else np.asarray(image)
            return self.detect_passed(tuple(a.shape[-2:]), keypoints)
        return super().detect(image, mask, cap)

    def detect_passed_batch(self, shape, kps, counts, cap=None):
        """shape = (h, w); kps [n,cap_in] structured, counts [n] -> (kps [n,cap], counts [n])."""
        h, w = (int(v) for v in shape)
        c = np.ascontiguousarray(counts, np.int32).reshape(-1)
        n = len(c)
        k = np.ascontiguousarray(kps, KP_DTYPE).reshape(n, -1)
        cap = int(cap or k.shape[1])
        out = np.zeros((n, cap), KP_DTYPE)
        oc = np.zeros(n, np.int32)
        self.ctx._check(self.ctx._lib.brisk_harris_detect_passed(self.ctx._h, self._h, n, w, h, _ptr(k), _ptr(c), k.shape[1],
                                                                 _ptr(out), _ptr(oc), cap))
        return out, oc

    def detect_passed(self, shape, keypoints):
        k = np.ascontiguousarray(keypoints, KP_DTYPE).reshape(1, -1)
        out, oc = self.detect_passed_batch(shape, k, [k.shape[1]])
        return out[0, :oc[0]].copy()


HarrisScaleSpaceFeatureDetector = ScaleSpaceFeatureDetector


class BriskFeature:
    """brisk::BriskFeature(octaves, uniformityRadius, absoluteThreshold=0, maxNumKpt, rotationInvariant=true,
    scaleInvariant=true, extractorVersion=briskV2): Harris scale-space detector + extractor in one
    object -- reference brisk/include/brisk/brisk-feature.h:54-114."""

    def __init__(self, octaves, uniformityRadius, absoluteThreshold=0.0, maxNumKpt=None, rotationInvariant=True,
                 scaleInvariant=True, extractorVersion=2, ctx=None):
        self.ctx = ctx or default_context()
        self.detector = ScaleSpaceFeatureDetector(octaves, uniformityRadius, absoluteThreshold, maxNumKpt, ctx=self.ctx)
        self.extractor = BriskDescriptorExtractor(rotationInvariant, scaleInvariant, extractorVersion, ctx=self.ctx)

    def detectAndCompute(self, image, mask=None, cap=65536, keypoints=None):
        """keypoints (useProvidedKeypoints = true): the detector re-filters them instead of detecting
        (brisk-feature.h:80-93), then the extractor runs."""
        if keypoints is not None and len(keypoints):
            return self.extractor.compute(image, self.detector.detect(image, keypoints=keypoints))
        kps, counts, desc = detect_and_compute_batch(self.detector, self.extractor, image, cap=cap)
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()


class BriskDescriptorExtractor:
    """brisk::BriskDescriptorExtractor(rotationInvariant, scaleInvariant, version, patternScale)."""

    briskV1, briskV2 = 1, 2
    kDescriptorLength = 384

    def __init__(self, rotationInvariant=True, scaleInvariant=True, version=2, patternScale=1.0, fname=None, ctx=None):
        self.ctx = ctx or default_context()
        self.rotationInvariance, self.scaleInvariance = bool(rotationInvariant), bool(scaleInvariant)
        self._h = C.c_void_p()
        f = fname.encode() if fname else None
        self.ctx._check(self.ctx._lib.brisk_extractor_create(self.ctx._h, int(self.rotationInvariance), int(self.scaleInvariance),
                                                              int(version), C.c_float(patternScale), f, C.byref(self._h)))

    def descriptorSize(self):
        return int(self.ctx._lib.brisk_extractor_descriptor_size(self._h))

    def descriptorType(self):
        return 0  # CV_8U

    def pattern(self):
        counts = np.zeros(4, np.int32)
        L = self.ctx._lib
        self.ctx._check(L.brisk_extractor_pattern(self._h, _ptr(counts), None, None, None, None, None))
        P, ns, nl, nb = (int(v) for v in counts)
        pts = np.zeros((64, 1024, P, 3), np.float32)
        scale_list = np.zeros(64, np.float32)
        size_list = np.zeros(64, np.uint32)
        sp = np.zeros((ns, 2), np.uint32)
        lp = np.zeros((nl, 4), np.int32)
        self.ctx._check(L.brisk_extractor_pattern(self._h, _ptr(counts), _ptr(pts), _ptr(scale_list), _ptr(size_list), _ptr(sp), _ptr(lp)))
        return dict(points=P, strings=nb, pts=pts, scale_list=scale_list, size_list=size_list, short_pairs=sp, long_pairs=lp)

    def compute_batch(self, images, kps, counts, cap=None):
        """kps [n,cap] / counts [n] in -> (kps, counts, desc [n,cap,descriptorSize]); inputs are not modified."""
        a, n, h, w, stride, fp = _frames(images)
        kps = np.ascontiguousarray(kps, KP_DTYPE).copy()
        counts = np.ascontiguousarray(counts, np.int32).copy()
        cap = kps.shape[1]
        desc = np.zeros((n, cap, self.descriptorSize()), np.uint8)
        self.ctx._check(self.ctx._lib.brisk_describe(self.ctx._h, self._h, _ptr(a), n, w, h, C.c_size_t(stride), C.c_size_t(fp),
                                                     _ptr(kps), _ptr(counts), int(cap), _ptr(desc)))
        return kps, counts, desc

    def compute(self, image, keypoints):
        """cv::Feature2D::compute for one image -> (surviving key points with angle, descriptors [m, size])."""
        k = np.ascontiguousarray(keypoints, KP_DTYPE)
        cap = max(len(k), 1)
        kk = np.zeros((1, cap), KP_DTYPE)
        kk[0, :len(k)] = k
        kps, counts, desc = self.compute_batch(image, kk, np.array([len(k)], np.int32))
        m = int(counts[0])
        return kps[0, :m].copy(), desc[0, :m].copy()

    def __del__(self):
        try:
            if self._h.value:
                self.ctx._lib.brisk_extractor_destroy(self._h)
        except Exception:
            pass


class HarrisFeatureDetector:
    """brisk::HarrisFeatureDetector(radius): the legacy single-scale Harris detector (harris-feature-detector.h:51-82)."""

    def __init__(self, radius, ctx=None):
        self.ctx = ctx or default_context()
        self._h = C.c_void_p()
        self.ctx._check(self.ctx._lib.brisk_harris_legacy_detector_create(self.ctx._h, C.c_double(radius), C.byref(self._h)))

    def set_corner_capacity(self, n):
        self.ctx._check(self.ctx._lib.brisk_detector_set_corner_capacity(self._h, int(n)))

    def detect_batch(self, images, cap=16384):
        a, n, h, w, stride, fp = _frames(images)
        kps = np.zeros((n, cap), KP_DTYPE)
        counts = np.zeros(n, np.int32)
        self.ctx._check(self.ctx._lib.brisk_detect(self.ctx._h, self._h, _ptr(a), n, w, h, C.c_size_t(stride), C.c_size_t(fp), None,
                                                   _ptr(kps), _ptr(counts), int(cap)))
        return kps, counts

    def detect(self, image, mask=None, cap=65536):
        kps, counts = self.detect_batch(image, cap=cap)
        return kps[0, :counts[0]].copy()

    def __del__(self):
        try:
            if self._h.value:
                self.ctx._lib.brisk_detector_destroy(self._h)
        except Exception:
            pass


class HarrisScoreCalculator:
    """brisk::HarrisScoreCalculator (harris-score-calculator.h:52-90): SetImage computes the integer Harris score map on the
    GPU; Score(int, int) / Score(double, double) read it the way the reference's inline accessors do; Get2dMaxima lists the
    8-neighbour maxima in raster order as (score, x, y)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()
        self._img = None
        self._scores = None

    def SetImage(self, img, initScores=True):
        self._img = np.ascontiguousarray(img, np.uint8)
        if initScores:
            h, w = self._img.shape
            self._scores = np.zeros((h, w), np.int32)
            self.ctx._check(self.ctx._lib.brisk_harris_scores(self.ctx._h, _ptr(self._img), w, h, C.c_size_t(w), 0, _ptr(self._scores), None, 0, None))

    def scores(self):
        return self._scores

    def Score(self, u, v):
        s = self._scores
        if isinstance(u, (int, np.integer)) and isinstance(v, (int, np.integer)):
            return int(s[v, u])
        ui, vi = int(u), int(v)
        if ui + 1 >= s.shape[1] or vi + 1 >= s.shape[0] or ui < 0 or vi < 0:
            return 0.0
        ru, rv = float(u) - float(ui), float(v) - float(vi)
        return (1.0 - rv) * ((1.0 - ru) * float(s[vi, ui]) + ru * float(s[vi, ui + 1])) + rv * ((1.0 - ru) * float(s[vi + 1, ui]) + ru * float(s[vi + 1, ui + 1]))

    def Get2dMaxima(self, absoluteThreshold=0, cap=1 << 16):
        h, w = self._img.shape
        while True:
            out = np.zeros((cap, 3), np.int32)
            n = C.c_int32(0)
            rc = self.ctx._lib.brisk_harris_scores(self.ctx._h, _ptr(self._img), w, h, C.c_size_t(w), int(absoluteThreshold), None, _ptr(out), cap, C.byref(n))
            if rc == BRISK_ERR_CAPACITY and n.value > cap:
                cap = n.value
                continue
            self.ctx._check(rc)
            return out[:n.value].copy()


def detect_and_compute_batch(detector, extractor, images, masks=None, cap=16384, out=None, allow_truncation=False, async_=False):
    """detect() + compute() without leaving the device.

    -> (kps [n,cap], counts [n], desc [n,cap,descriptorSize]).  `out` may hold
    preallocated (kps, counts, desc) numpy arrays or torch CUDA tensors.
    async_=True (brisk_detect_describe_async): the batch's last chunk stays in flight; `images` and `out` must stay
    alive and untouched until the call after next of the same shape has returned or ctx.sync() was called.
    """
    ctx = detector.ctx
    assert extractor.ctx is ctx
    a, n, h, w, stride, fp = _frames(images)
    m = _mask_frames(masks, (a, n, h, w, stride, fp))
    if out is None:
        kps = np.zeros((n, cap), KP_DTYPE)
        counts = np.zeros(n, np.int32)
        desc = np.zeros((n, cap, extractor.descriptorSize()), np.uint8)
    else:
        kps, counts, desc = out
    fn = ctx._lib.brisk_detect_describe_async if async_ else ctx._lib.brisk_detect_describe
    rc = fn(ctx._h, detector._h, extractor._h, _ptr(a), n, w, h, C.c_size_t(stride), C.c_size_t(fp),
            _ptr(m), _ptr(kps), _ptr(counts), int(cap), _ptr(desc))
    ctx._check(rc, allow_capacity=allow_truncation)
    return kps, counts, desc


class Hamming:
    """brisk::Hamming functor: popcount(a xor b)."""

    def __init__(self, ctx=None):
        self.ctx = ctx or default_context()

    def __call__(self, a, b, size=None):
        a = np.ascontiguousarray(a, np.uint8).reshape(1, -1)
        b = np.ascontiguousarray(b, np.uint8).reshape(1, -1)
        nb = a.shape[1] if size is None else int(size)
        d = np.zeros(1, np.int32)
        self.ctx._check(self.ctx._lib.brisk_hamming_distance(self.ctx._h, _ptr(a), _ptr(b), C.c_int64(1), nb, _ptr(d)))
        return int(d[0])

    def pairs(self, a, b):
        """Row-wise distances of two [n, bytes] arrays (counted over bytes // 16 whole 128-bit words, as the reference does)."""
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        assert a.ndim == 2 and a.shape == b.shape
        d = np.zeros(len(a), np.int32)
        self.ctx._check(self.ctx._lib.brisk_hamming_distance(self.ctx._h, _ptr(a), _ptr(b), C.c_int64(len(a)), a.shape[1], _ptr(d)))
        return d

    @staticmethod
    def PopcntofXORed(signature1, signature2, numberOf128BitWords, ctx=None):
        """Static form of the reference (hamming.h:79-91)."""
        return Hamming(ctx)(signature1, signature2, 16 * int(numberOf128BitWords))


class BruteForceMatcher:
    """brisk::BruteForceMatcher (reference brisk/include/brisk/brute-force-matcher.h:54-94): brute-force Hamming
    kNN / radius matching against a train collection, with the cv::DescriptorMatcher conventions the
    reference implements in brisk/src/brute-force-matcher.cc:80-214 (masks, several train images,
    compactResult, result order)."""

    def __init__(self, distance=None, ctx=None):
        self.ctx = ctx or (distance.ctx if distance is not None else default_context())
        self._train = []

    def isMaskSupported(self):
        return True  # brute-force-matcher.h:60-62

    def add(self, descriptors):
        """cv::DescriptorMatcher::add: one descriptor matrix, or a list of them (one per train image)."""
        if isinstance(descriptors, (list, tuple)):
            self._train.extend(np.ascontiguousarray(d, np.uint8) for d in descriptors)
        else:
            self._train.append(np.ascontiguousarray(descriptors, np.uint8))

    def clear(self):
        self._train = []

    def getTrainDescriptors(self):
        return self._train

    def empty(self):
        return not self._train

    def clone(self, emptyTrainData=False):
        m = BruteForceMatcher(ctx=self.ctx)
        if not emptyTrainData:
            m._train = [t.copy() for t in self._train]
        return m

    def knn(self, query, train, k, mask=None):
        """-> (idx [nq,k] int32, dist [nq,k] int32); query/train: numpy u8 [n, bytes] or torch CUDA tensors;
        mask: optional [nq, nt] u8, 0 = pair excluded.  Missing neighbours are -1."""
        nq, nb = query.shape
        nt = train.shape[0]
        if _is_torch(query):
            import torch
            idx = torch.empty((nq, k), dtype=torch.int32, device=query.device)
            dist = torch.empty((nq, k), dtype=torch.int32, device=query.device)
        else:
            query = np.ascontiguousarray(query, np.uint8)
            train = np.ascontiguousarray(train, np.uint8)
            idx = np.zeros((nq, k), np.int32)
            dist = np.zeros((nq, k), np.int32)
        if mask is None:
            rc = self.ctx._lib.brisk_hamming_knn(self.ctx._h, _ptr(query), C.c_int64(nq), _ptr(train), C.c_int64(nt), int(nb), int(k),
                                                 _ptr(idx), _ptr(dist))
        else:
            if not _is_torch(mask):
                mask = np.ascontiguousarray(mask, np.uint8)
            assert tuple(mask.shape) == (nq, nt)
            rc = self.ctx._lib.brisk_hamming_knn_masked(self.ctx._h, _ptr(query), C.c_int64(nq), _ptr(train), C.c_int64(nt), int(nb), int(k),
                                                        _ptr(mask), _ptr(idx), _ptr(dist))
        self.ctx._check(rc)
        return idx, dist

    def radius(self, query, train, maxDistance, mask=None, sort=True):
        """-> (offsets [nq+1] int64, idx, dist): the matches of query i are idx/dist[offsets[i]:offsets[i+1]], in the
        reference's order (sort=True) or in train order.  numpy in / numpy out."""
        query = np.ascontiguousarray(query, np.uint8)
        train = np.ascontiguousarray(train, np.uint8).reshape(-1, query.shape[1])
        nq, nb = query.shape
        nt = train.shape[0]
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            assert mask.shape == (nq, nt)
        offsets = np.zeros(nq + 1, np.int64)
        cap = max(1024, 4 * nq)
        while True:
            idx = np.zeros(cap, np.int32)
            dist = np.zeros(cap, np.int32)
            rc = self.ctx._lib.brisk_hamming_radius(self.ctx._h, _ptr(query), C.c_int64(nq), _ptr(train), C.c_int64(nt), int(nb),
                                                    C.c_float(maxDistance), _ptr(mask) if mask is not None else None, int(bool(sort)),
                                                    _ptr(offsets), _ptr(idx), _ptr(dist), C.c_int64(cap))
            if rc == -4 and offsets[nq] > cap:  # BRISK_ERR_CAPACITY: offsets[nq] is the size needed
                cap = int(offsets[nq])
                continue
            self.ctx._check(rc)
            n = int(offsets[nq])
            return offsets, idx[:n], dist[:n]

    # --- cv::DescriptorMatcher surface ---
    def _collection(self, query, trainDescriptors, masks):
        query = np.ascontiguousarray(query, np.uint8)
        trains = [np.ascontiguousarray(trainDescriptors, np.uint8)] if trainDescriptors is not None else self._train
        trains = [t.reshape(-1, query.shape[1]) for t in trains]
        start = np.concatenate([[0], np.cumsum([len(t) for t in trains])]).astype(np.int64)
        train = np.concatenate(trains) if trains else np.zeros((0, query.shape[1]), np.uint8)
        mask = None
        masked_out = np.zeros(len(query), bool)
        if masks is not None and not isinstance(masks, (list, tuple)):
            masks = [masks]
        if masks:
            assert len(masks) == len(trains)
            cols = []
            every_mask_given = True
            for m, t in zip(masks, trains):
                if m is None or np.size(m) == 0:
                    cols.append(np.ones((len(query), len(t)), np.uint8))
                    every_mask_given = False
                else:
                    cols.append((np.asarray(m).reshape(len(query), len(t)) != 0).astype(np.uint8))
            mask = np.ascontiguousarray(np.concatenate(cols, axis=1))
            if every_mask_given:  # cv::DescriptorMatcher::isMaskedOut
                masked_out = np.array([all(not c[q].any() for c in cols) for q in range(len(query))], bool)
        return query, trains, train, start, mask, masked_out

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=1, masks=None, compactResult=False):
        """cv::DescriptorMatcher::knnMatch -> list (per query) of (queryIdx, trainIdx, imgIdx, distance) tuples,
        as BruteForceMatcher::commonKnnMatchImpl returns them (brute-force-matcher.cc:80-162)."""
        query, trains, train, start, mask, masked_out = self._collection(queryDescriptors, trainDescriptors, masks)
        out = []
        if len(query) == 0:
            return out
        nonempty = [i for i, t in enumerate(trains) if len(t)]
        if not nonempty:
            return [[] for q in range(len(query)) if not (compactResult and masked_out[q])]
        idx, dist = self.knn(query, train, k, mask)
        for q in range(len(query)):
            if masked_out[q]:
                if not compactResult:
                    out.append([])
                continue
            cur = []
            for g, d in zip(idx[q], dist[q]):
                if g >= 0:
                    img = int(np.searchsorted(start, g, side="right") - 1)
                    cur.append((q, int(g - start[img]), img, float(d)))
                else:
                    # the reference keeps selecting once the real candidates have run out: every entry holds INT_MAX,
                    # minMaxLoc returns location 0 and INT_MAX (as double) is below the float it was just rounded to,
                    # so the last non-empty image wins (brute-force-matcher.cc:138-157)
                    cur.append((q, 0, nonempty[-1], 2147483648.0))
            if len(cur) > 16:
                # the reference's final std::sort (brute-force-matcher.cc:160): beyond 16 entries libstdc++'s introsort
                # permutes entries of equal distance
                t_ = np.array([m_[1] for m_ in cur], np.int32)
                i_ = np.array([m_[2] for m_ in cur], np.int32)
                d_ = np.array([m_[3] for m_ in cur], np.float32)
                self.ctx._check(self.ctx._lib.brisk_std_sort_matches(C.c_int64(len(cur)), _ptr(t_), _ptr(i_), _ptr(d_)))
                cur = [(q, int(a), int(b), float(c_)) for a, b, c_ in zip(t_, i_, d_)]
            out.append(cur)
        return out

    def radiusMatch(self, queryDescriptors, trainDescriptors=None, maxDistance=0.0, masks=None, compactResult=False):
        """cv::DescriptorMatcher::radiusMatch -> as knnMatch; BruteForceMatcher::commonRadiusMatchImpl
        (brute-force-matcher.cc:164-214): distance < maxDistance, sorted by distance the way std::sort leaves them."""
        query, trains, train, start, mask, masked_out = self._collection(queryDescriptors, trainDescriptors, masks)
        if len(query) == 0:
            return []
        if len(train) == 0:
            return [[] for q in range(len(query)) if not (compactResult and masked_out[q])]
        offsets, idx, dist = self.radius(query, train, maxDistance, mask, sort=True)
        out = []
        for q in range(len(query)):
            if masked_out[q]:
                if not compactResult:
                    out.append([])
                continue
            cur = []
            for g, d in zip(idx[offsets[q]:offsets[q + 1]], dist[offsets[q]:offsets[q + 1]]):
                img = int(np.searchsorted(start, g, side="right") - 1)
                cur.append((q, int(g - start[img]), img, float(d)))
            out.append(cur)
        return out

    def match(self, queryDescriptors, trainDescriptors=None, masks=None):
        """cv::DescriptorMatcher::match: knnMatch(k=1, compactResult=true), flattened."""
        return [m for ms in self.knnMatch(queryDescriptors, trainDescriptors, 1, masks, True) for m in ms]

    # --- train set sharded across GPUs: local keys, caller exchanges them (NCCL all-gather), merge ---
    def knn_keys(self, query, train_shard, k, global_offset, keys_out):
        nq, nb = query.shape
        self.ctx._check(self.ctx._lib.brisk_hamming_knn_keys(self.ctx._h, _ptr(query), C.c_int64(nq), _ptr(train_shard),
                                                             C.c_int64(train_shard.shape[0]), int(nb), int(k), C.c_int64(global_offset), _ptr(keys_out)))
        return keys_out

    def merge_keys(self, gathered_keys, n_shards, nq, k, idx_out, dist_out):
        self.ctx._check(self.ctx._lib.brisk_knn_merge_keys(self.ctx._h, _ptr(gathered_keys), int(n_shards), C.c_int64(nq), int(k),
                                                           _ptr(idx_out), _ptr(dist_out)))
        return idx_out, dist_out
