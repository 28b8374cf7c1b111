"""CPU oracle for Oat's tracking hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The product
(``oat_b200`` / ``liboatgpu.so``) never imports or links anything from here.

Three layers:

* ``oracle.lib``      -- ctypes binding of ``liboat_oracle.so`` (``oat_oracle.c``), the C
  restatement of the OpenCV algorithms behind each reference call site;
* ``oracle.synth``    -- numpy twin of the synthetic stream generator;
* ``oracle.cv2ref``   -- the reference's own call sequence executed with the real OpenCV
  (``cv2`` wheel) -- the pin the C restatement is checked against, and the CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboat_oracle.so")


def build(force: bool = False) -> str:
    """Compile oat_oracle.c with gcc (seconds). Returns the path of the shared object."""
    src = os.path.join(_HERE, "oat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "-B", "liboat_oracle.so"], check=True)
    return _SO


class MogParams(C.Structure):
    _fields_ = [
        ("history", C.c_int),
        ("nmixtures", C.c_int),
        ("var_threshold", C.c_float),
        ("var_threshold_gen", C.c_float),
        ("background_ratio", C.c_float),
        ("var_init", C.c_float),
        ("var_min", C.c_float),
        ("var_max", C.c_float),
        ("ct", C.c_float),
        ("detect_shadows", C.c_int),
        ("shadow_value", C.c_int),
        ("shadow_threshold", C.c_float),
    ]


class Detection(C.Structure):
    _fields_ = [
        ("position_valid", C.c_int32),
        ("n_components", C.c_int32),
        ("x", C.c_double),
        ("y", C.c_double),
        ("area", C.c_double),
    ]

    def as_tuple(self):
        return (bool(self.position_valid), self.x, self.y, self.area)


class Position(C.Structure):
    """The Position2D fields the position filters/combiners touch (lib/datatypes/Position2D.h:112-155)."""
    _fields_ = [
        ("position_valid", C.c_int32),
        ("velocity_valid", C.c_int32),
        ("heading_valid", C.c_int32),
        ("reserved", C.c_int32),
        ("x", C.c_double),
        ("y", C.c_double),
        ("vx", C.c_double),
        ("vy", C.c_double),
        ("hx", C.c_double),
        ("hy", C.c_double),
    ]

    def as_tuple(self):
        return (bool(self.position_valid), bool(self.velocity_valid), bool(self.heading_valid),
                self.x, self.y, self.vx, self.vy, self.hx, self.hy)


class ContourRec(C.Structure):
    _fields_ = [
        ("first_index", C.c_int32),
        ("npoints", C.c_int32),
        ("m00", C.c_double),
        ("m10", C.c_double),
        ("m01", C.c_double),
    ]


_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p, vp, sz, i, u32, dbl = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_double
    L.orc_synth_frame.argtypes = [u8p, sz, i, i, u32, u32]
    L.orc_mog_default_params.argtypes = [C.POINTER(MogParams)]
    L.orc_mog_create.argtypes = [i, i, C.POINTER(MogParams)]
    L.orc_mog_create.restype = vp
    L.orc_mog_destroy.argtypes = [vp]
    L.orc_mog_reset.argtypes = [vp]
    L.orc_mog_state.argtypes = [vp] + [C.POINTER(C.c_void_p)] * 4
    L.orc_mog_apply.argtypes = [vp, u8p, sz, u8p, sz, dbl]
    L.orc_mog_effective_rate.argtypes = [i, dbl, i]
    L.orc_mog_effective_rate.restype = dbl
    L.orc_zero_where_mask0.argtypes = [u8p, sz, u8p, sz, i, i, i]
    L.orc_hsv_tables.argtypes = [vp, vp]
    L.orc_bgr2hsv.argtypes = [u8p, sz, u8p, sz, i, i]
    L.orc_inrange3.argtypes = [u8p, sz, u8p, sz, i, i, C.POINTER(C.c_int * 3), C.POINTER(C.c_int * 3)]
    L.orc_erode_rect.argtypes = [u8p, sz, u8p, sz, i, i, i]
    L.orc_dilate_rect.argtypes = [u8p, sz, u8p, sz, i, i, i]
    L.orc_label8.argtypes = [u8p, sz, i, i, vp]
    L.orc_external_contours.argtypes = [u8p, sz, i, i, C.POINTER(ContourRec), i]
    L.orc_external_contours.restype = i
    L.orc_sift_contours.argtypes = [u8p, sz, i, i, dbl, dbl, C.POINTER(Detection)]
    L.orc_cell_moments.argtypes = [u8p, sz, i, i, vp, vp, vp, vp, i]
    L.orc_cell_moments.restype = i
    L.orc_bsub_create.argtypes = [i, i, i, dbl]
    L.orc_bsub_create.restype = vp
    L.orc_bsub_destroy.argtypes = [vp]
    L.orc_bsub_apply.argtypes = [vp, u8p, sz, u8p, sz]
    L.orc_kalman_create.argtypes = [dbl, dbl, dbl, dbl]
    L.orc_kalman_create.restype = vp
    L.orc_kalman_destroy.argtypes = [vp]
    L.orc_kalman_filter.argtypes = [vp, C.POINTER(Position)]
    L.orc_mean_combine.argtypes = [C.POINTER(Position), i, i, C.POINTER(Position)]
    _lib = L
    return L


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype=np.uint8):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


# ---- thin functional wrappers ------------------------------------------------------------

def synth_frame(rows: int, cols: int, seed: int, t: int) -> np.ndarray:
    out = np.empty((rows, cols, 3), np.uint8)
    _load().orc_synth_frame(_p(out), cols * 3, rows, cols, seed, t)
    return out


def default_mog_params() -> MogParams:
    p = MogParams()
    _load().orc_mog_default_params(C.byref(p))
    return p


class Mog2:
    """C restatement of cv::BackgroundSubtractorMOG2 with visible state (OpenCV layout)."""

    def __init__(self, rows: int, cols: int, params: MogParams | None = None):
        self.rows, self.cols = rows, cols
        self.params = params or default_mog_params()
        self.K = self.params.nmixtures
        self._h = _load().orc_mog_create(rows, cols, C.byref(self.params))

    def __del__(self):
        if getattr(self, "_h", None):
            _load().orc_mog_destroy(self._h)
            self._h = None

    def reset(self):
        _load().orc_mog_reset(self._h)

    def apply(self, bgr: np.ndarray, learning_rate: float) -> np.ndarray:
        bgr = _c(bgr)
        assert bgr.shape == (self.rows, self.cols, 3)
        mask = np.empty((self.rows, self.cols), np.uint8)
        _load().orc_mog_apply(self._h, _p(bgr), self.cols * 3, _p(mask), self.cols, learning_rate)
        return mask

    def state(self):
        """(modes_used u8[H,W], weight f32[H,W,K], variance f32[H,W,K], mean f32[H,W,K,3]) copies."""
        ptrs = [C.c_void_p() for _ in range(4)]
        _load().orc_mog_state(self._h, *[C.byref(p) for p in ptrs])
        n = self.rows * self.cols

        def arr(p, count, ct, dt, shape):
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), (count,)).astype(dt).reshape(shape).copy()

        K = self.K
        return (
            arr(ptrs[0], n, C.c_uint8, np.uint8, (self.rows, self.cols)),
            arr(ptrs[1], n * K, C.c_float, np.float32, (self.rows, self.cols, K)),
            arr(ptrs[2], n * K, C.c_float, np.float32, (self.rows, self.cols, K)),
            arr(ptrs[3], n * K * 3, C.c_float, np.float32, (self.rows, self.cols, K, 3)),
        )


def zero_where_mask0(frame: np.ndarray, mask: np.ndarray) -> np.ndarray:
    out = _c(frame).copy()
    mask = _c(mask)
    ch = 1 if out.ndim == 2 else out.shape[2]
    _load().orc_zero_where_mask0(_p(out), out.shape[1] * ch, _p(mask), mask.shape[1], out.shape[0], out.shape[1], ch)
    return out


def bgr2hsv(bgr: np.ndarray) -> np.ndarray:
    bgr = _c(bgr)
    out = np.empty_like(bgr)
    _load().orc_bgr2hsv(_p(bgr), bgr.shape[1] * 3, _p(out), bgr.shape[1] * 3, bgr.shape[0], bgr.shape[1])
    return out


def hsv_tables():
    s = np.zeros(256, np.int32)
    h = np.zeros(256, np.int32)
    _load().orc_hsv_tables(_p(s), _p(h))
    return s, h


def inrange3(img: np.ndarray, lo, hi) -> np.ndarray:
    img = _c(img)
    out = np.empty(img.shape[:2], np.uint8)
    lo_a = (C.c_int * 3)(*[int(v) for v in lo])
    hi_a = (C.c_int * 3)(*[int(v) for v in hi])
    _load().orc_inrange3(_p(img), img.shape[1] * 3, _p(out), img.shape[1], img.shape[0], img.shape[1],
                         C.byref(lo_a), C.byref(hi_a))
    return out


def erode_rect(mask: np.ndarray, k: int) -> np.ndarray:
    mask = _c(mask)
    out = np.empty_like(mask)
    _load().orc_erode_rect(_p(mask), mask.shape[1], _p(out), mask.shape[1], mask.shape[0], mask.shape[1], k)
    return out


def dilate_rect(mask: np.ndarray, k: int) -> np.ndarray:
    mask = _c(mask)
    out = np.empty_like(mask)
    _load().orc_dilate_rect(_p(mask), mask.shape[1], _p(out), mask.shape[1], mask.shape[0], mask.shape[1], k)
    return out


def label8(mask: np.ndarray) -> np.ndarray:
    mask = _c(mask)
    out = np.empty(mask.shape, np.int32)
    _load().orc_label8(_p(mask), mask.shape[1], mask.shape[0], mask.shape[1], _p(out))
    return out


def external_contours(mask: np.ndarray):
    """[(first_index, npoints, m00, m10, m01)] in raster order of the first pixel."""
    mask = _c(mask)
    cap = 4096
    while True:
        recs = (ContourRec * cap)()
        n = _load().orc_external_contours(_p(mask), mask.shape[1], mask.shape[0], mask.shape[1], recs, cap)
        if n <= cap:
            break
        cap = n
    return [(r.first_index, r.npoints, r.m00, r.m10, r.m01) for r in recs[:n]]


def sift_contours(mask: np.ndarray, min_area: float = 0.0, max_area: float = float(np.finfo(np.float64).max)):
    mask = _c(mask)
    d = Detection()
    _load().orc_sift_contours(_p(mask), mask.shape[1], mask.shape[0], mask.shape[1], min_area, max_area, C.byref(d))
    return d


def cell_moments(mask: np.ndarray):
    """Integer 2x2-cell sums per external component: (first_index, 2*m00, 6*m10, 6*m01)."""
    mask = _c(mask)
    cap = mask.size // 2 + 2
    fi = np.zeros(cap, np.int32)
    s00 = np.zeros(cap, np.int64)
    s10 = np.zeros(cap, np.int64)
    s01 = np.zeros(cap, np.int64)
    n = _load().orc_cell_moments(_p(mask), mask.shape[1], mask.shape[0], mask.shape[1], _p(fi), _p(s00), _p(s10),
                                 _p(s01), cap)
    return fi[:n], s00[:n], s10[:n], s01[:n]


class Bsub:
    def __init__(self, rows, cols, ch=3, alpha=0.0):
        self.rows, self.cols, self.ch = rows, cols, ch
        self._h = _load().orc_bsub_create(rows, cols, ch, alpha)

    def __del__(self):
        if getattr(self, "_h", None):
            _load().orc_bsub_destroy(self._h)
            self._h = None

    def apply(self, frame):
        frame = _c(frame)
        out = np.empty_like(frame)
        pitch = self.cols * self.ch
        _load().orc_bsub_apply(self._h, _p(frame), pitch, _p(out), pitch)
        return out


# ---- the whole hot path, restated -----------------------------------------------------------

DBL_MAX = float(np.finfo(np.float64).max)


class HsvParams:
    """HSVDetector options (src/positiondetector/HSVDetector.cpp:54-69; defaults HSVDetector.h:86-94, .cpp:42-43)."""

    def __init__(self, h=(0, 256), s=(0, 256), v=(0, 256), erode=0, dilate=10, area=(0.0, DBL_MAX)):
        self.h, self.s, self.v = tuple(h), tuple(s), tuple(v)
        self.erode, self.dilate = erode, dilate
        self.area = tuple(area)


def hsv_detect(hsv: np.ndarray, p: HsvParams):
    """HSVDetector::detectPosition (HSVDetector.cpp:142-173) -> (Detection, threshold mask)."""
    thr = inrange3(hsv, (p.h[0], p.s[0], p.v[0]), (p.h[1], p.s[1], p.v[1]))
    if p.erode > 0:
        thr = erode_rect(thr, p.erode)
    if p.dilate > 0:
        thr = dilate_rect(thr, p.dilate)
    return sift_contours(thr, p.area[0], p.area[1]), thr


class Tracker:
    """framefilt mog -> framefilt col -C HSV -> posidet hsv, C restatement end to end."""

    def __init__(self, rows, cols, mog_params: MogParams | None = None):
        self.mog = Mog2(rows, cols, mog_params)

    def track(self, bgr, learning_rate, p: HsvParams):
        fg = self.mog.apply(bgr, learning_rate)
        filt = zero_where_mask0(bgr, fg)
        hsv = bgr2hsv(filt)
        det, thr = hsv_detect(hsv, p)
        return det, dict(fgmask=fg, bgr=filt, hsv=hsv, thresh=thr)


# ---- posidet thresh / framefilt thresh / framefilt mask (numpy restatements; pinned against cv2 in tests) -------
def inrange1(grey: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """cv::inRange on a 1-channel u8 image with Oat's 0..256 option range (src/positiondetector/
    SimpleThreshold.cpp:169-172): inclusive both ends, bounds saturate to u8 so 256 == 255, lo > 255 passes nothing."""
    g = grey.astype(np.int32)
    return (((g >= lo) & (g <= hi)).astype(np.uint8)) * 255


def bgr2grey(bgr: np.ndarray) -> np.ndarray:
    """8-bit cv::COLOR_BGR2GRAY as the pinned OpenCV 4.13 computes it: 15-bit fixed point
    (3735 b + 19235 g + 9798 r + 16384) >> 15 (OpenCV 3.x used 14-bit coefficients; +-1 on ~0.3 % of colours)."""
    b = bgr[..., 0].astype(np.int32)
    g = bgr[..., 1].astype(np.int32)
    r = bgr[..., 2].astype(np.int32)
    return ((3735 * b + 19235 * g + 9798 * r + 16384) >> 15).astype(np.uint8)


def threshold_filter(frame: np.ndarray, i_min: int, i_max: int) -> np.ndarray:
    """Threshold::filter (src/framefilter/Threshold.cpp:67-81): grey inRange, then frame.setTo(0, thresh == 0)."""
    grey = bgr2grey(frame) if frame.ndim == 3 else frame
    keep = inrange1(grey, i_min, i_max) != 0
    out = frame.copy()
    out[~keep] = 0
    return out


def mask_filter(frame: np.ndarray, roi: np.ndarray) -> np.ndarray:
    """FrameMasker::filter (src/framefilter/FrameMasker.cpp:71-75): frame.setTo(0, roi == 0)."""
    out = frame.copy()
    out[roi == 0] = 0
    return out


def thresh_detect(grey: np.ndarray, t_min: int, t_max: int, p: "HsvParams"):
    """SimpleThreshold::detectPosition: inRange -> erode -> dilate -> siftContours. Returns (Detection, mask)."""
    m = inrange1(grey, t_min, t_max)
    if p.erode > 0:
        m = erode_rect(m, p.erode)
    if p.dilate > 0:
        m = dilate_rect(m, p.dilate)
    return sift_contours(m, p.area[0], p.area[1]), m


def blur_nonzero(mask: np.ndarray, k: int) -> np.ndarray:
    """(cv::blur(mask, k x k) != 0) for a 0/255 mask: box average with anchor k/2, BORDER_REFLECT_101, rounded to
    u8 -- non-zero iff the (multiplicity-counted) number c of set samples in the box has 510 c > k^2."""
    a = k // 2
    rows, cols = mask.shape
    p = np.pad((mask != 0).astype(np.int64), ((a, k - 1 - a), (a, k - 1 - a)), mode="reflect")
    cnt = np.zeros(mask.shape, np.int64)
    for dy in range(k):
        for dx in range(k):
            cnt += p[dy:dy + rows, dx:dx + cols]
    return ((510 * cnt > k * k).astype(np.uint8)) * 255


class DifferenceDetector:
    """DifferenceDetector::detectPosition (src/positiondetector/DifferenceDetector.cpp:118-173)."""

    def __init__(self, diff_threshold=10, blur=2, area=(0.0, DBL_MAX)):
        self.thr, self.blur, self.area = diff_threshold, blur, tuple(area)
        self.last = None

    def detect(self, grey: np.ndarray):
        if self.last is None:  # threshold_frame_ = frame.clone(): the raw frame is sifted (non-zero = object)
            m = ((grey != 0).astype(np.uint8)) * 255
        else:
            d = np.abs(grey.astype(np.int32) - self.last.astype(np.int32))
            m = ((d > self.thr).astype(np.uint8)) * 255
            if self.blur > 0:
                m = blur_nonzero(m, self.blur)
        self.last = grey.copy()
        return sift_contours(m, self.area[0], self.area[1]), m


# ---- posifilt kalman / posicom mean -------------------------------------
class Kalman2D:
    """KalmanFilter2D (src/positionfilter/KalmanFilter2D.cpp:95-200); defaults KalmanFilter2D.h:66-83."""

    def __init__(self, dt=0.02, timeout=0.0, sigma_accel=5.0, sigma_noise=0.0):
        self._h = _load().orc_kalman_create(dt, timeout, sigma_accel, sigma_noise)

    def __del__(self):
        if getattr(self, "_h", None):
            _load().orc_kalman_destroy(self._h)
            self._h = None

    def filter(self, valid: bool, x: float = 0.0, y: float = 0.0) -> Position:
        p = Position(position_valid=int(valid), x=x, y=y)
        _load().orc_kalman_filter(self._h, C.byref(p))
        return p


def mean_combine(sources, heading_anchor: int = -1) -> Position:
    """MeanPosition::combine (src/positioncombiner/MeanPosition.cpp:60-118)."""
    arr = (Position * len(sources))(*sources)
    out = Position()
    _load().orc_mean_combine(arr, len(sources), heading_anchor, C.byref(out))
    return out
