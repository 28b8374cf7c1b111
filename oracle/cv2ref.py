"""The reference's own call sequence on this path, executed with the real OpenCV (cv2 wheel).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the PIN for the C restatement in
oat_oracle.c and the CPU baseline bench.py times: a line-by-line Python transcription of

  BackgroundSubtractorMOG::filter   /root/reference/src/framefilter/BackgroundSubtractorMOG.cpp:124-125
  ColorConvert::filter              /root/reference/src/framefilter/ColorConvert.cpp:101-107
  HSVDetector::detectPosition       /root/reference/src/positiondetector/HSVDetector.cpp:142-173
  siftContours                      /root/reference/src/positiondetector/DetectorFunc.cpp:31-66
  BackgroundSubtractor::filter      /root/reference/src/framefilter/BackgroundSubtractor.cpp:87-100

calling the same cv:: functions.  Version skew: the reference targeted OpenCV 3.x; this wheel is
4.13 (same algorithms/defaults on this path; findContours no longer modifies its input).
"""
from __future__ import annotations

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover - cv2 is in the image, but keep the import soft
    cv2 = None

DBL_MAX = float(np.finfo(np.float64).max)


def available() -> bool:
    return cv2 is not None


class MogFilter:
    """framefilt mog (CPU branch)."""

    def __init__(self, learning_coeff: float = 0.0):
        self.bs = cv2.createBackgroundSubtractorMOG2()  # all defaults, ...MOG.cpp:83
        self.lr = learning_coeff
        self.mask = None

    def filter(self, frame: np.ndarray) -> np.ndarray:
        """In place, like the reference; returns frame."""
        self.mask = self.bs.apply(frame, self.mask, self.lr)
        frame[self.mask == 0] = 0  # frame.setTo(0, background_mask_ == 0)
        return frame


def col_hsv(frame: np.ndarray) -> np.ndarray:
    """framefilt col -C HSV."""
    return cv2.cvtColor(frame, cv2.COLOR_BGR2HSV)


class HsvDetector:
    """posidet hsv."""

    def __init__(self, h=(0, 256), s=(0, 256), v=(0, 256), erode=0, dilate=10, area=(0.0, DBL_MAX)):
        self.lo = (h[0], s[0], v[0])
        self.hi = (h[1], s[1], v[1])
        self.erode_el = cv2.getStructuringElement(cv2.MORPH_RECT, (erode, erode)) if erode > 0 else None
        self.dilate_el = cv2.getStructuringElement(cv2.MORPH_RECT, (dilate, dilate)) if dilate > 0 else None
        self.min_area, self.max_area = area
        self.thr = None

    def detect(self, hsv: np.ndarray):
        """-> (position_valid, x, y, area); self.thr holds the post-morphology mask."""
        thr = cv2.inRange(hsv, self.lo, self.hi)
        if self.erode_el is not None:
            thr = cv2.erode(thr, self.erode_el)
        if self.dilate_el is not None:
            thr = cv2.dilate(thr, self.dilate_el)
        self.thr = thr
        return sift_contours(thr, self.min_area, self.max_area)


def sift_contours(thr: np.ndarray, min_area=0.0, max_area=DBL_MAX):
    contours, _ = cv2.findContours(thr, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    object_area = 0.0
    valid, x, y = False, 0.0, 0.0
    for c in contours:
        m = cv2.moments(c)
        a = m["m00"]
        if a >= min_area and a < max_area and a > object_area:
            x = m["m10"] / a
            y = m["m01"] / a
            valid = True
            object_area = a
    return valid, x, y, object_area


class Bsub:
    """framefilt bsub."""

    def __init__(self, alpha: float = 0.0):
        self.alpha = alpha
        self.bg = None
        self.bgf = None

    def filter(self, frame: np.ndarray) -> np.ndarray:
        if self.bg is None:
            self.bg = frame.copy()
            self.bgf = frame.astype(np.float32)
        if self.alpha > 0.0:
            cv2.accumulateWeighted(frame, self.bgf, self.alpha)
            self.bg = np.clip(np.rint(self.bgf), 0, 255).astype(np.uint8)  # convertTo(CV_8U): cvRound + saturate
        return cv2.subtract(frame, self.bg)


class Pipeline:
    """mog -> col HSV -> hsv, preallocated, for parity and for CPU timing."""

    def __init__(self, learning_coeff=0.0, **hsv_kw):
        self.mog = cv2.createBackgroundSubtractorMOG2()
        self.lr = learning_coeff
        self.det = HsvDetector(**hsv_kw)
        self.mask = None
        self.zero = None
        self.filt = None
        self.hsv = None

    def step(self, frame: np.ndarray):
        if self.zero is None:
            self.zero = np.zeros_like(frame)
            self.filt = np.empty_like(frame)
            self.hsv = np.empty_like(frame)
            self.mask = np.empty(frame.shape[:2], np.uint8)
        self.mask = self.mog.apply(frame, self.mask, self.lr)
        # frame.setTo(0, mask == 0)  ==  copy frame where mask != 0 over zeros  
        self.filt[:] = 0
        cv2.copyTo(frame, self.mask, self.filt)
        cv2.cvtColor(self.filt, cv2.COLOR_BGR2HSV, dst=self.hsv)
        return self.det.detect(self.hsv)


class KalmanFilter2D:
    """KalmanFilter2D::filter on the real cv::KalmanFilter(4, 2, 0, CV_64F)
    (src/positionfilter/KalmanFilter2D.cpp:95-200).  Returns (valid, x, vx, y, vy)."""

    def __init__(self, dt=0.02, timeout=0.0, sigma_accel=5.0, sigma_noise=0.0):
        self.dt, self.sa, self.sn = dt, sigma_accel, sigma_noise
        self.thr = int(timeout / dt)
        self.kf = cv2.KalmanFilter(4, 2, 0, cv2.CV_64F)
        self.found, self.nf = False, 0
        self.meas = np.full((2, 1), 6.0)
        self.pred = np.full((4, 1), 6.0)

    def _init(self):
        dt, sa = self.dt, self.sa
        A = np.eye(4)
        A[0, 1] = dt
        A[2, 3] = dt
        self.kf.transitionMatrix = A
        H = np.zeros((2, 4))
        H[0, 0] = 1.0
        H[1, 2] = 1.0
        self.kf.measurementMatrix = H
        Q = np.zeros((4, 4))
        Q[0, 0] = Q[2, 2] = sa * sa * (dt * dt * dt * dt) / 4.0
        Q[0, 1] = Q[1, 0] = Q[2, 3] = Q[3, 2] = sa * sa * (dt * dt * dt) / 2.0
        Q[1, 1] = Q[3, 3] = sa * sa * (dt * dt)
        self.kf.processNoiseCov = Q
        self.kf.measurementNoiseCov = np.eye(2) * (self.sn * self.sn)
        self.kf.errorCovPre = np.eye(4) * 1000.0
        s0 = np.array([[self.meas[0, 0]], [0.0], [self.meas[1, 0]], [0.0]])
        self.kf.statePre = s0.copy()
        self.kf.statePost = s0.copy()

    def filter(self, valid, x=0.0, y=0.0):
        if valid:
            self.meas = np.array([[x], [y]], dtype=np.float64)
            self.nf = 0
            if not self.found:
                self._init()
            self.found = True
        else:
            self.nf += 1
        if self.nf >= self.thr:
            self.found = False
        if self.found:
            self.pred = self.kf.predict().copy()
            self.kf.correct(self.meas)
        return (self.found, self.pred[0, 0], self.pred[1, 0], self.pred[2, 0], self.pred[3, 0])
