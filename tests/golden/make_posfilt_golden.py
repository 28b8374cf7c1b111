"""Generates tests/golden/posfilt_golden.npz: KalmanFilter2D::filter (src/positionfilter/KalmanFilter2D.cpp:95-200)
executed on the REAL cv::KalmanFilter through the cv2 wheel (oracle/cv2ref.py: KalmanFilter2D) on the seeded
measurement streams of tests/golden/inputs.py.  The reference's own tests hold no vectors for it.

    python tests/golden/make_posfilt_golden.py     # needs cv2; run in the authoring container
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import cv2  # noqa: E402

import inputs  # noqa: E402
from oracle import cv2ref  # noqa: E402

out = {}
for name, (dt, timeout, sa, sn, seed, n) in inputs.KALMAN_CASES.items():
    kf = cv2ref.KalmanFilter2D(dt, timeout, sa, sn)
    rows = [tuple(float(v) for v in kf.filter(*m)) for m in inputs.kalman_track(seed, n)]
    out[f"kalman_{name}"] = np.array(rows, np.float64)  # valid, x, vx, y, vy
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(HERE, "posfilt_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
