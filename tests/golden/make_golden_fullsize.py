"""Full-size pins of the HEADLINE configurations (BASELINE.json configs 1-2), made with the REAL OpenCV (cv2
wheel) executing the reference's call sequence (oracle/cv2ref.py: BackgroundSubtractorMOG.cpp:124-125,
ColorConvert.cpp:101-107, HSVDetector.cpp:142-173, DetectorFunc.cpp:31-66) on the SURVEY.md 8(d) stream:

    1080p, -a 0.01 : 240 frames        1080p, -a 0 : 120 frames        4K, -a 0.01 : 40 frames

Per frame only a CRC32 of the foreground mask, of the HSV frame and of the post-morphology threshold mask plus the
detection (valid, x, y, area) are stored -- a few KB -- so that the CUDA path can be compared against cv2 itself
at full size on a box without cv2, and the C oracle at full size in the CPU suite.

    python tests/golden/make_golden_fullsize.py        # needs cv2; run in the authoring container (~1 min)
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import cv2  # noqa: E402

import inputs  # noqa: E402
from oracle import cv2ref  # noqa: E402

# name: (rows, cols, frames, learning rate)
FULLSIZE = {"1080p_a001": (1080, 1920, 240, 0.01), "1080p_a0": (1080, 1920, 120, 0.0), "4k_a001": (2160, 3840, 40, 0.01)}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def main():
    out = {}
    for name, (rows, cols, n, lr) in FULLSIZE.items():
        pipe = cv2ref.Pipeline(lr, **inputs.HSV_BAND)
        rec = np.zeros((n, 3), np.uint32)
        det = np.zeros((n, 4), np.float64)
        nfg = np.zeros(n, np.int64)
        for t in range(n):
            f = inputs.synth_frame(rows, cols, inputs.SEED, t)
            valid, x, y, area = pipe.step(f)
            rec[t] = (crc(pipe.mask), crc(pipe.hsv), crc(pipe.det.thr))
            det[t] = (float(valid), x, y, area)
            nfg[t] = int((pipe.mask == 255).sum())
        out[f"{name}_crc"] = rec
        out[f"{name}_det"] = det
        out[f"{name}_nfg"] = nfg
        print(name, "done; last detection", det[-1])
    np.savez_compressed(os.path.join(HERE, "golden_fullsize.npz"), **out)
    with open(os.path.join(HERE, "golden_fullsize.meta"), "w") as f:
        f.write(f"cv2 {cv2.__version__}\n")
    print("wrote", os.path.getsize(os.path.join(HERE, "golden_fullsize.npz")), "bytes")


if __name__ == "__main__":
    main()
