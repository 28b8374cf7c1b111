"""Generates tests/golden/*.npz with the REAL OpenCV (cv2 wheel, the library the reference links
for every operation on this path) executing the reference's call sequences (oracle/cv2ref.py).
The reference's own tests hold no vectors for this path, so these are the pins.

    python tests/golden/make_golden.py        # needs cv2; run in the authoring container

Inputs are regenerated from seeds by tests/golden/inputs.py (shared with the tests), so only
outputs are stored."""
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


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def main():
    out = {}
    meta = {"cv2": cv2.__version__}
    # 1. MOG2 masks on noisy streams
    for name, (rows, cols, n, sigma, seed, lr) in inputs.MOG_STREAMS.items():
        bs = cv2.createBackgroundSubtractorMOG2()
        masks = [bs.apply(f, None, lr) for f in inputs.noisy_stream(rows, cols, n, sigma, seed)]
        out[f"mog_{name}"] = np.stack(masks)
    # 2. whole chain on the synthetic stream
    for name, (rows, cols, n, lr) in inputs.CHAINS.items():
        pipe = cv2ref.Pipeline(lr, **inputs.HSV_BAND)
        fg, thr, det, hcrc = [], [], [], []
        for t in range(n):
            f = inputs.synth_frame(rows, cols, inputs.SEED, t)
            valid, x, y, area = pipe.step(f)
            fg.append(pipe.mask.copy())
            thr.append(pipe.det.thr.copy())
            hcrc.append(crc(pipe.hsv))
            det.append((float(valid), x, y, area))
        out[f"chain_{name}_fg"] = np.stack(fg)
        out[f"chain_{name}_thr"] = np.stack(thr)
        out[f"chain_{name}_det"] = np.array(det, np.float64)
        out[f"chain_{name}_hsvcrc"] = np.array(hcrc, np.uint32)
    # 3. BGR->HSV on random colours + the table extremes
    bgr = inputs.hsv_colours()
    out["hsv_out"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
    # 4. inRange + morphology
    hsvimg = inputs.hsv_image()
    for i, (lo, hi) in enumerate(inputs.INRANGE_CASES):
        out[f"inrange_{i}"] = cv2.inRange(hsvimg, lo, hi)
    m = inputs.morph_mask()
    for k in inputs.MORPH_K:
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
        out[f"dilate_{k}"] = cv2.dilate(m, el)
        out[f"erode_{k}"] = cv2.erode(m, el)
    # 5. contours: per mask the cv2 contour list (first pixel index, m00, m10, m01) in cv2 order
    for name, mask in inputs.contour_masks().items():
        cs, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        rec = []
        for c in cs:
            mm = cv2.moments(c)
            x, y = c[0][0]
            rec.append((int(y) * mask.shape[1] + int(x), mm["m00"], mm["m10"], mm["m01"]))
        out[f"contours_{name}"] = np.array(rec, np.float64).reshape(-1, 4)
        out[f"sift_{name}"] = np.array(cv2ref.sift_contours(mask), np.float64)
        _, lab = cv2.connectedComponents(mask, connectivity=8)
        out[f"ncc_{name}"] = np.array([lab.max()], np.int64)
    # 6. bsub (alpha = 0 is exact in every cv2 build)
    b = cv2ref.Bsub(0.0)
    out["bsub_a0"] = np.stack([b.filter(f) for f in inputs.bsub_frames()])
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    with open(os.path.join(HERE, "golden.meta"), "w") as f:
        f.write(f"cv2 {cv2.__version__}\n")
    print("wrote", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes", meta)


if __name__ == "__main__":
    main()
