"""numpy twin of the synthetic stream generator; bit-identical to
``orc_synth_frame`` (oat_oracle.c) and to the CUDA generator ``oat_synth_frame``.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np

DISC_BGR = (40, 220, 60)  # -> HSV (57, 209, 220)


def fmix32(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32, copy=True)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def disc_centre(rows: int, cols: int, t: int):
    return cols // 4 + (7 * t) % (cols // 2), rows // 3 + (4 * t) % (rows // 3)


def disc_radius(rows: int) -> int:
    return rows // 20


def frame(rows: int, cols: int, seed: int, t: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        kbg = fmix32(np.array([(seed ^ 0x9E3779B9) & 0xFFFFFFFF], np.uint32))[0]
        knz = fmix32(np.array([(seed + 0x7F4A7C15 * (t + 1)) & 0xFFFFFFFF], np.uint32))[0]
        idx = np.arange(rows * cols * 3, dtype=np.uint32)
        bg = 40 + (fmix32(idx ^ kbg) % np.uint32(81)).astype(np.int32)
        nz = (fmix32(idx ^ knz) % np.uint32(7)).astype(np.int32) - 3
    img = (bg + nz).astype(np.uint8).reshape(rows, cols, 3)
    if t != 0:
        cx, cy = disc_centre(rows, cols, t)
        r = disc_radius(rows)
        yy, xx = np.mgrid[0:rows, 0:cols]
        img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = DISC_BGR
    return img
