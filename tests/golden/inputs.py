"""Deterministic inputs of the golden fixtures (shared by make_golden.py and the tests).
Pure numpy; no cv2, no oracle."""
import numpy as np

SEED = 1000
HSV_BAND = dict(h=(40, 80), s=(100, 256), v=(100, 256))

# name: (rows, cols, frames, noise sigma, seed, learning rate)
MOG_STREAMS = {
    "frozen": (64, 96, 16, 3.0, 1, 0.0),
    "slow": (64, 96, 24, 3.0, 2, 0.01),
    "fast": (48, 80, 24, 10.0, 3, 0.1),
    "auto": (37, 53, 24, 8.0, 4, -1.0),
    "reinit": (32, 40, 4, 2.0, 5, 1.0),
}
# name: (rows, cols, frames, learning rate)
CHAINS = {"a0": (120, 160, 12, 0.0), "a001": (120, 160, 16, 0.01), "odd": (99, 150, 8, 0.01)}
INRANGE_CASES = [((0, 0, 0), (256, 256, 256)), ((40, 100, 100), (80, 256, 256)), ((256, 0, 0), (256, 256, 256)),
                 ((80, 0, 0), (40, 256, 256)), ((0, 10, 20), (255, 10, 20)), ((5, 5, 5), (200, 100, 50))]
MORPH_K = [1, 2, 3, 4, 5, 10, 11, 33]


def noisy_stream(rows, cols, nframes, sigma, seed):
    rng = np.random.default_rng(seed)
    bg = rng.integers(30, 200, (rows, cols, 3)).astype(np.float32)
    for t in range(nframes):
        f = bg + rng.normal(0, sigma, bg.shape) if sigma > 0 else bg.copy()
        x0 = (5 * t) % max(cols - 12, 1)
        y0 = (3 * t) % max(rows - 12, 1)
        f[y0:y0 + 12, x0:x0 + 12] = (30, 220, 60)
        xs = (cols - 20 - 4 * t) % max(cols - 16, 1)
        f[rows // 2:rows // 2 + 10, xs:xs + 16] = bg[rows // 2:rows // 2 + 10, xs:xs + 16] * 0.7
        yield np.clip(np.rint(f), 0, 255).astype(np.uint8)


def _fmix32(h):
    h = h.astype(np.uint32, copy=True)
    h ^= h >> np.uint32(16)
    h *= np.uint32(0x85EBCA6B)
    h ^= h >> np.uint32(13)
    h *= np.uint32(0xC2B2AE35)
    h ^= h >> np.uint32(16)
    return h


def synth_frame(rows, cols, seed, t):
    """The SURVEY.md 8(d) stream (same arithmetic as oracle/synth.py and the CUDA generator)."""
    with np.errstate(over="ignore"):
        kbg = _fmix32(np.array([(seed ^ 0x9E3779B9) & 0xFFFFFFFF], np.uint32))[0]
        knz = _fmix32(np.array([(seed + 0x7F4A7C15 * (t + 1)) & 0xFFFFFFFF], np.uint32))[0]
        idx = np.arange(rows * cols * 3, dtype=np.uint32)
        bg = 40 + (_fmix32(idx ^ kbg) % np.uint32(81)).astype(np.int32)
        nz = (_fmix32(idx ^ knz) % np.uint32(7)).astype(np.int32) - 3
    img = (bg + nz).astype(np.uint8).reshape(rows, cols, 3)
    if t != 0:
        cx, cy = cols // 4 + (7 * t) % (cols // 2), rows // 3 + (4 * t) % (rows // 3)
        r = rows // 20
        yy, xx = np.mgrid[0:rows, 0:cols]
        img[(xx - cx) ** 2 + (yy - cy) ** 2 <= r * r] = (40, 220, 60)
    return img


def hsv_colours():
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (255, 256, 3)).astype(np.uint8)
    edge = np.array([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 0, 0], [0, 1, 0],
                     [0, 0, 1], [255, 255, 0], [0, 255, 255], [255, 0, 255], [128, 128, 128], [254, 255, 255],
                     [17, 17, 16], [200, 100, 100], [100, 200, 100]], np.uint8)
    row = np.zeros((1, 256, 3), np.uint8)
    row[0, :16] = edge
    row[0, 16:] = rng.integers(0, 4, (240, 3)).astype(np.uint8) * 85
    return np.concatenate([a, row], 0)


def hsv_image():
    rng = np.random.default_rng(8)
    img = rng.integers(0, 256, (60, 100, 3)).astype(np.uint8)
    img[..., 0] %= 180
    img[10:30, 10:40] = (60, 200, 210)
    return img


def blobs(rows, cols, n, rmax, seed, holes=True):
    rng = np.random.default_rng(seed)
    m = np.zeros((rows, cols), np.uint8)
    yy, xx = np.mgrid[0:rows, 0:cols]
    for _ in range(n):
        cy, cx, r = rng.integers(0, rows), rng.integers(0, cols), rng.integers(1, rmax + 1)
        m[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
        if holes and r > 3 and rng.random() < 0.5:
            m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 2) ** 2] = 0
            if r > 8 and rng.random() < 0.5:
                m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 4) ** 2] = 255
    return m


def morph_mask():
    m = blobs(120, 200, 10, 20, seed=99)
    m[0:3, 0:3] = 255
    m[60, 100] = 255
    return m


def contour_masks():
    rows, cols = 80, 140
    z = np.zeros((rows, cols), np.uint8)
    c = {}
    c["empty"] = z.copy()
    c["full"] = np.full((rows, cols), 255, np.uint8)
    m = z.copy(); m[10, 10] = 255; c["single_pixel"] = m
    m = z.copy(); m[20, 5:100] = 255; m[5:70, 120] = 255; c["lines"] = m
    m = z.copy(); m[10:15, 10:17] = 255; c["rect_5x7"] = m
    m = z.copy(); m[10:12, 10:12] = 255; m[12:14, 12:14] = 255; c["diag_blocks"] = m
    m = z.copy(); m[10:40, 10:40] = 255; m[15:35, 15:35] = 0; m[22:28, 22:28] = 255; c["ring_with_island"] = m
    m = z.copy(); m[5:75, 5:135] = 255; m[10:70, 10:130] = 0; m[15:65, 15:125] = 255; m[20:60, 20:120] = 0
    m[30:50, 40:100] = 255; m[35:45, 50:90] = 0; c["nested_rings"] = m
    m = z.copy(); m[10:20, 10:20] = 255; m[10:20, 60:70] = 255; m[40:50, 30:40] = 255; c["area_tie"] = m
    m = z.copy(); m[0:10, 0:10] = 255; m[rows - 6:rows, cols - 9:cols] = 255; m[0:4, 60:90] = 255
    c["touching_borders"] = m
    m = np.full((rows, cols), 255, np.uint8); m[20:30, 20:30] = 0; m[0, 50] = 0; m[40:42, cols - 1] = 0
    c["full_with_holes"] = m
    yy, xx = np.mgrid[0:rows, 0:cols]
    m = z.copy(); m[(yy + xx) % 2 == 0] = 255; c["checkerboard"] = m
    m = z.copy(); m[30:34, 28:36] = 255; m[31:33, 31:33] = 0; c["word_boundary_hole"] = m
    for s in range(6):
        c[f"blobs{s}"] = blobs([64, 33, 100][s % 3], [96, 47, 131][s % 3], 12, 14, s)
    for i, d in enumerate([0.1, 0.5, 0.9]):
        c[f"noise{i}"] = (np.random.default_rng(50 + i).random((90, 150)) < d).astype(np.uint8) * 255
    return c


def bsub_frames():
    rng = np.random.default_rng(0)
    return [rng.integers(0, 256, (40, 56, 3)).astype(np.uint8) for _ in range(6)]


# ---- posifilt kalman / posicom mean ------------------------------------
# name -> (dt, timeout, sigma_accel, sigma_noise, seed, samples)
KALMAN_CASES = {
    "default": (0.02, 0.0, 5.0, 0.0, 11, 60),          # Oat defaults: timeout 0 -> never valid
    "t1": (0.02, 1.0, 5.0, 0.0, 12, 300),
    "noisy": (1.0 / 30.0, 0.5, 50.0, 3.5, 13, 300),
    "stiff": (0.01, 0.2, 1.0, 1.0, 14, 300),
    "degenerate": (0.02, 1.0, 0.0, 0.0, 15, 120),       # singular innovation covariance
}


def kalman_track(seed, n):
    """Seeded measurement stream [(valid, x, y)]: a random walk with drop-outs, incl. gaps longer than
    any timeout above (re-initialisation) and an invalid lead-in."""
    rng = np.random.default_rng(seed)
    x, y = 320.0, 240.0
    out = []
    for t in range(n):
        x += rng.normal() * 3.0 + 1.0
        y += rng.normal() * 3.0 - 0.5
        if t < 3:
            valid = False
        elif (t // 50) % 3 == 2:
            valid = rng.random() > 0.93
        else:
            valid = rng.random() > 0.2
        out.append((bool(valid), float(x), float(y)))
    return out
