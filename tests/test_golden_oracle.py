"""The CPU oracle against the committed golden vectors (made with the real OpenCV by
tests/golden/make_golden.py).  CPU only; does not need cv2."""
import numpy as np
import pytest

import oracle
from golden_util import crc, inputs, load

G = load()
TOL = 1e-6


@pytest.mark.parametrize("name", list(inputs.MOG_STREAMS))
def test_mog_masks(name):
    rows, cols, n, sigma, seed, lr = inputs.MOG_STREAMS[name]
    orc = oracle.Mog2(rows, cols)
    for t, f in enumerate(inputs.noisy_stream(rows, cols, n, sigma, seed)):
        assert np.array_equal(orc.apply(f, lr), G[f"mog_{name}"][t]), f"{name} frame {t}"


@pytest.mark.parametrize("name", list(inputs.CHAINS))
def test_chain(name):
    rows, cols, n, lr = inputs.CHAINS[name]
    trk = oracle.Tracker(rows, cols)
    hp = oracle.HsvParams(**inputs.HSV_BAND)
    for t in range(n):
        f = inputs.synth_frame(rows, cols, inputs.SEED, t)
        assert np.array_equal(f, oracle.synth_frame(rows, cols, inputs.SEED, t))
        d, eg = trk.track(f, lr, hp)
        assert np.array_equal(eg["fgmask"], G[f"chain_{name}_fg"][t])
        assert np.array_equal(eg["thresh"], G[f"chain_{name}_thr"][t])
        assert crc(eg["hsv"]) == G[f"chain_{name}_hsvcrc"][t]
        valid, x, y, area = G[f"chain_{name}_det"][t]
        assert bool(d.position_valid) == bool(valid)
        assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL and abs(d.area - area) <= TOL


def test_hsv_inrange_morph():
    assert np.array_equal(oracle.bgr2hsv(inputs.hsv_colours()), G["hsv_out"])
    img = inputs.hsv_image()
    for i, (lo, hi) in enumerate(inputs.INRANGE_CASES):
        assert np.array_equal(oracle.inrange3(img, lo, hi), G[f"inrange_{i}"])
    m = inputs.morph_mask()
    for k in inputs.MORPH_K:
        assert np.array_equal(oracle.dilate_rect(m, k), G[f"dilate_{k}"])
        assert np.array_equal(oracle.erode_rect(m, k), G[f"erode_{k}"])


@pytest.mark.parametrize("name", list(inputs.contour_masks()))
def test_contours(name):
    mask = inputs.contour_masks()[name]
    want = G[f"contours_{name}"]  # cv2 order = reverse raster
    got = oracle.external_contours(mask)
    assert [g[0] for g in got] == [int(w[0]) for w in want][::-1]
    for g, w in zip(got, want[::-1]):
        assert abs(g[2] - w[1]) <= TOL and abs(g[3] - w[2]) <= TOL and abs(g[4] - w[3]) <= TOL
    fi, s00, s10, s01 = oracle.cell_moments(mask)  # the identity the CUDA path uses
    assert list(fi) == [g[0] for g in got]
    for i, w in enumerate(want[::-1]):
        assert s00[i] == round(2 * w[1]) and abs(s10[i] / 6 - w[2]) <= TOL and abs(s01[i] / 6 - w[3]) <= TOL
    valid, x, y, area = G[f"sift_{name}"]
    d = oracle.sift_contours(mask)
    assert bool(d.position_valid) == bool(valid) and abs(d.area - area) <= TOL
    assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL
    lab = oracle.label8(mask)
    assert len(np.unique(lab[lab >= 0])) == int(G[f"ncc_{name}"][0])


def test_bsub():
    b = oracle.Bsub(40, 56, 3, 0.0)
    for t, f in enumerate(inputs.bsub_frames()):
        assert np.array_equal(b.apply(f), G["bsub_a0"][t])


# ---- full size: the C restatement against the cv2 pins of the headline configurations -------------------
FULLSIZE = {"1080p_a001": (1080, 1920, 24, 0.01), "1080p_a0": (1080, 1920, 12, 0.0), "4k_a001": (2160, 3840, 5, 0.01)}


@pytest.mark.parametrize("name", list(FULLSIZE))
def test_oracle_matches_cv2_pins_at_full_size(name):
    """tests/golden/golden_fullsize.npz (cv2, the reference's call sequence, 1080p / 4K): the oracle reproduces the first
    frames' foreground mask, HSV frame, threshold mask (CRC32) and detection -- so the GPU-vs-oracle parity tests at
    full size stand on a pinned oracle."""
    import os
    import zlib

    G2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_fullsize.npz"))
    rows, cols, n, lr = FULLSIZE[name]
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**inputs.HSV_BAND)
    for t in range(n):
        d, eg = orc.track(inputs.synth_frame(rows, cols, inputs.SEED, t), lr, op)
        want = G2[f"{name}_crc"][t]
        assert zlib.crc32(np.ascontiguousarray(eg["fgmask"]).tobytes()) == want[0], t
        assert zlib.crc32(np.ascontiguousarray(eg["hsv"]).tobytes()) == want[1], t
        assert zlib.crc32(np.ascontiguousarray(eg["thresh"]).tobytes()) == want[2], t
        valid, x, y, area = G2[f"{name}_det"][t]
        assert bool(d.position_valid) == bool(valid)
        assert abs(d.x - x) <= 1e-6 and abs(d.y - y) <= 1e-6 and abs(d.area - area) <= 1e-6
