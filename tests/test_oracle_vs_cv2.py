"""Pins the CPU oracle (oracle/oat_oracle.c) against the real OpenCV (cv2 wheel): the reference
delegates all arithmetic on this path to OpenCV and ships no golden vectors for it,
so the library the reference links is the pin.  CPU only; skipped if cv2 is not importable
(the committed fixtures in tests/golden/ then carry the pin, see test_golden.py)."""
import numpy as np
import pytest

import oracle
from oracle import cv2ref, synth

cv2 = pytest.importorskip("cv2")


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


def test_mog2_defaults():
    """SURVEY A1."""
    bs = cv2.createBackgroundSubtractorMOG2()
    p = oracle.default_mog_params()
    assert (bs.getHistory(), bs.getNMixtures()) == (p.history, p.nmixtures)
    assert bs.getVarThreshold() == p.var_threshold and bs.getVarThresholdGen() == p.var_threshold_gen
    assert np.float32(bs.getBackgroundRatio()) == np.float32(p.background_ratio)
    assert (bs.getVarInit(), bs.getVarMin(), bs.getVarMax()) == (p.var_init, p.var_min, p.var_max)
    assert np.float32(bs.getComplexityReductionThreshold()) == np.float32(p.ct)
    assert bs.getDetectShadows() and bs.getShadowValue() == p.shadow_value
    assert bs.getShadowThreshold() == p.shadow_threshold


@pytest.mark.parametrize("lr,sigma", [(0.0, 3.0), (0.01, 3.0), (0.1, 8.0), (-1.0, 8.0), (0.3, 20.0), (1.0, 0.0)])
def test_mog2_mask_bit_exact(lr, sigma):
    """the restatement reproduces cv2's mask bit for bit (incl. pruning, replacement, shadows)."""
    rows, cols = 96, 128
    bs = cv2.createBackgroundSubtractorMOG2()
    orc = oracle.Mog2(rows, cols)
    shadow_seen = fg_seen = 0
    for t, f in enumerate(noisy_stream(rows, cols, 40, sigma, seed=11)):
        want = bs.apply(f, None, lr)
        got = orc.apply(f, lr)
        assert np.array_equal(got, want), f"frame {t}: {(got != want).sum()} px differ"
        shadow_seen += int((want == 127).sum())
        fg_seen += int((want == 255).sum())
    assert shadow_seen > 0
    if 0 <= lr <= 0.1:
        assert fg_seen > 0
    if lr in (0.1, 0.3):
        assert orc.state()[0].max() >= 3  # several modes were live


def test_mog2_on_synthetic_stream():
    rows, cols = 240, 320
    for lr in (0.0, 0.01):
        bs = cv2.createBackgroundSubtractorMOG2()
        orc = oracle.Mog2(rows, cols)
        for t in range(20):
            f = oracle.synth_frame(rows, cols, 1000, t)
            assert np.array_equal(orc.apply(f, lr), bs.apply(f, None, lr))
        if lr == 0.0:
            assert (orc.state()[0] == 1).all()  # A3


def test_synth_twins():
    for t in (0, 1, 9):
        assert np.array_equal(oracle.synth_frame(90, 120, 1000, t), synth.frame(90, 120, 1000, t))


def test_bgr2hsv_all_colours():
    """SURVEY A7."""
    v = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(oracle.bgr2hsv(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV))


def test_inrange_semantics():
    """SURVEY A8."""
    ramp = np.arange(256, dtype=np.uint8)
    img = np.stack([ramp, ramp, ramp], -1).reshape(16, 16, 3)
    for lo, hi in [((0, 0, 0), (256, 256, 256)), ((40, 100, 100), (80, 256, 256)), ((256, 0, 0), (256, 256, 256)),
                   ((80, 0, 0), (40, 256, 256)), ((0, 10, 20), (255, 10, 20)), ((5, 5, 5), (5, 5, 5))]:
        assert np.array_equal(oracle.inrange3(img, lo, hi), cv2.inRange(img, lo, hi)), (lo, hi)


def blobs(rows, cols, n, rmax, seed):
    rng = np.random.default_rng(seed)
    m = np.zeros((rows, cols), np.uint8)
    yy, xx = np.mgrid[0:rows, 0:cols]
    for _ in range(n):
        cy, cx, r = rng.integers(0, rows), rng.integers(0, cols), rng.integers(1, rmax + 1)
        m[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
        if r > 3 and rng.random() < 0.5:
            m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 2) ** 2] = 0
            if r > 8 and rng.random() < 0.5:
                m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 4) ** 2] = 255
    return m


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 10, 11, 33, 50])
def test_morphology(k):
    """SURVEY A9 (window, anchor of even kernels, ignored border, in-place == out-of-place)."""
    m = blobs(120, 200, 10, 20, seed=k)
    m[0:3, 0:3] = 255
    m[60, 100] = 255
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (k, k))
    assert np.array_equal(oracle.dilate_rect(m, k), cv2.dilate(m, el))
    assert np.array_equal(oracle.erode_rect(m, k), cv2.erode(m, el))


def cv2_contour_list(mask):
    cs, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    out = []
    for c in cs:
        mm = cv2.moments(c)
        x, y = c[0][0]
        out.append((int(y) * mask.shape[1] + int(x), mm["m00"], mm["m10"], mm["m01"]))
    return out  # cv2 order


@pytest.mark.parametrize("seed", range(12))
def test_external_contours_and_moments(seed):
    """same contours (by first pixel), reverse raster order, same moments; and the
    2x2-cell identity the CUDA path uses equals cv::moments of the traced border."""
    rows, cols = [(64, 96), (33, 47), (100, 131), (90, 150)][seed % 4]
    m = blobs(rows, cols, 14, 14, seed) if seed < 8 else \
        ((np.random.default_rng(seed).random((rows, cols)) < [0.1, 0.4, 0.6, 0.9][seed - 8]).astype(np.uint8) * 255)
    want = cv2_contour_list(m)
    got = oracle.external_contours(m)
    assert [g[0] for g in got] == [w[0] for w in want][::-1]
    for g, w in zip(got, want[::-1]):
        assert abs(g[2] - w[1]) < 1e-9 and abs(g[3] - w[2]) < 1e-6 and abs(g[4] - w[3]) < 1e-6
    fi, s00, s10, s01 = oracle.cell_moments(m)
    assert list(fi) == [g[0] for g in got]
    for i, g in enumerate(got):
        assert s00[i] == round(2 * g[2]) and abs(s10[i] / 6.0 - g[3]) < 1e-6 and abs(s01[i] / 6.0 - g[4]) < 1e-6
    o = oracle.sift_contours(m)
    valid, x, y, area = cv2ref.sift_contours(m)
    assert bool(o.position_valid) == valid and abs(o.area - area) < 1e-9
    if valid:
        assert abs(o.x - x) < 1e-9 and abs(o.y - y) < 1e-9


def test_contour_facts():
    """SURVEY A11-A13."""
    z = np.zeros((80, 140), np.uint8)
    m = z.copy(); m[10:15, 10:17] = 255
    assert oracle.sift_contours(m).area == 24.0 == cv2ref.sift_contours(m)[3]
    m = z.copy(); m[10, 10] = 255
    assert not oracle.sift_contours(m).position_valid and not cv2ref.sift_contours(m)[0]
    m = z.copy(); m[10:12, 10:12] = 255; m[12:14, 12:14] = 255
    o = oracle.sift_contours(m)
    assert o.n_components == 1 and o.area == 2.0 == cv2ref.sift_contours(m)[3]
    m = z.copy(); m[10:40, 10:40] = 255; m[15:35, 15:35] = 0; m[22:28, 22:28] = 255
    o = oracle.sift_contours(m)
    assert o.n_components == 1 and o.area == 841.0 == cv2ref.sift_contours(m)[3]
    m = z.copy(); m[10:20, 10:20] = 255; m[10:20, 60:70] = 255; m[40:50, 30:40] = 255
    o = oracle.sift_contours(m)
    assert (o.x, o.y) == (34.5, 44.5) == cv2ref.sift_contours(m)[1:3]
    m = z.copy(); m[10:20, 10:20] = 255; m[10:20, 60:70] = 255  # same rows: right-hand one wins
    assert oracle.sift_contours(m).x == 64.5 == cv2ref.sift_contours(m)[1]


def test_bsub():
    """bit-exact for alpha = 0; alpha > 0 matches cv2's non-SIMD path
    (setUseOptimized(False)); the optimised path differs by last-bit rounding of the float background."""
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (40, 56, 3)).astype(np.uint8) for _ in range(10)]
    a, b = oracle.Bsub(40, 56, 3, 0.0), cv2ref.Bsub(0.0)
    for f in frames:
        assert np.array_equal(a.apply(f), b.filter(f))
    had = cv2.useOptimized()
    cv2.setUseOptimized(False)  # the plain C++ loop `dst = src*a + dst*(1-a)`; cv2's SIMD path rounds differently
    try:
        a, b = oracle.Bsub(40, 56, 3, 0.05), cv2ref.Bsub(0.05)
        for f in frames:
            assert np.array_equal(a.apply(f), b.filter(f))
    finally:
        cv2.setUseOptimized(had)


@pytest.mark.parametrize("lr", [0.0, 0.01])
def test_whole_chain_vs_cv2(lr):
    """mog -> col HSV -> hsv, oracle vs the reference's call sequence; known answer for -a 0 (A18)."""
    rows, cols = 240, 320
    band = dict(h=(40, 80), s=(100, 256), v=(100, 256))
    orc = oracle.Tracker(rows, cols)
    ref = cv2ref.Pipeline(lr, **band)
    for t in range(25):
        f = oracle.synth_frame(rows, cols, 1000, t)
        o, eg = orc.track(f, lr, oracle.HsvParams(**band))
        valid, x, y, area = ref.step(f)
        assert np.array_equal(eg["fgmask"], ref.mask) and np.array_equal(eg["hsv"], ref.hsv)
        assert np.array_equal(eg["thresh"], ref.det.thr)
        assert bool(o.position_valid) == valid and abs(o.area - area) < 1e-9
        assert abs(o.x - x) < 1e-9 and abs(o.y - y) < 1e-9
        if lr == 0.0 and t >= 1:
            cx, cy = synth.disc_centre(rows, cols, t)
            assert (o.x, o.y) == (cx + 0.5, cy + 0.5) and o.n_components == 1


def test_thresh_and_mask_restatements_match_cv2():
    """posidet thresh / framefilt thresh / framefilt mask: the numpy restatements against the cv2 calls the
    reference makes (SimpleThreshold.cpp:169-172, Threshold.cpp:67-81, FrameMasker.cpp:71-75)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    grey = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(oracle.bgr2grey(bgr), grey)
    # exhaustive check of the fixed-point grey formula on a colour cube slice
    cube = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(256), np.arange(0, 256, 3), indexing="ij"), -1).astype(np.uint8)
    assert np.array_equal(oracle.bgr2grey(cube), cv2.cvtColor(cube.reshape(-1, 1, 3), cv2.COLOR_BGR2GRAY).reshape(cube.shape[:3]))
    for lo, hi in [(0, 256), (100, 200), (200, 100), (256, 256), (0, 0), (255, 256), (37, 37)]:
        want = cv2.inRange(grey, lo, hi)
        assert np.array_equal(oracle.inrange1(grey, lo, hi), want), (lo, hi)
        t = bgr.copy()
        t[want == 0] = 0  # frame.setTo(0, thresh == 0)
        assert np.array_equal(oracle.threshold_filter(bgr, lo, hi), t)
    roi = (rng.random((64, 96)) < 0.6).astype(np.uint8) * 200
    assert np.array_equal(oracle.mask_filter(bgr, roi), cv2.bitwise_and(bgr, bgr, mask=roi))


def test_difference_detector_restatement_matches_cv2():
    """posidet diff: the numpy restatement against the reference's cv2 call sequence
    (src/positiondetector/DifferenceDetector.cpp:154-173 + DetectorFunc.cpp:31-66), incl. the even-kernel
    anchor / reflected border of cv::blur and the raw-first-frame quirk."""
    rng = np.random.default_rng(11)
    rows, cols = 60, 83
    for k in (0, 1, 2, 3, 4, 5, 8, 10, 15, 22):
        det = oracle.DifferenceDetector(diff_threshold=10, blur=k)
        last = None
        for t in range(6):
            grey = np.clip(rng.normal(100, 3, (rows, cols)), 0, 255).astype(np.uint8)
            if t == 0:
                grey[rng.random((rows, cols)) < 0.3] = 0  # so that the raw first frame is not one big blob
            y0, x0 = int(rng.integers(0, rows - 8)), int(rng.integers(0, cols - 8))
            grey[y0:y0 + 8, x0:x0 + 8] = 220
            if t % 2:
                grey[0:3, 0:2] = 250  # activity in the corner: the reflected border of an even box matters there
            # reference sequence
            if last is None:
                thr = grey.copy()
            else:
                thr = cv2.absdiff(grey, last)
                _, thr = cv2.threshold(thr, 10, 255, cv2.THRESH_BINARY)
                if k > 0:
                    thr = cv2.blur(thr, (k, k))
            last = grey.copy()
            want = cv2ref.sift_contours(thr)
            got, m = det.detect(grey)
            assert np.array_equal(m != 0, thr != 0), (k, t)
            assert bool(got.position_valid) == bool(want[0]), (k, t)
            if want[0]:
                assert abs(got.x - want[1]) < 1e-9 and abs(got.y - want[2]) < 1e-9 and abs(got.area - want[3]) < 1e-9, (k, t)


@pytest.mark.parametrize("shape", [(40, 64), (64, 96), (100, 131)])
def test_external_contours_on_structure_scenes(shape):
    """The geometries the band pre-labelling tests use on the GPU (tests/test_tail_model.py: rings, rings with a disc
    inside, U shapes, worms, speckle, shapes on the frame border): the oracle's contours against cv2's, so that what
    those tests compare with is pinned on exactly such masks."""
    from test_tail_model import _scene

    rows, cols = shape
    rng = np.random.default_rng(rows + cols)
    for trial in range(20):
        m = _scene(rng, rows, cols).astype(np.uint8) * 255
        if not m.any():
            continue
        want = cv2_contour_list(m)
        got = oracle.external_contours(m)
        assert [g[0] for g in got] == [w[0] for w in want][::-1], trial
        for g, w in zip(got, want[::-1]):
            assert abs(g[2] - w[1]) < 1e-9 and abs(g[3] - w[2]) < 1e-6 and abs(g[4] - w[3]) < 1e-6, trial
        o = oracle.sift_contours(m)
        valid, x, y, area = cv2ref.sift_contours(m)
        assert bool(o.position_valid) == valid and abs(o.area - area) < 1e-9, trial
        if valid:
            assert abs(o.x - x) < 1e-9 and abs(o.y - y) < 1e-9, trial
