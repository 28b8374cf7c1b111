"""GPU parity: BGR->HSV, inRange, erode/dilate, labelling and contour moments (through the C ABI)
vs the CPU oracle.  Reference: src/framefilter/ColorConvert.cpp:101-107,
src/positiondetector/HSVDetector.cpp:142-173, src/positiondetector/DetectorFunc.cpp:31-66."""
import numpy as np
import pytest

import oat_b200
import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-6  # (x, y) and area within 1e-6 of cv::moments(contour)


def check_detection(d, o, ctx_msg=""):
    assert bool(d.position_valid) == bool(o.position_valid), ctx_msg
    assert d.n_components == o.n_components, ctx_msg
    assert abs(d.area - o.area) <= TOL, ctx_msg
    if o.position_valid:
        assert abs(d.x - o.x) <= TOL and abs(d.y - o.y) <= TOL, f"{ctx_msg}: ({d.x},{d.y}) vs ({o.x},{o.y})"


def test_bgr2hsv_all_colours(ctx):
    """All 2^24 BGR triples, bit-exact."""
    v = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    got = oat_b200.color_convert_hsv(ctx, bgr)
    want = oracle.bgr2hsv(bgr)
    assert np.array_equal(got, want)


def blobs(rows, cols, n, rmax, seed, holes=True):
    rng = np.random.default_rng(seed)
    m = np.zeros((rows, cols), np.uint8)
    yy, xx = np.mgrid[0:rows, 0:cols]
    for _ in range(n):
        cy, cx, r = rng.integers(0, rows), rng.integers(0, cols), rng.integers(1, rmax + 1)
        m[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 255
        if holes and r > 3 and rng.random() < 0.5:
            m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 2) ** 2] = 0
            if r > 8 and rng.random() < 0.5:  # island inside the hole
                m[(yy - cy) ** 2 + (xx - cx) ** 2 <= (r // 4) ** 2] = 255
    return m


def run_sift(ctx, mask, erode=0, dilate=0, area=(0.0, oat_b200.DBL_MAX)):
    rows, cols = mask.shape
    det = oat_b200.HSVDetector(ctx, rows, cols, oat_b200.HsvParams.make(erode=erode, dilate=dilate, area=area))
    d, thr, lab = det.sift_contours(mask, want_thresh=True, want_labels=True)  # labels: the unbounded multi-launch tail
    d_fast, thr_fast, _ = det.sift_contours(mask, want_thresh=True)            # no labels: the one-launch tail (or its replay)
    det.close()
    # oracle: same morphology, then sift
    om = (mask != 0).astype(np.uint8) * 255
    if erode > 0:
        om = oracle.erode_rect(om, erode)
    if dilate > 0:
        om = oracle.dilate_rect(om, dilate)
    o = oracle.sift_contours(om, area[0], area[1])
    assert np.array_equal(thr, om), "post-morphology mask differs"
    assert np.array_equal(lab, oracle.label8(om)), "component labels differ"
    check_detection(d, o)
    assert np.array_equal(thr_fast, om), "post-morphology mask of the one-launch tail differs"
    check_detection(d_fast, o)
    return d, o


@pytest.mark.parametrize("shape", [(64, 96), (33, 47), (100, 131), (7, 300), (240, 320)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_sift_random_blobs(ctx, shape, seed):
    run_sift(ctx, blobs(shape[0], shape[1], 12, 14, seed))


@pytest.mark.parametrize("density", [0.02, 0.2, 0.5, 0.8, 0.98])
def test_sift_random_noise(ctx, density):
    rng = np.random.default_rng(int(density * 100))
    m = (rng.random((90, 150)) < density).astype(np.uint8) * 255
    run_sift(ctx, m)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 10, 11, 33, 50])
def test_morphology_kernel_sizes(ctx, k):
    m = blobs(120, 200, 10, 20, seed=k)
    m[0:3, 0:3] = 255  # corner (border is ignored)
    m[60, 100] = 255   # impulse
    run_sift(ctx, m, dilate=k)
    run_sift(ctx, m, erode=k)
    run_sift(ctx, m, erode=max(1, k // 2), dilate=k)


def test_sift_structures(ctx):
    rows, cols = 80, 140
    cases = {}
    z = np.zeros((rows, cols), np.uint8)
    cases["empty"] = z.copy()
    cases["full"] = np.full((rows, cols), 255, np.uint8)
    m = z.copy(); m[10, 10] = 255
    cases["single_pixel"] = m
    m = z.copy(); m[20, 5:100] = 255; m[5:70, 120] = 255
    cases["lines"] = m
    m = z.copy(); m[10:15, 10:17] = 255
    cases["rect_5x7"] = m  # m00 = 24
    m = z.copy(); m[10:12, 10:12] = 255; m[12:14, 12:14] = 255
    cases["diag_blocks"] = m  # one contour, m00 = 2 (A11)
    m = z.copy(); m[10:40, 10:40] = 255; m[15:35, 15:35] = 0; m[22:28, 22:28] = 255
    cases["ring_with_island"] = m  # one external contour (A12)
    m = z.copy(); m[5:75, 5:135] = 255; m[10:70, 10:130] = 0; m[15:65, 15:125] = 255; m[20:60, 20:120] = 0
    m[30:50, 40:100] = 255; m[35:45, 50:90] = 0
    cases["nested_rings"] = m
    m = z.copy(); m[10:20, 10:20] = 255; m[10:20, 60:70] = 255; m[40:50, 30:40] = 255
    cases["area_tie"] = m  # raster-last wins (A13)
    m = z.copy(); m[0:10, 0:10] = 255; m[rows - 6:rows, cols - 9:cols] = 255; m[0:4, 60:90] = 255
    cases["touching_borders"] = m
    m = np.full((rows, cols), 255, np.uint8); m[20:30, 20:30] = 0; m[0, 50] = 0; m[40:42, cols - 1] = 0
    cases["full_with_holes_and_border_notches"] = m
    m = z.copy(); m[::2, ::2] = 255
    cases["isolated_grid"] = m
    m = z.copy(); yy, xx = np.mgrid[0:rows, 0:cols]; m[(yy + xx) % 2 == 0] = 255
    cases["checkerboard"] = m  # one 8-connected component, every background pixel a separate hole or exterior
    m = z.copy(); m[30:34, 28:36] = 255; m[31:33, 31:33] = 0  # hole straddling a word boundary
    cases["word_boundary_hole"] = m
    m = z.copy(); m[5:60, 20:110] = 255; m[10:55, 25:105] = 0; m[12:20, 63:66] = 255; m[30:50, 30:100] = 255; m[35:45, 60:70] = 0
    cases["islands_in_big_hole"] = m
    for name, mask in cases.items():
        d, o = run_sift(ctx, mask)
        if name == "rect_5x7":
            assert d.area == 24.0 and d.x == 13.0 and d.y == 12.0
        if name == "diag_blocks":
            assert d.n_components == 1 and d.area == 2.0
        if name == "ring_with_island":
            assert d.n_components == 1 and d.area == 841.0
        if name == "area_tie":
            assert d.n_components == 3 and (d.x, d.y) == (34.5, 44.5)
        if name in ("empty", "single_pixel", "lines", "isolated_grid"):
            assert not d.position_valid


def test_sift_area_gate(ctx):
    m = np.zeros((60, 100), np.uint8)
    m[5:10, 5:10] = 255      # area 16
    m[20:40, 20:50] = 255    # area 19*29 = 551
    m[45:55, 60:75] = 255    # area 9*14 = 126
    d, _ = run_sift(ctx, m)
    assert d.area == 551.0
    d, _ = run_sift(ctx, m, area=(0.0, 551.0))  # max is exclusive
    assert d.area == 126.0
    d, _ = run_sift(ctx, m, area=(126.0, 127.0))  # min is inclusive
    assert d.area == 126.0
    d, _ = run_sift(ctx, m, area=(600.0, 700.0))
    assert not d.position_valid and d.n_components == 3


@pytest.mark.parametrize("shape", [(48, 64), (61, 77)])
def test_hsv_detect_random_frames(ctx, shape):
    rows, cols = shape
    rng = np.random.default_rng(rows)
    hsv = rng.integers(0, 256, (rows, cols, 3)).astype(np.uint8)
    hsv[..., 0] %= 180
    hsv[10:30, 10:40] = (60, 200, 210)
    for (h, s, v, e, dl) in [((40, 80), (100, 256), (100, 256), 0, 10), ((0, 256), (0, 256), (0, 256), 0, 10),
                             ((50, 70), (0, 255), (128, 256), 3, 5), ((80, 40), (0, 256), (0, 256), 0, 0),
                             ((0, 90), (256, 256), (0, 256), 0, 3), ((0, 89), (0, 127), (0, 256), 2, 0)]:
        p = oat_b200.HsvParams.make(h=h, s=s, v=v, erode=e, dilate=dl)
        det = oat_b200.HSVDetector(ctx, rows, cols, p)
        d, thr, lab = det.detect(hsv, want_thresh=True, want_labels=True)
        det.close()
        o, othr = oracle.hsv_detect(hsv, oracle.HsvParams(h, s, v, e, dl))
        assert np.array_equal(thr, othr)
        assert np.array_equal(lab, oracle.label8(othr))
        check_detection(d, o, f"{h}{s}{v} e{e} d{dl}")


def test_bsub_parity(ctx):
    rows, cols = 40, 56
    rng = np.random.default_rng(5)
    for alpha in (0.0, 0.05, 0.5):
        for ch in (3, 1):
            shape = (rows, cols, 3) if ch == 3 else (rows, cols)
            gpu = oat_b200.BackgroundSubtractor(ctx, rows, cols, ch, alpha)
            orc = oracle.Bsub(rows, cols, ch, alpha)
            for t in range(12):
                f = rng.integers(0, 256, shape).astype(np.uint8)
                assert np.array_equal(gpu.filter(f), orc.apply(f)), f"alpha={alpha} ch={ch} t={t}"
            gpu.close()


@pytest.mark.parametrize("band", [(0, 256), (100, 200), (200, 100), (256, 256), (180, 256)])
@pytest.mark.parametrize("morph", [(0, 0), (0, 10), (3, 5)])
def test_posidet_thresh_parity(ctx, band, morph):
    """posidet thresh (src/positiondetector/SimpleThreshold.cpp:169-182): GREY inRange -> erode -> dilate -> siftContours."""
    rows, cols = 96, 140
    rng = np.random.default_rng(band[0] * 7 + morph[1])
    grey = np.clip(rng.normal(90, 30, (rows, cols)), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[:rows, :cols]
    for cy, cx, r in [(30, 40, 14), (60, 100, 9), (80, 20, 6)]:
        grey[(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 230
    det = oat_b200.HSVDetector(ctx, rows, cols, oat_b200.HsvParams.make(erode=morph[0], dilate=morph[1]))
    d, thr, lab = det.thresh_detect(grey, band[0], band[1], want_thresh=True, want_labels=True)
    o, om = oracle.thresh_detect(grey, band[0], band[1], oracle.HsvParams(erode=morph[0], dilate=morph[1]))
    assert np.array_equal(thr, om)
    assert np.array_equal(lab, oracle.label8(om))
    check_detection(d, o)
    with pytest.raises(oat_b200.OatError):
        det.thresh_detect(grey, 0, 300)
    det.close()


@pytest.mark.parametrize("shape", [(48, 64, 3), (37, 53, 3), (40, 72)])
def test_framefilt_thresh_and_mask_parity(ctx, shape):
    """framefilt thresh (Threshold.cpp:67-81) and framefilt mask (FrameMasker.cpp:71-75): bit-exact."""
    rng = np.random.default_rng(shape[1])
    frame = rng.integers(0, 256, shape, dtype=np.uint8)
    for lo, hi in [(0, 256), (64, 192), (192, 64), (256, 256), (0, 0)]:
        assert np.array_equal(oat_b200.threshold_filter(ctx, frame, lo, hi), oracle.threshold_filter(frame, lo, hi)), (lo, hi)
    roi = (rng.random(shape[:2]) < 0.5).astype(np.uint8) * 255
    assert np.array_equal(oat_b200.mask_filter(ctx, frame, roi), oracle.mask_filter(frame, roi))


@pytest.mark.parametrize("shape,n,rmax,dilate", [((1080, 1920), 6, 90, 10), ((1080, 1920), 60, 40, 0), ((2160, 3840), 5, 300, 10),
                                                 ((720, 1000), 25, 60, 4), ((1080, 1920), 400, 12, 0)])
def test_sift_full_size_masks(ctx, shape, n, rmax, dilate):
    """Frame-sized masks with holes, islands in holes and many blobs: one-launch tail (wide and narrow regions,
    the run table filling up, overflow -> replay) and the multi-launch tail, both against the oracle."""
    run_sift(ctx, blobs(shape[0], shape[1], n, rmax, seed=n + rmax), dilate=dilate)


@pytest.mark.parametrize("blur", [0, 1, 2, 3, 4, 7, 10, 22])
@pytest.mark.parametrize("shape", [(60, 83), (64, 128)])
def test_posidet_diff_parity(ctx, blur, shape):
    """posidet diff (src/positiondetector/DifferenceDetector.cpp:118-173): absdiff -> threshold -> blur -> siftContours,
    with the raw-first-frame quirk and cv::blur's even-box reflected border (activity in the corner)."""
    rows, cols = shape
    rng = np.random.default_rng(blur * 31 + rows)
    det = oat_b200.DifferenceDetector(ctx, rows, cols, diff_threshold=10, blur=blur)
    orc = oracle.DifferenceDetector(diff_threshold=10, blur=blur)
    for t in range(7):
        grey = np.clip(rng.normal(100, 3, (rows, cols)), 0, 255).astype(np.uint8)
        if t == 0:
            grey[rng.random((rows, cols)) < 0.3] = 0
        y0, x0 = int(rng.integers(0, rows - 8)), int(rng.integers(0, cols - 8))
        grey[y0:y0 + 8, x0:x0 + 8] = 220
        if t % 2:
            grey[0:3, 0:2] = 250
        d, thr = det.detect(grey, want_thresh=True)
        o, om = orc.detect(grey)
        assert np.array_equal(thr, om), f"sifted mask differs at t={t}: {(thr != om).sum()} px"
        check_detection(d, o, f"t={t}")
    det.close()
    with pytest.raises(oat_b200.OatError):
        oat_b200.DifferenceDetector(ctx, rows, cols, blur=23).detect(np.zeros((rows, cols), np.uint8))
