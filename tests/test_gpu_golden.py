"""The CUDA path (through the C ABI) against the committed golden vectors made with the real
OpenCV -- no oracle in the loop."""
import numpy as np
import pytest

import oat_b200
from golden_util import crc, inputs, load

pytestmark = pytest.mark.gpu
G = load()
TOL = 1e-6


@pytest.mark.parametrize("name", list(inputs.MOG_STREAMS))
def test_mog_masks(ctx, name):
    rows, cols, n, sigma, seed, lr = inputs.MOG_STREAMS[name]
    mog = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols)
    for t, f in enumerate(inputs.noisy_stream(rows, cols, n, sigma, seed)):
        out, mask = mog.apply(f, learning_rate=lr)
        want = G[f"mog_{name}"][t]
        assert np.array_equal(mask, want), f"{name} frame {t}"
        ref = f.copy()
        ref[want == 0] = 0
        assert np.array_equal(out, ref)
    mog.close()


@pytest.mark.parametrize("name", list(inputs.CHAINS))
def test_chain(ctx, name):
    rows, cols, n, lr = inputs.CHAINS[name]
    trk = oat_b200.Tracker(ctx, rows, cols, lr, oat_b200.HsvParams.make(**inputs.HSV_BAND))
    for t in range(n):
        f = inputs.synth_frame(rows, cols, inputs.SEED, t)
        assert np.array_equal(ctx.synth_frame(rows, cols, inputs.SEED, t), f)
        d, eg = trk.track(f, egress=("fgmask", "hsv", "thresh"))
        assert np.array_equal(eg["fgmask"], G[f"chain_{name}_fg"][t])
        assert np.array_equal(eg["thresh"], G[f"chain_{name}_thr"][t])
        assert crc(eg["hsv"]) == G[f"chain_{name}_hsvcrc"][t]
        valid, x, y, area = G[f"chain_{name}_det"][t]
        assert bool(d.position_valid) == bool(valid)
        assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL and abs(d.area - area) <= TOL
    trk.close()


def test_hsv_inrange_morph(ctx):
    assert np.array_equal(oat_b200.color_convert_hsv(ctx, inputs.hsv_colours()), G["hsv_out"])
    img = inputs.hsv_image()
    rows, cols = img.shape[:2]
    for i, (lo, hi) in enumerate(inputs.INRANGE_CASES):
        p = oat_b200.HsvParams.make(h=(lo[0], hi[0]), s=(lo[1], hi[1]), v=(lo[2], hi[2]), erode=0, dilate=0)
        det = oat_b200.HSVDetector(ctx, rows, cols, p)
        _, thr, _ = det.detect(img, want_thresh=True)
        det.close()
        assert np.array_equal(thr, G[f"inrange_{i}"])
    m = inputs.morph_mask()
    for k in inputs.MORPH_K:
        for op in ("dilate", "erode"):
            p = oat_b200.HsvParams.make(erode=k if op == "erode" else 0, dilate=k if op == "dilate" else 0)
            det = oat_b200.HSVDetector(ctx, m.shape[0], m.shape[1], p)
            _, thr, _ = det.sift_contours(m, want_thresh=True)
            det.close()
            assert np.array_equal(thr, G[f"{op}_{k}"]), f"{op} {k}"


@pytest.mark.parametrize("name", list(inputs.contour_masks()))
def test_contours(ctx, name):
    mask = inputs.contour_masks()[name]
    det = oat_b200.HSVDetector(ctx, mask.shape[0], mask.shape[1], oat_b200.HsvParams.make(dilate=0))
    d, _, lab = det.sift_contours(mask, want_labels=True)
    det.close()
    valid, x, y, area = G[f"sift_{name}"]
    assert d.n_components == len(G[f"contours_{name}"])
    assert bool(d.position_valid) == bool(valid) and abs(d.area - area) <= TOL
    assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL
    assert (lab >= 0).sum() == (mask != 0).sum()
    assert len(np.unique(lab[lab >= 0])) == int(G[f"ncc_{name}"][0])
    firsts = sorted(int(w[0]) for w in G[f"contours_{name}"])
    assert set(firsts) <= set(np.unique(lab[lab >= 0]).tolist())  # every external contour starts a component


def test_bsub(ctx):
    b = oat_b200.BackgroundSubtractor(ctx, 40, 56, 3, 0.0)
    for t, f in enumerate(inputs.bsub_frames()):
        assert np.array_equal(b.filter(f), G["bsub_a0"][t])
    b.close()
