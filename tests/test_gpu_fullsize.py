"""The HEADLINE configurations at full size against OpenCV itself.

* committed pins (tests/golden/golden_fullsize.npz, made by tests/golden/make_golden_fullsize.py with the cv2 wheel
  running the reference's call sequence): per-frame CRC32 of the foreground mask, the HSV frame and the
  post-morphology threshold mask + the detection, 1080p -a 0.01 (240 frames), 1080p -a 0 (120 frames), 4K -a 0.01
  (40 frames) -- the CUDA path must reproduce every one of them through the C ABI;
* live, when cv2 is importable on the box: a different stream seed through cv2ref.Pipeline and the CUDA path side
  by side (masks compared byte for byte, not by CRC);
* the resident engine on the same frames: its detections must equal the pinned ones too."""
import os
import zlib

import numpy as np
import pytest

import oat_b200
from golden_util import inputs

pytestmark = pytest.mark.gpu
TOL = 1e-6
HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_fullsize.npz"))
FULLSIZE = {"1080p_a001": (1080, 1920, 240, 0.01), "1080p_a0": (1080, 1920, 120, 0.0), "4k_a001": (2160, 3840, 40, 0.01)}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


@pytest.mark.parametrize("name", list(FULLSIZE))
def test_headline_config_matches_cv2_pins(ctx, name):
    rows, cols, n, lr = FULLSIZE[name]
    hp = oat_b200.HsvParams.make(**inputs.HSV_BAND)
    trk = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    buf = ctx.alloc(rows * cols * 3)
    bufs = []
    for t in range(n):
        ctx.synth_frame(rows, cols, inputs.SEED, t, out=buf)  # bit-identical to inputs.synth_frame (test_synth_generator_matches_oracle)
        d, eg = trk.track(buf, egress=("fgmask", "hsv", "thresh"))
        want = G[f"{name}_crc"][t]
        assert crc(eg["fgmask"]) == want[0], f"{name}: foreground mask differs from cv2 at t={t}"
        assert int((eg["fgmask"] == 255).sum()) == int(G[f"{name}_nfg"][t])
        assert crc(eg["hsv"]) == want[1], f"{name}: HSV frame differs from cv2 at t={t}"
        assert crc(eg["thresh"]) == want[2], f"{name}: threshold mask differs from cv2 at t={t}"
        valid, x, y, area = G[f"{name}_det"][t]
        assert bool(d.position_valid) == bool(valid), t
        assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL and abs(d.area - area) <= TOL, t
    trk.close()
    # the same stream as a device-resident clip through the resident engine (one launch per 32 frames)
    m = min(n, 96)
    for t in range(m):
        b = ctx.alloc(rows * cols * 3)
        ctx.synth_frame(rows, cols, inputs.SEED, t, out=b)
        bufs.append(b)
    trk = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=64)
    dets = trk.run_clip(bufs)
    assert trk.tail_stats()["clip_frames"] == m - 1
    for t in range(m):
        valid, x, y, area = G[f"{name}_det"][t]
        d = dets[t]
        assert bool(d.position_valid) == bool(valid), t
        assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL and abs(d.area - area) <= TOL, t
    trk.close()
    for b in bufs:
        b.free()
    buf.free()


@pytest.mark.parametrize("lr", [0.01, 0.0])
def test_headline_config_live_against_cv2(ctx, lr):
    cv2ref = pytest.importorskip("oracle.cv2ref")
    if not cv2ref.available():
        pytest.skip("cv2 is not importable on this box")
    rows, cols, n, seed = 1080, 1920, 200, 4321
    hp = oat_b200.HsvParams.make(**inputs.HSV_BAND)
    trk = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    pipe = cv2ref.Pipeline(lr, **inputs.HSV_BAND)
    for t in range(n):
        f = ctx.synth_frame(rows, cols, seed, t)
        d, eg = trk.track(f, egress=("fgmask", "thresh", "bgr"))
        valid, x, y, area = pipe.step(f)
        assert np.array_equal(eg["fgmask"], pipe.mask), f"foreground mask differs from cv2 at t={t}"
        assert np.array_equal(eg["bgr"], pipe.filt), f"filtered frame differs from cv2 at t={t}"
        assert np.array_equal(eg["thresh"], pipe.det.thr), f"threshold mask differs from cv2 at t={t}"
        assert bool(d.position_valid) == bool(valid)
        assert abs(d.x - x) <= TOL and abs(d.y - y) <= TOL and abs(d.area - area) <= TOL
    trk.close()
