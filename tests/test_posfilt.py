"""posifilt kalman + posicom mean: the oracle against the golden vectors made with the
real cv::KalmanFilter and against a literal transcription of MeanPosition::combine; the CUDA epilogue (through
the C ABI) against the golden vectors and the oracle, stand-alone and fused behind the tracker.
fp64 throughout; tolerance 1e-9 relative (cv::solve(DECOMP_SVD) vs a closed-form 2x2 pseudo-inverse)."""
import math
import os

import numpy as np
import pytest

import oat_b200
import oracle
from golden_util import HERE, inputs

G = np.load(os.path.join(HERE, "golden", "posfilt_golden.npz"))
TOL = 1e-9


def close(a, b):
    return abs(a - b) <= TOL * max(1.0, abs(a), abs(b))


def check_row(p, row, t, seen_valid):
    valid, x, vx, y, vy = row
    assert bool(p.position_valid) == bool(valid) and bool(p.velocity_valid) == bool(valid), t
    if seen_valid:  # before the first prediction the reference's output is unspecified (and flagged invalid)
        assert close(p.x, x) and close(p.vx, vx) and close(p.y, y) and close(p.vy, vy), (t, p.x, x, p.y, y)


def py_mean(sources, anchor):
    """MeanPosition::combine transcribed literally (src/positioncombiner/MeanPosition.cpp:60-118)."""
    md = 1.0 / len(sources)
    o = dict(pv=True, vv=True, hv=True, x=0.0, y=0.0, vx=0.0, vy=0.0, hx=0.0, hy=0.0)
    for p in sources:
        if p.position_valid:
            o["x"] += md * p.x
            o["y"] += md * p.y
        else:
            o["pv"] = False
        if p.velocity_valid:
            o["vx"] += md * p.vx
            o["vy"] += md * p.vy
        else:
            o["vv"] = False
        if anchor >= 0:
            if o["pv"]:
                o["hx"] += p.x - sources[anchor].x
                o["hy"] += p.y - sources[anchor].y
            else:
                o["hv"] = False
        elif p.heading_valid:
            o["hx"] += p.hx
            o["hy"] += p.hy
        else:
            o["hv"] = False
    if o["hv"]:
        mag = math.sqrt(o["hx"] ** 2 + o["hy"] ** 2)
        o["hx"], o["hy"] = (o["hx"] / mag, o["hy"] / mag) if mag > 0 else (math.nan, math.nan)
    return o


def random_sources(rng, n, cls):
    out = []
    for _ in range(n):
        h = rng.normal(size=2)
        h /= np.linalg.norm(h)
        out.append(cls(position_valid=int(rng.random() > 0.15), velocity_valid=int(rng.random() > 0.15),
                       heading_valid=int(rng.random() > 0.15), x=rng.uniform(0, 640), y=rng.uniform(0, 480),
                       vx=rng.normal() * 20, vy=rng.normal() * 20, hx=h[0], hy=h[1]))
    return out


def same_mean(p, o):
    assert (bool(p.position_valid), bool(p.velocity_valid), bool(p.heading_valid)) == (o["pv"], o["vv"], o["hv"])
    for k in ("x", "y", "vx", "vy"):
        assert close(getattr(p, k), o[k]), k
    if o["hv"] and not math.isnan(o["hx"]):
        assert close(p.hx, o["hx"]) and close(p.hy, o["hy"])


# ---- CPU: the oracle is pinned ---------------------------------------------------------------
@pytest.mark.parametrize("name", list(inputs.KALMAN_CASES))
def test_oracle_kalman_matches_cv_kalmanfilter_golden(name):
    dt, timeout, sa, sn, seed, n = inputs.KALMAN_CASES[name]
    k = oracle.Kalman2D(dt, timeout, sa, sn)
    seen = False
    for t, (m, row) in enumerate(zip(inputs.kalman_track(seed, n), G[f"kalman_{name}"])):
        p = k.filter(*m)
        seen |= bool(row[0])
        check_row(p, row, t, seen)
    if name == "default":
        assert not G["kalman_default"][:, 0].any()  # timeout 0: KalmanFilter2D.cpp:113-117 never validates


def test_oracle_kalman_matches_cv2_live():
    cv2ref = pytest.importorskip("oracle.cv2ref")
    if not cv2ref.available():
        pytest.skip("cv2 not importable")
    rng = np.random.default_rng(5)
    for trial in range(10):
        dt, to = float(rng.choice([0.02, 0.005, 0.1])), float(rng.choice([0.0, 0.3, 2.0]))
        sa, sn = float(rng.choice([0.0, 5.0, 80.0])), float(rng.choice([0.0, 0.5, 4.0]))
        a, b = oracle.Kalman2D(dt, to, sa, sn), cv2ref.KalmanFilter2D(dt, to, sa, sn)
        seen = False
        for t, m in enumerate(inputs.kalman_track(100 + trial, 150)):
            row = b.filter(*m)
            seen |= bool(row[0])
            check_row(a.filter(*m), row, t, seen)


@pytest.mark.parametrize("anchor", [-1, 0, 2])
def test_oracle_mean_combine(anchor):
    rng = np.random.default_rng(7 + anchor)
    for trial in range(200):
        n = int(rng.integers(1, 6)) if anchor < 2 else int(rng.integers(3, 6))
        src = random_sources(rng, n, oracle.Position)
        same_mean(oracle.mean_combine(src, anchor), py_mean(src, anchor))


# ---- GPU: the CUDA epilogue ----------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(inputs.KALMAN_CASES))
def test_gpu_kalman_golden_and_oracle(ctx, name):
    dt, timeout, sa, sn, seed, n = inputs.KALMAN_CASES[name]
    f = oat_b200.PositionFilter(ctx, 1, dict(dt=dt, timeout=timeout, sigma_accel=sa, sigma_noise=sn))
    k = oracle.Kalman2D(dt, timeout, sa, sn)
    seen = False
    for t, (m, row) in enumerate(zip(inputs.kalman_track(seed, n), G[f"kalman_{name}"])):
        p = f.apply([oat_b200.Position(position_valid=int(m[0]), x=m[1], y=m[2])])[0]
        seen |= bool(row[0])
        check_row(p, row, t, seen)
        o = k.filter(*m)
        check_row(p, (o.position_valid, o.x, o.vx, o.y, o.vy), t, seen)
    f.reset()  # a reset filter replays the stream identically
    seen = False
    for t, (m, row) in enumerate(zip(inputs.kalman_track(seed, n), G[f"kalman_{name}"])):
        p = f.apply([oat_b200.Position(position_valid=int(m[0]), x=m[1], y=m[2])])[0]
        seen |= bool(row[0])
        check_row(p, row, t, seen)
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("anchor", [-1, 0, 2])
def test_gpu_mean_combine(ctx, anchor):
    rng = np.random.default_rng(70 + anchor)
    for n in ([1, 2, 3, 5, 8] if anchor < 2 else [3, 4, 8]):
        f = oat_b200.PositionFilter(ctx, n, None, combine_mean=True, heading_anchor=anchor)
        for trial in range(25):
            src = random_sources(rng, n, oat_b200.Position)
            osrc = [oracle.Position(**{k: getattr(s, k) for k, _ in oracle.Position._fields_}) for s in src]
            got = f.apply(src)
            same_mean(got, py_mean(src, anchor))
            o = oracle.mean_combine(osrc, anchor)
            assert (got.position_valid, got.velocity_valid, got.heading_valid) == (o.position_valid, o.velocity_valid, o.heading_valid)
        f.close()


@pytest.mark.gpu
def test_gpu_two_colour_graph(ctx):
    """examples/mouse-track/two-color-det.sh: two detectors -> kalman each -> mean with a heading anchor."""
    kp = dict(dt=0.02, timeout=0.5, sigma_accel=5.0, sigma_noise=1.0)
    f = oat_b200.PositionFilter(ctx, 2, kp, combine_mean=True, heading_anchor=0)
    ks = [oracle.Kalman2D(0.02, 0.5, 5.0, 1.0) for _ in range(2)]
    a, b = inputs.kalman_track(31, 200), inputs.kalman_track(32, 200)
    for t in range(200):
        got = f.apply([oat_b200.Position(position_valid=int(m[0]), x=m[1], y=m[2]) for m in (a[t], b[t])])
        want = oracle.mean_combine([k.filter(*m) for k, m in zip(ks, (a[t], b[t]))], 0)
        assert (got.position_valid, got.velocity_valid, got.heading_valid) == (want.position_valid, want.velocity_valid, want.heading_valid)
        if t > 10:
            for key in ("x", "y", "vx", "vy"):
                assert close(getattr(got, key), getattr(want, key)), (t, key)
            if want.heading_valid and not math.isnan(want.hx):
                assert close(got.hx, want.hx) and close(got.hy, want.hy)
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pipelined", [False, True])
def test_gpu_tracker_epilogue(ctx, pipelined):
    """The filter fused behind the tracker: every frame's detection feeds it on the device in frame order --
    also with tails of consecutive frames overlapping and with the first frame's tail overflowing (replay)."""
    rows, cols, n = 240, 320, 40
    band = dict(h=(40, 80), s=(100, 256), v=(100, 256))
    trk = oat_b200.Tracker(ctx, rows, cols, 0.01, oat_b200.HsvParams.make(**band))
    f = oat_b200.PositionFilter(ctx, 1, dict(dt=1 / 30, timeout=0.3, sigma_accel=20.0, sigma_noise=0.5))
    trk.attach_posfilt(f)
    orc, k = oracle.Tracker(rows, cols), oracle.Kalman2D(1 / 30, 0.3, 20.0, 0.5)
    frames = [oracle.synth_frame(rows, cols, 1000, t) for t in range(n)]
    bufs = [ctx.alloc(rows * cols * 3) for _ in range(n)]
    for b, fr in zip(bufs, frames):
        b.upload(fr)
    got = []
    if pipelined:
        depth = 4
        for t in range(n):
            if t >= depth:
                got.append(trk.collect_position())
            trk.submit(bufs[t])
        while len(got) < n:
            got.append(trk.collect_position())
    else:
        for t in range(n):
            trk.submit(bufs[t])
            got.append(trk.collect_position())
    seen = False
    for t, (d, p) in enumerate(got):
        o, _ = orc.track(frames[t], 0.01, oracle.HsvParams(**band))
        assert bool(d.position_valid) == bool(o.position_valid) and abs(d.x - o.x) < 1e-6 and abs(d.y - o.y) < 1e-6
        w = k.filter(bool(o.position_valid), o.x, o.y)
        seen |= bool(w.position_valid)
        check_row(p, (w.position_valid, w.x, w.vx, w.y, w.vy), t, seen)
    assert seen
    trk.attach_posfilt(None)
    f.close()
    trk.close()


@pytest.mark.gpu
def test_gpu_run_clip_matches_per_frame(ctx):
    """oat_tracker_run_clip (native pipelining loop) == frame-by-frame track(), detections and filtered positions."""
    rows, cols, n = 240, 320, 30
    band = dict(h=(40, 80), s=(100, 256), v=(100, 256))
    kp = dict(dt=1 / 30, timeout=0.3, sigma_accel=20.0, sigma_noise=0.5)
    bufs = [ctx.alloc(rows * cols * 3) for _ in range(n)]
    for t, b in enumerate(bufs):
        ctx.synth_frame(rows, cols, 1000, t, out=b)
    res = []
    for mode in ("clip", "frames"):
        trk = oat_b200.Tracker(ctx, rows, cols, 0.01, oat_b200.HsvParams.make(**band), ring_depth=4)
        f = oat_b200.PositionFilter(ctx, 1, kp)
        trk.attach_posfilt(f)
        if mode == "clip":
            dets, poss = trk.run_clip(bufs, depth=4, positions=True)
        else:
            dets, poss = [], []
            for b in bufs:
                trk.submit(b)
                d, p = trk.collect_position()
                dets.append(d)
                poss.append(p)
        res.append([(d.position_valid, d.x, d.y, d.area, p.position_valid, p.x, p.y, p.vx, p.vy) for d, p in zip(dets, poss)])
        trk.attach_posfilt(None)
        f.close()
        trk.close()
    assert res[0] == res[1]
    assert any(r[4] for r in res[0])
