"""Seeded randomised parity of the whole fused path against the CPU oracle: random geometry (tight and
ragged widths, so both tile orders of the pipelined kernel and the generic kernel run), learning rates
(frozen / live / automatic / reset), noise levels (one to five live modes, pruning, replacement), HSV
bands, erode/dilate sizes and area gates, host and device-resident frames, sync and pipelined calls.
Everything must match bit for bit (masks) / to 1e-6 (moments)."""
import numpy as np
import pytest

import oat_b200
import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-6


def stream(rng, rows, cols, n, sigma, nblobs):
    bg = rng.integers(20, 200, (rows, cols, 3)).astype(np.float32)
    pos = rng.uniform(0, 1, (nblobs, 2))
    vel = rng.uniform(-0.05, 0.05, (nblobs, 2))
    rad = rng.integers(2, max(3, min(rows, cols) // 5), nblobs)
    col = rng.integers(0, 256, (nblobs, 3))
    yy, xx = np.mgrid[:rows, :cols]
    for t in range(n):
        f = bg + (rng.normal(0, sigma, bg.shape) if sigma > 0 else 0)
        for b in range(nblobs):
            if t == 0:
                continue
            cy, cx = (pos[b] + vel[b] * t) % 1.0 * (rows, cols)
            f[(yy - cy) ** 2 + (xx - cx) ** 2 <= rad[b] ** 2] = col[b]
        yield np.clip(np.rint(f), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("seed", range(120))
def test_fuzz_tracker(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    rows = int(rng.integers(24, 160))
    cols = int(rng.choice([32, 64, 96, 128, 160, 256, 320, 40, 100, 72, 37, 150, 33]))
    lr = float(rng.choice([0.0, 0.0, 0.01, 0.05, 0.2, 0.5, -1.0]))
    sigma = float(rng.choice([0.0, 2.0, 6.0, 15.0]))
    band = dict(h=tuple(sorted(rng.integers(0, 257, 2).tolist())), s=(int(rng.integers(0, 120)), 256), v=(int(rng.integers(0, 120)), 256))
    erode, dilate = int(rng.choice([0, 0, 2, 3])), int(rng.choice([0, 3, 4, 10]))
    area = (float(rng.choice([0.0, 4.0, 30.0])), float(rng.choice([oat_b200.DBL_MAX, 500.0, 5000.0])))
    hp = oat_b200.HsvParams.make(erode=erode, dilate=dilate, area=area, **band)
    op = oracle.HsvParams(erode=erode, dilate=dilate, area=area, **band)
    use_lr = max(lr, 0.0)
    trk = oat_b200.Tracker(ctx, rows, cols, adaptation_coeff=use_lr, hsv=hp, ring_depth=3)
    orc = oracle.Tracker(rows, cols)
    dev = ctx.alloc(rows * cols * 3)
    frames = list(stream(rng, rows, cols, 18, sigma, int(rng.integers(1, 4))))
    pending = []
    for t, f in enumerate(frames):
        o, oeg = orc.track(f, lr, op)
        mode = t % 3
        if mode == 0:  # synchronous, host frame, every egress
            d, eg = trk.track(f, egress=("bgr", "fgmask", "hsv", "thresh"), learning_rate=lr)
            for k in ("fgmask", "bgr", "hsv", "thresh"):
                assert np.array_equal(eg[k], oeg[k]), f"seed {seed} t={t}: {k} differs ({(eg[k] != oeg[k]).sum()} px)"
            got = [d]
        elif mode == 1:  # pipelined, device-resident frame (two frames in flight with the next one)
            dev2 = ctx.alloc(rows * cols * 3).upload(f)
            trk.submit(dev2, learning_rate=lr)
            pending.append((o, dev2))
            continue
        else:  # pipelined, host frame
            trk.submit(f, learning_rate=lr)
            pending.append((o, None))
            got = []
            want = []
            for po, _ in pending:
                got.append(trk.collect())
                want.append(po)
            pending = []
            for d, po in zip(got, want):
                assert bool(d.position_valid) == bool(po.position_valid), f"seed {seed} t={t}"
                assert d.n_components == po.n_components
                assert abs(d.x - po.x) <= TOL and abs(d.y - po.y) <= TOL and abs(d.area - po.area) <= TOL
            continue
        d = got[0]
        assert bool(d.position_valid) == bool(o.position_valid), f"seed {seed} t={t}"
        assert d.n_components == o.n_components
        assert abs(d.x - o.x) <= TOL and abs(d.y - o.y) <= TOL and abs(d.area - o.area) <= TOL
    m, w, v, mu = trk.state()
    om, ow, ov, omu = orc.mog.state()
    assert np.array_equal(m, om), f"seed {seed}: live-mode counts differ"
    live = np.arange(w.shape[2])[None, None, :] < om[:, :, None]
    assert np.array_equal(w.view(np.uint32)[live], ow.view(np.uint32)[live])
    assert np.array_equal(v.view(np.uint32)[live], ov.view(np.uint32)[live])
    assert np.array_equal(mu.view(np.uint32)[np.broadcast_to(live[..., None], mu.shape)],
                          omu.view(np.uint32)[np.broadcast_to(live[..., None], omu.shape)])
    trk.close()
