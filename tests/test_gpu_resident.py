"""The resident clip engine (oat_tracker_run_clip / oat_tracker_run_clips on device-resident frames): ONE launch of
the fused kernel + ONE launch of the tail server per chunk of frames, frame-to-frame ordering tile by tile on
the device.  Everything it produces -- detections, filtered positions, the whole GMM state -- must be bit-identical
to the frame-by-frame synchronous path (which tests/test_gpu_tracker.py pins to the oracle), for every geometry
class, for frames smaller than one CTA's pipeline, across chunk and launch boundaries, and over long runs."""
import os

import numpy as np
import pytest

import oat_b200
import oracle

pytestmark = pytest.mark.gpu
HSV_BAND = dict(h=(40, 80), s=(100, 256), v=(100, 256))
TOL = 1e-6


def _det(d):
    return (d.position_valid, d.n_components, d.x, d.y, d.area)


def _frames(ctx, rows, cols, seed, n, pitch=None):
    """n synthetic frames of one stream in device memory (t = 0..n-1); pitch: row pitch in bytes (default tight)."""
    bufs = []
    for t in range(n):
        if pitch is None:
            b = ctx.alloc(rows * cols * 3)
            ctx.synth_frame(rows, cols, seed, t, out=b)
        else:
            f = oracle.synth_frame(rows, cols, seed, t)
            padded = np.zeros((rows, pitch), np.uint8)
            padded[:, :cols * 3] = f.reshape(rows, cols * 3)
            b = ctx.alloc(rows * pitch)
            b.upload(padded)
        bufs.append(b)
    return bufs


def _state_equal(a, b):
    """Mode counts and every LIVE mode (weight, variance, mean) bit-identical; dead slots are unspecified."""
    am, aw, av, amu = a.state()
    bm, bw, bv, bmu = b.state()
    assert np.array_equal(am, bm)
    live = np.arange(aw.shape[2])[None, None, :] < am[:, :, None]
    assert np.array_equal(aw.view(np.uint32)[live], bw.view(np.uint32)[live])
    assert np.array_equal(av.view(np.uint32)[live], bv.view(np.uint32)[live])
    amu, bmu = amu.reshape(am.shape + (aw.shape[2], 3)), bmu.reshape(bm.shape + (bw.shape[2], 3))
    assert np.array_equal(amu.view(np.uint32)[live], bmu.view(np.uint32)[live])


@pytest.mark.parametrize("shape", [(120, 160), (240, 320), (64, 64), (20, 32), (480, 640)])
@pytest.mark.parametrize("lr", [0.0, 0.02])
@pytest.mark.parametrize("ring", [2, 8, 64])
def test_clip_engine_equals_frame_by_frame(ctx, shape, lr, ring):
    """Small frames: fewer tiles than CTAs, frames smaller than one CTA's ring of stages (a CTA's next tile depends on a
    tile in its own pipeline), chunks of 1 / 4 / 32 frames."""
    rows, cols = shape
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    n = 41
    bufs = _frames(ctx, rows, cols, 1000, n)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=ring)
    got = [_det(d) for d in a.run_clip(bufs, depth=4)]
    assert a.tail_stats()["clip_frames"] == n - 1, "the resident engine did not take the clip"
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in bufs]
    assert got == want
    _state_equal(a, b)
    # and against the oracle directly
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**HSV_BAND)
    for t in range(n):
        o, _ = orc.track(oracle.synth_frame(rows, cols, 1000, t), lr, op)
        d = got[t]
        assert bool(d[0]) == bool(o.position_valid) and d[1] == o.n_components
        assert abs(d[2] - o.x) <= TOL and abs(d[3] - o.y) <= TOL and abs(d[4] - o.area) <= TOL
    assert a.live_modes() == int(orc.mog.state()[0].sum())
    a.close()
    b.close()


@pytest.mark.parametrize("cols", [80, 144, 1008])
def test_clip_engine_ragged_widths(ctx, cols):
    """cols % 32 != 0 with 16-byte aligned rows: the producer lane stages the BGR bytes with one bulk copy per row
    segment (no per-thread loads), dynamic scheduler and all."""
    rows, lr, n = 90, 0.03, 19
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    assert (3 * cols) % 16 == 0 and cols % 32 != 0
    bufs = _frames(ctx, rows, cols, 1003, n)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=8)
    got = [_det(d) for d in a.run_clip(bufs, depth=4)]
    assert a.tail_stats()["clip_frames"] == n - 1
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**HSV_BAND)
    for t in range(n):
        o, _ = orc.track(oracle.synth_frame(rows, cols, 1003, t), lr, op)
        d = got[t]
        assert bool(d[0]) == bool(o.position_valid) and d[1] == o.n_components
        assert abs(d[2] - o.x) <= TOL and abs(d[3] - o.y) <= TOL and abs(d[4] - o.area) <= TOL
    om, ow, ov, omu = orc.mog.state()
    gm, gw, gv, gmu = a.state()
    assert np.array_equal(gm, om)
    live = np.arange(gw.shape[2])[None, None, :] < om[:, :, None]
    assert np.array_equal(gw.view(np.uint32)[live], ow.view(np.uint32)[live])
    assert np.array_equal(gv.view(np.uint32)[live], ov.view(np.uint32)[live])
    a.close()


@pytest.mark.parametrize("cols,pitch", [(100, 304), (1000, 3008), (68, 208)])
def test_pitched_frames_take_the_staged_kernel(ctx, cols, pitch):
    """Row pitch > 3*cols (what host frames become in the staging buffer, and what a 1000-column frame needs to have
    16-byte aligned rows): egress, masks and state against the oracle, frame by frame."""
    rows, lr = 75, 0.02
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    trk = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**HSV_BAND)
    for t in range(12):
        f = oracle.synth_frame(rows, cols, 1000, t)
        d, eg = trk.track(f, egress=("bgr", "fgmask", "hsv", "thresh"))  # host frame: staged with a 16-byte aligned pitch
        o, oeg = orc.track(f, lr, op)
        for k in ("fgmask", "bgr", "hsv", "thresh"):
            assert np.array_equal(eg[k], oeg[k]), f"{k} differs at t={t}"
        assert _det(d)[:2] == (o.position_valid, o.n_components)
        assert abs(d.x - o.x) <= TOL and abs(d.y - o.y) <= TOL and abs(d.area - o.area) <= TOL
    trk.close()
    # the same stream as a pitched device-resident clip through the resident engine
    bufs = _frames(ctx, rows, cols, 1000, 12, pitch=pitch)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=8)
    ptrs = oat_b200.frame_pointers(bufs)
    import ctypes as C
    out = (oat_b200.Detection * 12)()
    oat_b200._ck(oat_b200.lib().oat_tracker_run_clip(a._h, ptrs, 12, pitch, lr, C.byref(a.hsv), 4, out, None))
    orc = oracle.Tracker(rows, cols)
    for t in range(12):
        o, _ = orc.track(oracle.synth_frame(rows, cols, 1000, t), lr, op)
        assert bool(out[t].position_valid) == bool(o.position_valid)
        assert abs(out[t].x - o.x) <= TOL and abs(out[t].y - o.y) <= TOL and abs(out[t].area - o.area) <= TOL
    assert a.tail_stats()["clip_frames"] == 11
    a.close()


def test_clip_engine_soak_1080p(ctx):
    """The benchmark's own configuration, long: 2000 frames 1080p, -a 0.01, chunks of 32 frames (62 chunk boundaries,
    every one of them overlapped tile by tile with its predecessor) against the synchronous frame-by-frame path:
    every detection and the whole GMM state bit-identical."""
    rows, cols, lr, n = 1080, 1920, 0.01, int(os.environ.get("OAT_SOAK_FRAMES", "2000"))  # (shortened under compute-sanitizer)
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    R = 24
    bufs = _frames(ctx, rows, cols, 1000, R + 1)
    seq = [bufs[0]] + [bufs[1 + i % R] for i in range(n - 1)]
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=64)
    got = [_det(d) for d in a.run_clip(seq)]
    assert a.tail_stats()["clip_frames"] == n - 1
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in seq]
    assert got == want
    assert a.live_modes() == b.live_modes()
    _state_equal(a, b)
    a.close()
    b.close()


def test_clip_engine_interleaved_streams(ctx):
    """Three independent 1080p streams interleaved in ONE queue (oat_tracker_run_clips), 400 frames each: per-stream
    detections and state equal each stream's own synchronous run; then the same again continuing on the per-frame
    path (a resident launch followed by chained single-frame launches)."""
    rows, cols, lr, n, S = 1080, 1920, 0.02, 400, 3
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    R = 10
    bufs = [_frames(ctx, rows, cols, 1000 + s, R + 1) for s in range(S)]
    seqs = [[bufs[s][0]] + [bufs[s][1 + i % R] for i in range(n - 1)] for s in range(S)]
    trks = [oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=16) for _ in range(S)]
    for s in range(S):  # the first frame of every model goes frame by frame
        trks[s].submit(seqs[s][0])
        trks[s].collect()
    got = oat_b200.Tracker.run_clips(trks, [[seqs[s][i] for s in range(S)] for i in range(1, n)])
    assert all(t.tail_stats()["clip_frames"] == n - 1 for t in trks)
    tails = [[] for _ in range(S)]
    for i in range(5):  # ... and a few more frames through submit/collect, chained to the resident launch
        for s in range(S):
            trks[s].submit(seqs[s][1 + i])
        for s in range(S):
            tails[s].append(_det(trks[s].collect()))
    for s in range(S):
        ref = oat_b200.Tracker(ctx, rows, cols, lr, hp)
        want = [_det(ref.track(f)[0]) for f in seqs[s]]
        assert [_det(got[i][s]) for i in range(n - 1)] == want[1:], s
        want_tail = [_det(ref.track(seqs[s][1 + i])[0]) for i in range(5)]
        assert tails[s] == want_tail, s
        _state_equal(trks[s], ref)
        ref.close()
        trks[s].close()


def test_clip_engine_fused_only_advances_the_model_identically(ctx):
    rows, cols, lr, n, S = 480, 640, 0.05, 30, 2
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    bufs = [_frames(ctx, rows, cols, 1000 + s, n) for s in range(S)]
    trks = [oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=8) for _ in range(S)]
    for s in range(S):
        trks[s].submit(bufs[s][0])
        trks[s].collect()
    assert oat_b200.Tracker.run_clips(trks, [[bufs[s][i] for s in range(S)] for i in range(1, n)], fused_only=True) is None
    for s in range(S):
        ref = oat_b200.Tracker(ctx, rows, cols, lr, hp)
        for f in bufs[s]:
            ref.track(f)
        _state_equal(trks[s], ref)
        # the tracker is usable afterwards
        assert _det(trks[s].track(bufs[s][3])[0]) == _det(ref.track(bufs[s][3])[0])
        ref.close()
        trks[s].close()


def test_clip_engine_with_position_filter(ctx):
    rows, cols, lr, n = 240, 320, 0.01, 40
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    kal = dict(dt=1 / 30, timeout=0.3, sigma_accel=20.0, sigma_noise=0.5)
    bufs = _frames(ctx, rows, cols, 1000, n)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=16)
    pa = oat_b200.PositionFilter(ctx, 1, kal)
    a.attach_posfilt(pa)
    dets, poss = a.run_clip(bufs, positions=True)
    assert a.tail_stats()["clip_frames"] == n - 1
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=2)
    pb = oat_b200.PositionFilter(ctx, 1, kal)
    b.attach_posfilt(pb)
    for t in range(n):
        b.submit(bufs[t])
        d, p = b.collect_position()
        assert _det(d) == _det(dets[t])
        q = poss[t]
        assert (p.position_valid, p.x, p.y, p.vx, p.vy) == (q.position_valid, q.x, q.y, q.vx, q.vy), t
    for trk, f in ((a, pa), (b, pb)):
        trk.attach_posfilt(None)
        f.close()
        trk.close()


def test_clip_engine_overflowing_masks_are_replayed(ctx):
    """A band that passes everything on a noisy stream: the threshold mask is the (speckled) foreground mask, far more
    runs than the one-launch tail's table holds.  Those frames are replayed through the unbounded path after their
    chunk; detections must still equal the synchronous path."""
    from test_gpu_mog import noisy_stream

    rows, cols, lr, n = 480, 640, 0.05, 20
    hp = oat_b200.HsvParams.make(h=(0, 256), s=(0, 256), v=(1, 256), dilate=0)
    host = list(noisy_stream(rows, cols, n, 15.0, seed=3))
    bufs = []
    for f in host:
        b = ctx.alloc(rows * cols * 3)
        b.upload(f)
        bufs.append(b)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=8)
    got = [_det(d) for d in a.run_clip(bufs)]
    st = a.tail_stats()
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in bufs]
    assert got == want
    assert st["clip_frames"] > 0 and st["replays"] > 0, st
    _state_equal(a, b)
    a.close()
    b.close()


def test_clip_engine_hands_over_to_generic_kernel_on_busy_stream(ctx):
    """The census still drives the kernel choice: a stream that leaves the fast path on most pixels leaves the resident
    engine after the chunk that noticed, and the rest of the clip runs frame by frame on the generic kernel."""
    from test_gpu_mog import noisy_stream

    rows, cols, lr, n = 96, 128, 0.3, 60
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    host = list(noisy_stream(rows, cols, n, 20.0, seed=7))
    bufs = []
    for f in host:
        b = ctx.alloc(rows * cols * 3)
        b.upload(f)
        bufs.append(b)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=8)
    got = [_det(d) for d in a.run_clip(bufs)]
    st = a.tail_stats()
    assert st["generic_frames"] > 0 and 0 < st["clip_frames"] < n - 1, st
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**HSV_BAND)
    for t in range(n):
        o, _ = orc.track(host[t], lr, op)
        d = got[t]
        assert bool(d[0]) == bool(o.position_valid) and abs(d[2] - o.x) <= TOL and abs(d[3] - o.y) <= TOL and abs(d[4] - o.area) <= TOL
    assert a.live_modes() == int(orc.mog.state()[0].sum())
    a.close()


@pytest.mark.parametrize("shape", [(480, 640), (360, 480), (720, 1280)])
def test_clip_engine_close_dependencies_stress(ctx, shape):
    """Frames of about one grid's worth of tiles: consecutive frames of the stream are in flight TOGETHER, so nearly
    every tile load really waits for the previous frame's publication of that tile (at 1080p and above the previous
    frame's tile was published long before).  This is the regime that exposes a publication that is not a proper
    release (an unfenced flag store behind completed bulk stores fails here about once in 30 clips).  Repeated
    clips, whole-state comparison against the synchronous path."""
    rows, cols = shape
    lr, n, reps = 0.05, 160, int(os.environ.get("OAT_STRESS_REPS", "12"))  # (fewer under compute-sanitizer)
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    R = 16
    bufs = _frames(ctx, rows, cols, 1000, R + 1)
    seq = [bufs[0]] + [bufs[1 + i % R] for i in range(n - 1)]
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in seq]
    for rep in range(reps):
        a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=(64, 16, 8)[rep % 3])
        got = [_det(d) for d in a.run_clip(seq)]
        assert got == want, rep
        _state_equal(a, b)
        a.close()
    b.close()


def _blob_scene(rows, cols, nblobs, nframes, seed=1000):
    """Background + `nblobs` shapes of the target colour wandering over the cells of a grid: discs, rings (a hole that must
    join its contour) and rings with a disc inside (a contour inside a hole: its own external contour)."""
    bg = oracle.synth_frame(rows, cols, seed, 0)
    gx = max(1, int(round((nblobs * cols / rows) ** 0.5)))
    gy = (nblobs + gx - 1) // gx
    cw, ch = cols // gx, rows // gy
    r = max(4, min(ch // 5, cw // 5, rows // 30))
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    d2 = xx ** 2 + yy ** 2
    out = [bg]
    for t in range(nframes):
        f = bg.copy()
        for k in range(nblobs):
            i, j = k % gx, k // gx
            cx = i * cw + r + (t * 37 + k * 11) % max(1, cw - 2 * r)
            cy = j * ch + r + (t * 23 + k * 7) % max(1, ch - 2 * r)
            win = f[cy - r:cy + r + 1, cx - r:cx + r + 1]
            keep = win.copy()
            shape = d2 <= (r - k % 3) ** 2
            if k % 3 == 1:
                shape &= d2 >= (r // 2) ** 2                      # ring
            elif k % 3 == 2:
                shape &= (d2 >= (r // 2) ** 2) | (d2 <= (r // 4) ** 2)  # ring with a disc inside its hole
            win[shape] = (40, 220, 60)
            win[~shape] = keep[~shape]
        out.append(f)
    return out


@pytest.mark.parametrize("nblobs", [8, 60, 200])
def test_clip_engine_many_blobs_are_labelled_on_the_device(ctx, nblobs):
    """Blobs spread over the frame: the mask's bounding region is the whole frame (259 KB at 1080p), far more than the
    labelling CTA's shared memory.  8 and 60 blobs: the run table still fits there, the mask is read -- and its holes
    filled -- IN PLACE; 200 blobs: table and contour count outgrow shared memory, the tail server labels the frame a
    second time in its global-memory scratch area.  Either way on the device, no host replay, and the same result as the
    synchronous path (which replays through the unbounded multi-kernel tail)."""
    rows, cols, lr = 1080, 1920, 0.01
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    host = _blob_scene(rows, cols, nblobs, 7)
    bufs = []
    for f in host:
        b = ctx.alloc(rows * cols * 3)
        b.upload(f)
        bufs.append(b)
    seq = [bufs[0]] + [bufs[1 + i % 7] for i in range(40)]
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=16)
    got = [_det(d) for d in a.run_clip(seq)]
    st = a.tail_stats()
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in seq]
    assert got == want
    assert want[-1][1] >= nblobs - nblobs // 8, want[-1]   # the scene really has that many contours (rings count once)
    assert st["clip_frames"] == 40
    assert st["replays"] <= 1, st               # (the first frame's whole-image blob goes frame by frame and is replayed)
    # and against the oracle on the last frames' masks: the tail's contour semantics (holes, nested contours)
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(**HSV_BAND)
    for t, f in enumerate([host[0]] + [host[1 + i % 7] for i in range(3)]):
        o, _ = orc.track(f, lr, op)
        d = got[t]
        assert bool(d[0]) == bool(o.position_valid) and d[1] == o.n_components, (t, d, o.n_components)
        assert abs(d[2] - o.x) <= TOL and abs(d[3] - o.y) <= TOL and abs(d[4] - o.area) <= TOL
    a.close()
    b.close()
    for x in bufs:
        x.free()


# ---- streaming use of the engine (oat_tracker_stream_*): what `oat posidet track --pipeline` drives ----------------
def _stream_all(trk, frames, copy, flush_every=None, poll_every=3, pitch=None):
    """Push every frame (flush / poll on a schedule), then drain; returns the detections in order."""
    got = []
    for i, f in enumerate(frames):
        trk.stream_push(f, copy=copy, pitch=pitch)
        if copy:
            trk.stream_wait_ingest()
        if flush_every and (i + 1) % flush_every == 0:
            trk.stream_flush()
        if poll_every and (i + 1) % poll_every == 0:
            got += trk.stream_poll()
    while sum(trk.stream_pending()) > 0:
        got += trk.stream_poll(block=True)
    return [_det(d) for d in got]


@pytest.mark.parametrize("shape", [(120, 160), (480, 640), (100, 1000)])
@pytest.mark.parametrize("lr", [0.0, 0.02])
@pytest.mark.parametrize("ring,flush_every", [(2, None), (8, None), (8, 1), (64, 5), (64, None)])
def test_stream_equals_frame_by_frame(ctx, shape, lr, ring, flush_every):
    """Device frames read in place: full chunks, short chunks (flush after every frame / every 5), polls in between --
    detections and the whole GMM state equal the synchronous path's."""
    rows, cols = shape
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    n = 75
    pitch = None if (cols * 3) % 16 == 0 else (cols * 3 + 15) // 16 * 16  # (in place: the engine wants 16-byte aligned rows)
    bufs = _frames(ctx, rows, cols, 1000, n, pitch=pitch)
    a = oat_b200.Tracker(ctx, rows, cols, lr, hp, ring_depth=ring)
    got = _stream_all(a, bufs, copy=False, flush_every=flush_every, pitch=pitch)
    assert a.tail_stats()["clip_frames"] == n - 1, "the resident engine did not take the stream"
    b = oat_b200.Tracker(ctx, rows, cols, lr, hp)
    want = []
    for f in bufs:  # the per-frame path, one frame in flight
        b.submit(f, pitch=pitch)
        want.append(_det(b.collect()))
    assert got == want
    _state_equal(a, b)
    a.close()
    b.close()


@pytest.mark.parametrize("source", ["host", "pinned", "device"])
@pytest.mark.parametrize("cols", [160, 1000])
def test_stream_with_staged_frames(ctx, source, cols):
    """OAT_STREAM_COPY: ONE buffer is overwritten with the next frame as soon as wait_ingest returns (what a lock-step
    SOURCE does with its shared frame) -- pageable host, pinned host and device memory."""
    rows = 120
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    n = 50
    frames = [oracle.synth_frame(rows, cols, 7, t) for t in range(n)]
    a = oat_b200.Tracker(ctx, rows, cols, 0.02, hp, ring_depth=8)
    if source == "host":
        shared = np.empty((rows, cols, 3), np.uint8)
    elif source == "pinned":
        pin = oat_b200.PinnedArray((rows, cols, 3))
        shared = pin.array
    else:
        shared = ctx.alloc(rows * cols * 3)
    got = []
    for t in range(n):
        if source == "device":
            shared.upload(frames[t])
        else:
            shared[...] = frames[t]
        a.stream_push(shared, copy=True)
        a.stream_wait_ingest()
        if t % 2:
            a.stream_flush()
        got += a.stream_poll()
    while sum(a.stream_pending()) > 0:
        got += a.stream_poll(block=True)
    b = oat_b200.Tracker(ctx, rows, cols, 0.02, hp)
    want = [_det(b.track(f)[0]) for f in frames]
    assert [_det(d) for d in got] == want
    _state_equal(a, b)
    a.close()
    b.close()


def test_stream_mixed_parameters_and_positions(ctx):
    """A learning rate that changes mid-stream closes the chunk; with a position filter attached poll returns the
    filtered positions, equal to the per-frame path's; the other entry points refuse while frames are in flight."""
    rows, cols = 240, 320
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    n = 40
    bufs = _frames(ctx, rows, cols, 1000, n)
    lrs = [0.02 if t < 17 else 0.05 if t < 30 else 0.0 for t in range(n)]
    a = oat_b200.Tracker(ctx, rows, cols, 0.02, hp, ring_depth=16)
    fa = oat_b200.PositionFilter(ctx, 1, kalman=dict(dt=0.02, timeout=1.0, sigma_accel=5.0, sigma_noise=1.0))
    a.attach_posfilt(fa)
    dets, poss = [], []
    for t in range(n):
        a.stream_push(bufs[t], learning_rate=lrs[t])
        if t == 5:
            with pytest.raises(oat_b200.OatError):
                a.submit(bufs[t])
        d, p = a.stream_poll(positions=True)
        dets += d
        poss += p
    while sum(a.stream_pending()) > 0:
        d, p = a.stream_poll(block=True, positions=True)
        dets += d
        poss += p
    b = oat_b200.Tracker(ctx, rows, cols, 0.02, hp)
    fb = oat_b200.PositionFilter(ctx, 1, kalman=dict(dt=0.02, timeout=1.0, sigma_accel=5.0, sigma_noise=1.0))
    b.attach_posfilt(fb)
    for t in range(n):
        b.submit(bufs[t], learning_rate=lrs[t])
        d, p = b.collect_position()
        assert _det(dets[t]) == _det(d)
        assert (poss[t].position_valid, poss[t].x, poss[t].y, poss[t].vx, poss[t].vy) == (p.position_valid, p.x, p.y, p.vx, p.vy)
    _state_equal(a, b)
    a.close()
    b.close()


# ---- band pre-labelling (tail_fast.cuh: band_prelabel): the run table of a busy mask arrives merged inside every band ----
def _structure_scene(rows, cols, nframes, seed):
    """Frames whose target-coloured pixels form structures that stress everything a band cannot decide alone: worms and
    rings that cross band borders (rows that are multiples of 32), U shapes whose inside reaches the exterior only
    through another band, holes with islands, nested rings, speckle (many short runs), shapes on the frame border."""
    rng = np.random.default_rng(seed)
    bg = oracle.synth_frame(rows, cols, 1000, 0)
    yy, xx = np.mgrid[0:rows, 0:cols]
    out = [bg]
    dens = max(1, min(6, rows * cols // 25000))  # (a frame that is mostly foreground leaves the resident engine: keep it sparse)
    for t in range(nframes):
        m = np.zeros((rows, cols), bool)
        for _ in range(dens):  # rings and nested rings, any size, clipped by the frame
            cy, cx, r = rng.integers(0, rows), rng.integers(0, cols), rng.integers(6, max(8, rows // 3))
            d2 = (yy - cy) ** 2 + (xx - cx) ** 2
            w = rng.integers(1, 4)
            m |= (d2 <= r * r) & (d2 >= (r - w) ** 2)
            if rng.integers(2):
                m |= d2 <= (r // 3) ** 2
            if rng.integers(2):
                m |= (d2 <= (r // 2) ** 2) & (d2 >= (r // 2 - 1) ** 2)
        for _ in range(dens):  # U / C shapes: a frame of a rectangle with one side open
            y0, x0 = rng.integers(0, rows - 8), rng.integers(0, cols - 8)
            h, w = rng.integers(6, max(8, rows // 2)), rng.integers(6, max(8, cols // 3))
            y1, x1 = min(rows - 1, y0 + h), min(cols - 1, x0 + w)
            box = np.zeros((rows, cols), bool)
            box[y0:y1 + 1, x0:x1 + 1] = True
            box[y0 + 2:y1 - 1, x0 + 2:x1 - 1] = False
            side = rng.integers(4)
            if side == 0:
                box[y0:y0 + 2, x0 + 3:x1 - 2] = False
            elif side == 1:
                box[y1 - 1:y1 + 1, x0 + 3:x1 - 2] = False
            elif side == 2:
                box[y0 + 3:y1 - 2, x0:x0 + 2] = False
            else:
                box[y0 + 3:y1 - 2, x1 - 1:x1 + 1] = False
            m |= box
        for _ in range(max(1, dens - 2)):  # worms
            y, x = int(rng.integers(0, rows)), int(rng.integers(0, cols))
            for _ in range(rows // 2):
                m[max(0, y - 1):y + 1, max(0, x - 1):x + 1] = True
                y = int(np.clip(y + rng.integers(-2, 3), 0, rows - 1))
                x = int(np.clip(x + rng.integers(-2, 3), 0, cols - 1))
        ph, pw = (40, 60) if dens > 2 else (12, 20)
        py, px = rng.integers(0, max(1, rows - ph)), rng.integers(0, max(1, cols - pw))  # a speckled patch
        m[py:py + ph, px:px + pw] |= rng.random(m[py:py + ph, px:px + pw].shape) < 0.3
        if t % 3 == 0:  # something on every frame border
            m[0, cols // 4:cols // 2] = True
            m[rows - 1, cols // 3:cols // 2] = True
            m[rows // 4:rows // 2, 0] = True
            m[rows // 3:rows // 2, cols - 1] = True
        f = bg.copy()
        f[m] = (40, 220, 60)
        out.append(f)
    return out


@pytest.mark.parametrize("shape", [(96, 128), (250, 336), (480, 640), (1080, 1920)])
@pytest.mark.parametrize("dilate", [0, 3])
def test_prelabelled_bands_equal_the_synchronous_tail(shape, dilate):
    """The tail server with the bands' pre-labelled pool forced on (a private context: the switch is read when it is
    created) against the synchronous per-frame path, which labels from the mask; then against the oracle's contours."""
    rows, cols = shape
    lr, n = 0.01, 12
    os.environ["OAT_B200_FORCE_PRELABEL"] = "1"
    try:
        c2 = oat_b200.Context(0)
    finally:
        del os.environ["OAT_B200_FORCE_PRELABEL"]
    hp = oat_b200.HsvParams.make(dilate=dilate, **HSV_BAND)
    host = _structure_scene(rows, cols, n, seed=rows + dilate)
    bufs = []
    for f in host:
        b = c2.alloc(rows * cols * 3)
        b.upload(f)
        bufs.append(b)
    a = oat_b200.Tracker(c2, rows, cols, lr, hp, ring_depth=8)
    got = [_det(d) for d in a.run_clip(bufs)]
    st = a.tail_stats()
    b = oat_b200.Tracker(c2, rows, cols, lr, hp)
    want = [_det(b.track(f)[0]) for f in bufs]
    assert got == want
    assert st["generic_frames"] <= 1 and st["clip_frames"] == n, st   # (the scene stayed on the resident engine ...)
    assert st["prelabelled"] == 1 and st["status"] == 0, st           # (... and its last frame took the bands' table)
    orc = oracle.Tracker(rows, cols)
    op = oracle.HsvParams(dilate=dilate, **HSV_BAND)
    for t in range(4 if rows > 500 else n + 1):
        o, _ = orc.track(host[t], lr, op)
        d = got[t]
        assert bool(d[0]) == bool(o.position_valid) and d[1] == o.n_components, (t, d, o.n_components)
        assert abs(d[2] - o.x) <= TOL and abs(d[3] - o.y) <= TOL and abs(d[4] - o.area) <= TOL
    # ... and against OpenCV itself where the box has it: findContours(RETR_EXTERNAL) + moments on every frame
    from oracle import cv2ref
    if cv2ref.available():
        pipe = cv2ref.Pipeline(lr, dilate=dilate, **HSV_BAND)
        for t, f in enumerate(host):
            valid, x, y, area = pipe.step(f)
            d = got[t]
            assert bool(d[0]) == bool(valid), t
            assert abs(d[2] - x) <= TOL and abs(d[3] - y) <= TOL and abs(d[4] - area) <= TOL, (t, d, x, y, area)
    a.close()
    b.close()
    for x in bufs:
        x.free()
    c2.close()
