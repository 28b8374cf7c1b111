"""GPU parity: the fused MOG2 kernel (through the C ABI) vs the CPU oracle -- bit-exact masks,
filtered frames and GMM state.  Reference: src/framefilter/BackgroundSubtractorMOG.cpp:114-127."""
import numpy as np
import pytest

import oat_b200
import oracle

pytestmark = pytest.mark.gpu


def noisy_stream(rows, cols, nframes, sigma, seed):
    """Static background + gaussian noise + a moving bright/dark blob pair (exercises new modes,
    pruning, mode replacement and shadows)."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(30, 200, (rows, cols, 3)).astype(np.float32)
    for t in range(nframes):
        f = bg + rng.normal(0, sigma, bg.shape) if sigma > 0 else bg.copy()
        x0 = (5 * t) % max(cols - 12, 1)
        y0 = (3 * t) % max(rows - 12, 1)
        f[y0:y0 + 12, x0:x0 + 12] = (30, 220, 60)
        # a "shadow": darkened copy of the background
        xs = (cols - 20 - 4 * t) % max(cols - 16, 1)
        f[rows // 2:rows // 2 + 10, xs:xs + 16] = bg[rows // 2:rows // 2 + 10, xs:xs + 16] * 0.7
        yield np.clip(np.rint(f), 0, 255).astype(np.uint8)


def assert_state_equal(gpu_state, orc_state):
    gm, gw, gv, gmu = gpu_state
    om, ow, ov, omu = orc_state
    np.testing.assert_array_equal(gm, om)
    K = gw.shape[2]
    live = np.arange(K)[None, None, :] < om[:, :, None]
    # bitwise comparison of live modes only (dead slots are unspecified)
    assert np.array_equal(gw.view(np.uint32)[live], ow.view(np.uint32)[live])
    assert np.array_equal(gv.view(np.uint32)[live], ov.view(np.uint32)[live])
    live3 = np.broadcast_to(live[..., None], gmu.shape)
    assert np.array_equal(gmu.view(np.uint32)[live3], omu.view(np.uint32)[live3])


@pytest.mark.parametrize("shape", [(48, 64), (37, 53), (40, 100)])
@pytest.mark.parametrize("lr,sigma", [(0.0, 3.0), (0.01, 3.0), (0.1, 8.0), (-1.0, 8.0), (0.3, 20.0), (1.0, 0.0)])
def test_mog_mask_state_bit_exact(ctx, shape, lr, sigma):
    rows, cols = shape
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols)
    orc = oracle.Mog2(rows, cols)
    for t, frame in enumerate(noisy_stream(rows, cols, 30, sigma, seed=rows * 1000 + cols)):
        out, mask = gpu.apply(frame, learning_rate=lr)
        omask = orc.apply(frame, lr)
        assert np.array_equal(mask, omask), f"mask differs at frame {t}: {(mask != omask).sum()} px"
        assert np.array_equal(out, oracle.zero_where_mask0(frame, omask)), f"filtered frame differs at {t}"
        if t % 7 == 0 or t == 29:
            assert_state_equal(gpu.state(), orc.state())
    total = gpu.live_modes()
    assert total == int(orc.state()[0].sum())
    gpu.close()


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_mog_nmixtures(ctx, K):
    rows, cols = 32, 64
    p = oat_b200.default_mog_params()
    p.nmixtures = K
    op = oracle.default_mog_params()
    op.nmixtures = K
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols, params=p)
    orc = oracle.Mog2(rows, cols, op)
    for t, frame in enumerate(noisy_stream(rows, cols, 25, 15.0, seed=K)):
        _, mask = gpu.apply(frame, learning_rate=0.2)
        assert np.array_equal(mask, orc.apply(frame, 0.2)), f"K={K} frame {t}"
    assert_state_equal(gpu.state(), orc.state())
    gpu.close()


def test_mog_no_shadow_param(ctx):
    rows, cols = 32, 64
    p = oat_b200.default_mog_params()
    p.detect_shadows = 0
    op = oracle.default_mog_params()
    op.detect_shadows = 0
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols, params=p)
    orc = oracle.Mog2(rows, cols, op)
    for frame in noisy_stream(rows, cols, 10, 5.0, seed=3):
        _, mask = gpu.apply(frame, learning_rate=0.05)
        assert np.array_equal(mask, orc.apply(frame, 0.05))
        assert set(np.unique(mask)) <= {0, 255}
    gpu.close()


def test_mog_first_frame_all_shadow_and_reset(ctx):
    """the first apply() returns an all-127 mask; reset makes the next frame first again."""
    rows, cols = 24, 32
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols)
    f = np.full((rows, cols, 3), 90, np.uint8)
    _, m = gpu.apply(f)
    assert (m == 127).all()
    _, m = gpu.apply(f)
    assert (m == 0).all()
    gpu.reset()
    _, m = gpu.apply(f)
    assert (m == 127).all()
    # a black pixel has den == 0 in the shadow test -> foreground on the first frame
    gpu.reset()
    f[0, 0] = 0
    _, m = gpu.apply(f)
    assert m[0, 0] == 255 and (m.ravel()[1:] == 127).all()
    gpu.close()


def test_mog_filter_in_place_and_pitched_device(ctx):
    """filter() mutates the frame in place like FrameFilter::filter(cv::Mat&); device-resident
    input with a padded pitch takes the same path."""
    import ctypes as C
    rows, cols = 30, 44
    frames = list(noisy_stream(rows, cols, 6, 4.0, seed=9))
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols, adaptation_coeff=0.02)
    orc = oracle.Mog2(rows, cols)
    for f in frames[:3]:
        want = oracle.zero_where_mask0(f, orc.apply(f, 0.02))
        g = f.copy()
        gpu.filter(g)
        assert np.array_equal(g, want)
    # pitched device input / output
    pitch = 3 * cols + 20  # 152: 4-byte aligned -> vector path
    dev_in = ctx.alloc(pitch * rows)
    dev_out = ctx.alloc(pitch * rows)
    dev_mask = ctx.alloc(64 * rows)
    for f in frames[3:]:
        padded = np.zeros((rows, pitch), np.uint8)
        padded[:, :3 * cols] = f.reshape(rows, -1)
        dev_in.upload(padded)
        oat_b200._ck(oat_b200.lib().oat_mog_apply(gpu._h, C.c_void_p(dev_in.ptr), pitch, C.c_void_p(dev_out.ptr), pitch,
                                                  C.c_void_p(dev_mask.ptr), 64, 0.02))
        om = orc.apply(f, 0.02)
        got = dev_out.download((rows, pitch))[:, :3 * cols].reshape(rows, cols, 3)
        gm = dev_mask.download((rows, 64))[:, :cols]
        assert np.array_equal(gm, om)
        assert np.array_equal(got, oracle.zero_where_mask0(f, om))
    gpu.close()


def test_mog_unaligned_pitch_scalar_path(ctx):
    """An odd input pitch forces the 1-pixel-per-thread kernel; results must not change."""
    import ctypes as C
    rows, cols = 20, 36
    pitch = 3 * cols + 1
    gpu = oat_b200.BackgroundSubtractorMOG(ctx, rows, cols)
    orc = oracle.Mog2(rows, cols)
    for f in noisy_stream(rows, cols, 8, 6.0, seed=2):
        padded = np.zeros((rows, pitch), np.uint8)
        padded[:, :3 * cols] = f.reshape(rows, -1)
        mask = np.empty((rows, cols), np.uint8)
        oat_b200._ck(oat_b200.lib().oat_mog_apply(gpu._h, padded.ctypes.data_as(C.c_void_p), pitch, None, 0,
                                                  mask.ctypes.data_as(C.c_void_p), cols, 0.05))
        assert np.array_equal(mask, orc.apply(f, 0.05))
    gpu.close()
