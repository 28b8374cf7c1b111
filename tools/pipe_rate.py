import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oat_b200
rows, cols = 1080, 1920
ctx = oat_b200.Context(0)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
R = 32
frames = [ctx.alloc(rows * cols * 3) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, t, out=b)
for alpha in (0.0, 0.01):
    trk = oat_b200.Tracker(ctx, rows, cols, alpha, hp, ring_depth=4)
    trk.track(frames[0])
    for i in range(50):
        trk.track(frames[1 + i % R])
    ctx.sync(); N = 2000; out = 0; t0 = time.perf_counter()
    for i in range(N):
        trk.submit(frames[1 + i % R]); out += 1
        if out == 4:
            trk.collect(); out -= 1
    while out:
        trk.collect(); out -= 1
    dt = time.perf_counter() - t0
    print("alpha", alpha, "us/frame %.2f" % (1e6 * dt / N), os.environ.get("OAT_B200_NO_TRACK", ""), os.environ.get("OAT_B200_NO_OVERLAP", ""))
    trk.close()
