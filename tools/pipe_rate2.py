import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oat_b200
rows, cols = 1080, 1920
ctx = oat_b200.Context(0)
R = 32
frames = [ctx.alloc(rows * cols * 3) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, t, out=b)
for name, hp in (("band", oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))), ("empty-mask", oat_b200.HsvParams.make(h=(200, 210), s=(100, 256), v=(100, 256)))):
    for alpha, depth in ((0.0, 8), (0.01, 8)):
        trk = oat_b200.Tracker(ctx, rows, cols, alpha, hp, ring_depth=depth)
        trk.track(frames[0])
        for i in range(50):
            trk.track(frames[1 + i % R])
        ctx.sync(); N = 2000; out = 0; t0 = time.perf_counter()
        for i in range(N):
            trk.submit(frames[1 + i % R]); out += 1
            if out == depth:
                trk.collect(); out -= 1
        while out:
            trk.collect(); out -= 1
        dt = time.perf_counter() - t0
        st = trk.tail_stats()
        print(name, "depth", depth, "alpha", alpha, "us/frame %.2f" % (1e6 * dt / N), [(st["cyc"][i] - st["cyc"][0]) & 0xffffffff for i in range(1, 8)], st["nodes"])
        trk.close()
