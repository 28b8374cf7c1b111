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
trk = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=8)
trk.track(frames[0])
for i in range(50):
    trk.track(frames[1 + i % R])
ctx.sync()
# host cost of submit when the GPU queue is short (8 in flight), then collect
N = 400
ts = 0.0; tc = 0.0
t_all0 = time.perf_counter()
out = 0
for i in range(N):
    t0 = time.perf_counter(); trk.submit(frames[1 + i % R]); ts += time.perf_counter() - t0
    out += 1
    if out == 8:
        t0 = time.perf_counter(); trk.collect(); tc += time.perf_counter() - t0
        out -= 1
while out:
    trk.collect(); out -= 1
t_all = time.perf_counter() - t_all0
print(f"submit host {1e6*ts/N:.1f} us/call, collect host {1e6*tc/(N-7):.1f} us/call, total {1e6*t_all/N:.1f} us/frame")
