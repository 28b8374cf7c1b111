import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oat_b200
rows, cols = 1080, 1920
ctx = oat_b200.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
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
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 400
    e0.record(st)
    for i in range(N):
        trk.submit_fused_only(frames[1 + i % R])
    e1.record(st)
    ctx.sync()
    print("alpha", alpha, "fused-only back-to-back, one stream, hot L2: %.2f us/launch" % (1e3 * e0.elapsed_time(e1) / N))
    trk.close()
