"""Where does a 1080p frame's time go once launches overlap?  (a) the fused kernel alone, same model, launched
back to back (chained) -- GPU time per launch by events and host time per enqueue; (b) host cost of
submit/collect in the pipelined loop.  Experiments only, not the judged bench."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oat_b200
res = sys.argv[1] if len(sys.argv) > 1 else "1080p"
rows, cols = {"1080p": (1080, 1920), "4k": (2160, 3840)}[res]
ctx = oat_b200.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
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
N = 2000
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    t0 = time.perf_counter()
    a.record(st)
    for i in range(N):
        trk.submit_fused_only(frames[1 + (50 + i) % R])
    b.record(st)
    th = time.perf_counter() - t0
    ctx.sync()
    print(f"{res} fused-only same-model chain: GPU {1e3*a.elapsed_time(b)/N:.2f} us/launch, host enqueue {1e6*th/N:.2f} us/launch")
# (a') the same with an event record after every launch (what the tracker does for the tail's stream)
evs = [torch.cuda.Event() for _ in range(64)]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ctx.sync()
a.record(st)
for i in range(N):
    trk.submit_fused_only(frames[1 + (50 + i) % R])
    evs[i % 64].record(st)
b.record(st)
ctx.sync()
print(f"{res} fused-only chain + event record per launch: GPU {1e3*a.elapsed_time(b)/N:.2f} us/launch")
for depth in (4, 8):
    ts = tc = 0.0
    out = 0
    M = 2000
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0 = time.perf_counter()
    a.record(st)
    for i in range(M):
        t0 = time.perf_counter(); trk.submit(frames[1 + i % R]); ts += time.perf_counter() - t0
        out += 1
        if out == depth:
            t0 = time.perf_counter(); trk.collect(); tc += time.perf_counter() - t0
            out -= 1
    while out:
        trk.collect(); out -= 1
    b.record(st)
    t_all = time.perf_counter() - t_all0
    ctx.sync()
    print(f"{res} pipelined depth {depth}: submit host {1e6*ts/M:.1f} us/call, collect host {1e6*tc/(M-depth+1):.1f} us/call, wall {1e6*t_all/M:.1f} us/frame, GPU {1e3*a.elapsed_time(b)/M:.2f} us/frame")
for depth in (4, 8):
    clip = [frames[1 + i % R] for i in range(2000)]
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(st)
    trk.run_clip(clip, depth=depth)
    b.record(st)
    tw = time.perf_counter() - t0
    ctx.sync()
    print(f"{res} run_clip depth {depth}: wall {1e6*tw/2000:.1f} us/frame, GPU {1e3*a.elapsed_time(b)/2000:.2f} us/frame")
