"""Host cost per frame (the host was co-critical at 15.5 us of CPU per frame):
(a) the resident engine -- oat_tracker_run_clip on device-resident frames: process CPU time and wall time per frame;
(b) the per-frame path -- submit / collect called frame by frame (what a host-fed stream pays), CPU time per call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oat_b200

res = sys.argv[1] if len(sys.argv) > 1 else "1080p"
rows, cols = {"1080p": (1080, 1920), "4k": (2160, 3840)}[res]
ctx = oat_b200.Context(0)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
R = 32
frames = [ctx.alloc(rows * cols * 3) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, t, out=b)
trk = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=64)
trk.submit(frames[0])
trk.collect()
N = 4096
clip = oat_b200.frame_pointers([frames[1 + i % R] for i in range(N)])
trk.run_clip(clip)
for rep in range(3):
    ctx.sync()
    ctx.clip_host_stats()
    t0 = time.perf_counter()
    trk.run_clip(clip)
    dt = time.perf_counter() - t0
    busy, wait, nf = ctx.clip_host_stats()
    print(f"{res} resident engine, {N} frames: wall {1e6 * dt / N:.2f} us/frame ({N / dt:.0f} frames/s); host thread: "
          f"{busy / nf:.2f} us/frame working (descriptors, 2 launches per 32 frames, reading results), {wait / nf:.2f} us/frame waiting for chunks")
# how much of that is work: time the host needs to set up and launch one chunk (32 frames) with nothing to wait for
trk2 = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=8)
trk2.submit(frames[0])
trk2.collect()
M = 2000
ts = tc = 0.0
out = 0
ctx.sync()
t_all0 = time.perf_counter()
for i in range(M):
    t0 = time.perf_counter()
    trk2.submit(frames[1 + i % R])
    ts += time.perf_counter() - t0
    out += 1
    if out == 8:
        t0 = time.perf_counter()
        trk2.collect()
        tc += time.perf_counter() - t0
        out -= 1
while out:
    trk2.collect()
    out -= 1
t_all = time.perf_counter() - t_all0
print(f"{res} per-frame path depth 8: submit {1e6 * ts / M:.1f} us/call, collect {1e6 * tc / (M - 7):.1f} us/call, wall {1e6 * t_all / M:.1f} us/frame")
