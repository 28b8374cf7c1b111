"""Where a clip's time goes: wall time of oat_tracker_run_clip (device-resident frames, resident engine) against the
clip length -- the intercept is the fixed cost of a call (descriptor set-up, two launches, the last frame's detect
tail, the completion wait), the slope the steady-state time per frame -- and the host CPU time per frame."""
import sys
import time

sys.path.insert(0, ".")
import oat_b200

rows, cols = (1080, 1920) if len(sys.argv) < 2 or sys.argv[1] == "1080p" else (2160, 3840)
ctx = oat_b200.Context(0)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
R = 32
frames = [ctx.alloc(rows * cols * 3) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, t, out=b)
trk = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=64)
trk.submit(frames[0])
trk.collect()
trk.run_clip(oat_b200.frame_pointers([frames[1 + i % R] for i in range(64)]))
print(f"{cols}x{rows}: frames  wall_us  us/frame  cpu_us/frame")
for n in (1, 2, 4, 8, 16, 20, 32, 64, 128, 512, 2048):
    clip = oat_b200.frame_pointers([frames[1 + i % R] for i in range(n)])
    best, cpu = None, None
    for rep in range(7):
        ctx.sync()
        c0 = time.process_time()
        t0 = time.perf_counter()
        trk.run_clip(clip)
        dt = time.perf_counter() - t0
        dc = time.process_time() - c0
        if best is None or dt < best:
            best, cpu = dt, dc
    print(f"  {n:6d} {best * 1e6:9.1f} {best * 1e6 / n:9.2f} {cpu * 1e6 / n:9.2f}")
trk.close()
ctx.close()
