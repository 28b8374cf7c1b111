#!/usr/bin/env python
"""The detect tail under load (diagnostic): a scene with N blobs per frame through the resident engine -- frames/s, and
the labelling CTA's SM-clock stamps of the last frame (start, ticket, staged, counted, filled, merged, holes, end).
usage: tail_probe.py [--blobs 60] [--res 1080p] [--frames 256]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import oat_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blobs", type=int, nargs="+", default=[1, 8, 60, 200])
ap.add_argument("--res", default="1080p")
ap.add_argument("--frames", type=int, default=256)
args = ap.parse_args()
rows, cols = bench.WORKLOADS[args.res]
ctx = oat_b200.Context(0)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
for nb in args.blobs:
    host = bench.multi_blob_frames(rows, cols, nb, 24)
    bufs = []
    for f in host:
        b = ctx.alloc(rows * cols * 3)
        b.upload(f)
        bufs.append(b)
    trk = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=64)
    trk.submit(bufs[0])
    trk.collect()
    clip = oat_b200.frame_pointers([bufs[1 + i % 24] for i in range(args.frames)])
    trk.run_clip(clip)
    ctx.sync()
    t0 = time.perf_counter()
    d = trk.run_clip(clip)[-1]
    dt = time.perf_counter() - t0
    s = trk.tail_stats()
    cyc = s["cyc"]
    names = ["ticket", "extents", "counted", "staged+filled", "merged", "holes", "end"]
    deltas = [((cyc[i + 1] - cyc[i]) & 0xffffffff) / 1965.0 for i in range(7)]
    print(f"{nb:4d} blobs: {args.frames / dt:9.0f} frames/s ({1e6 * dt / args.frames:6.1f} us/frame); last frame: {d.n_components} components, "
          f"{s['nodes']} run-table entries, status {s['status']}, replays {s['replays']}; labelling CTA (us): "
          + ", ".join(f"{n} {v:.1f}" for n, v in zip(names, deltas)) + f"  total {sum(deltas):.1f}")
    # the same clip through the fused kernel alone (no detect tail): what the tail is being measured against
    trk2 = oat_b200.Tracker(ctx, rows, cols, 0.01, hp, ring_depth=64)
    trk2.submit(bufs[0])
    trk2.collect()
    oat_b200.Tracker.run_clips([trk2], clip, fused_only=True)
    ctx.sync()
    t0 = time.perf_counter()
    oat_b200.Tracker.run_clips([trk2], clip, fused_only=True)
    ctx.sync()
    dt2 = time.perf_counter() - t0
    print(f"            fused kernel alone on this scene: {args.frames / dt2:9.0f} frames/s ({1e6 * dt2 / args.frames:6.1f} us/frame), "
          f"mean live modes {trk2.live_modes() / (rows * cols):.3f}")
    trk2.close()
    trk.close()
    for b in bufs:
        b.free()
