#!/usr/bin/env python
"""Driver for ncu captures of the two resident kernels (not the judged bench): S independent streams x N frames through
ONE launch of the resident engine, fused kernel only (--fused-only: the launch ncu profiles as mog_stream_kernel) or with
the tail server (tail_stream_kernel).  Short queues on purpose: ncu --set full replays a launch ~40 times.
OAT_B200_LIB selects a library build."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oat_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", default="1080p")
ap.add_argument("--alpha", type=float, default=0.01)
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--frames", type=int, default=4, help="frames per stream per launch (ring = 2x)")
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--fused-only", action="store_true")
args = ap.parse_args()
rows, cols = {"1080p": (1080, 1920), "4k": (2160, 3840), "480p": (480, 640), "1mp": (1000, 1000)}[args.res]
pitch = (3 * cols + 15) // 16 * 16
ctx = oat_b200.Context(0)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
R = 16
frames = [ctx.alloc(rows * pitch) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, t, out=b, pitch=pitch)
S = args.streams
trks = [oat_b200.Tracker(ctx, rows, cols, args.alpha, hp, ring_depth=2 * args.frames) for _ in range(S)]
for t_ in trks:
    t_.submit(frames[0], pitch=pitch)
    t_.collect()
k = 0
for _ in range(args.launches + 2):  # two warm-up launches, then the ones to capture (ncu -s skips the warm-up)
    clip = oat_b200.frame_pointers([frames[1 + (k + i) % R] for i in range(args.frames) for _ in range(S)])
    k += args.frames
    oat_b200.Tracker.run_clips(trks, clip, fused_only=args.fused_only, pitch=pitch)
ctx.sync()
m = sum(t_.live_modes() for t_ in trks) / (S * rows * cols)
print(f"{args.res}: {S} streams x {args.frames} frames per launch, mean live modes {m:.4f}, "
      f"algorithmic bytes per launch {(5 + 40 * m) * rows * cols * S * args.frames / 1e6:.1f} MB (detect-only)")
