#!/usr/bin/env python
"""Kernel-level timing helper (experiments; not the judged bench): times the fused MOG kernel and
the whole frame for a workload, L2 flushed between frames.  OAT_B200_LIB selects a library build."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oat_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--res", default="1080p")
ap.add_argument("--alpha", type=float, default=0.01)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--noflush", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--static", action="store_true", help="blob-free frames only (every pixel takes the fast path)")
args = ap.parse_args()
rows, cols = {"1080p": (1080, 1920), "4k": (2160, 3840), "480p": (480, 640)}[args.res]
ctx = oat_b200.Context(0)
st = torch.cuda.ExternalStream(ctx.stream)
hp = oat_b200.HsvParams.make(h=(40, 80), s=(100, 256), v=(100, 256))
R = 32
frames = [ctx.alloc(rows * cols * 3) for _ in range(R + 1)]
for t, b in enumerate(frames):
    ctx.synth_frame(rows, cols, 1000, 0 if args.static else t, out=b)
trk = oat_b200.Tracker(ctx, rows, cols, args.alpha, hp)
trk.track(frames[0])
for i in range(30):
    trk.track(frames[1 + i % R])
trk.profile(True)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
for i in range(args.steps):
    if not args.noflush:
        ctx.flush_l2()
    ev[i][0].record(st)
    trk.submit(frames[1 + (30 + i) % R])
    ev[i][1].record(st)
    trk.collect()
ctx.sync()
tot = sorted(a.elapsed_time(b) for a, b in ev)
ts = trk.tail_stats()
print('slow groups in last frame:', ts['slow_groups'], '=', ts['slow_groups'] / (rows * cols / 4) * 100, '% of groups; generic-kernel frames:', ts['generic_frames'])
c = ts["cyc"]
print("tail:", {k: ts[k] for k in ("status", "nodes", "replays", "fast")}, "label-CTA cycles since start:",
      [(c[i] - c[0]) & 0xffffffff for i in range(1, 8)])
kms, n = trk.profile_read()
mbar = trk.live_modes() / (rows * cols)
balg = 8 + 40 * mbar
print(f"{args.tag or os.environ.get('OAT_B200_LIB','default')}: {args.res} a={args.alpha} flush={not args.noflush} "
      f"frame median {1e3*tot[len(tot)//2]:.1f} us mean {1e3*sum(tot)/len(tot):.1f} us | fused kernel {1e3*kms:.1f} us "
      f"mbar {mbar:.3f} -> {balg*rows*cols/kms/1e6:.0f} GB/s")
