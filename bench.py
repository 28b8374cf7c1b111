#!/usr/bin/env python
"""bench.py -- frames/s of the MOG + HSV-detect hot path on synthetic 1080p / 4K streams.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1080p|4k|1mp] [--alpha A]
  python bench.py --impl reference ...        # the reference's CPU path (OpenCV) on the host cores

A "step" is ONE frame of ONE video stream through the whole path (framefilt mog -> framefilt col
-C HSV -> posidet hsv: GMM update, zero background, BGR->HSV, inRange, dilate, contour moments,
largest-blob centroid) -- the unit BASELINE.json's metric counts.  Measurements per run:

* value    : device-resident input frames (a ring of distinct synthetic frames in HBM, larger than
             L2), K frames through oat_tracker_run_clip: the RESIDENT engine -- per chunk of up to 32
             frames one launch of the fused kernel and one launch of the tail server work through a
             queue of frame descriptors on the device, no host work per frame; one CUDA-event pair on
             the library's stream around the K frames (every detection is back on the host inside it);
* e2e      : the same frames in pinned HOST memory through the public C-ABI (oat_tracker_submit /
             collect): the H2D copy of every frame and the D2H read of every detection are inside
             the timed region (wall clock, sync on both sides);
* roofline : the resident fused MOG+HSV+threshold kernel alone, 8 independent streams interleaved
             in its queue (8 GMM states > L2, so every frame streams its planes from HBM): launch
             duration from a CUDA-event pair around each launch on the launching stream, algorithmic
             bytes = (5 + 40*m) B/px x pixels x frames in the launch, m = mean live GMM modes
             measured in this run (detect-only mode), against the measured HBM copy
             bandwidth in MEASURED_PEAKS.json;
* extra keys: cold single-frame latency, 8 streams per GPU with detect tails (BASELINE configs 3-4; at 4K
             this is config 4's per-GPU share), the drop-in `framefilt mog` with frame egress
             (8 + 40*m B/px), a multi-blob scene for the detect tail.

N > 1 (torchrun): one independent stream per rank/GPU, no data-path collective ("weak" scaling);
the job time is the max over ranks.  Prints ONE JSON line on rank 0.

Timed region of `value`: W warm-up steps, barrier + synchronize, ONE more untimed chunk of 32 steps, then exactly K
steps between a CUDA-event pair on the library's stream, barrier + synchronize.  The extra chunk is there because a
barrier leaves the GPU idle until the slowest rank has arrived, and a 20-step region (0.3 ms) that starts on a GPU
just back from idling measures the wake-up: 17.1 us per step at N = 2 against 15.3 us with the chunk and 15.1 us at
N = 1, on every rank alike (OAT_BENCH_NO_REWARM=1 leaves it out; DESIGN.md 6, config.timing in the JSON line).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {"1080p": (1080, 1920), "4k": (2160, 3840), "480p": (480, 640), "1mp": (1000, 1000),
             "1mp-tight": (1000, 1024)}  # (diagnostic: the 1000x1000 stream's blob on rows that are whole tiles)
HSV_BAND = dict(h=(40, 80), s=(100, 256), v=(100, 256))
SEED = 1000
METRIC = "frames/s MOG+HSV detect"
FALLBACK_HBM_GBS = 6650.0
CHUNK = 32  # frames per launch of the resident engine (ring of 64 slots per tracker)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=list(WORKLOADS))
    ap.add_argument("--alpha", type=float, default=0.01, help="framefilt mog --adaptation-coeff")
    ap.add_argument("--ring", type=int, default=32, help="distinct synthetic frames cycled as input")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--depth", type=int, default=8, help="frames in flight in the per-frame submit/collect runs (e2e)")
    ap.add_argument("--loop", default="native", choices=["native", "python"],
                    help="value: oat_tracker_run_clip (resident engine) or per-frame submit/collect called from this script")
    ap.add_argument("--streams", type=int, default=8,
                    help="extra measurement: aggregate frames/s of this many independent streams on each GPU (BASELINE configs 3-4); 0 = skip")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra keys (cold latency, multi-stream, egress, multi-blob, config 4)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def workload_name(args):
    r, c = WORKLOADS[args.workload]
    return (f"{c}x{r} synthetic single-blob stream, framefilt mog -a {args.alpha:g} + col HSV + posidet hsv "
            f"(H40-80 S100-256 V100-256, dilate 10)")


def frame_pitch(cols):
    """Row pitch of the device-resident input frames: tight when that keeps rows 16-byte aligned, else padded to it."""
    return (3 * cols + 15) // 16 * 16


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own call sequence on the host cores (oracle/cv2ref.py runs the same
# cv:: functions the reference calls, in its order; the reference cannot be compiled here).
# preallocated frames, warm-up, best of 3 passes, all cores and 1 core,
# S streams = S processes x cores/S threads.
# ---------------------------------------------------------------------------------------------
def _cpu_pass(workload, alpha, ring, nframes, warmup, threads, passes):
    import oracle
    from oracle import cv2ref

    rows, cols = WORKLOADS[workload]
    base = [oracle.synth_frame(rows, cols, SEED, t) for t in range(min(ring, 16))]
    if cv2ref.available():
        import cv2

        cv2.setNumThreads(threads)
        pipe = cv2ref.Pipeline(alpha, **HSV_BAND)
        step = lambda f: pipe.step(f)  # noqa: E731
        note = f"cv2 {cv2.__version__} calls in the reference's order (oracle/cv2ref.py), {threads} thread(s)"
    else:  # C restatement, single thread
        threads = 1
        trk = oracle.Tracker(rows, cols)
        hp = oracle.HsvParams(**HSV_BAND)
        step = lambda f: trk.track(f, alpha, hp)  # noqa: E731
        note = "C restatement oracle/oat_oracle.c, 1 thread (cv2 not importable)"
    for i in range(warmup):
        step(base[i % len(base)].copy())
    frames = [base[i % len(base)].copy() for i in range(min(nframes, 64))]
    best = None
    for _ in range(max(1, passes)):
        t0 = time.perf_counter()
        for i in range(nframes):
            step(frames[i % len(frames)])
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return nframes / best, best, threads, note


def _cpu_worker(q, workload, alpha, ring, nframes, warmup, threads, passes):
    q.put(_cpu_pass(workload, alpha, ring, nframes, warmup, threads, passes))


def cpu_pipeline_fps(args, nframes, warmup, threads=None, passes=1, streams=1):
    """-> (aggregate frames/s, seconds, threads used in total, note).  streams > 1: that many processes, each with
    cores/streams OpenCV threads, timed concurrently (aggregate = streams * nframes / slowest)."""
    cores = os.cpu_count() or 1
    if streams <= 1:
        return _cpu_pass(args.workload, args.alpha, args.ring, nframes, warmup, threads or cores, passes)
    import multiprocessing as mp

    per = max(1, cores // streams)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_cpu_worker, args=(q, args.workload, args.alpha, args.ring, nframes, warmup, per, passes))
             for _ in range(streams)]
    for p_ in procs:
        p_.start()
    res = [q.get() for _ in procs]
    for p_ in procs:
        p_.join()
    slowest = max(r[1] for r in res)
    return streams * nframes / slowest, slowest, per * streams, f"{streams} processes x {per} thread(s); {res[0][3]}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(max(1, args.gpus))))
    rows, cols = WORKLOADS[args.workload]
    # N GPUs serve N streams: the CPU arm runs N streams too (N processes x cores/N threads)
    fps, dt, cores, note = cpu_pipeline_fps(args, args.steps, max(args.warmup, 3), passes=1, streams=world)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": fps,
        "unit": "frames/s",
        "mpix_per_s": fps * rows * cols / 1e6,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "l2": "n/a (CPU)", "streams": world},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{world} stream(s) x {args.steps} frames, one frame per step; {note}"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
            }
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_traffic(workload, alpha):
    """Per-FRAME DRAM bytes of the fused kernel from the committed ncu --set full capture (profiles/traffic.json);
    not measured in this run."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return d.get(f"{workload}_a{alpha:g}")
    except Exception:
        return None


def multi_blob_frames(rows, cols, nblobs, nframes):
    """A busy scene for the detect tail: the stream's static background + `nblobs` discs of the target colour, one per
    cell of a grid, each wandering over its whole cell from frame to frame (a pixel is covered ~7 % of the time, so an
    adapting MOG model keeps calling the discs foreground however long the stream runs).  Host numpy arrays."""
    import numpy as np

    import oracle

    bg = oracle.synth_frame(rows, cols, SEED, 0)
    gx = max(1, int(round((nblobs * cols / rows) ** 0.5)))
    gy = (nblobs + gx - 1) // gx
    cw, ch = cols // gx, rows // gy
    r = max(3, min(ch // 6, cw // 6, rows // 40))
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    out = []
    for t in range(nframes):
        f = bg.copy()
        k = 0
        for j in range(gy):
            for i in range(gx):
                if k >= nblobs:
                    break
                rr = r - (k % 3)
                cx = i * cw + r + (t * 37 + k * 11) % max(1, cw - 2 * r)
                cy = j * ch + r + (t * 23 + k * 7) % max(1, ch - 2 * r)
                win = f[cy - r:cy + r + 1, cx - r:cx + r + 1]
                win[xx ** 2 + yy ** 2 <= rr * rr] = (40, 220, 60)
                k += 1
        out.append(f)
    return [bg] + out


def run_b200(args):
    import torch

    import oat_b200
    from oat_b200 import sharding

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        # one rank per GPU on shared host cores: give every rank its own slice of them
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except Exception:
            pass
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available() or oat_b200.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    rows, cols = WORKLOADS[args.workload]
    npx = rows * cols
    pitch = frame_pitch(cols)
    fbytes = rows * pitch
    K, W = args.steps, args.warmup
    hp = oat_b200.HsvParams.make(**HSV_BAND)
    ctx = oat_b200.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    # one independent stream per rank: stream s -> GPU s mod G, no data-path collective
    seed = sharding.stream_seed(SEED, sharding.streams_for_rank(world, rank, world)[0])

    # ---- input ring: distinct synthetic frames, resident in HBM before timing starts --------
    R = max(2, args.ring)
    dev_frames = [ctx.alloc(fbytes) for _ in range(R)]
    # frame 0 is blob-free and becomes the first model; the ring then cycles t = 1..R
    f0 = ctx.alloc(fbytes)
    ctx.synth_frame(rows, cols, seed, 0, out=f0, pitch=pitch)
    for i, b in enumerate(dev_frames):
        ctx.synth_frame(rows, cols, seed, i + 1, out=b, pitch=pitch)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def new_tracker(ring_depth=2 * CHUNK):
        t_ = oat_b200.Tracker(ctx, rows, cols, args.alpha, hp, ring_depth=ring_depth)
        t_.submit(f0, pitch=pitch)
        t_.collect()
        return t_

    # ---- value: whole-frame throughput, device-resident input --------------------------------------
    # Inputs (ring of R distinct frames, R*6.2 MB at 1080p) are larger than L2, so no flush is needed
    # between steps; the stream's own GMM state is re-read every frame, exactly as in production.
    DEPTH = args.depth
    trk = new_tracker()

    def make_clip(n, start):
        return oat_b200.frame_pointers([dev_frames[(start + i) % R] for i in range(n)])

    det_out = (oat_b200.Detection * max(K, 4 * CHUNK, 1000, max(W, 2 * CHUNK + 8)))()  # (no per-frame Python work in the timed call)

    def run_value(n, start, clip=None):
        if clip is not None:  # the resident engine: one launch of each kernel per chunk of frames
            return trk.run_clip(clip, depth=DEPTH, pitch=pitch, out=det_out)[n - 1]
        out = 0
        last = None
        for i in range(n):
            trk.submit(dev_frames[(start + i) % R], pitch=pitch)
            out += 1
            if out == DEPTH:
                last = trk.collect()
                out -= 1
        while out:
            last = trk.collect()
            out -= 1
        return last

    native = args.loop == "native"
    # warm-up (untimed): the W frames asked for, topped up to two full chunks so that every ring slot, both chunk
    # buffers and the clocks have seen the workload before the timed region starts
    WU = max(W, 2 * CHUNK + 8)
    run_value(WU, 0, make_clip(WU, 0) if native else None)
    modes_before = trk.live_modes() / npx
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = ev(), ev()
    clip = make_clip(K, W) if native else None
    run_value(4 * CHUNK, WU, make_clip(4 * CHUNK, WU) if native else None)  # (untimed) the GPU is busy right up to the barrier
    barrier()
    # A barrier leaves every GPU idle until the slowest rank has arrived (milliseconds at N > 1), and what follows an
    # idle GPU starts slowly: measured at N = 2, 20 steps, 17.1 us per step right after the barrier, 18.1 us after
    # 2 ms more of idling, 15.1 us when the GPU is busy up to the start (N = 1, where the barrier is immediate).  A
    # 20-step region is 0.3 ms, so that wake-up would be a tenth of it.  One more untimed chunk of warm-up steps runs
    # between the barrier and the timed steps: every rank does the same, at the same time, and the events bracket
    # exactly the K timed steps.
    if os.environ.get("OAT_BENCH_NO_REWARM") is None:
        run_value(CHUNK, WU + 4 * CHUNK, make_clip(CHUNK, WU + 4 * CHUNK) if native else None)
    ctx.clip_host_stats()
    launches0 = ctx.kernel_launches
    cpu0 = time.process_time()
    wall0 = time.perf_counter()
    e0.record(stream)
    last = run_value(K, W, clip)  # returns with every detection on the host
    e1.record(stream)
    if native:
        last = oat_b200.Detection.from_buffer_copy(last)  # (det_out is reused by the runs below)
    host_s = time.perf_counter() - wall0
    cpu_s = time.process_time() - cpu0
    busy_us, wait_us, clip_n = ctx.clip_host_stats()
    barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.kernel_launches - launches0
    total_ms = e0.elapsed_time(e1)
    modes_after = trk.live_modes() / npx
    mbar = 0.5 * (modes_before + modes_after)
    tail = trk.tail_stats()
    # a longer run of the same thing, so that the clock sampler (NVML, 2 ms period) sees the GPU under this load
    if K < 1000:
        run_value(1000, W + K, make_clip(1000, W + K) if native else None)
    sampler.stop()

    extras = {}
    if not args.no_extras:
        # ---- cold single-frame latency: L2 flushed, one frame in flight (submit -> collect) -------------
        NL = 100
        lat = []
        for i in range(NL):
            ctx.flush_l2()
            a_, b_ = ev(), ev()
            a_.record(stream)
            trk.submit(dev_frames[(W + K + i) % R], pitch=pitch)
            trk.collect()
            b_.record(stream)
            lat.append((a_, b_))
        barrier()
        lat_ms = sorted(x.elapsed_time(y) for x, y in lat)
        extras["cold_frame"] = {"latency_ms": lat_ms[len(lat_ms) // 2],
                                "note": "median submit->collect of one frame in flight, L2 flushed before it (fused kernel + "
                                        "detect tail + result read-back, serialised, per-frame launches)"}
    trk.close()

    # ---- roofline of the resident fused kernel ---------------------------------------------------------
    # S independent streams of the same workload interleaved in ONE queue of the resident kernel (fused kernel ONLY:
    # nothing else runs on the GPU); the S states (S * ~45 MB live at 1080p) exceed L2, so every frame streams its
    # planes from HBM without an explicit flush.  Every launch is bracketed by a CUDA-event pair on the launching
    # stream inside the library (oat_ctx_profile_resident).
    S = 8
    trks = [new_tracker() for _ in range(S)]
    NFR = CHUNK  # frames per stream per launch

    def fused_clip(start, n):
        return oat_b200.frame_pointers([dev_frames[(start + i) % R] for i in range(n) for _ in range(S)])

    oat_b200.Tracker.run_clips(trks, fused_clip(0, NFR), fused_only=True, pitch=pitch)  # warm-up: one whole launch
    m0 = sum(t_.live_modes() for t_ in trks) / (S * npx)
    fc = fused_clip(NFR, 2 * NFR)
    barrier()
    ctx.profile_resident(True)
    oat_b200.Tracker.run_clips(trks, fc, fused_only=True, pitch=pitch)
    kern_total_ms, kern_n, kern_frames = ctx.profile_resident_read()
    ctx.profile_resident(False)
    barrier()
    m1 = sum(t_.live_modes() for t_ in trks) / (S * npx)
    mbar_roof = 0.5 * (m0 + m1)
    kern_ms = kern_total_ms / max(1, kern_n)            # average launch duration
    frames_per_launch = kern_frames / max(1, kern_n)
    frame_ms = kern_total_ms / max(1, kern_frames)      # time per frame inside the resident kernel

    # ---- many independent streams per GPU (each its own GMM state), with detect tails -------------------
    if args.streams > 1 and not args.no_extras:
        S2 = args.streams
        mt = trks[:S2] if S2 <= S else trks + [new_tracker() for _ in range(S2 - S)]
        nfr = max(CHUNK, min(4 * CHUNK, K // S2))
        mc = oat_b200.frame_pointers([dev_frames[(7 + i) % R] for i in range(nfr) for _ in range(S2)])
        oat_b200.Tracker.run_clips(mt, oat_b200.frame_pointers([dev_frames[i % R] for i in range(2 * CHUNK) for _ in range(S2)]), pitch=pitch)
        m0_, m1_ = ev(), ev()
        barrier()
        m0_.record(stream)
        oat_b200.Tracker.run_clips(mt, mc, pitch=pitch)
        m1_.record(stream)
        barrier()
        multi_ms = sharding.max_over_ranks([m0_.elapsed_time(m1_)], dist, dev)[0]
        extras["multi_stream"] = {
            "streams_per_gpu": S2, "value": world * S2 * nfr / (multi_ms * 1e-3), "unit": "frames/s",
            "note": f"{S2} independent streams per GPU ({S2 * 45 * npx / 1e6:.0f} MB of live GMM state per GPU) interleaved in one "
                    f"queue of the resident engine, detect tails included, {nfr} frames per stream"}
        for t_ in mt[S:]:
            t_.close()
    for t_ in trks:
        t_.close()

    if not args.no_extras:
        # ---- the drop-in `framefilt mog` with frame egress: 8 + 40*m B/px --------------------------
        mogs = [oat_b200.BackgroundSubtractorMOG(ctx, rows, cols, args.alpha) for _ in range(S)]
        outs = [ctx.alloc(fbytes) for _ in range(2)]
        for m_ in mogs:
            m_.apply_async(f0, outs[0], pitch=pitch)
        for i in range(6):
            for m_ in mogs:
                m_.apply_async(dev_frames[i % R], outs[i & 1], pitch=pitch)
        NE = 6
        g0, g1 = ev(), ev()
        barrier()
        mm0 = sum(m_.live_modes() for m_ in mogs) / (S * npx)
        g0.record(stream)
        for i in range(NE):
            for m_ in mogs:
                m_.apply_async(dev_frames[(6 + i) % R], outs[i & 1], pitch=pitch)
        g1.record(stream)
        barrier()
        mm1 = sum(m_.live_modes() for m_ in mogs) / (S * npx)
        eg_ms = g0.elapsed_time(g1) / (NE * S)
        b_eg = 8.0 + 40.0 * 0.5 * (mm0 + mm1)
        peak_, _ = load_peak()
        extras["framefilt_mog_egress"] = {
            "kernel_ms": eg_ms, "algorithmic_bytes_per_px": b_eg, "achieved_gbs": b_eg * npx / (eg_ms * 1e-3) / 1e9,
            "frac": b_eg * npx / (eg_ms * 1e-3) / 1e9 / peak_, "frames_per_s": 1e3 / eg_ms,
            "note": f"oat_mog_apply_async (the drop-in framefilt mog: MOG2 + setTo(0), filtered BGR frame written to HBM), {S} models "
                    f"round-robin (states > L2), one launch per frame, {NE * S} back-to-back launches between one CUDA-event pair"}
        for m_ in mogs:
            m_.close()
        for o_ in outs:
            o_.free()

        # ---- the detect tail under load: ~60 blobs per frame -------------------------------------------------------------
        try:
            import numpy as np

            host = multi_blob_frames(rows, cols, 60, 24)
            mb = []
            for f in host:
                b_ = ctx.alloc(fbytes)
                pad = np.zeros((rows, pitch), np.uint8)
                pad[:, :cols * 3] = f.reshape(rows, cols * 3)
                b_.upload(pad)
                mb.append(b_)
            tb = oat_b200.Tracker(ctx, rows, cols, args.alpha, hp, ring_depth=2 * CHUNK)
            tb.submit(mb[0], pitch=pitch)
            tb.collect()
            nb = max(2 * CHUNK, min(K, 512))
            bc = oat_b200.frame_pointers([mb[1 + i % (len(mb) - 1)] for i in range(nb)])
            for _ in range(3):  # (untimed: the stream's tail-load estimate settles, its pre-labelling pools get allocated)
                tb.run_clip(bc, pitch=pitch)
            b0, b1 = ev(), ev()
            barrier()
            b0.record(stream)
            dl = tb.run_clip(bc, pitch=pitch)[-1]
            b1.record(stream)
            barrier()
            ts_ = tb.tail_stats()
            extras["multi_blob"] = {"value": nb / (b0.elapsed_time(b1) * 1e-3), "unit": "frames/s", "blobs_per_frame": 60,
                                    "components_last": dl.n_components, "run_table_entries": ts_["nodes"], "replays": ts_["replays"],
                                    "note": "same path, every frame carries ~60 discs of the target colour (the detect tail, not the fused kernel, has the work)"}
            tb.close()
            for b_ in mb:
                b_.free()
        except Exception as e:  # pragma: no cover
            extras["multi_blob"] = {"error": repr(e)}

    # ---- e2e: pinned host frames through submit/collect, copies inside the timed region ------
    HR = min(R, 8)
    pin = oat_b200.PinnedArray((HR, rows, pitch))
    for i in range(HR):
        ctx.memcpy(pin.ptr + i * fbytes, dev_frames[i], fbytes)
    trk2 = oat_b200.Tracker(ctx, rows, cols, args.alpha, hp, ring_depth=DEPTH)
    trk2.submit(f0, pitch=pitch)
    trk2.collect()
    for i in range(max(W, 2 * DEPTH + 4)):  # warm-up (untimed): every ring slot has staged a frame, pipelined like the timed loop
        trk2.submit(pin.ptr + (i % HR) * fbytes, pitch=pitch)
        if i >= DEPTH - 1:
            trk2.collect()
    for _ in range(DEPTH - 1):
        trk2.collect()
    barrier()
    if os.environ.get("OAT_BENCH_NO_REWARM") is None:  # (untimed: the GPU idled in the barrier, see config.timing)
        for i in range(DEPTH):
            trk2.submit(pin.ptr + (i % HR) * fbytes, pitch=pitch)
            trk2.collect()
    t0 = time.perf_counter()
    outstanding = 0
    for i in range(K):
        trk2.submit(pin.ptr + (i % HR) * fbytes, pitch=pitch)
        outstanding += 1
        if outstanding == DEPTH:
            trk2.collect()
            outstanding -= 1
    while outstanding:
        trk2.collect()
        outstanding -= 1
    ctx.sync()
    e2e_local = time.perf_counter() - t0
    barrier()
    e2e_s = time.perf_counter() - t0
    trk2.close()

    # ---- config 4's per-GPU share: 8 concurrent 4K streams on this GPU (64 x 4K over 8 GPUs) ----------------------
    if not args.no_extras and args.workload == "1080p":
        try:
            r4, c4 = WORKLOADS["4k"]
            for b in dev_frames:
                b.free()
            dev_frames = []
            R4 = 8
            fr4 = [ctx.alloc(r4 * c4 * 3) for _ in range(R4 + 1)]
            for i, b in enumerate(fr4):
                ctx.synth_frame(r4, c4, seed, i, out=b)
            t4 = [oat_b200.Tracker(ctx, r4, c4, args.alpha, hp, ring_depth=16) for _ in range(8)]
            for t_ in t4:
                t_.submit(fr4[0])
                t_.collect()
            n4 = 24
            # warm-up: two chunks, so that both halves of the engine's buffers exist before the timed region
            oat_b200.Tracker.run_clips(t4, oat_b200.frame_pointers([fr4[1 + i % R4] for i in range(16) for _ in range(8)]))
            c4p = oat_b200.frame_pointers([fr4[1 + (3 + i) % R4] for i in range(n4) for _ in range(8)])
            q_all = []
            for _ in range(3):  # three passes, each between its own event pair; the median is reported, all are listed
                q0, q1 = ev(), ev()
                barrier()
                q0.record(stream)
                oat_b200.Tracker.run_clips(t4, c4p)
                q1.record(stream)
                barrier()
                q_all.append(sharding.max_over_ranks([q0.elapsed_time(q1)], dist, dev)[0])
            q_ms = sorted(q_all)[1]
            extras["config4_8x4k_per_gpu"] = {
                "streams": 8 * world, "streams_per_gpu": 8, "value": world * 8 * n4 / (q_ms * 1e-3), "unit": "frames/s",
                "mpix_per_s": world * 8 * n4 * r4 * c4 / (q_ms * 1e-3) / 1e6, "ms_per_pass": q_all,
                "note": f"BASELINE config 4 (64 concurrent 4K streams over 8 GPUs): 8 independent 3840x2160 streams per GPU in one queue of "
                        f"the resident engine, detect tails included, {n4} frames per stream per pass (median of 3 passes, max over ranks each); "
                        f"whole-job aggregate over {world} GPU(s)"}
            for t_ in t4:
                t_.close()
        except Exception as e:  # pragma: no cover
            extras["config4_8x4k_per_gpu"] = {"error": repr(e)}

    # ---- reduce over ranks: the job took as long as its slowest rank -------------------------
    per_rank_ms = [total_ms]
    if dist is not None:
        t_all = torch.zeros(world, dtype=torch.float64, device=dev)
        t_all[rank] = total_ms
        dist.all_reduce(t_all)
        per_rank_ms = [float(x) for x in t_all.tolist()]
    total_ms, e2e_s, kern_ms, frame_ms, host_s = sharding.max_over_ranks([total_ms, e2e_s, kern_ms, frame_ms, host_s], dist, dev)
    launches = sharding.sum_over_ranks(launches, dist, dev)
    if rank == 0:
        peak, peak_src = load_peak()
        fps = world * K / (total_ms * 1e-3)
        # detect-only fused mode (no frame egress) moves 5 + 40*m bytes per pixel
        # (a frozen model, -a 0, is compared but never written back: 4 + 20*m read, nothing but the threshold bits written)
        b_alg = 5.0 + 40.0 * mbar_roof if args.alpha != 0 else 4.0 + 20.0 * mbar_roof
        achieved = b_alg * npx / (frame_ms * 1e-3) / 1e9 if frame_ms > 0 else 0.0
        traffic = load_traffic(args.workload, args.alpha)
        line = {
            "metric": METRIC,
            "value": fps,
            "unit": "frames/s",
            "mpix_per_s": fps * npx / 1e6,
            "n_gpus": world,
            "steps": K,
            "warmup": W,
            "ms_per_step": total_ms / K,
            "ms_per_step_per_rank": [t / K for t in per_rank_ms],
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32",
            "data": "synthetic",
            "config": {
                "workload": workload_name(args),
                "streams_per_gpu": 1,
                "input": f"ring of {R} distinct device-resident frames ({R * fbytes / 1e6:.0f} MB)",
                "l2": f"no flush: the inputs ({R * fbytes / 1e6:.0f} MB ring) are larger than the 126 MB L2; one CUDA-event pair on the "
                      f"launching stream around the {K} frames; the stream's own GMM state (~{45 * npx / 1e6:.0f} MB) stays L2-resident between "
                      "its frames, as it does in production",
                "mean_live_modes": mbar,
                "loop": (f"oat_tracker_run_clip: resident engine, chunks of {CHUNK} frames, one fused-kernel launch + one tail-server launch per chunk, "
                         "no host work per frame" if native else "submit/collect called per frame from this script"),
                "parallelism": f"{world} independent stream(s), one per GPU, no collective",
                "timing": (f"{W} warm-up steps (topped up to {WU}), barrier + synchronize, one more untimed chunk of {CHUNK} steps "
                           "(a GPU that idled in the barrier starts slowly: DESIGN.md 6), then exactly the timed steps between one "
                           "CUDA-event pair per rank, barrier + synchronize, max over ranks"
                           if os.environ.get("OAT_BENCH_NO_REWARM") is None else
                           f"{W} warm-up steps (topped up to {WU}), barrier + synchronize, the timed steps between one CUDA-event pair per rank, "
                           "barrier + synchronize, max over ranks"),
            },
            "roofline": {
                "bound": "hbm",
                "kernel": "mog_stream_kernel<5,false,true>" if cols % 32 == 0 else "mog_stream_kernel<5,false,false>",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "peak_source": peak_src,
                "algorithmic_bytes_per_px": b_alg,
                "algorithmic_bytes_formula": "5 + 40*m (BGR 3 + count 1+1 + state 20*m read and written)" if args.alpha != 0 else "4 + 20*m (frozen model: read only)",
                "algorithmic_bytes_per_launch": b_alg * npx * frames_per_launch,
                "kernel_ms": kern_ms,
                "frames_per_launch": frames_per_launch,
                "ms_per_frame": frame_ms,
                "kernel_launches_timed": kern_n,
                "how": f"{kern_n} launches of the resident fused kernel (fused kernel only), each working through a queue of {frames_per_launch:.0f} "
                       f"frames = {S} independent streams x {NFR} frames interleaved; a CUDA-event pair on the launching stream around each launch; "
                       f"{S} GMM states (resident {S * 45 * npx / 1e6:.0f} MB live planes) > L2, no flush; consecutive launches overlap tile by tile, "
                       "so the brackets partition the timeline",
                "mean_live_modes": mbar_roof,
                "traffic": traffic * frames_per_launch if traffic else None,
                "traffic_source": "profiles/traffic.json (ncu --set full capture, DRAM bytes per frame x frames per launch); not measured in this run",
            },
            "e2e": {
                "value": world * K / e2e_s,
                "unit": "frames/s",
                "h2d_bytes_per_step": fbytes,
                "d2h_bytes_per_step": 88,
                "note": f"pinned host frames via oat_tracker_submit/collect, ring depth {DEPTH}, wall clock (rank 0 alone: {K / e2e_local:.0f} frames/s)"
                        + ("" if os.environ.get("OAT_BENCH_NO_REWARM") is not None else f"; {DEPTH} untimed frames between the opening barrier and the timed ones (config.timing)"),
            },
            "host": {"busy_us_per_frame": busy_us / max(1, clip_n), "wait_us_per_frame": wait_us / max(1, clip_n),
                     "call_us_per_frame": 1e6 * host_s / K, "cpu_us_per_frame": 1e6 * cpu_s / K,
                     "note": "rank 0, the timed oat_tracker_run_clip call: time the calling thread spent working (descriptors, launches, reading results) "
                             "and waiting for chunks (it spins in cudaEventSynchronize, so process CPU time ~ wall time), per frame"},
            "tail": {"one_launch": bool(tail["fast"]), "run_table_entries": tail["nodes"], "replays": tail["replays"],
                     "resident_frames": tail["clip_frames"]},
            "gpu_launches": launches,
            "clocks": sampler.summary(),
            "wall_s_timed_region": wall,
            "last_detection": list(last.as_tuple()) if last is not None else None,
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            try:
                nf = args.cpu_frames or (240 if args.workload != "4k" else 80)
                cfps, cdt, cores, note = cpu_pipeline_fps(args, nf, 20, passes=3)
                line["cpu_baseline"] = {"value": cfps, "unit": "frames/s", "cores": cores, "kind": "port",
                                        "sample": f"{nf} frames of the same stream, best of 3 passes ({cdt:.1f} s); {note}"}
                c1, d1, _, _ = cpu_pipeline_fps(args, max(30, nf // 4), 5, threads=1, passes=3)
                line["cpu_baseline"]["one_thread"] = {"value": c1, "unit": "frames/s", "cores": 1,
                                                      "sample": f"{max(30, nf // 4)} frames, best of 3 passes ({d1:.1f} s), cv2.setNumThreads(1)"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
