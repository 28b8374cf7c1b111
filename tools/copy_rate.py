#!/usr/bin/env python
"""Copy-engine rates that decide how frames are staged (diagnostic, cuda-python): linear vs strided H2D from pinned
memory, strided D2D re-pitch, and the host cost of issuing each.  usage: copy_rate.py"""
import time

from cuda.bindings import runtime as rt


def ck(r):
    if isinstance(r, tuple):
        err, *rest = r
        assert err == rt.cudaError_t.cudaSuccess, err
        return rest[0] if len(rest) == 1 else rest
    assert r == rt.cudaError_t.cudaSuccess, r


def timed(stream, fn, n=200):
    e0, e1 = ck(rt.cudaEventCreate()), ck(rt.cudaEventCreate())
    for _ in range(10):
        fn()
    ck(rt.cudaStreamSynchronize(stream))
    ck(rt.cudaEventRecord(e0, stream))
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    host = (time.perf_counter() - t0) / n
    ck(rt.cudaEventRecord(e1, stream))
    ck(rt.cudaStreamSynchronize(stream))
    ms = ck(rt.cudaEventElapsedTime(e0, e1))
    return ms * 1e3 / n, host * 1e6


ck(rt.cudaSetDevice(0))
stream = ck(rt.cudaStreamCreate())
K = rt.cudaMemcpyKind
for name, rows, cols in (("1mp", 1000, 1000), ("1080p", 1080, 1920), ("1000x1080p-ragged", 1080, 1000)):
    rb = 3 * cols
    tight = (rb + 15) // 16 * 16
    h = ck(rt.cudaMallocHost(rb * rows))
    d_raw = ck(rt.cudaMalloc(rb * rows))
    d_al = ck(rt.cudaMalloc(tight * rows))
    res = {}
    res["H2D linear"] = timed(stream, lambda: ck(rt.cudaMemcpyAsync(d_raw, h, rb * rows, K.cudaMemcpyHostToDevice, stream)))
    res["H2D strided (pitch %d -> %d)" % (rb, tight)] = timed(stream, lambda: ck(rt.cudaMemcpy2DAsync(d_al, tight, h, rb, rb, rows, K.cudaMemcpyHostToDevice, stream)))
    res["D2D linear"] = timed(stream, lambda: ck(rt.cudaMemcpyAsync(d_al, d_raw, rb * rows, K.cudaMemcpyDeviceToDevice, stream)))
    res["D2D strided (pitch %d -> %d)" % (rb, tight)] = timed(stream, lambda: ck(rt.cudaMemcpy2DAsync(d_al, tight, d_raw, rb, rb, rows, K.cudaMemcpyDeviceToDevice, stream)))
    print(f"{name}: {rows} x {cols} x 3 = {rb * rows / 1e6:.2f} MB")
    for k, (dev_us, host_us) in res.items():
        print(f"   {k:38s} {dev_us:8.2f} us on the device ({rb * rows / dev_us / 1e3:7.1f} GB/s), {host_us:6.2f} us of host time per call")
    ck(rt.cudaFreeHost(h))
    ck(rt.cudaFree(d_raw))
    ck(rt.cudaFree(d_al))

# two copy streams: does a second DMA stream hide the per-copy start-up gaps of back-to-back H2D copies?
rows, cols = 1080, 1920
nb = rows * cols * 3
hs = [ck(rt.cudaMallocHost(nb)) for _ in range(4)]
ds = [ck(rt.cudaMalloc(nb)) for _ in range(4)]
s2 = ck(rt.cudaStreamCreate())
for label, streams in (("one stream", [stream]), ("two streams, frames alternating", [stream, s2])):
    e0, e1 = ck(rt.cudaEventCreate()), ck(rt.cudaEventCreate())
    for st in streams:
        ck(rt.cudaStreamSynchronize(st))
    n = 400
    t0 = time.perf_counter()
    for i in range(n):
        ck(rt.cudaMemcpyAsync(ds[i % 4], hs[i % 4], nb, K.cudaMemcpyHostToDevice, streams[i % len(streams)]))
    for st in streams:
        ck(rt.cudaStreamSynchronize(st))
    dt = time.perf_counter() - t0
    print(f"1080p H2D, {label}: {n * nb / dt / 1e9:6.1f} GB/s ({1e6 * dt / n:6.1f} us per frame)")
for label, parts in (("halves on two streams", 2),):
    n = 400
    t0 = time.perf_counter()
    for i in range(n):
        for k in range(parts):
            off = k * (nb // parts)
            ck(rt.cudaMemcpyAsync(ds[i % 4] + off, hs[i % 4] + off, nb // parts, K.cudaMemcpyHostToDevice, [stream, s2][k]))
    ck(rt.cudaStreamSynchronize(stream))
    ck(rt.cudaStreamSynchronize(s2))
    dt = time.perf_counter() - t0
    print(f"1080p H2D, {label}: {n * nb / dt / 1e9:6.1f} GB/s ({1e6 * dt / n:6.1f} us per frame)")

# both directions at once (a frame filter moves every frame up AND down): do H2D and D2H in flight together each get the link?
hs2 = [ck(rt.cudaMallocHost(nb)) for _ in range(4)]
for label, up, down in (("H2D alone", True, False), ("D2H alone", False, True), ("H2D + D2H together", True, True)):
    ck(rt.cudaStreamSynchronize(stream))
    ck(rt.cudaStreamSynchronize(s2))
    n = 400
    t0 = time.perf_counter()
    for i in range(n):
        if up:
            ck(rt.cudaMemcpyAsync(ds[i % 4], hs[i % 4], nb, K.cudaMemcpyHostToDevice, stream))
        if down:
            ck(rt.cudaMemcpyAsync(hs2[i % 4], ds[(i + 2) % 4], nb, K.cudaMemcpyDeviceToHost, s2))
    ck(rt.cudaStreamSynchronize(stream))
    ck(rt.cudaStreamSynchronize(s2))
    dt = time.perf_counter() - t0
    print(f"1080p {label}: {n * nb * (int(up) + int(down)) / dt / 1e9:6.1f} GB/s in total ({1e6 * dt / n:6.1f} us per frame)")
