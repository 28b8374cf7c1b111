"""End-to-end dataflow through real shared memory on the GPU box, one OS process per node exactly like an
Oat graph (README.md:125-149):

    oat frameserve synth raw | oat framefilt mog raw filt | oat framefilt col filt hsv -C HSV |
    oat posidet hsv hsv pos -H .. -S .. -V .. | oat posisock std pos

and the fused one-component form (oat posidet track raw pos).  Positions (JSON lines, the byte format of
PositionCout + serializePosition) are compared with the CPU oracle on the same synthetic frames."""
import json
import os
import subprocess

import pytest

import oracle
from graph_util import wait_sources

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oat_b200", "bin")
ROWS, COLS, N = 240, 320, 14
BAND = dict(h=(40, 80), s=(100, 256), v=(100, 256))
HSV_ARGS = ["-H", "[40,80]", "-S", "[100,256]", "-V", "[100,256]"]


def oracle_positions(lr):
    trk = oracle.Tracker(ROWS, COLS)
    hp = oracle.HsvParams(**BAND)
    out = []
    for t in range(N):
        o, _ = trk.track(oracle.synth_frame(ROWS, COLS, 1000, t), lr, hp)
        out.append(o)
    return out


def used_addresses(names, argvs):
    """the addresses of `names` the graph uses (the last one is the socket's): each needs its SOURCE attached before
    frames flow"""
    flat = [a for argv in argvs for a in argv]
    return [n for n in names if n in flat or n == names[-1]]


def run_graph(tag, nodes, serve_args=(), rows=ROWS, cols=COLS, n=N, server=None):
    """nodes: list of argv lists; consumers are started first, the frame server last (examples/*/*.sh).
    server: None = `frameserve synth`, else a function (raw address) -> argv of the frame server."""
    names = [f"oatb200pipe_{tag}_{n}" for n in ("raw", "filt", "hsv", "pos")]
    subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)
    procs = []
    try:
        sock = subprocess.Popen([os.path.join(BIN, "oat-posisock"), "std", names[3]], stdout=subprocess.PIPE, text=True)
        for argv in nodes(names):
            procs.append(subprocess.Popen([os.path.join(BIN, argv[0])] + argv[1:], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                          text=True))
        wait_sources(*used_addresses(names, nodes(names)))
        if server is None:
            serve = subprocess.Popen([os.path.join(BIN, "oat-frameserve"), "synth", names[0], "--rows", str(rows), "--cols", str(cols),
                                      "--num-samples", str(n), "--fps", "100"] + list(serve_args))
        else:
            serve = subprocess.Popen([os.path.join(BIN, "oat-frameserve")] + server(names[0]) + list(serve_args))
        out, _ = sock.communicate(timeout=120)
        assert serve.wait(timeout=30) == 0
        for p in procs:
            so, se = p.communicate(timeout=30)
            assert p.returncode == 0, (p.args, so, se)  # END propagates downstream; every node exits 0
        return [json.loads(line) for line in out.splitlines() if line.strip()]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)


def check(positions, want):
    assert len(positions) == N
    for t, (p, o) in enumerate(zip(positions, want)):
        assert p["tick"] == t + 1  # the frame server's Sample travels frame -> frame -> position
        assert p["usec"] == 10000 * (t + 1)
        assert p["pos_ok"] == bool(o.position_valid), t
        if o.position_valid:
            assert abs(p["pos_xy"][0] - o.x) < 1e-4 and abs(p["pos_xy"][1] - o.y) < 1e-4, (t, p, o.x, o.y)  # 5 decimals in JSON


@pytest.mark.parametrize("lr", [0.0, 0.05])
def test_three_component_chain(lr):
    def nodes(n):
        return [["oat-posidet", "hsv", n[2], n[3]] + HSV_ARGS,
                ["oat-framefilt", "col", n[1], n[2], "-C", "HSV"],
                ["oat-framefilt", "mog", n[0], n[1], "-a", str(lr)]]

    check(run_graph(f"chain{int(lr * 100)}", nodes), oracle_positions(lr))


def test_device_resident_chain():
    """The north star's device-resident Frame variant: the frame server generates frames ON the GPU and
    publishes them through a CUDA IPC handle in the shared frame header (memory kind DEVICE); mog and col
    publish device frames too (--device-sink), so pixels never touch host memory between the four processes."""
    def nodes(n):
        return [["oat-posidet", "hsv", n[2], n[3]] + HSV_ARGS,
                ["oat-framefilt", "col", n[1], n[2], "-C", "HSV", "--device-sink"],
                ["oat-framefilt", "mog", n[0], n[1], "-a", "0.05", "--device-sink"]]

    check(run_graph("devchain", nodes, serve_args=("--device",)), oracle_positions(0.05))


def test_fused_tracker_component():
    def nodes(n):
        return [["oat-posidet", "track", n[0], n[3], "-A", "0.05"] + HSV_ARGS]

    check(run_graph("fused", nodes), oracle_positions(0.05))


def test_config0_bsub_chain_analytic():
    """BASELINE config 0: 640x480, 120 frames, frameserve -> framefilt bsub ->
    framefilt col -C HSV -> posidet hsv -H [40,80] -S [100,256] -V [90,256] through real shared memory.  The
    first frame becomes the background (no position); afterwards the difference image is the disc alone and
    the centroid is EXACTLY the disc centre + 0.5."""
    from oracle import synth

    rows, cols, n = 480, 640, 120

    def nodes(nm):
        return [["oat-posidet", "hsv", nm[2], nm[3], "-H", "[40,80]", "-S", "[100,256]", "-V", "[90,256]"],
                ["oat-framefilt", "col", nm[1], nm[2], "-C", "HSV"],
                ["oat-framefilt", "bsub", nm[0], nm[1]]]

    got = run_graph("cfg0", nodes, rows=rows, cols=cols, n=n)
    _check_config0(got, rows, cols, n)


def _check_config0(got, rows, cols, n):
    from oracle import synth

    assert len(got) == n
    assert got[0]["pos_ok"] is False
    for t in range(1, n):
        cx, cy = synth.disc_centre(rows, cols, t)
        assert got[t]["tick"] == t + 1 and got[t]["pos_ok"] is True, (t, got[t])
        assert got[t]["pos_xy"] == [cx + 0.5, cy + 0.5], (t, got[t], cx, cy)


def test_hsv_requires_hsv_source():
    """posidet hsv on a BGR source: 'Maybe use oat-framefilt col?' and exit -1 (Source.h:300-313)."""
    name = "oatb200pipe_color"
    subprocess.run([os.path.join(BIN, "oat-clean"), name, name + "_pos"], capture_output=True)
    det = subprocess.Popen([os.path.join(BIN, "oat-posidet"), "hsv", name, name + "_pos"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           text=True)
    wait_sources(name)
    serve = subprocess.Popen([os.path.join(BIN, "oat-frameserve"), "synth", name, "--rows", "60", "--cols", "80", "--num-samples", "3"])
    so, se = det.communicate(timeout=60)
    serve.wait(timeout=30)
    assert det.returncode == 255 and "Maybe use oat-framefilt col?" in se
    subprocess.run([os.path.join(BIN, "oat-clean"), name, name + "_pos"], capture_output=True)


def test_two_colour_graph_with_kalman_and_mean():
    """examples/mouse-track/two-color-det.sh shape: two detectors on one frame SOURCE -> posifilt kalman each ->
    posicom mean -> posisock; seven processes, positions compared with the oracle's
    Tracker -> Kalman2D -> mean_combine on the same frames."""
    tag = "oatb200pipe_2c_"
    names = [tag + n for n in ("raw", "pa", "pb", "ka", "kb", "pos")]
    raw, pa, pb, ka, kb, pos = names
    subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)
    kargs = ["--dt", "0.01", "-T", "0.05", "-a", "20", "-n", "0.5"]
    argvs = [["oat-posicom", "mean", ka, kb, pos],
             ["oat-posifilt", "kalman", pa, ka] + kargs,
             ["oat-posifilt", "kalman", pb, kb] + kargs,
             ["oat-posidet", "track", raw, pa, "-A", "0.05"] + HSV_ARGS,
             ["oat-posidet", "track", raw, pb, "-A", "0.0"] + HSV_ARGS]
    procs = []
    try:
        sock = subprocess.Popen([os.path.join(BIN, "oat-posisock"), "std", pos], stdout=subprocess.PIPE, text=True)
        for argv in argvs:
            procs.append(subprocess.Popen([os.path.join(BIN, argv[0])] + argv[1:], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        wait_sources(raw, n=2)
        wait_sources(pa, pb, ka, kb, pos)
        serve = subprocess.Popen([os.path.join(BIN, "oat-frameserve"), "synth", raw, "--rows", str(ROWS), "--cols", str(COLS),
                                  "--num-samples", str(N), "--fps", "100"])
        out, _ = sock.communicate(timeout=120)
        assert serve.wait(timeout=30) == 0
        for p in procs:
            so, se = p.communicate(timeout=30)
            assert p.returncode == 0, (p.args, so, se)
        got = [json.loads(line) for line in out.splitlines() if line.strip()]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)
    hp = oracle.HsvParams(**BAND)
    trk = [oracle.Tracker(ROWS, COLS), oracle.Tracker(ROWS, COLS)]
    kal = [oracle.Kalman2D(0.01, 0.05, 20.0, 0.5), oracle.Kalman2D(0.01, 0.05, 20.0, 0.5)]
    assert len(got) == N
    n_valid = 0
    for t in range(N):
        f = oracle.synth_frame(ROWS, COLS, 1000, t)
        filt = []
        for k, (tr, lr) in zip(kal, zip(trk, (0.05, 0.0))):
            o, _ = tr.track(f, lr, hp)
            filt.append(k.filter(bool(o.position_valid), o.x, o.y))
        w = oracle.mean_combine(filt, -1)
        p = got[t]
        assert p["pos_ok"] == bool(w.position_valid) and p["vel_ok"] == bool(w.velocity_valid), (t, p)
        if w.position_valid:
            n_valid += 1
            assert abs(p["pos_xy"][0] - w.x) < 1e-4 and abs(p["pos_xy"][1] - w.y) < 1e-4, (t, p, w.x, w.y)
            assert abs(p["vel_xy"][0] - w.vx) < 1e-4 and abs(p["vel_xy"][1] - w.vy) < 1e-4, (t, p, w.vx, w.vy)
    assert n_valid >= N - 3


@pytest.mark.parametrize("device", [False, True])
def test_fused_tracker_component_pipelined(device):
    """--pipeline 4: up to four frames in flight on the GPU behind the SOURCE; the SINK still sees one position per frame,
    in order, each carrying its own frame's Sample -- identical to the synchronous component."""
    def nodes(n):
        return [["oat-posidet", "track", n[0], n[3], "-A", "0.05", "--pipeline", "4"] + HSV_ARGS]

    check(run_graph("fusedp" + ("d" if device else "h"), nodes, serve_args=(("--device",) if device else ())), oracle_positions(0.05))


def test_frameserve_test_through_mog_component(tmp_path):
    """The reference's perf-protocol graph (test/perf/framefilt-mog.sh): `frameserve test` -> `framefilt mog`, here with a
    reader behind it: a static image gives an all-shadow first frame (kept) and all-background frames afterwards (zeroed)."""
    import numpy as np

    img = oracle.synth_frame(ROWS, COLS, 1000, 2)
    path = tmp_path / "img.npy"
    np.save(path, img)
    names = ["oatb200pipe_tf_raw", "oatb200pipe_tf_filt"]
    subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)
    reader = subprocess.Popen([os.path.join(BIN, "shmemdf_test"), "dump-frames", names[1]], stdout=subprocess.PIPE, text=True)
    mog = subprocess.Popen([os.path.join(BIN, "oat-framefilt"), "mog", names[0], names[1], "-a", "0.0"], stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True)
    wait_sources(*names)
    serve = subprocess.run([os.path.join(BIN, "oat-frameserve"), "test", names[0], "-f", str(path), "-n", "5"], capture_output=True, text=True,
                           timeout=120)
    assert serve.returncode == 0, serve.stderr
    out, _ = reader.communicate(timeout=60)
    assert mog.wait(timeout=30) == 0
    lines = [ln.split() for ln in out.splitlines() if ln.strip()]
    assert [int(l[0]) for l in lines] == [1, 2, 3, 4, 5]

    def fnv(buf):
        h = 1469598103934665603
        for b in buf:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return h

    assert int(lines[0][6]) == fnv(img.tobytes())                       # first frame: mask all 127 -> the frame is kept
    assert all(int(l[6]) == fnv(bytes(img.size)) for l in lines[1:])    # frozen model, static image: background -> zeros
    subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)


def _write_clip(path, rows, cols, n, seed=1000):
    import numpy as np

    clip = np.stack([oracle.synth_frame(rows, cols, seed, t) for t in range(n)])
    np.save(path, clip)
    return clip


@pytest.mark.parametrize("device", [False, True])
def test_config0_from_a_clip_file(tmp_path, device):
    """BASELINE config 0 as worded: a 640x480 test CLIP served by `frameserve file` (FileReader.cpp:103-131; the clip is a
    lossless .npy) -> framefilt bsub -> framefilt col -C HSV -> posidet hsv, analytic answer as above.  With --device
    the clip is uploaded to HBM once and published in place (device_offset), and bsub / col hand device frames on."""
    rows, cols, n = 480, 640, 60
    path = tmp_path / "clip.npy"
    _write_clip(path, rows, cols, n)
    ds = ["--device-sink"] if device else []

    def nodes(nm):
        return [["oat-posidet", "hsv", nm[2], nm[3], "-H", "[40,80]", "-S", "[100,256]", "-V", "[90,256]"],
                ["oat-framefilt", "col", nm[1], nm[2], "-C", "HSV"] + ds,
                ["oat-framefilt", "bsub", nm[0], nm[1]] + ds]

    got = run_graph("cfg0f" + ("d" if device else "h"), nodes, serve_args=(("--device",) if device else ()),
                    server=lambda raw: ["file", raw, "-f", str(path), "-r", "100"])
    _check_config0(got, rows, cols, n)


@pytest.mark.parametrize("device_sink", [False, True])
def test_buffer_component_in_front_of_the_tracker(tmp_path, device_sink):
    """frameserve file -> `oat buffer frame` (FIFO in HBM, src/buffer/FrameBuffer.cpp:56-116) -> posidet track --pipeline 4:
    every frame comes out once, in order, with its own Sample; positions equal the oracle's.  (The FIFO holds the whole
    clip: the tracker spends its first half second creating its model while the server already runs.)"""
    path = tmp_path / "clip.npy"
    _write_clip(path, ROWS, COLS, N)
    names = ["oatb200pipe_buf_" + k + ("d" if device_sink else "h") for k in ("raw", "fifo", "pos")]
    subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)
    procs = []
    try:
        sock = subprocess.Popen([os.path.join(BIN, "oat-posisock"), "std", names[2]], stdout=subprocess.PIPE, text=True)
        procs.append(subprocess.Popen([os.path.join(BIN, "oat-posidet"), "track", names[1], names[2], "-A", "0.05", "--pipeline", "4"] + HSV_ARGS,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
        procs.append(subprocess.Popen([os.path.join(BIN, "oat-buffer"), "frame", names[0], names[1], "--capacity", "32"] +
                                      (["--device-sink"] if device_sink else []), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
        wait_sources(*names)
        serve = subprocess.run([os.path.join(BIN, "oat-frameserve"), "file", names[0], "-f", str(path), "-r", "100"], capture_output=True,
                               text=True, timeout=120)
        assert serve.returncode == 0, serve.stderr
        out, _ = sock.communicate(timeout=120)
        for p in procs:
            _, se = p.communicate(timeout=60)
            assert p.returncode == 0, (p.args, se)
            assert "Buffer overrun" not in se
        check([json.loads(line) for line in out.splitlines() if line.strip()], oracle_positions(0.05))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        subprocess.run([os.path.join(BIN, "oat-clean")] + names, capture_output=True)


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("shape", [(240, 320), (250, 1000)])
def test_streaming_tracker_behind_a_free_running_clip(tmp_path, device, shape):
    """`frameserve file` WITHOUT --fps (free-running, as the reference's perf protocol runs its server) -> posidet track
    --pipeline 64: the component drives the streaming resident engine (chunks of up to 32 frames; with --device the clip
    is persistent in HBM and read in place, from shm every frame is staged H2D; 1000 columns: rows that are not a
    multiple of 16 bytes are re-pitched on the device).  Every frame comes out once, in order, with its own Sample,
    and every position equals the oracle's."""
    rows, cols = shape
    n = 150
    path = tmp_path / "clip.npy"
    _write_clip(path, rows, cols, n)
    trk = oracle.Tracker(rows, cols)
    hp = oracle.HsvParams(**BAND)
    want = [trk.track(oracle.synth_frame(rows, cols, 1000, t), 0.05, hp)[0] for t in range(n)]

    def nodes(nm):
        return [["oat-posidet", "track", nm[0], nm[3], "-A", "0.05", "--pipeline", "64"] + HSV_ARGS]

    got = run_graph(f"strm{rows}" + ("d" if device else "h"), nodes, serve_args=(("--device",) if device else ()),
                    server=lambda raw: ["file", raw, "-f", str(path)])
    assert [p["tick"] for p in got] == list(range(1, n + 1))
    n_valid = 0
    for t, (p, o) in enumerate(zip(got, want)):
        assert p["pos_ok"] == bool(o.position_valid), t
        if o.position_valid:
            n_valid += 1
            assert abs(p["pos_xy"][0] - o.x) < 1e-4 and abs(p["pos_xy"][1] - o.y) < 1e-4, (t, p, o.x, o.y)
    assert n_valid >= n - 2
