"""Host side of the drop-in (oat_b200/host, C++17): the reference's shmemdf transport tests
(test/shmemdf/*_test.cpp) re-expressed in oat_b200/host/shmemdf_test.cpp, and the CLI / configuration
contract of the two hot-path executables (src/framefilter/main.cpp, src/positiondetector/main.cpp,
lib/utility/TOMLSanitize.h).  None of this needs a GPU: option errors surface before the device is touched."""
import os
import subprocess

import pytest

from graph_util import wait_sources

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oat_b200", "bin")


@pytest.fixture(scope="module")
def host_bin():
    # liboatgpu.so must exist for the link step of the two GPU executables
    import oat_b200

    oat_b200.build()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oat_b200", "host")], check=True)
    return BIN


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, timeout=60, **kw)


def test_shmemdf_transport_suite(host_bin):
    r = run([os.path.join(host_bin, "shmemdf_test"), os.path.join(host_bin, "oat-frameserve")])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


def test_dispatcher_and_usage(host_bin):
    r = run([os.path.join(host_bin, "oat"), "framefilt", "--help"])
    assert r.returncode == 0 and "framefilt TYPE SOURCE SINK" in r.stdout
    r = run([os.path.join(host_bin, "oat"), "posidet", "--version"])
    assert r.returncode == 0 and "Position Detector" in r.stdout
    r = run([os.path.join(host_bin, "oat"), "nonsense"])
    assert r.returncode != 0


@pytest.mark.parametrize("args,msg", [
    (["oat-framefilt", "bogus", "a", "b"], "invalid TYPE"),
    (["oat-framefilt", "mog", "a"], "a SINK must be specified"),
    (["oat-framefilt", "mog"], "a SOURCE must be specified"),
    (["oat-framefilt", "mog", "a", "b", "-a", "1.5"], "out of bounds"),       # getNumericValue range check, ...MOG.cpp:87-88
    (["oat-framefilt", "mog", "a", "b", "--nonsense", "1"], "unrecognised option"),
    (["oat-framefilt", "col", "a", "b"], "pixel color must be specified"),
    (["oat-posidet", "hsv", "a", "b", "-H", "[40,300]"], "Values of h-thresh should be between 0 and 256."),
    (["oat-posidet", "hsv", "a", "b", "-a", "[10,5]"], "Max area should be larger than min area."),
    (["oat-posidet", "hsv", "a", "b", "-S", "[1,2,3]"], "2 elements"),
    (["oat-posidet", "kalman", "a", "b"], "invalid TYPE"),
    (["oat-posidet", "diff", "a", "b", "-a", "[9,3]"], "Max area should be larger than min area."),
    (["oat-posidet", "thresh", "a", "b", "-T", "[10,999]"], "Values of thresh should be between 0 and 256."),
    (["oat-framefilt", "thresh", "a", "b", "-I", "[-1,5]"], "Values of intensity should be between 0 and 256."),
    (["oat-framefilt", "mask", "a", "b", "-m", "/nonexistent.pgm"], "could not be read"),
    (["oat-posifilt", "homography", "a", "b"], "invalid TYPE"),
    (["oat-posifilt", "kalman", "a"], "a SINK name must be specified"),
    (["oat-posifilt", "kalman", "a", "b", "--dt", "-1"], "out of bounds"),
    (["oat-posifilt", "kalman", "a", "b", "--nonsense", "1"], "unrecognised option"),
    (["oat-posicom", "median", "a", "b", "c"], "invalid TYPE"),
    (["oat-posicom", "mean", "a", "b"], "At least two SOURCES and a SINK must be specified."),   # PositionCombiner.cpp:41-42
    (["oat-posicom", "mean", "a", "b", "c", "-h", "2"], "out of bounds"),                          # MeanPosition.cpp:52-54
])
def test_cli_errors_exit_minus_one(host_bin, args, msg):
    """Every exception is caught in main -> 'name: message' on stderr -> return -1 (main.cpp:278-295)."""
    r = run([os.path.join(host_bin, args[0])] + args[1:])
    assert r.returncode == 255, (r.returncode, r.stdout, r.stderr)
    assert msg in r.stderr, r.stderr


def test_toml_config_table(host_bin, tmp_path):
    """-c FILE KEY selects a table; unknown keys are rejected (checkKeys); CLI beats TOML."""
    cfg = tmp_path / "config.toml"
    cfg.write_text('[hsv]\nh-thresh = [40, 80]\ns-thresh = [100, 256]\nerode = 0\ndilate = 10\n\n'
                   '[stale]\nh_thresholds = [1, 2]\n\n[bad]\nh-thresh = [40, 999]\n')
    exe = os.path.join(host_bin, "oat-posidet")
    r = run([exe, "hsv", "a", "b", "-c", str(cfg), "stale"])
    assert r.returncode == 255 and "Unknown configuration key 'h_thresholds'" in r.stderr
    r = run([exe, "hsv", "a", "b", "-c", str(cfg), "missing"])
    assert r.returncode == 255 and "No configuration table named 'missing'" in r.stderr
    r = run([exe, "hsv", "a", "b", "-c", str(cfg), "bad"])
    assert r.returncode == 255 and "between 0 and 256" in r.stderr
    # CLI beats TOML (getValue, TOMLSanitize.h:175-183): with a good -H on the command line the bad table value
    # is never used; the component then creates its CUDA context and blocks in connect() waiting for a SINK, and SIGINT
    # ends it cleanly (0).  On a box without a GPU it stops at the context instead (there is no CPU fallback).
    import signal
    import time

    p = subprocess.Popen([exe, "hsv", "oatb200test_nosrc", "oatb200test_nosink", "-c", str(cfg), "bad", "-H", "[40,80]"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    time.sleep(0.5)
    p.send_signal(signal.SIGINT)
    out, err = p.communicate(timeout=10)
    assert p.returncode == 0 or (p.returncode == 255 and "no CUDA device" in err), (p.returncode, out, err)
    assert "between 0 and 256" not in err
    subprocess.run([os.path.join(host_bin, "oat-clean"), "oatb200test_nosrc", "oatb200test_nosink"], capture_output=True)


def test_position_egress_formats(host_bin, tmp_path):
    """posisock std: JSON lines with serializePosition's keys (lib/datatypes/Position2D.h:169-233) and a .npy file
    with the recorder's header + 82-byte packed records (lib/datatypes/Position2D.cpp:24-96, src/recorder/Format.cpp:35-93)
    that numpy loads with the reference's dtype."""
    import json

    import numpy as np

    for mode in ("json", "npy"):
        addr = f"oatb200test_pos_{mode}"
        npy = tmp_path / "pos.npy"
        sock_args = [os.path.join(host_bin, "oat-posisock"), "std", addr] + (["--npy", str(npy)] if mode == "npy" else [])
        sock = subprocess.Popen(sock_args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        emit = subprocess.Popen([os.path.join(host_bin, "shmemdf_test"), "emit-positions", addr, "7"])
        out, err = sock.communicate(timeout=60)
        assert emit.wait(timeout=30) == 0 and sock.returncode == 0, err
        if mode == "json":
            recs = [json.loads(ln) for ln in out.splitlines() if ln.strip()]
            assert [r["tick"] for r in recs] == list(range(1, 8))
            assert [r["usec"] for r in recs] == [20000 * k for k in range(1, 8)]
            assert all(list(r) [:4] == ["tick", "usec", "unit", "pos_ok"] for r in recs)
            assert recs[1]["pos_ok"] is True and recs[1]["pos_xy"] == [11.5, 0.125]
            assert recs[0]["pos_ok"] is False and "pos_xy" not in recs[0]
        else:
            a = np.load(npy)
            assert a.shape == (7,) and a.dtype.itemsize == 82
            assert a["tick"].tolist() == list(range(1, 8)) and a["usec"].tolist() == [20000 * k for k in range(1, 8)]
            assert a["pos_ok"].tolist() == [0, 1, 1, 0, 1, 1, 0]
            assert np.allclose(a["pos_xy"][:, 0], 10.5 + np.arange(7)) and np.allclose(a["pos_xy"][:, 1], 0.125 * np.arange(7))
        subprocess.run([os.path.join(host_bin, "oat-clean"), addr], capture_output=True)


def _fnv1a(buf: bytes) -> int:
    h = 1469598103934665603
    for b in buf:
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("fmt", ["npy", "ppm", "pgm-as-bgr", "npy-as-grey"])
def test_frameserve_test_serves_a_static_image(host_bin, tmp_path, fmt):
    """`oat frameserve test SINK -f IMAGE -n N` (src/frameserver/TestFrame.cpp:37-128): the static image N times, tick
    1..N, frame period from --fps, BGR unless -C GREY; read back by a plain Source<Frame>."""
    import numpy as np

    rng = np.random.default_rng(11)
    rows, cols, n = 37, 52, 6
    bgr = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
    grey = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
    args, want, ch, color = [], None, 3, 2
    if fmt == "npy":
        path = tmp_path / "img.npy"
        np.save(path, bgr)
        want = bgr
    elif fmt == "ppm":  # PPM is RGB on disk
        path = tmp_path / "img.ppm"
        path.write_bytes(b"P6\n# a comment\n%d %d\n255\n" % (cols, rows) + bgr[:, :, ::-1].tobytes())
        want = bgr
    elif fmt == "pgm-as-bgr":  # a grey file served as BGR: the three channels carry the grey value
        path = tmp_path / "img.pgm"
        path.write_bytes(b"P5 %d %d 255\n" % (cols, rows) + grey.tobytes())
        want = np.repeat(grey[:, :, None], 3, axis=2)
    else:  # a colour file served as GREY (-C GREY): OpenCV's BGR -> grey weights
        path = tmp_path / "img.npy"
        np.save(path, bgr)
        args = ["-C", "GREY"]
        b, g, r = (bgr[:, :, i].astype(np.int64) for i in range(3))
        want = ((b * 1868 + g * 9617 + r * 4899 + (1 << 13)) >> 14).astype(np.uint8)
        ch, color = 1, 1
    addr = f"oatb200test_tf_{fmt.replace('-', '_')}"
    subprocess.run([os.path.join(host_bin, "oat-clean"), addr], capture_output=True)
    reader = subprocess.Popen([os.path.join(host_bin, "shmemdf_test"), "dump-frames", addr], stdout=subprocess.PIPE, text=True)
    wait_sources(addr)
    serve = run([os.path.join(host_bin, "oat-frameserve"), "test", addr, "-f", str(path), "-n", str(n), "-r", "500"] + args)
    out, _ = reader.communicate(timeout=60)
    assert serve.returncode == 0, serve.stderr
    lines = [ln.split() for ln in out.splitlines() if ln.strip()]
    assert [int(l[0]) for l in lines] == list(range(1, n + 1))
    assert [int(l[1]) for l in lines] == [2000 * k for k in range(1, n + 1)]
    for l in lines:
        assert (int(l[2]), int(l[3]), int(l[4]), int(l[5])) == (rows, cols, ch, color)
        assert int(l[6]) == _fnv1a(np.ascontiguousarray(want).tobytes())
    subprocess.run([os.path.join(host_bin, "oat-clean"), addr], capture_output=True)


@pytest.mark.parametrize("args,msg", [
    (["oat-frameserve", "bogus", "a"], "invalid TYPE"),
    (["oat-frameserve", "test"], "a SINK must be specified"),
    (["oat-frameserve", "test", "oatb200test_tf_err"], "Required configuration key 'test-image'"),
    (["oat-frameserve", "test", "oatb200test_tf_err", "-f", "/nonexistent.ppm"], "could not be read"),
    (["oat-frameserve", "test", "oatb200test_tf_err", "-f", "/dev/null", "-C", "HSV"], "Invalid color format"),
    (["oat-frameserve", "test", "oatb200test_tf_err", "-f", "x", "-n", "0"], "out of bounds"),
])
def test_frameserve_cli_errors(host_bin, args, msg):
    r = run([os.path.join(host_bin, args[0])] + args[1:])
    assert r.returncode == 255, (r.returncode, r.stdout, r.stderr)
    assert msg in r.stderr, r.stderr


def test_frameserve_file_serves_a_clip(host_bin, tmp_path):
    """`oat frameserve file SINK -f CLIP -r FPS [--roi]` (src/frameserver/FileReader.cpp:36-131): every frame of the clip
    once, in order, then end of stream; --roi crops."""
    import numpy as np

    rng = np.random.default_rng(3)
    clip = rng.integers(0, 256, (5, 30, 44, 3), dtype=np.uint8)
    path = tmp_path / "clip.npy"
    np.save(path, clip)
    for tag, args, frames in (("full", [], clip), ("roi", ["--roi", "[4,2,20,16]"], clip[:, 2:18, 4:24])):
        addr = f"oatb200test_file_{tag}"
        subprocess.run([os.path.join(host_bin, "oat-clean"), addr], capture_output=True)
        reader = subprocess.Popen([os.path.join(host_bin, "shmemdf_test"), "dump-frames", addr], stdout=subprocess.PIPE, text=True)
        serve = run([os.path.join(host_bin, "oat-frameserve"), "file", addr, "-f", str(path), "-r", "250"] + args)
        out, _ = reader.communicate(timeout=60)
        assert serve.returncode == 0, serve.stderr
        lines = [ln.split() for ln in out.splitlines() if ln.strip()]
        assert [int(l[0]) for l in lines] == [1, 2, 3, 4, 5]
        assert [int(l[1]) for l in lines] == [4000 * k for k in range(1, 6)]
        for t, l in enumerate(lines):
            assert (int(l[2]), int(l[3]), int(l[4])) == frames.shape[1:]
            assert int(l[6]) == _fnv1a(np.ascontiguousarray(frames[t]).tobytes())
        subprocess.run([os.path.join(host_bin, "oat-clean"), addr], capture_output=True)
    r = run([os.path.join(host_bin, "oat-frameserve"), "file", "oatb200test_file_err", "-f", str(path), "--roi", "[40,0,20,16]"])
    assert r.returncode == 255 and "ROI must fit" in r.stderr
    r = run([os.path.join(host_bin, "oat-buffer"), "pos2D", "a", "b"])
    assert r.returncode == 255 and "invalid TYPE" in r.stderr
    r = run([os.path.join(host_bin, "oat-buffer"), "frame", "a"])
    assert r.returncode == 255 and "a SINK must be specified" in r.stderr
