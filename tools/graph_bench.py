"""The reference's own performance protocol (test/perf/framefilt-mog.sh:1-3, test/perf/results.md:12-18) through the
drop-in components and REAL shared memory: `oat frameserve test` pushes N copies of one static image, free-running,
through the listening component(s); the figure is N / wall time of the frame server, exactly as `time oat frameserve
test raw -f IMAGE -c test.toml test` measured it (573 fps for cv::cuda MOG on a GTX 970, 75.7 fps for CPU MOG2 on an
i7-5600U, 1000 x 1 MP frames).

    python tools/graph_bench.py [--frames 1000] [--out FILE.json]

Graphs: framefilt mog alone (the published number's graph); frameserve -> mog -> col HSV -> posidet hsv (the
reference-shaped chain); frameserve -> posidet track (the fused component), synchronous and with --pipeline 8; each
with host frames (page-locked shm, H2D/D2H per component) and with device-resident frames (--device /
--device-sink: CUDA IPC, pixels never leave the GPU)."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN = os.path.join(ROOT, "oat_b200", "bin")
HSV = ["-H", "[40,80]", "-S", "[100,256]", "-V", "[100,256]"]


def make_image(path, rows, cols):
    import oracle

    np.save(path, oracle.synth_frame(rows, cols, 1000, 3))  # one frame of the tracking stream (with its blob)


def run(tag, image, n, consumers, device, last_is_position):
    names = {k: f"oatb200gb_{tag}_{k}" for k in ("raw", "filt", "hsv", "pos")}
    subprocess.run([os.path.join(BIN, "oat-clean")] + list(names.values()), capture_output=True)
    procs = []
    try:
        sock = None
        sock_out = open(f"/tmp/oatb200_graph_bench/{tag}.positions", "w+")  # (a pipe would fill up and stall the graph)
        if last_is_position:
            sock = subprocess.Popen([os.path.join(BIN, "oat-posisock"), "std", names["pos"]], stdout=sock_out, stderr=subprocess.PIPE, text=True,
                                    env=dict(os.environ, OAT_B200_TIMING="1"))
        env = dict(os.environ, OAT_B200_TIMING="1")
        for argv in consumers(names):
            procs.append(subprocess.Popen([os.path.join(BIN, argv[0])] + argv[1:], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env))
        time.sleep(3.0)  # the reference's scripts sleep too: every consumer has its CUDA context and waits in connect()
        t0 = time.perf_counter()
        try:
            serve = subprocess.run([os.path.join(BIN, "oat-frameserve"), "test", names["raw"], "-f", image, "-n", str(n)] +
                                   (["--device"] if device else []), capture_output=True, text=True, timeout=90, env=env)
        except subprocess.TimeoutExpired:
            # say who is still there and what they wrote before giving up on this graph
            state = []
            for p in procs:
                alive = p.poll() is None
                if alive:
                    p.kill()
                _, se = p.communicate(timeout=10)
                state.append((p.args[0].split("/")[-1], p.args[1], "alive" if alive else f"rc={p.returncode}", se[-300:]))
            if sock is not None:
                sock.kill()
            return {"frames": n, "error": "frame server timed out", "components": state}
        wall = time.perf_counter() - t0
        assert serve.returncode == 0, serve.stderr
        npos = None
        marks = None  # token times of the LAST component of the graph
        if sock is not None:
            _, se = sock.communicate(timeout=120)
            sock_out.seek(0)
            npos = len([ln for ln in sock_out.read().splitlines() if ln.strip()])
            marks = [ln for ln in se.splitlines() if "tokens out at:" in ln]
        stages, ends = [], []
        for p in procs:
            _, se = p.communicate(timeout=120)
            assert p.returncode == 0, (p.args, se)
            stages += [ln for ln in se.splitlines() if "per frame (us)" in ln]
            ends += [float(ln.rsplit(" ", 1)[1]) for ln in se.splitlines() if "end of stream at" in ln]
            if marks is None:
                marks = [ln for ln in se.splitlines() if "tokens out at:" in ln]
        # the rate once the graph is warm: from the 100th token out of the last component to its last one
        warm = None
        if marks:
            tk = dict((k, float(v)) for k, v in (w.replace("last(", "").replace(")", "").split(":") for w in marks[0].split("tokens out at:")[1].split()))
            nl = max(int(k) for k in tk)
            if "100" in tk and nl > 100 and tk[str(nl)] > tk["100"]:
                warm = (nl - 100) / (tk[str(nl)] - tk["100"])
        # steady state: from the server's first frame to the moment the last token left the last component (the
        # server's own start-up -- for --device its CUDA context -- and the components' teardown are not in it)
        started = [float(ln.rsplit(" ", 1)[1]) for ln in serve.stderr.splitlines() if "serving started at" in ln]
        steady = max(ends) - started[0] if ends and started else None
        return {"frames": n, "wall_s": wall, "fps": n / wall, "steady_s": steady, "steady_fps": n / steady if steady else None,
                "warm_fps": warm, "positions": npos, "stages": stages}
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        subprocess.run([os.path.join(BIN, "oat-clean")] + list(names.values()), capture_output=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--out", default="")
    ap.add_argument("--mps", action="store_true", help="run the graphs under the CUDA MPS daemon (the processes of a graph share the GPU "
                    "concurrently instead of time-slicing it between their contexts)")
    ap.add_argument("--only", default="", help="comma-separated graph name prefixes to run (default: all)")
    args = ap.parse_args()
    mps = None
    if args.mps:
        os.makedirs("/tmp/oatb200_mps/log", exist_ok=True)
        os.environ["CUDA_MPS_PIPE_DIRECTORY"] = "/tmp/oatb200_mps"
        os.environ["CUDA_MPS_LOG_DIRECTORY"] = "/tmp/oatb200_mps/log"
        mps = subprocess.run(["nvidia-cuda-mps-control", "-d"], capture_output=True, text=True, timeout=30)
        print("MPS daemon:", mps.returncode, mps.stdout.strip(), mps.stderr.strip(), flush=True)
    try:
        _main(args)
    finally:
        if args.mps:
            subprocess.run(["nvidia-cuda-mps-control"], input="quit\n", capture_output=True, text=True, timeout=30)


def _main(args):
    want = [w for w in args.only.split(",") if w]

    def sel(name):
        return not want or any(name.startswith(w) for w in want)

    res = {"protocol": "N static frames from `oat-frameserve test`, free-running, through real shm; fps = N / wall time of the frame server "
                       "process (test/perf/results.md:12-18: `time oat frameserve test ...`; with --device that includes creating its CUDA "
                       "context); steady_fps = N / (first frame served -> last token out of the last component); warm_fps = rate between the 100th and "
                       "the last token out of the last component (the first frames pay for model allocation, lazy module loading and the GPU's "
                       "clock ramp: 10 ms to 1 s from run to run); consumers started 3 s earlier",
           "reference_published": {"framefilt mog, cv::cuda MOG, GTX 970 (results.md:34-37)": 573.0,
                                   "framefilt mog, CPU MOG2, i7-5600U (results.md:91-95)": 75.7,
                                   "posidet hsv, GTX 970 box CPU path (results.md:55-58)": 214.0}}
    tmp = "/tmp/oatb200_graph_bench"
    os.makedirs(tmp, exist_ok=True)
    for wl, (rows, cols) in {"1mp": (1000, 1000), "1080p": (1080, 1920)}.items():
        img = os.path.join(tmp, f"{wl}.npy")
        make_image(img, rows, cols)
        r = {}
        for dev in (False, True):
            k = "device" if dev else "host"
            ds = ["--device-sink"] if dev else []
            if sel("mog"):
                r[f"mog_{k}"] = run(f"{wl}m{k}", img, args.frames, lambda n: [["oat-framefilt", "mog", n["raw"], n["filt"], "-a", "0.01"] + ds], dev, False)
            if sel("chain"):
                r[f"chain_{k}"] = run(f"{wl}c{k}", img, args.frames, lambda n: [
                ["oat-posidet", "hsv", n["hsv"], n["pos"]] + HSV,
                ["oat-framefilt", "col", n["filt"], n["hsv"], "-C", "HSV"] + ds,
                ["oat-framefilt", "mog", n["raw"], n["filt"], "-a", "0.01"] + ds], dev, True)
            if sel("track_" + k):
                r[f"track_{k}"] = run(f"{wl}t{k}", img, args.frames, lambda n: [["oat-posidet", "track", n["raw"], n["pos"], "-A", "0.01"] + HSV], dev, True)
            if sel("track_pipeline8"):
                r[f"track_pipeline8_{k}"] = run(f"{wl}p{k}", img, args.frames, lambda n: [
                ["oat-posidet", "track", n["raw"], n["pos"], "-A", "0.01", "--pipeline", "8"] + HSV], dev, True)
            # the streaming resident engine behind the lock-step SOURCE: chunks of 32 frames, long enough a run to see its rate
            if sel("track_pipeline64"):
                r[f"track_pipeline64_{k}"] = run(f"{wl}q{k}", img, args.frames * 20, lambda n: [
                ["oat-posidet", "track", n["raw"], n["pos"], "-A", "0.01", "--pipeline", "64"] + HSV], dev, True)
        res[wl] = r
        for k, v in r.items():
            if "error" in v:
                print(f"{wl:6s} {k:24s} FAILED: {v}", flush=True)
            else:
                print(f"{wl:6s} {k:24s} {v['fps']:10.1f} fps by the protocol's clock ({v['frames']} frames, `time oat-frameserve` {v['wall_s']:.3f} s), "
                      f"{v['steady_fps'] or 0:10.1f} fps first frame served -> last token out ({v['steady_s'] or 0:.4f} s), "
                      f"{v['warm_fps'] or 0:10.1f} fps WARM (100th -> last token out of the last component), positions {v['positions']}", flush=True)
                for ln in v["stages"]:
                    print("         " + ln, flush=True)
    # frameserve alone (no listener), the protocol's own overhead
    img = os.path.join(tmp, "1mp.npy")
    subprocess.run([os.path.join(BIN, "oat-clean"), "oatb200gb_alone"], capture_output=True)
    t0 = time.perf_counter()
    subprocess.run([os.path.join(BIN, "oat-frameserve"), "test", "oatb200gb_alone", "-f", img, "-n", str(args.frames)], check=True)
    res["frameserve_alone_s"] = time.perf_counter() - t0
    print("frameserve alone:", res["frameserve_alone_s"], "s")
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
