"""N>1 host logic on CPU: world size 2 over gloo (streams shard one per GPU, no data-path
collective; ranks only exchange timings and positions).  The per-stream compute here is the CPU oracle --
as the checker's stand-in for the device -- so the test pins that (a) every stream is served by exactly one
rank, (b) sharded results equal the single-process results, (c) job time = max over ranks, (d) the
reference arm of bench.py prints exactly one JSON line under a 2-rank launch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["OAT_ROOT"])
import torch.distributed as dist
import oracle
from oat_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["OAT_PORT"],
                        rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
rows, cols, nframes, nstreams = 60, 80, 6, 5
hp = oracle.HsvParams(h=(40, 80), s=(100, 256), v=(100, 256))
local = {}
for s in sharding.streams_for_rank(nstreams, rank, world):
    trk = oracle.Tracker(rows, cols)
    res = []
    for t in range(nframes):
        o, _ = trk.track(oracle.synth_frame(rows, cols, sharding.stream_seed(1000, s), t), 0.0, hp)
        res.append((bool(o.position_valid), o.x, o.y))
    local[s] = res
times = sharding.max_over_ranks([10.0 + rank, 5.0 - rank], dist)
frames = sharding.sum_over_ranks(len(local) * nframes, dist)
merged = sharding.gather_positions(local, dist)
if rank == 0:
    print(json.dumps({"times": times, "frames": frames, "streams": sorted(merged), "pos": {str(k): v for k, v in merged.items()}}))
dist.barrier()
dist.destroy_process_group()
'''


def test_stream_sharding_partition():
    from oat_b200 import sharding

    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.streams_for_rank(64, r, world)
        assert sorted(seen) == list(range(64))
        assert all(s % world == r for r in range(world) for s in sharding.streams_for_rank(64, r, world))
    with pytest.raises(ValueError):
        sharding.streams_for_rank(8, 2, 2)


def test_two_ranks_gloo(tmp_path):
    import oracle

    oracle.build()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", OAT_ROOT=ROOT, OAT_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    assert outs[1][0].strip() == ""  # only rank 0 reports
    got = json.loads(outs[0][0].strip().splitlines()[-1])
    assert got["times"] == [11.0, 5.0]  # max over ranks
    assert got["frames"] == 30 and got["streams"] == [0, 1, 2, 3, 4]
    # sharded == single process
    hp = oracle.HsvParams(h=(40, 80), s=(100, 256), v=(100, 256))
    for s in range(5):
        trk = oracle.Tracker(60, 80)
        for t in range(6):
            o, _ = trk.track(oracle.synth_frame(60, 80, 1000 + s, t), 0.0, hp)
            v, x, y = got["pos"][str(s)][t]
            assert v == bool(o.position_valid) and x == o.x and y == o.y


def test_reference_arm_two_rank_launch():
    """bench.py --impl reference under a 2-rank launch: rank 0 alone runs and prints ONE JSON line."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29999")
    lines = []
    for rank in range(2):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
                            "--warmup", "1", "--workload", "480p"], env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)),
                           capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stderr[-2000:]
        lines.append([ln for ln in r.stdout.splitlines() if ln.strip()])
    assert lines[1] == []
    assert len(lines[0]) == 1
    d = json.loads(lines[0][0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["cores"] >= 1 and d["n_gpus"] == 2
