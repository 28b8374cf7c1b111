"""Multi-GPU host logic of the path: video streams are independent units -- each has its
own GMM state and yields its own position -- so stream ``s`` goes to GPU ``s mod G`` and there is NO
data-path collective.  The reference exposes the same knob per process (``--gpu-index``,
src/framefilter/BackgroundSubtractorMOG.cpp:61-63, :92-111).  The only cross-rank traffic is timing
(max over ranks) and the <=100-byte positions gathered by the host; both go through torch.distributed
(NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def streams_for_rank(n_streams: int, rank: int, world: int) -> list[int]:
    """Stream ids served by ``rank``: s -> GPU (s mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [s for s in range(n_streams) if s % world == rank]


def stream_seed(base_seed: int, stream: int) -> int:
    """Seed of synthetic stream ``s`` (seed_s = 1000 + s)."""
    return base_seed + stream


def max_over_ranks(values, dist=None, device="cpu"):
    """The job took as long as its slowest rank: element-wise MAX all-reduce of per-rank timings."""
    vals = [float(v) for v in values]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    import torch

    t = torch.tensor(vals, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(value, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return int(value)
    import torch

    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def gather_positions(local: dict, dist=None):
    """local: {stream id: [(valid, x, y), ...]} on every rank -> merged dict on rank 0 (None elsewhere)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == 0 else None
    dist.gather_object(local, out, dst=0)
    if rank != 0:
        return None
    merged = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError(f"stream {k} was served by two ranks")
            merged[k] = v
    return merged
