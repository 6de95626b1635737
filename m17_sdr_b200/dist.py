"""Channel sharding across the GPUs of one box (one process per GPU, torch.distributed).

Channels are independent: each rank owns a contiguous channel range, keeps its state blobs and
IQ resident on its own GPU, and the data path needs NO collective.  The only optional exchange is a gather of
the fixed-size decoded-frame records and an all-reduce of the per-channel counters (NCCL over NVLink on GPUs,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(nchan_total, rank, world):
    """Contiguous range [c0, c1) of channels owned by `rank`: [g*C/G, (g+1)*C/G)."""
    if not (0 <= rank < world) or nchan_total < 0:
        raise ValueError("bad rank/world/nchan")
    return (rank * nchan_total) // world, ((rank + 1) * nchan_total) // world


def owner_of(channel, nchan_total, world):
    """Rank that owns a global channel index (inverse of shard_range)."""
    if not (0 <= channel < nchan_total):
        raise ValueError("channel out of range")
    for r in range(world):
        c0, c1 = shard_range(nchan_total, r, world)
        if c0 <= channel < c1:
            return r
    raise AssertionError


def reduce_stats(stats):
    """stats: int64 [nchan_local][8] per-channel counters -> int64 [8] job-wide totals on every rank."""
    tot = stats.sum(0)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tot)
    return tot


def gather_records(frames, nframes, nchan_total, dst=0):
    """frames: uint8 [nchan_local][cap][64], nframes: int32 [nchan_local].  Returns on `dst` the job-wide
    (frames [nchan_total][cap][64], nframes [nchan_total]) in global channel order, elsewhere (None, None)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return frames, nframes
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(nchan_total, r, world) for r in range(world)]
    assert frames.shape[0] == sizes[rank][1] - sizes[rank][0]
    cap = frames.shape[1]
    mx = max(c1 - c0 for (c0, c1) in sizes)          # dist.gather needs equal shapes: pad ragged shards
    pad = mx - frames.shape[0]
    if pad:
        frames = torch.cat([frames, frames.new_zeros((pad, cap, 64))], 0)
        nframes = torch.cat([nframes, nframes.new_zeros((pad,))], 0)
    if rank == dst:
        fl = [torch.empty((mx, cap, 64), dtype=frames.dtype, device=frames.device) for _ in sizes]
        nl = [torch.empty((mx,), dtype=nframes.dtype, device=nframes.device) for _ in sizes]
    else:
        fl = nl = None
    dist.gather(frames.contiguous(), fl, dst=dst)
    dist.gather(nframes.contiguous(), nl, dst=dst)
    if rank != dst:
        return None, None
    return (torch.cat([f[: c1 - c0] for f, (c0, c1) in zip(fl, sizes)], 0),
            torch.cat([n[: c1 - c0] for n, (c0, c1) in zip(nl, sizes)], 0))
