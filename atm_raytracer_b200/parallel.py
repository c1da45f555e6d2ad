"""Column-block sharding of one panorama across the GPUs of a box.

Pixels are independent (fast.rs:52-92): stage A is per column, stage B per row, stage C per pixel.
Each rank renders the contiguous column block [g*W/G, (g+1)*W/G) with the same C-ABI call
(atmrt_params.x0/x1); there is no collective in the march. The only exchanges are the one-off
broadcast of the packed terrain (NCCL over NVLink) and the final gather of the image / metadata
shards to rank 0, which interleaves them into the row-major [y][x] image.

Every function takes torch tensors and works on any torch.distributed backend (nccl on the GPU box,
gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def column_blocks(width, world):
    """[(x0, x1)] per rank: contiguous blocks g*W/G .. (g+1)*W/G (integer division, like the
    reference's integer pixel arithmetic); every column belongs to exactly one rank."""
    return [(g * width // world, (g + 1) * width // world) for g in range(world)]


def shard_params(params, rank, world):
    """Copy of ``params`` restricted to this rank's column block."""
    q = type(params).from_buffer_copy(params)
    q.x0, q.x1 = column_blocks(params.width, world)[rank]
    return q


def broadcast_terrain(packed, src=0, group=None):
    """Broadcast the packed device terrain (uint8 tensor, atmrt_pack_terrain layout) from ``src``."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(packed, src=src, group=group)
    return packed


def gather_columns(shard, width, dst=0, group=None):
    """Gather [H, wl_g, ...] column shards into the full [H, W, ...] tensor on ``dst`` (None elsewhere).

    NCCL has no native ragged gather, so shards are padded to the widest block; with W divisible by
    the world size (every BASELINE config) no padding is added."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return shard
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    blocks = column_blocks(width, world)
    wmax = max(b - a for a, b in blocks)
    h = shard.shape[0]
    if shard.shape[1] != wmax:
        pad = torch.zeros((h, wmax - shard.shape[1]) + tuple(shard.shape[2:]), dtype=shard.dtype, device=shard.device)
        shard = torch.cat([shard, pad], dim=1)
    shard = shard.contiguous()
    bufs = [torch.empty_like(shard) for _ in range(world)] if rank == dst else None
    dist.gather(shard, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty((h, width) + tuple(shard.shape[2:]), dtype=shard.dtype, device=shard.device)
    for (a, b), buf in zip(blocks, bufs):
        full[:, a:b] = buf[:, : b - a]
    return full


def reduce_stats(stats, device, group=None):
    """Sum the additive counters of atmrt_stats over ranks (ray steps, trace points, hits ...)."""
    keys = ["ray_steps", "trace_points", "pixels_hit", "step_overflows", "terrain_samples", "path_steps", "kernel_launches"]
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = dict(stats)
    for k, v in zip(keys, t.tolist()):
        out[k] = int(v)
    return out
