"""The N>1 host path on CPU: world_size-2 (and 3) gloo groups exercise the column-block partition, the
terrain broadcast and the gather that interleaves the shards into the row-major image on rank 0."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from atm_raytracer_b200 import abi, parallel


def test_column_blocks_partition():
    for w, g in [(16384, 8), (1920, 8), (641, 3), (7, 8), (640, 1)]:
        b = parallel.column_blocks(w, g)
        assert b[0][0] == 0 and b[-1][1] == w
        assert all(a[1] == c[0] for a, c in zip(b[:-1], b[1:]))
        sizes = [x1 - x0 for x0, x1 in b]
        assert max(sizes) - min(sizes) <= 1
    p = abi.Params()
    p.width, p.height, p.x0, p.x1 = 1000, 10, 0, 1000
    q = parallel.shard_params(p, 2, 3)
    assert (q.x0, q.x1) == (666, 1000) and (p.x0, p.x1) == (0, 1000)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, width, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = 5
        full_rgb = torch.arange(h * width * 3, dtype=torch.int64).reshape(h, width, 3).to(torch.uint8)
        full_meta = torch.arange(h * width * 4, dtype=torch.float64).reshape(h, width, 4)
        x0, x1 = parallel.column_blocks(width, world)[rank]
        rgb = parallel.gather_columns(full_rgb[:, x0:x1].contiguous(), width)
        meta = parallel.gather_columns(full_meta[:, x0:x1].contiguous(), width)
        packed = torch.arange(1000, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.zeros(1000, dtype=torch.uint8)
        parallel.broadcast_terrain(packed)
        stats = {"ray_steps": 10 + rank, "trace_points": 1, "pixels_hit": rank, "step_overflows": 0, "terrain_samples": 100,
                 "path_steps": 7, "kernel_launches": 9, "n_terrain": 2000}
        red = parallel.reduce_stats(stats, torch.device("cpu"))
        ok = bool((packed == torch.arange(1000, dtype=torch.int64).to(torch.uint8)).all())
        ok = ok and red["ray_steps"] == sum(10 + r for r in range(world)) and red["n_terrain"] == 2000
        if rank == 0:
            ok = ok and torch.equal(rgb, full_rgb) and torch.equal(meta, full_meta)
        else:
            ok = ok and rgb is None and meta is None
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width", [(2, 64), (2, 63), (3, 100)])
def test_gather_and_broadcast_over_gloo(world, width):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, width, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True) for r in range(world)]
