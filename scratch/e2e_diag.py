import sys, time, torch, numpy as np
sys.path.insert(0, ".")
import bench
from atm_raytracer_b200 import runtime, parallel
cfg, params, terrain, objects, textures = bench.build_workload("c5", 1.0, True)
dev = torch.device("cuda", 0)
ctx = runtime.Context(0)
nbytes = ctx.packed_bytes(terrain)
packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
pinned = [torch.from_numpy(p).pin_memory() for _, p in terrain.tiles]
tp = runtime.Terrain([(d, t_.numpy()) for (d, _), t_ in zip(terrain.tiles, pinned)])
ctx.pack_terrain(tp, packed.data_ptr()); ctx.bind_terrain(terrain, packed.data_ptr()); ctx.set_params(params); ctx.set_objects([], [])
H, W = params.height, params.width
rgb = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
host_rgb = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
s = torch.cuda.Stream(device=dev); torch.cuda.set_stream(s)
def T(f, n=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("pack_terrain ms", T(lambda: ctx.pack_terrain(tp, packed.data_ptr())))
print("render (no meta) ms", T(lambda: ctx.render_device(rgb.data_ptr(), 0, 0, s.cuda_stream)))
print("d2h rgb ms", T(lambda: host_rgb.copy_(rgb, non_blocking=True)))
def full():
    ctx.pack_terrain(tp, packed.data_ptr()); ctx.render_device(rgb.data_ptr(), 0, 0, s.cuda_stream); host_rgb.copy_(rgb, non_blocking=True); torch.cuda.synchronize()
print("full ms", T(full))
