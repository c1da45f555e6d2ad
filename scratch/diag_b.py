import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import scene
from atm_raytracer_b200 import runtime
ctx = runtime.Context(0)
p, terrain, _, _ = scene("c5", 0.02)   # 328 x 82, N_t = 16000
ctx.set_terrain(terrain); ctx.set_objects([])
W, H = p.width, p.height
print("size", W, H, "fov", p.fov, "vertical fov", p.fov * H / W)
# rows per tilt window: vertical fov = fov*H/W ; use 32 rows => one warp
p.height = 32
for tilt in (-40, -20, -10, -5, -2, 0, 1, 2, 3, 5, 10, 20, 24, 30, 40):
    for vf in (0.7,):  # degrees spanned by the 32 rows (c5 spacing: 90/4096*32 = 0.7)
        p.tilt = float(tilt)
        p.fov = vf * p.width / p.height
        ctx.set_params(p)
        ctx.render(meta=False, steps=False); ctx.stage_times()
        ctx.render(meta=False, steps=False)
        st = ctx.stage_times()
        n = [len(ctx.path(y)["dist"]) for y in (0, 16, 31)]
        e = [float(np.nanmax(ctx.path(y)["elev"])) for y in (0, 31)]
        nan = [int(np.isnan(ctx.path(y)["elev"]).sum()) for y in (0, 31)]
        print(f"tilt {tilt:4d} paths {st['ms_paths']:.3f} ms  n {n} top {e} nan {nan}")
