import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import scene
import oracle
from atm_raytracer_b200 import runtime
ctx = runtime.Context(0)
for name, sc in (("c2", 0.05), ("c5", 0.004)):
    p, terrain, objects, textures = scene(name, sc)
    ctx.set_terrain(terrain); ctx.set_params(p); ctx.set_objects([])
    ctx.render(meta=False, steps=False)
    worst = 0.0; worstl = 0.0
    for y in range(0, p.height, max(1, p.height // 16)):
        g = ctx.path(y); w = oracle.path_cache(p, terrain.tiles, y)
        n = len(g["dist"]); ok = ~np.isnan(w["elev"][:n])
        if ok.any():
            worst = max(worst, np.abs(g["elev"][ok] - w["elev"][:n][ok]).max())
            worstl = max(worstl, np.abs(g["path_length"][ok] - w["path_length"][:n][ok]).max())
    print(name, p.width, p.height, "max |d elev| vs oracle [m]:", worst, " max |d path_length| [m]:", worstl)
