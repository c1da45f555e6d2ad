import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import scene
import oracle
from atm_raytracer_b200 import runtime
p, terrain, objects, textures = scene("c4", 0.2)
ctx = runtime.Context(0)
ctx.set_terrain(terrain); ctx.set_params(p); ctx.set_objects(objects, textures)
got = ctx.render()
want = oracle.render(p, terrain.tiles, objects, textures, max_points=24)
pts, cnt = ctx.render_trace(24)
bad = np.argwhere(got["steps"] != want["steps"])
print("nbad", len(bad), "count mismatches", (cnt != want["counts"]).sum())
for y, x in bad[:8]:
    print("pixel", y, x, "steps", got["steps"][y, x], want["steps"][y, x], "counts", cnt[y, x], want["counts"][y, x])
    n = max(cnt[y,x], want["counts"][y,x])
    for i in range(min(n, 24)):
        g, w = pts[y, x, i], want["points"][y, x, i]
        print("   g", g["step"], g["is_terrain"], g["distance"], g["elevation"], g["color"], "| w", w["step"], w["is_terrain"], w["distance"], w["elevation"], w["color"])
ys = np.unique(bad[:,0]); xs = np.unique(bad[:,1])
print("rows", ys[:20], "cols", xs[:40])
