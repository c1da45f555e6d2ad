# Where k_march's time goes on c4: variants of the scene, timed with the library's own per-kernel events.
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atm_raytracer_b200 import runtime, config, scenes

def run(label, alpha=None, with_objects=True, march_mode=0, reps=4):
    cfg, grid = scenes.make_scene("c4", scale=1.0)
    terrain = runtime.Terrain.from_arrays(scenes.terrain_arrays(grid))
    p = config.into_params(cfg); objects, textures = config.lower_objects(cfg)
    if alpha is not None: p.terrain_alpha = alpha
    if not with_objects: objects, textures = [], []
    c = runtime.Context(0); c.set_terrain(terrain); c.set_params(p); c.set_objects(objects, textures); c.set_march_mode(march_mode)
    c.render(meta=False, steps=False); c.stage_times()
    for _ in range(reps): r = c.render(meta=False, steps=False)
    kt = c.kernel_times(); st = r["stats"]
    print(label, {k: round(v, 3) for k, v in kt.items()}, "trace_points", st["trace_points"], "ray_steps", st["ray_steps"], flush=True)
    c.close()

for mode in (0, 2):
    print("march mode", mode, "(0: crossing march, 2: hierarchical march)")
    run("c4 as is            ", march_mode=mode)
    run("c4 no objects a=.5  ", with_objects=False, march_mode=mode)
    run("c4 objects a=1      ", alpha=1.0, march_mode=mode)
run("c4 brute            ", march_mode=1)
