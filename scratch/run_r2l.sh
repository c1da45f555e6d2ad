# round 2, call L (2 GPUs): regular-terrain fast path in get_elev; real 2-GPU validation of the group API and of the torchrun bench
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2l_tests.log
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2l_n1.json 2> gpurun_out/r2l_var.err
python - > gpurun_out/r2l_group2.log 2>&1 <<'PY'
# the group API on two REAL GPUs against the single-GPU render
import numpy as np, time
from atm_raytracer_b200 import runtime, config, scenes
for name, scale in (("c4", 0.2), ("c5", 0.25)):
    cfg, grid = scenes.make_scene(name, scale=scale)
    terrain = runtime.Terrain.from_arrays(scenes.terrain_arrays(grid))
    p = config.into_params(cfg); objects, textures = config.lower_objects(cfg)
    c = runtime.Context(0); c.set_terrain(terrain); c.set_params(p); c.set_objects(objects, textures); want = c.render(); c.close()
    g = runtime.Group(2); g.set_terrain(terrain); g.set_params(p); g.set_objects(objects, textures)
    got = g.render(); t0 = time.perf_counter(); g.set_terrain(terrain); got = g.render(); dt = time.perf_counter() - t0; g.close()
    same = all(np.array_equal(got[k], want[k]) for k in ("rgb", "steps")) and all(np.array_equal(got["meta"][f], want["meta"][f], equal_nan=True) for f in ("lat", "lon", "elevation", "distance"))
    print(name, scale, p.width, p.height, "2-GPU group == 1-GPU context:", same, "stats", got["stats"]["ray_steps"] == want["stats"]["ray_steps"], "frame+terrain %.2f ms" % (dt * 1e3))
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2l_n2.json 2> gpurun_out/r2l_n2.err; echo "n2 rc $?"
tail -2 gpurun_out/r2l_n2.err
