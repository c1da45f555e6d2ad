# round 2, call X (2 GPUs): torchrun bench at N = 2 with the host-side standby barrier, the reference arm under torchrun, group == context on two real GPUs
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r3i_n2.json 2> gpurun_out/r3i_n2.err; echo "n2 rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r3i_ref2.json 2> gpurun_out/r3i_ref2.err; echo "ref2 rc $?"
python - > gpurun_out/r3i_group2.log 2>&1 <<'PY'
import numpy as np, time
from atm_raytracer_b200 import runtime, config, scenes
for name, scale, gen in (("c4", 0.2, None), ("c5", 0.25, None), ("c2", 0.1, "InterpolatingRectilinear")):
    cfg, grid = scenes.make_scene(name, scale=scale)
    if gen: cfg["output"]["generator"] = gen; cfg["view"]["frame"]["fov"] = 12.0
    terrain = runtime.Terrain.from_arrays(scenes.terrain_arrays(grid))
    p = config.into_params(cfg); objects, textures = config.lower_objects(cfg)
    c = runtime.Context(0); c.set_terrain(terrain); c.set_params(p); c.set_objects(objects, textures); want = c.render(); c.close()
    g = runtime.Group(2); g.set_terrain(terrain); g.set_params(p); g.set_objects(objects, textures)
    got = g.render(); t0 = time.perf_counter(); g.set_terrain(terrain); got = g.render(); dt = time.perf_counter() - t0; g.close()
    same = all(np.array_equal(got[k], want[k]) for k in ("rgb", "steps")) and all(np.array_equal(got["meta"][f], want["meta"][f], equal_nan=True) for f in ("lat", "lon", "elevation", "distance"))
    print(name, scale, gen or "Fast", p.width, p.height, "2-GPU group == 1-GPU context:", same, "frame+terrain %.2f ms" % (dt * 1e3))
PY
cat gpurun_out/r3i_group2.log; tail -2 gpurun_out/r3i_n2.err
