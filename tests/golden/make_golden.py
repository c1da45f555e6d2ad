"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).

The reference (Rust) cannot be built or imported in this image and ships no golden vectors for this
path, so these fixtures are oracle outputs on seeded synthetic inputs: they pin the oracle against
drift (tests/test_golden.py, CPU) and give the GPU tests a checker that needs nothing but the
committed files. The analytic known answers that anchor the oracle itself live in
tests/test_oracle_known_answers.py.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {"c1": 0.075, "c2": 0.025, "c3_flat": 0.0125, "c3_sph": 0.0125, "c4": 0.025,
         # the other earth models and the Rectilinear generator
         "c3_wgs84": 0.0125, "c3_azeq": 0.0125, "c4_obsae": 0.02, "rect_c2": 0.02, "rect_c4": 0.02,
         # f4 and the Spline temperature functions
         "interp_c2": 0.025, "interp_c4": 0.025, "spline_c2": 0.025}


def tiles_digest(terrain):
    h = hashlib.sha256()
    for d, posts in terrain.tiles:
        h.update(np.ascontiguousarray(posts).tobytes())
    return h.hexdigest()


def main():
    import oracle
    from conftest import scene

    only = sys.argv[1:]
    for name, scale in CASES.items():
        if only and name not in only:
            continue
        p, terrain, objects, textures = scene(name, scale)
        r = oracle.render(p, terrain.tiles, objects, textures, max_points=12)
        cols = [0, p.width // 2, p.width - 1]
        rows = [0, p.height // 2, p.height - 1]
        if p.generator != 0:  # no image-aligned caches: the probes below describe the Fast generator's
            q = type(p).from_buffer_copy(p)
            q.generator = 0
        else:
            q = p
        tc = [oracle.terrain_cache(q, terrain.tiles, x, objects) for x in cols]
        pc = [oracle.path_cache(q, terrain.tiles, y) for y in rows]
        n = min(len(c["dist"]) for c in pc)
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            scale=scale, width=p.width, height=p.height, tiles_sha256=tiles_digest(terrain),
            rgb=r["rgb"], meta=r["meta"], steps=r["steps"], counts=r["counts"], points=r["points"],
            ray_steps=r["stats"]["ray_steps"], trace_points=r["stats"]["trace_points"],
            cols=cols, rows=rows,
            t_lat=np.stack([c["lat"] for c in tc]), t_lon=np.stack([c["lon"] for c in tc]), t_elev=np.stack([c["elev"] for c in tc]),
            t_normal=np.stack([c["normal"] for c in tc]), t_close=np.stack([c["close"] for c in tc]),
            p_dist=np.stack([c["dist"][:n] for c in pc]), p_elev=np.stack([c["elev"][:n] for c in pc]),
            p_len=np.stack([c["path_length"][:n] for c in pc]),
        )
        print(name, p.width, p.height, "hit", int((r["counts"] > 0).sum()), "max points", int(r["counts"].max()))


if __name__ == "__main__":
    main()
