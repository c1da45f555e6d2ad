"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).

CPU: the oracle still reproduces them (pins the checker against drift).
GPU: the CUDA path matches them within the north-star tolerances, without running the oracle."""
import os

import numpy as np
import pytest

from conftest import scene

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["c1", "c2", "c3_flat", "c3_sph", "c4", "c3_wgs84", "c3_azeq", "c4_obsae", "rect_c2", "rect_c4", "interp_c2", "interp_c4", "spline_c2"]


def load(name):
    g = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    p, terrain, objects, textures = scene(name, float(g["scale"]))
    assert (p.width, p.height) == (int(g["width"]), int(g["height"]))
    import hashlib

    h = hashlib.sha256()
    for _, posts in terrain.tiles:
        h.update(np.ascontiguousarray(posts).tobytes())
    assert h.hexdigest() == str(g["tiles_sha256"]), "synthetic terrain changed: regenerate the golden files"
    return g, p, terrain, objects, textures


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(oracle_lib, name):
    g, p, terrain, objects, textures = load(name)
    r = oracle_lib.render(p, terrain.tiles, objects, textures, max_points=12)
    np.testing.assert_array_equal(r["steps"], g["steps"])
    np.testing.assert_array_equal(r["counts"], g["counts"])
    np.testing.assert_array_equal(r["rgb"], g["rgb"])
    for f in ("lat", "lon", "elevation", "distance"):
        np.testing.assert_allclose(r["meta"][f], g["meta"][f], rtol=1e-12, atol=1e-9, equal_nan=True)
    assert r["stats"]["ray_steps"] == int(g["ray_steps"]) and r["stats"]["trace_points"] == int(g["trace_points"])
    for i, x in enumerate(g["cols"]):
        t = oracle_lib.terrain_cache(p, terrain.tiles, int(x), objects)
        np.testing.assert_allclose(t["elev"], g["t_elev"][i], rtol=0, atol=1e-9)
        np.testing.assert_allclose(t["normal"], g["t_normal"][i], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(t["close"], g["t_close"][i])
    for i, y in enumerate(g["rows"]):
        c = oracle_lib.path_cache(p, terrain.tiles, int(y))
        n = g["p_elev"].shape[1]
        np.testing.assert_allclose(c["elev"][:n], g["p_elev"][i], rtol=1e-12, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_golden(ctx, name):
    from test_gpu_parity import compare_render

    g, p, terrain, objects, textures = load(name)
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    ctx.set_objects(objects, textures)
    got = ctx.render()
    want = {"rgb": g["rgb"], "meta": g["meta"], "steps": g["steps"]}
    compare_render(got, want, f"golden-{name}", finish_moves_frac=0.01 if objects else 0.001)
    if name.startswith(("rect_", "interp_")):  # these generators keep no image-aligned caches to probe
        return
    for i, x in enumerate(g["cols"]):
        t = ctx.terrain_profile(int(x))
        np.testing.assert_allclose(t["lat"], g["t_lat"][i], rtol=1e-13)
        np.testing.assert_allclose(t["lon"], g["t_lon"][i], rtol=1e-13)
        np.testing.assert_allclose(t["elev"], g["t_elev"][i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(t["normal"], g["t_normal"][i], rtol=0, atol=1e-7)
    for i, y in enumerate(g["rows"]):
        c = ctx.path(int(y))
        n = min(len(c["elev"]), g["p_elev"].shape[1])
        np.testing.assert_allclose(c["dist"][:n], g["p_dist"][i][:n], rtol=1e-14)
        np.testing.assert_allclose(c["elev"][:n], g["p_elev"][i][:n], rtol=1e-9, atol=1e-5)
        np.testing.assert_allclose(c["path_length"][:n], g["p_len"][i][:n], rtol=1e-12, atol=1e-5)
