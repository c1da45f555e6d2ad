"""The InterpolatingRectilinear generator (generators/interpolating_rectilinear.rs): the oracle's
restatement against hand-checkable properties on the CPU, the device path against the oracle on the GPU."""
import numpy as np
import pytest

from atm_raytracer_b200 import abi
from conftest import scene


def small_scene(name="c2", scale=0.04, tilt=-1.0, fov=12.0):
    p, terrain, objects, textures = scene(name, scale)
    p.generator = abi.GENERATOR_INTERPOLATING_RECTILINEAR
    p.tilt, p.fov = tilt, fov
    return p, terrain, objects, textures


def test_grid_steps_follow_the_finest_pixel_spacing(oracle_lib):
    """gen_fov_data: 1.5 x the smallest angular distance between neighbouring pixels, never below fov / width / 3."""
    p, terrain, _, _ = small_scene()
    el, az = oracle_lib.pixel_angles(p)
    # the blended ResultPixel angles reproduce the pixel's own ray to within the bilinear weights' rounding
    q = abi.Params.from_buffer_copy(p)
    q.generator = abi.GENERATOR_RECTILINEAR
    el_r, az_r = oracle_lib.pixel_angles(q)
    np.testing.assert_allclose(el, el_r, atol=1e-9)
    # ... except where the two grid directions straddle north: the reference wraps each grid azimuth into [0, 360)
    # before it blends them (Cache::get_pixel, :97-103), so 359.99 and 0.01 blend to something near 180
    d = (az - az_r + 180.0) % 360.0 - 180.0
    away = np.minimum(az_r % 360.0, 360.0 - az_r % 360.0) > 1.6 * p.fov / p.width  # one grid step
    assert away.mean() > 0.8 and (~away).any()
    np.testing.assert_allclose(d[away], 0.0, atol=1e-9)
    assert np.abs(d[~away]).max() > 1.0


def test_oracle_interpolating_close_to_rectilinear(oracle_lib):
    """On smooth terrain the blend of the four surrounding grid pixels lands next to the pixel's own Rectilinear march."""
    p, terrain, _, _ = small_scene()
    got = oracle_lib.render(p, terrain.tiles)
    q = abi.Params.from_buffer_copy(p)
    q.generator = abi.GENERATOR_RECTILINEAR
    want = oracle_lib.render(q, terrain.tiles)
    hit_g, hit_w = ~np.isnan(got["meta"]["distance"]), ~np.isnan(want["meta"]["distance"])
    assert (hit_g != hit_w).mean() < 0.02  # silhouettes: a blend needs enough of its four corners
    both = hit_g & hit_w
    rel = np.abs(got["meta"]["distance"][both] / want["meta"]["distance"][both] - 1.0)
    assert np.median(rel) < 0.03 and np.mean(rel < 0.2) > 0.95  # (a pixel of this small image spans 0.16 degrees)
    diff = np.abs(got["rgb"].astype(int) - want["rgb"].astype(int)).max(axis=-1)
    assert np.mean(diff <= 8) > 0.9
    assert (got["steps"] == 0).all()


def test_oracle_interpolating_column_blocks_and_strides(oracle_lib):
    """The grid steps come from the whole image: a column block, and a strided subset, give the full render's pixels."""
    p, terrain, _, _ = small_scene(scale=0.03)
    full = oracle_lib.render(p, terrain.tiles)
    q = abi.Params.from_buffer_copy(p)
    q.x0, q.x1 = p.width // 3, p.width // 3 + 7
    part = oracle_lib.render(q, terrain.tiles)
    np.testing.assert_array_equal(part["rgb"], full["rgb"][:, q.x0:q.x1])
    np.testing.assert_array_equal(part["meta"]["distance"], full["meta"]["distance"][:, q.x0:q.x1])
    sub = oracle_lib.render(p, terrain.tiles, stride_x=3, stride_y=2)
    np.testing.assert_array_equal(sub["rgb"], full["rgb"][::2, ::3])


def test_blend_rules_on_a_step_edge(oracle_lib):
    """A pixel whose four corners are split between sky and ground keeps the trace point only on the ground's side of the
    half-way lines (interpolate_trace_points, :274-345): along the skyline of a flat plateau the blended silhouette is the
    grid's, moved by half a grid step at most."""
    p, terrain, _, _ = small_scene(tilt=0.0)
    got = oracle_lib.render(p, terrain.tiles, max_points=4)
    counts = got["counts"]
    assert counts.max() >= 1 and (counts == 0).any()
    # every emitted point is finite and its colour class survives the blend
    pts = got["points"]
    m = np.arange(pts.shape[-1])[None, None, :] < np.minimum(counts, pts.shape[-1])[..., None]
    assert np.isfinite(pts["distance"][m]).all() and (pts["is_terrain"][m] == 1).all()


# ---------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,scale,tilt", [("c2", 0.06, -1.0), ("c3_flat", 0.03, -0.5), ("c4", 0.06, -1.0), ("c1", 0.2, -2.0)])
def test_device_matches_oracle(ctx, oracle_lib, name, scale, tilt):
    from test_gpu_parity import compare_render

    p, terrain, objects, textures = small_scene(name, scale, tilt=tilt)
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    ctx.set_objects(objects, textures)
    got = ctx.render()
    want = oracle_lib.render(p, terrain.tiles, objects, textures)
    assert (got["steps"] == 0).all()
    # blended trace points: the grid's own parity (1e-6 relative) carries through the convex weights
    compare_render(got, want, "interpolating-" + name, finish_moves_frac=0.0)
    assert got["stats"]["pixels_hit"] == want["stats"]["pixels_hit"] or abs(got["stats"]["pixels_hit"] - want["stats"]["pixels_hit"]) <= max(2, got["rgb"].shape[0] * got["rgb"].shape[1] // 1000)
    assert got["stats"]["ray_steps"] >= 0.99 * want["stats"]["ray_steps"]  # the device renders the whole bounding grid, the oracle the points in use
    el, az = ctx.pixel_angles()
    el_o, az_o = oracle_lib.pixel_angles(p)
    np.testing.assert_allclose(el, el_o, atol=1e-9)
    d = (az - az_o + 180.0) % 360.0 - 180.0
    assert np.mean(np.abs(d) < 1e-9) > 0.995  # (a pixel on a grid line at north can fall to the other side by an ulp)


@pytest.mark.gpu
def test_device_column_blocks_equal_the_full_image(ctx):
    p, terrain, objects, textures = small_scene("c2", 0.05)
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    ctx.set_objects([])
    full = ctx.render()
    q = abi.Params.from_buffer_copy(p)
    q.x0, q.x1 = p.width // 4, p.width // 4 + 19
    ctx.set_params(q)
    part = ctx.render()
    np.testing.assert_array_equal(part["rgb"], full["rgb"][:, q.x0:q.x1])
    np.testing.assert_array_equal(part["meta"]["distance"], full["meta"]["distance"][:, q.x0:q.x1])
    # and the Fast generator still renders afterwards with the image's own angles
    r = abi.Params.from_buffer_copy(p)
    r.generator = abi.GENERATOR_FAST
    ctx.set_params(r)
    fast = ctx.render()
    assert fast["steps"].max() > 0 and fast["rgb"].shape == full["rgb"].shape


@pytest.mark.gpu
def test_trace_lists_of_the_blend(ctx, oracle_lib):
    p, terrain, objects, textures = small_scene("c4", 0.05)
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    ctx.set_objects(objects, textures)
    pts, cnt = ctx.render_trace(max_points=6)
    want = oracle_lib.render(p, terrain.tiles, objects, textures, max_points=6)
    assert np.mean(cnt == want["counts"]) > 0.995
    same = cnt == want["counts"]
    m = same[..., None] & (np.arange(6)[None, None, :] < np.minimum(cnt, 6)[..., None])
    for f in ("distance", "elevation", "lat", "lon"):  # (grazing crossings amplify the 1e-6 m noise of the ray altitude: compare_render's 0.1 %)
        a, b = pts[f][m], want["points"][f][m]
        assert np.mean(np.abs(a - b) <= 1e-6 + 2e-6 * np.abs(b)) > 0.998
    assert (pts["is_terrain"][m] == want["points"]["is_terrain"][m]).mean() > 0.999
