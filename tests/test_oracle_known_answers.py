"""Known-answer tests that pin the CPU oracle.

The reference ships no golden vectors for this path and cannot be built here, so the oracle is
anchored on answers that do not depend on the restatement being right by construction: analytic
geometry, published standard-atmosphere / Ciddor values, hand arithmetic.
"""
import math

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config, runtime
from conftest import ramp_tile

R = 6371000.0
SPH, FLAT = abi.EARTH_SPHERICAL, abi.EARTH_FLAT_DISTORTED


# ---- 2. bilinear sampler -------------------------------------------------------------------
def test_bilinear_exact_on_planar_ramp(oracle_lib):
    n = 121
    posts = ramp_tile(45, 5, n)
    tiles = runtime.Terrain.from_arrays([(45, 5, posts)]).tiles
    rng = np.random.default_rng(1)
    lat = 45 + rng.uniform(0, 1, 500)
    lon = 5 + rng.uniform(0, 1, 500)
    e = oracle_lib.get_elev(tiles, lat, lon)
    want = 3 * (lon - 5) * (n - 1) - 2 * (lat - 45) * (n - 1) + 500
    np.testing.assert_allclose(e, want, rtol=0, atol=1e-8)


def test_bilinear_edges_and_outside(oracle_lib):
    n = 121
    posts = ramp_tile(45, 5, n)
    tiles = runtime.Terrain.from_arrays([(45, 5, posts)]).tiles
    # corners and the max edge (fix-up int -= 1, frac += 1)
    lat = np.array([45.0, 45.0, 46.0 - 1e-12, 45.5])
    lon = np.array([5.0, 6.0 - 1e-12, 5.0, 5.5])
    e = oracle_lib.get_elev(tiles, lat, lon)
    np.testing.assert_allclose(e, [500.0, 500 + 3 * 120, 500 - 2 * 120, 500 + 180 - 120], atol=1e-6)
    # outside the tile and outside coverage -> None (NaN here); key is floor(lat), floor(lon)
    out = oracle_lib.get_elev(tiles, np.array([44.999, 46.5, 45.5, 45.5]), np.array([5.5, 5.5, 4.999, 6.5]))
    assert np.isnan(out).all()


def test_negative_posts_and_tile_lookup(oracle_lib):
    a = np.full((121, 121), -40, np.int16)
    b = np.full((121, 121), 77, np.int16)
    tiles = runtime.Terrain.from_arrays([(-1, -1, a), (0, 0, b)]).tiles
    e = oracle_lib.get_elev(tiles, np.array([-0.5, 0.5, -0.5]), np.array([-0.5, 0.5, 0.5]))
    assert e[0] == -40.0 and e[1] == 77.0 and np.isnan(e[2])


# ---- 3. directional calculators ------------------------------------------------------------
def test_spherical_calc_identities(oracle_lib):
    lat, lon = oracle_lib.coords_at_dist(SPH, R, 0.0, 0.0, 0.0, np.array([0.0, math.pi * R / 2, math.pi * R / 4]))
    np.testing.assert_allclose(lat, [0.0, 90.0, 45.0], atol=1e-9)
    lat, lon = oracle_lib.coords_at_dist(SPH, R, 0.0, 10.0, 90.0, np.array([1000.0, 250000.0]))
    np.testing.assert_allclose(lat, 0.0, atol=1e-12)
    np.testing.assert_allclose(lon, 10.0 + np.degrees(np.array([1000.0, 250000.0]) / R), rtol=1e-13)


def test_spherical_calc_haversine_round_trip(oracle_lib):
    rng = np.random.default_rng(2)
    for _ in range(20):
        lat0, lon0, az = rng.uniform(-60, 60), rng.uniform(-170, 170), rng.uniform(0, 360)
        d = rng.uniform(10.0, 400e3, 16)
        lat, lon = oracle_lib.coords_at_dist(SPH, R, lat0, lon0, az, d)
        p1, p2 = np.radians(lat0), np.radians(lat)
        dl = np.radians(lon - lon0)
        a = np.sin((p2 - p1) / 2) ** 2 + np.cos(p1) * np.cos(p2) * np.sin(dl / 2) ** 2
        back = 2 * R * np.arcsin(np.sqrt(a))
        np.testing.assert_allclose(back, d, rtol=1e-9, atol=1e-6)


def test_flat_distorted_calc_closed_form(oracle_lib):
    d = np.array([0.0, 50.0, 1e5])
    lat, lon = oracle_lib.coords_at_dist(FLAT, 0.0, 45.0, 6.0, 30.0, d)
    deg = 1e7 / 90
    np.testing.assert_allclose(lat, 45.0 + math.cos(math.radians(30)) * d / deg, rtol=1e-15)
    np.testing.assert_allclose(lon, 6.0 + math.sin(math.radians(30)) * d / deg / math.cos(math.radians(45)), rtol=1e-15)


def test_world_directions_orthonormal(oracle_lib):
    for model in (SPH, FLAT):
        n, e, u = oracle_lib.world_directions(model, R, 37.0, -122.0)
        m = np.stack([n, e, u])
        np.testing.assert_allclose(m @ m.T, np.eye(3), atol=1e-15)
        np.testing.assert_allclose(np.cross(e, n), u, atol=1e-15)  # east x north = up
    # as_cartesian: spherical radius, flat polar projection
    np.testing.assert_allclose(np.linalg.norm(oracle_lib.as_cartesian(SPH, R, 12.0, 34.0, 500.0)), R + 500.0, rtol=1e-15)
    v = oracle_lib.as_cartesian(FLAT, 0.0, 80.0, 90.0, 123.0)
    np.testing.assert_allclose(v, [0.0, 10 * 1e7 / 90, 123.0], atol=1e-6)


def test_light_direction_matches_host_mirror(oracle_lib):
    for model in (SPH, FLAT):
        a = oracle_lib.light_dir(model, R, 45.05, 6.0, 30.0, 45.0, 20.0)
        b = config.light_direction(model, 45.05, 6.0, 30.0, 45.0, 20.0)
        np.testing.assert_allclose(a, b, atol=1e-15)
        assert abs(np.linalg.norm(a) - 1.0) < 1e-15


# ---- 6. atmosphere + Ciddor ------------------------------------------------------------------
def test_us76_table_values(oracle_lib):
    h = np.array([0.0, 5000.0, 11000.0, 20000.0, 32000.0])
    t, p, n = oracle_lib.atmosphere(abi.us_76(), 530e-9, h)
    np.testing.assert_allclose(t, [288.15, 255.65, 216.65, 216.65, 228.65], atol=1e-9)
    # US Standard Atmosphere 1976 (geopotential altitude) table pressures
    np.testing.assert_allclose(p, [101325.0, 54019.9, 22632.1, 5474.89, 868.019], rtol=3e-6)
    assert np.all(np.diff(n) < 0) and 1.00027 < n[0] < 1.00029


def test_ciddor_dispersion_spot_values(oracle_lib):
    # standard dry air, 15 C, 101325 Pa, 450 ppm CO2: (n-1)*1e8 from Ciddor's dispersion formula
    assert abs((oracle_lib.air_index(633e-9, 101325.0, 288.15) - 1) - 2.76530e-4) < 5e-9
    assert abs((oracle_lib.air_index(530e-9, 101325.0, 288.15) - 1) - 2.78252e-4) < 5e-9
    # humid air is optically thinner than dry air at the same p, T
    assert oracle_lib.air_index(530e-9, 101325.0, 293.15, 1.0) < oracle_lib.air_index(530e-9, 101325.0, 293.15, 0.0)
    # n-1 scales ~ p/T
    r = (oracle_lib.air_index(530e-9, 50000.0, 250.0) - 1) / (oracle_lib.air_index(530e-9, 100000.0, 250.0) - 1)
    assert abs(r - 0.5) < 1e-3


def test_custom_atmosphere_fixed_points(oracle_lib):
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure = 100.0, 100000.0
    a.temperature_altitude, a.temperature = 500.0, 280.0
    a.n_functions = 2
    a.fn_gradient[0], a.fn_gradient[1] = -0.01, 0.002
    a.fn_start_altitude[1] = 1000.0
    t, p, _ = oracle_lib.atmosphere(a, 530e-9, np.array([100.0, 500.0, 1000.0, 1500.0]))
    np.testing.assert_allclose(t, [284.0, 280.0, 275.0, 276.0], atol=1e-12)
    assert p[0] == 100000.0 and np.all(np.diff(p) < 0)
    # hydrostatic: dp/dh = -p g M / (R T)
    hh = np.array([700.0 - 0.5, 700.0 + 0.5])
    tt, pp, _ = oracle_lib.atmosphere(a, 530e-9, hh)
    slope = (pp[1] - pp[0]) / 1.0
    want = -pp.mean() * 9.80665 * 0.0289644 / (8.31432 * tt.mean())
    assert abs(slope / want - 1) < 1e-6


# ---- 4/5. ray paths --------------------------------------------------------------------------
def test_straight_ray_closed_forms(oracle_lib):
    us = abi.us_76()
    x, h = oracle_lib.ray_path(us, 530e-9, 1, R, 1, 100.0, 2.0, 50.0, 100)  # flat
    np.testing.assert_allclose(x, 50.0 * np.arange(1, 101))
    np.testing.assert_allclose(h, 100.0 + x * math.tan(math.radians(2.0)), rtol=1e-14)
    x, h = oracle_lib.ray_path(us, 530e-9, 0, R, 1, 100.0, 0.0, 50.0, 2000)  # spherical, horizontal
    np.testing.assert_allclose(h, (R + 100.0) / np.cos(x / R) - R, rtol=0, atol=1e-6)
    # tangent height: a ray cast downward at the horizon dip grazes the sphere at R*acos(R/(R+h))
    h0 = 1000.0
    dip = math.degrees(math.acos(R / (R + h0)))
    x, h = oracle_lib.ray_path(us, 530e-9, 0, R, 1, h0, -dip, 50.0, 4000)
    k = int(np.argmin(h))
    assert abs(h[k]) < 0.1 and abs(x[k] - R * math.acos(R / (R + h0))) <= 50.0


def test_rk4_fourth_order_convergence(oracle_lib):
    us = abi.us_76()

    def h_end(step):
        n = int(round(200000.0 / step))
        x, h = oracle_lib.ray_path(us, 530e-9, 0, R, 0, 10.0, 0.2, step, n)
        assert abs(x[-1] - 200000.0) < 1e-6
        return h[-1]

    # The ray stays below the 11 km kink of US-76; steps are huge on purpose so that the RK4
    # truncation error dominates the ~1e-6 m noise of the finite-difference dn/dh.
    ref = h_end(250.0)
    e = [abs(h_end(s) - ref) for s in (100000.0, 50000.0, 25000.0)]
    assert 8.0 < e[0] / e[1] < 32.0 and 8.0 < e[1] / e[2] < 32.0


def test_refraction_coefficient_band(oracle_lib):
    us = abi.us_76()
    # horizontal ray at sea level: compare the refracted and straight heights after 20 km
    _, hr = oracle_lib.ray_path(us, 530e-9, 0, R, 0, 2.0, 0.0, 50.0, 400)
    xs, hs = oracle_lib.ray_path(us, 530e-9, 0, R, 1, 2.0, 0.0, 50.0, 400)
    # h ~ x^2/(2R) * (1-k)
    k = 1.0 - (hr[-1] - 2.0) / (hs[-1] - 2.0)
    assert 0.13 < k < 0.18
    # flat-earth refracted ray bends DOWN by ~k x^2/(2R)
    xf, hf = oracle_lib.ray_path(us, 530e-9, 1, R, 0, 2.0, 0.0, 50.0, 400)
    assert abs((2.0 - hf[-1]) / (k * xf[-1] ** 2 / (2 * R)) - 1) < 0.05


def test_path_cache_termination_and_length(oracle_lib):
    from conftest import scene

    p, terrain, _, _ = scene("c1", 0.05)
    pc = oracle_lib.path_cache(p, terrain.tiles, p.height // 2)
    # elements: x = 0, step, 2 step, ...; loop ends one element after the previous state passed max
    n = len(pc["dist"])
    assert pc["dist"][0] == 0.0 and pc["path_length"][0] == 0.0
    assert pc["dist"][n - 2] > p.max_distance >= pc["dist"][n - 3]
    assert np.all(np.diff(pc["path_length"]) > 0)
    # a steeply downward row stops at h < -1000
    pc = oracle_lib.path_cache(p, terrain.tiles, p.height - 1)
    assert pc["elev"][-2] < -1000.0 <= pc["elev"][-3]


# ---- 7/8. objects -----------------------------------------------------------------------------
def _frustum(r1, r2, height, lat=0.0, lon=0.0, a=1.0):
    o = abi.Object()
    o.kind = abi.OBJECT_FRUSTUM
    o.latitude, o.longitude = lat, lon
    o.r1, o.r2, o.height = r1, r2, height
    o.color[0], o.color[1], o.color[2], o.color[3] = 0.2, 0.4, 0.6, a
    return o


def _flat_coords(x, y, z):
    """Inverse of the FlatDistorted as_cartesian around the object at lat 0, lon 0 placed on +x axis:
    as_cartesian = ((90-lat)*DEG*cos lon, (90-lat)*DEG*sin lon, elev)."""
    deg = 1e7 / 90
    r = math.hypot(x, y)
    return [90.0 - r / deg, math.degrees(math.atan2(y, x)), z]


def test_cylinder_side_hits_hand_computed(oracle_lib):
    # FlatDistorted: object at lat 0 lon 0 -> cartesian (90*DEG, 0, 0), up = +z.
    deg = 1e7 / 90
    cx = 90 * deg
    o = _frustum(10.0, 10.0, 50.0)
    p1 = _flat_coords(cx - 30.0, 0.0, 20.0)
    p2 = _flat_coords(cx + 30.0, 0.0, 20.0)
    props, normals, colors = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2)
    np.testing.assert_allclose(props, [20.0 / 60.0, 40.0 / 60.0], atol=1e-7)
    np.testing.assert_allclose(normals[0], [-1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(normals[1], [1, 0, 0], atol=1e-6)
    np.testing.assert_allclose(colors[0], [0.2, 0.4, 0.6, 1.0])
    # above the top: no side hit
    p1[2] = p2[2] = 60.0
    props, _, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2)
    assert len(props) == 0


def test_cylinder_caps_and_half_open_interval(oracle_lib):
    deg = 1e7 / 90
    cx = 90 * deg
    o = _frustum(10.0, 10.0, 50.0)
    # vertical-ish segment through the top cap from above
    p1 = _flat_coords(cx + 1.0, 0.0, 70.0)
    p2 = _flat_coords(cx + 1.0, 0.0, 30.0)
    props, normals, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2)
    np.testing.assert_allclose(props, [0.5], atol=1e-9)
    np.testing.assert_allclose(normals[0], [0, 0, 1], atol=1e-12)
    # t in [0,1): a segment ENDING exactly on the cap plane does not hit it, one STARTING there does
    p2e = _flat_coords(cx + 1.0, 0.0, 50.0)
    props, _, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2e)
    assert len(props) == 0
    props, normals, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p2e, p1)
    assert len(props) == 1 and props[0] == 0.0


def test_cone_normal_tilt_and_apex(oracle_lib):
    deg = 1e7 / 90
    cx = 90 * deg
    o = _frustum(10.0, 0.0, 10.0)  # 45-degree cone
    p1 = _flat_coords(cx - 30.0, 0.0, 5.0)
    p2 = _flat_coords(cx + 30.0, 0.0, 5.0)
    props, normals, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2)
    np.testing.assert_allclose(props, [25.0 / 60.0, 35.0 / 60.0], atol=1e-7)
    s = math.sqrt(0.5)
    np.testing.assert_allclose(normals[0], [-s, 0, s], atol=1e-6)
    np.testing.assert_allclose(normals[1], [s, 0, s], atol=1e-6)
    # above the apex nothing is hit
    p1[2] = p2[2] = 10.5
    props, _, _ = oracle_lib.check_collision(o, None, 0.0, FLAT, 0.0, p1, p2)
    assert len(props) == 0


def _billboard(width, height, tw, th):
    o = abi.Object()
    o.kind = abi.OBJECT_BILLBOARD
    o.width, o.height = width, height
    o.texture_width, o.texture_height = tw, th
    return o


def test_billboard_centre_texel_alpha_and_edges(oracle_lib):
    deg = 1e7 / 90
    cx = 90 * deg
    tex = np.zeros((4, 4, 4), np.uint8)
    tex[..., 0] = 200
    tex[..., 3] = 255
    tex[0, :, 3] = 0  # top row transparent
    o = _billboard(40.0, 20.0, 4, 4)
    p1 = _flat_coords(cx - 30.0, 0.0, 10.0)
    p2 = _flat_coords(cx + 30.0, 0.0, 10.0)
    props, normals, colors = oracle_lib.check_collision(o, tex, 0.0, FLAT, 0.0, p1, p2)
    np.testing.assert_allclose(props, [0.5], atol=1e-9)
    np.testing.assert_allclose(colors[0], [200 / 255, 0, 0, 1.0], atol=1e-12)
    assert abs(abs(normals[0][0]) - 1.0) < 1e-9  # camera-facing plane
    # top of the billboard samples the transparent row: (v flipped) alpha 0 at y -> height
    p1[2] = p2[2] = 19.9
    _, _, colors = oracle_lib.check_collision(o, tex, 0.0, FLAT, 0.0, p1, p2)
    assert colors[0][3] == 0.0
    # half-open: y == height misses, x == +w/2 misses, x == -w/2 hits
    p1[2] = p2[2] = 20.0
    assert len(oracle_lib.check_collision(o, tex, 0.0, FLAT, 0.0, p1, p2)[0]) == 0
    # lateral extent (lat/lon round trip makes exact edges fuzzy: probe 1 mm either side)
    for dy, hit in ((20.001, 0), (-20.001, 0), (19.999, 1), (-19.999, 1)):
        a = _flat_coords(cx - 30.0, dy, 10.0)
        b = _flat_coords(cx + 30.0, dy, 10.0)
        assert len(oracle_lib.check_collision(o, tex, 0.0, FLAT, 0.0, a, b)[0]) == hit, dy


# ---- 9. colouring + compositing ---------------------------------------------------------------
def _params(**kw):
    p = abi.Params()
    p.coloring = abi.COLORING_SHADING
    p.palette = abi.PALETTE_IMPROVED
    p.ambient_light = 1.0  # brightness 1 regardless of the normal
    p.water_level = 0.0
    p.terrain_alpha = 1.0
    p.simple_max_distance = 100000.0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _tp(elev=100.0, alpha=1.0, terrain=True, rgb=(0, 0, 0), dist=1000.0, plen=1000.0, normal=(0, 0, 1)):
    import oracle

    t = np.zeros(1, oracle.TRACE_DTYPE)
    t["elevation"], t["distance"], t["path_length"] = elev, dist, plen
    t["normal"] = normal
    t["color"] = (*rgb, alpha)
    t["is_terrain"] = 1 if terrain else 0
    return t


def test_compositing_hand_arithmetic(oracle_lib):
    p = _params()
    # no trace points: sky colour of the Improved palette after truncation
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, np.zeros(0, oracle_lib.TRACE_DTYPE)), [58, 104, 140])
    # single opaque terrain hit below 300 m: green (0.4,0.8,0.3)*255 truncated, unchanged by add()
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0)), [102, 204, 76])
    # water
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(0.0)), [58, 104, 140])
    # alpha 0.5 object over an opaque terrain: truncation after every term
    pts = np.concatenate([_tp(50.0, 0.5, False, (1.0, 0.0, 0.0)), _tp(100.0)])
    c1 = np.array([255, 0, 0])
    r1 = np.floor((0 / 255 + c1 / 255 * 0.5) * 255)
    c2 = np.array([102, 204, 76])
    r2 = np.floor((r1 / 255 + c2 / 255 * 0.5) * 255)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, pts), r2.astype(np.uint8))
    # translucent chain that never saturates: the sky fills the rest
    pts = np.concatenate([_tp(50.0, 0.5, False, (1.0, 1.0, 1.0)), _tp(50.0, 0.5, False, (0.0, 0.0, 0.0))])
    sky = np.array([58, 104, 140])
    r = np.floor((0 + 255 / 255 * 0.5) * 255)
    r = np.floor((r / 255 + 0.0) * 255)
    want = np.floor((r / 255 + sky / 255 * 0.25) * 255)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, pts), want.astype(np.uint8))


def test_identity_requantisation_all_values(oracle_lib):
    # ((c/255)*255) as u8 == c for every u8: an opaque object keeps its colour
    p = _params()
    for c in range(256):
        got = oracle_lib.draw_pixel(p, _tp(10.0, 1.0, False, (c / 255, c / 255, c / 255)))
        assert got[0] in (c, c - 1) and got[0] == int((c / 255) * 255.0)


def test_shading_brightness_and_palette(oracle_lib):
    p = _params(ambient_light=0.4)
    p.light_dir[0], p.light_dir[1], p.light_dir[2] = 0.0, 0.0, 1.0
    # normal == light: brightness 1; normal perpendicular: ambient only; facing away: clamped to 0
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0, normal=(0, 0, 1))), [102, 204, 76])
    want = np.floor(np.array([0.4, 0.8, 0.3]) * 0.4 * 255)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0, normal=(1, 0, 0))), want.astype(np.uint8))
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0, normal=(0, 0, -1))), want.astype(np.uint8))
    # palette interpolation midway between 300 and 1000 m
    p = _params()
    mid = np.floor((np.array([0.77, 0.84, 0.4]) * 0.5 + np.array([0.4, 0.8, 0.3]) * 0.5) * 255)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(650.0)), mid.astype(np.uint8))
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(5000.0)), np.floor(np.array([0.85, 0.92, 0.95]) * 255).astype(np.uint8))
    # Legacy palette
    p = _params(palette=abi.PALETTE_LEGACY)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0)), [0, 255, 0])
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, np.zeros(0, oracle_lib.TRACE_DTYPE)), [28, 28, 28])


def test_simple_colors_and_fog(oracle_lib):
    p = _params(coloring=abi.COLORING_SIMPLE)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, np.zeros(0, oracle_lib.TRACE_DTYPE)), [28, 28, 28])
    # water at distance ratio 0.5: mul = 0.7
    got = oracle_lib.draw_pixel(p, _tp(0.0, dist=50000.0))
    np.testing.assert_array_equal(got, [0, int(128 * 0.7), int(255 * 0.7)])
    # land at elevation 0+ : h = 120 (green), v = 0.9, s = 1 at distance 0
    got = oracle_lib.draw_pixel(p, _tp(1e-9, dist=0.0))
    assert got[1] == int(0.9 * 255) and got[0] <= 1 and got[2] == 0
    # fog: full fog at infinite path length, def colour = fog colour
    p = _params(fog_enabled=1, fog_distance=1000.0)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, np.zeros(0, oracle_lib.TRACE_DTYPE)), [160, 160, 160])
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0, plen=1e9)), [160, 160, 160])
    f = 1 - math.exp(-1.0)
    want = np.floor(np.array([102, 204, 76]) * (1 - f) + 160 * f)
    np.testing.assert_array_equal(oracle_lib.draw_pixel(p, _tp(100.0, plen=1000.0)), want.astype(np.uint8))


# ---- generator-level known answers --------------------------------------------------------------
def test_pixel_angles_i16_centring(oracle_lib):
    p = abi.Params()
    p.width, p.height, p.fov, p.direction, p.tilt = 641, 480, 30.0, 350.0, 1.0
    d, e = oracle_lib.ray_angles(p)
    assert d[320] == 350.0 and d[0] == 350.0 + (-320 / 641) * 30.0  # odd width: integer division
    assert e[240] == 1.0 and e[0] == 1.0 - ((-240 / 480) * 30.0) / (641 / 480)


def test_result_pixel_angles_wrap_once(oracle_lib):
    """ResultPixel.azimuth is get_ray_dir(x) wrapped ONCE into [0, 360) (fast.rs:67-72); elevation_angle = get_ray_elev(y)."""
    p = abi.Params()
    p.width, p.height, p.x0, p.x1 = 640, 480, 0, 640
    p.fov, p.direction, p.tilt = 30.0, 350.0, 1.0
    el, az = oracle_lib.pixel_angles(p)
    d, e = oracle_lib.ray_angles(p)
    assert az[0, 0] == 335.0 and az[5, 320] == 350.0
    assert az[0, 639] == d[639] - 360.0 and 0.0 <= az.min() and az.max() < 360.0  # 364.95... wraps down
    np.testing.assert_array_equal(el[:, 17], e)
    p.direction = -10.0
    _, az = oracle_lib.pixel_angles(p)
    assert az[0, 0] == 335.0 and az[0, 320] == 350.0 and az[0, 639] == -10.0 + (319 / 640) * 30.0
    p.direction = 730.0  # wrapped once only, like the reference
    _, az = oracle_lib.pixel_angles(p)
    assert az[0, 320] == 370.0
    # Rectilinear: the pixel's own angles (rectilinear.rs:78-116); the centre pixel looks along (tilt, direction)
    p.direction, p.tilt, p.generator = 20.0, -3.0, abi.GENERATOR_RECTILINEAR
    el, az = oracle_lib.pixel_angles(p)
    assert abs(el[240, 320] + 3.0) < 1e-12 and abs(az[240, 320] - 20.0) < 1e-12
    assert az[240, 0] < 20.0 - 14.0 and el[0, 320] > 0.0


def test_horizon_on_a_smooth_sphere(oracle_lib):
    """Straight rays over sea-level (no tiles => elevation 0): a pixel row hits iff its ray dips
    below the geometric horizon, and the hit distance follows the sphere intersection."""
    p = abi.Params()
    p.latitude, p.longitude = 10.0, 20.0
    p.altitude.kind, p.altitude.value = abi.ALT_ABSOLUTE, 100.0
    p.fov, p.max_distance, p.simulation_step = 0.08, 100000.0, 50.0  # vertical fov = fov*H/W = 8 degrees
    p.earth_model, p.radius, p.straight_rays = SPH, R, 1
    p.atmosphere, p.wavelength = abi.us_76(), 530e-9
    p.terrain_alpha, p.coloring, p.palette, p.ambient_light = 1.0, abi.COLORING_SHADING, abi.PALETTE_IMPROVED, 0.4
    p.light_dir[2] = 1.0
    p.width, p.height, p.x0, p.x1 = 8, 800, 0, 8
    r = oracle_lib.render(p, [])
    _, elev = oracle_lib.ray_angles(p)
    dip = -math.degrees(math.acos(R / (R + 100.0)))
    hit = ~np.isnan(r["meta"]["distance"][:, 0])
    # rows strictly below the dip hit, rows above never do (one pixel row = 0.01 degrees of slack)
    assert hit[elev < dip - 0.011].all() and not hit[elev > dip + 0.001].any()
    # hit distance: (R+h) cos(a) / cos(phi + a) = R  =>  phi = acos((R+h) cos a / R) - |a|  (downward a<0)
    rows = np.where(hit)[0][::40]
    for y in rows:
        a = math.radians(elev[y])
        phi = math.acos((R + 100.0) * math.cos(a) / R)
        want = R * (-a - phi)
        assert abs(r["meta"]["distance"][y, 0] - want) < 0.05 + 1e-3 * want  # chordal interpolation inside a 50 m step
        assert abs(r["meta"]["elevation"][y, 0]) < 1e-9
    # ray-step accounting: a miss consumes all N_t - 1 zip steps
    assert r["steps"][~hit].min() == r["stats"]["n_terrain"] - 1
