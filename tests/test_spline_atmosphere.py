"""Spline temperature functions (README.md:296-323 of the reference): the oracle's lowering against closed forms, the
YAML front ends, and -- on the GPU -- the device atmosphere, the g(h) table and whole ray paths against the oracle."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config

ROOT = Path(__file__).resolve().parent.parent
G, M, R = 9.80665, 0.0289644, 8.31432


def spline_function(a, i, start, points, boundary=abi.SPLINE_NATURAL, values=(0.0, 0.0)):
    a.fn_kind[i], a.fn_start_altitude[i] = abi.FUNCTION_SPLINE, start
    a.fn_boundary[i] = boundary
    a.fn_boundary_values[i][0], a.fn_boundary_values[i][1] = values
    a.fn_first_point[i], a.fn_n_points[i] = a.n_spline_points, len(points)
    for h, t in points:
        a.spline_points[a.n_spline_points][0], a.spline_points[a.n_spline_points][1] = h, t
        a.n_spline_points += 1


def ducting_atmosphere():
    """A surface inversion as a Spline between two Linear functions -- the README's own use of Spline."""
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure = 0.0, 101325.0
    a.n_functions = 3
    a.fn_gradient[0] = -0.0065
    spline_function(a, 1, 100.0, [(100.0, 288.0), (160.0, 291.5), (240.0, 293.0), (400.0, 291.0), (900.0, 287.5)],
                    abi.SPLINE_DERIVATIVES, (-0.0065, -0.0065))
    a.fn_start_altitude[2], a.fn_gradient[2] = 900.0, -0.0065
    return a


def test_spline_through_a_straight_line_is_the_linear_law(oracle_lib):
    """Points on T = 288.15 - 0.0065 h with matching end derivatives: the spline is that line, so the quadrature of the
    hydrostatic integral must reproduce the closed-form power law of a Linear function."""
    lapse = -0.0065
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure = 0.0, 101325.0
    a.n_functions = 1
    knots = [0.0, 700.0, 1500.0, 4000.0, 11000.0]
    spline_function(a, 0, 0.0, [(h, 288.15 + lapse * h) for h in knots], abi.SPLINE_DERIVATIVES, (lapse, lapse))
    h = np.concatenate([np.linspace(-400.0, 12000.0, 500), knots])
    t, p, n = oracle_lib.atmosphere(a, 530e-9, h)
    np.testing.assert_allclose(t, 288.15 + lapse * h, rtol=1e-14)
    np.testing.assert_allclose(p, 101325.0 * ((288.15 + lapse * h) / 288.15) ** (-G * M / (R * lapse)), rtol=2e-13)
    tu, pu, nu = oracle_lib.atmosphere(abi.us_76(), 530e-9, h[h < 11000.0])
    np.testing.assert_allclose(n[h < 11000.0] - 1.0, nu - 1.0, rtol=1e-12)


def test_natural_spline_closed_form(oracle_lib):
    """Three points, natural ends: the one interior second derivative is 3 (s1 - s0) / (h0 + h1)."""
    pts = [(0.0, 290.0), (100.0, 292.0), (300.0, 289.0)]
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure = 50.0, 100000.0
    a.n_functions = 1
    spline_function(a, 0, 0.0, pts)
    s0, s1 = (292.0 - 290.0) / 100.0, (289.0 - 292.0) / 200.0
    m1 = 3.0 * (s1 - s0) / (100.0 + 200.0)

    def want(h):
        if h < 100.0:
            x, d, y0, y1, z0, z1 = h, 100.0, 290.0, 292.0, 0.0, m1
        else:
            x, d, y0, y1, z0, z1 = h - 100.0, 200.0, 292.0, 289.0, m1, 0.0
        return y0 + x * ((y1 - y0) / d - d * (2 * z0 + z1) / 6) + x * x * z0 / 2 + x ** 3 * (z1 - z0) / (6 * d)

    h = np.array([-50.0, 0.0, 30.0, 99.999, 100.0, 100.001, 250.0, 300.0, 420.0])
    t, p, _ = oracle_lib.atmosphere(a, 530e-9, h)
    np.testing.assert_allclose(t, [want(x) for x in h], rtol=1e-14)
    assert t[1] == 290.0 and t[4] == 292.0 and abs(t[7] - 289.0) < 1e-12
    # hydrostatic everywhere, also in the continued end segments and across the knot
    for h0 in (-40.0, 50.0, 100.0, 280.0, 400.0):
        hh = np.array([h0 - 0.5, h0, h0 + 0.5])
        tt, pp, _ = oracle_lib.atmosphere(a, 530e-9, hh)
        assert abs((pp[2] - pp[0]) / (-pp[1] * G * M / (R * tt[1])) - 1.0) < 1e-7
    _, p50, _ = oracle_lib.atmosphere(a, 530e-9, np.array([50.0]))
    assert p50[0] == 100000.0


@pytest.mark.parametrize("boundary,values", [(abi.SPLINE_DERIVATIVES, (0.01, -0.004)), (abi.SPLINE_SECOND_DERIVATIVES, (2e-5, -1e-5)),
                                             (abi.SPLINE_NATURAL, (0.0, 0.0))])
def test_boundary_conditions_and_smoothness(oracle_lib, boundary, values):
    pts = [(100.0, 288.0), (110.0, 285.0), (120.0, 291.0), (150.0, 290.0)]
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure = 0.0, 101325.0
    a.n_functions = 1
    spline_function(a, 0, 0.0, pts, boundary, values)
    T = lambda h: oracle_lib.atmosphere(a, 530e-9, np.asarray(h, float))[0]
    np.testing.assert_allclose(T([h for h, _ in pts]), [t for _, t in pts], rtol=1e-14)
    e = 1e-3
    d1 = lambda h: (T([h + e])[0] - T([h - e])[0]) / (2 * e)
    d2 = lambda h: (T([h + e])[0] - 2 * T([h])[0] + T([h - e])[0]) / (e * e)
    for knot in (110.0, 120.0):  # C2 at the interior points
        assert abs(d1(knot - 2 * e) - d1(knot + 2 * e)) < 5e-3  # |T"| 4e-3 m
        assert abs(d2(knot - 2 * e) - d2(knot + 2 * e)) < 2e-2
    if boundary == abi.SPLINE_DERIVATIVES:
        assert abs(d1(100.0) - values[0]) < 1e-6 and abs(d1(150.0) - values[1]) < 1e-6
    else:
        assert abs(d2(100.0) - values[0]) < 1e-4 and abs(d2(150.0) - values[1]) < 1e-4


def test_linear_neighbours_join_the_spline(oracle_lib):
    a = ducting_atmosphere()
    e = 1e-9
    h = np.array([100.0 - e, 100.0, 900.0 - e, 900.0, 0.0, 50.0, 2000.0])
    t, p, n = oracle_lib.atmosphere(a, 530e-9, h)
    assert abs(t[0] - t[1]) < 1e-10 and abs(t[2] - t[3]) < 1e-10 and t[1] == 288.0
    assert abs(p[0] / p[1] - 1) < 1e-12 and abs(p[2] / p[3] - 1) < 1e-12
    np.testing.assert_allclose(t[4:], [288.0 + 0.0065 * 100.0, 288.0 + 0.0065 * 50.0, 287.5 - 0.0065 * 1100.0], rtol=1e-14)
    assert p[4] == 101325.0
    hh = np.linspace(-200.0, 3000.0, 3201)
    tt, pp, nn = oracle_lib.atmosphere(a, 530e-9, hh)
    assert np.all(np.diff(pp) < 0)
    assert tt[(hh > 100) & (hh < 240)].max() > 292.9  # the inversion is there
    # the temperature fixed point is ignored beside a Spline (README.md:318-323)
    b = ducting_atmosphere()
    b.temperature_altitude, b.temperature = 0.0, 250.0
    np.testing.assert_array_equal(oracle_lib.atmosphere(b, 530e-9, hh)[0], tt)


def test_invalid_splines_are_rejected(oracle_lib):
    a = abi.AtmosphereDef()
    a.pressure_altitude, a.pressure, a.n_functions = 0.0, 101325.0, 1
    spline_function(a, 0, 0.0, [(0.0, 288.0)])
    with pytest.raises(Exception):
        oracle_lib.atmosphere(a, 530e-9, np.array([0.0]))
    b = abi.AtmosphereDef()
    b.pressure_altitude, b.pressure, b.n_functions = 0.0, 101325.0, 1
    spline_function(b, 0, 0.0, [(0.0, 288.0), (0.0, 280.0)])
    with pytest.raises(Exception):
        oracle_lib.atmosphere(b, 530e-9, np.array([0.0]))


README_YAML = """
atmosphere:
    pressure:
        altitude: 0.0
        pressure: 101325
    first_temperature_function:
        Linear:
            gradient: -0.0065
    next_functions:
        - altitude: 100.0
          function:
            Spline:
                boundary_condition:
                    Derivatives:
                        - -0.0065
                        - 0.0
                points:
                    -
                        - 100.0
                        - 288.0
                    -
                        - 110.0
                        - 285.0
                    -
                        - 120.0
                        - 291.0
"""


def test_readme_example_parses(oracle_lib):
    import yaml

    a = config.atmosphere_def(yaml.safe_load(README_YAML)["atmosphere"])
    assert a.n_functions == 2 and a.fn_kind[0] == abi.FUNCTION_LINEAR and a.fn_kind[1] == abi.FUNCTION_SPLINE
    assert a.fn_boundary[1] == abi.SPLINE_DERIVATIVES and list(a.fn_boundary_values[1]) == [-0.0065, 0.0]
    assert a.n_spline_points == 3 and list(a.spline_points[2]) == [120.0, 291.0]
    t, _, _ = oracle_lib.atmosphere(a, 530e-9, np.array([0.0, 100.0, 110.0, 120.0]))
    np.testing.assert_allclose(t, [288.65, 288.0, 285.0, 291.0], rtol=1e-14)
    with pytest.raises(config.ConfigError):
        config.atmosphere_def({"pressure": {"altitude": 0, "pressure": 1e5}, "first_temperature_function": {"Linear": {"gradient": 0.0}}})
    with pytest.raises(config.ConfigError):
        config.atmosphere_def({"pressure": {"altitude": 0, "pressure": 1e5},
                               "first_temperature_function": {"Spline": {"points": [[0.0, 288.0]]}}})


def test_output_atm_with_a_spline(oracle_lib, tmp_path):
    """The oracle's `output-atm` on the README atmosphere: the spline's own points come back in the table."""
    import yaml

    a = config.atmosphere_def(yaml.safe_load(README_YAML)["atmosphere"])
    text = oracle_lib.output_atm(a, min_alt=100.0, max_alt=120.0, step=10.0)
    rows = [line.split() for line in text.strip().splitlines() if line and line[0].isdigit()]
    temps = [float(r[1]) for r in rows]
    np.testing.assert_allclose(temps[:3], [288.0, 285.0, 291.0], rtol=1e-14)


# ---------------------------------------------------------------------------------------------
# device
# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_atmosphere_and_table_with_a_spline(ctx, oracle_lib):
    from test_gpu_parity import scene

    p, terrain, _, _ = scene("c2", 0.04)
    p.atmosphere = ducting_atmosphere()
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    rng = np.random.default_rng(11)
    h = np.concatenate([rng.uniform(-500.0, 5000.0, 20000), [100.0, 160.0, 240.0, 400.0, 900.0], 240.0 + rng.uniform(-3, 3, 200)])
    t, pr, n = ctx.atmosphere_probe(h)
    to, po, no = oracle_lib.atmosphere(p.atmosphere, p.wavelength, h)
    # (library and oracle solve the spline's tridiagonal system in different orders: the cubics agree to a few ulp)
    np.testing.assert_allclose(t, to, rtol=2e-15)
    np.testing.assert_allclose(pr, po, rtol=1e-13)
    np.testing.assert_allclose(n - 1.0, no - 1.0, rtol=1e-11)
    gt, gl, served = ctx.refraction_probe(h, with_pieces=True)
    assert not np.isnan(gt).any()
    knots = np.array([100.0, 160.0, 240.0, 400.0, 900.0])
    away = np.abs(h[:, None] - knots[None, :]).min(axis=1) > 0.02
    rel = gt[away] / gl[away] - 1.0
    assert np.abs(rel).max() < 2e-5 and abs(rel.mean()) < 1e-7, (np.abs(rel).max(), rel.mean())


@pytest.mark.gpu
@pytest.mark.parametrize("flat", [False, True])
def test_ray_paths_through_a_spline_inversion(ctx, oracle_lib, flat):
    from test_gpu_parity import PATH_ATOL, compare_render, scene

    p, terrain, _, _ = scene("c3_flat" if flat else "c2", 0.04)
    p.atmosphere = ducting_atmosphere()
    p.tilt, p.fov = 0.2, 6.0
    ctx.set_terrain(terrain)
    ctx.set_params(p)
    ctx.set_objects([])
    got = ctx.render()
    rows = (0, p.height // 3, p.height // 2, p.height - 1)
    table = {y: ctx.path(y) for y in rows}
    for y in rows:
        g, w = table[y], oracle_lib.path_cache(p, terrain.tiles, y)
        k = len(g["dist"])
        np.testing.assert_allclose(g["elev"], w["elev"][:k], rtol=1e-9, atol=PATH_ATOL)
        np.testing.assert_allclose(g["path_length"], w["path_length"][:k], rtol=1e-11, atol=1e-9)
    compare_render(got, oracle_lib.render(p, terrain.tiles), "spline-" + ("flat" if flat else "sph"))
    ctx.set_path_mode(1)
    try:
        ctx.render(meta=False, steps=False)
        for y in rows:
            g, w = ctx.path(y), table[y]
            np.testing.assert_allclose(g["elev"], w["elev"], rtol=1e-9, atol=PATH_ATOL)
    finally:
        ctx.set_path_mode(0)


@pytest.mark.gpu
def test_gen_executable_accepts_the_readme_atmosphere(tmp_path):
    """`atm-raytracer output-atm <README atmosphere>` through the C++ host's own YAML reader (`-c` is --celsius there)."""
    cfg = tmp_path / "atm.yaml"
    cfg.write_text(README_YAML)
    exe = ROOT / "atm_raytracer_b200" / "atm-raytracer"
    out = subprocess.run([str(exe), "output-atm", str(cfg), "--min-alt", "100", "--max-alt", "120", "--step", "10"],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    rows = [line.split() for line in out.stdout.strip().splitlines() if line and line[0].isdigit()]
    np.testing.assert_allclose([float(r[1]) for r in rows[:3]], [288.0, 285.0, 291.0], rtol=1e-14)
