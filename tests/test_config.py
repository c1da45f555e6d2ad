"""Host-side mirror of the reference's `gen` configuration surface (generator/params.rs)."""
import math

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config

YAML = """
scene:
  terrain_folder: ./dted
  terrain_alpha: 0.5
  objects:
    - position: {latitude: 45.1, longitude: 6.0, altitude: {Relative: 0.0}}
      shape: {Cylinder: {radius: 60.0, height: 900.0}}
      color: {r: 0.9, g: 0.1, b: 0.1}
    - position: {latitude: 45.2, longitude: 6.1, altitude: {Absolute: 1500.0}}
      shape: {Cone: {radius: 250.0, height: 1500.0}}
      color: {r: 0.1, g: 0.2, b: 0.3, a: 0.5}
    - position: {latitude: 45.3, longitude: 6.2, altitude: {Relative: 5.0}}
      shape: {Frustum: {r1: 500.0, r2: 200.0, height: 2500.0}}
      color: {r: 1.0, g: 1.0, b: 1.0}
view:
  position:
    latitude: 45.05
    longitude: 6.0
    altitude:
      Absolute: 1800.0
  frame: {direction: 12.0, tilt: -1.0, fov: 10.0, max_distance: 200000.0}
  coloring:
    Shading: {water_level: 3.0, ambient_light: 0.3, light_zenith_angle: 30.0, light_dir: 90.0, palette: Legacy}
  fog_distance: 80000.0
earth_shape: FlatDistorted
straight_rays: false
simulation_step: 25.0
output: {file: out.png, width: 1920, height: 1080}
"""


def test_defaults_match_the_reference():
    p = config.into_params(config.default_config())
    assert (p.width, p.height, p.x0, p.x1) == (640, 480, 0, 640)
    assert (p.fov, p.max_distance, p.simulation_step) == (30.0, 150000.0, 50.0)
    assert p.altitude.kind == abi.ALT_RELATIVE and p.altitude.value == 1.0
    assert p.earth_model == abi.EARTH_SPHERICAL and p.radius == 6371000.0
    assert p.wavelength == 530e-9 and p.straight_rays == 0 and p.terrain_alpha == 1.0
    assert p.coloring == abi.COLORING_SHADING and p.palette == abi.PALETTE_IMPROVED and p.ambient_light == 0.4
    assert p.fog_enabled == 0
    a = p.atmosphere  # AtmosphereDef::us_76()
    assert a.n_functions == 7 and a.pressure == 101325.0 and a.temperature == 288.15 and a.humidity == 0.0
    # default light: zenith 45 deg towards the viewer's back (light_dir 0 => -front*sin + up*cos)
    np.testing.assert_allclose(np.linalg.norm(list(p.light_dir)), 1.0, atol=1e-15)


def test_yaml_schema_and_lowering(tmp_path):
    f = tmp_path / "c.yaml"
    f.write_text(YAML)
    cfg = config.parse_config(str(f))
    p = config.into_params(cfg)
    assert p.earth_model == abi.EARTH_FLAT_DISTORTED and p.simulation_step == 25.0
    assert (p.latitude, p.longitude, p.direction, p.tilt, p.fov) == (45.05, 6.0, 12.0, -1.0, 10.0)
    assert p.altitude.kind == abi.ALT_ABSOLUTE and p.altitude.value == 1800.0
    assert p.terrain_alpha == 0.5 and p.fog_enabled == 1 and p.fog_distance == 80000.0
    assert p.palette == abi.PALETTE_LEGACY and p.water_level == 3.0 and p.ambient_light == 0.3
    assert (p.width, p.height) == (1920, 1080) and cfg["output"]["file"] == "out.png"
    assert cfg["scene"]["terrain_folder"] == "./dted"
    objs, tex = config.lower_objects(cfg)
    assert [o.kind for o in objs] == [abi.OBJECT_FRUSTUM] * 3 and tex == [None] * 3
    assert (objs[0].r1, objs[0].r2, objs[0].height) == (60.0, 60.0, 900.0)      # Cylinder: r1 == r2
    assert (objs[1].r1, objs[1].r2, objs[1].color[3]) == (250.0, 0.0, 0.5)       # Cone: r2 == 0
    assert (objs[2].r1, objs[2].r2) == (500.0, 200.0) and objs[2].color[3] == 1.0  # default alpha 1
    assert objs[1].altitude.kind == abi.ALT_ABSOLUTE and objs[2].altitude.value == 5.0


def test_cli_overrides_yaml_with_the_reference_units(tmp_path):
    f = tmp_path / "c.yaml"
    f.write_text(YAML)
    cfg = config.read_config(["-c", str(f), "-m", "100", "-R", "7000", "-w", "320", "-h", "200", "-e", "12", "-s", "--step", "50",
                              "-d", "-30", "-l", "46", "-g", "7", "-f", "45", "-i", "2.5", "--output", "x.png", "--output-meta", "x.dat",
                              "-t", "/data/terrain"])
    p = config.into_params(cfg)
    assert p.max_distance == 100e3                     # --maxdist is km on the CLI, metres in YAML
    assert p.earth_model == abi.EARTH_SPHERICAL and p.radius == 7000e3   # --radius is km
    assert (p.width, p.height) == (320, 200)            # -h is --height as in the reference
    assert p.altitude.kind == abi.ALT_RELATIVE and p.altitude.value == 12.0   # -e = height above terrain
    assert p.straight_rays == 1 and p.simulation_step == 50.0
    assert (p.direction, p.latitude, p.longitude, p.fov, p.tilt) == (-30.0, 46.0, 7.0, 45.0, 2.5)
    assert cfg["output"]["file"] == "x.png" and cfg["output"]["file_metadata"] == "x.dat"
    assert cfg["scene"]["terrain_folder"] == "/data/terrain"
    cfg = config.read_config(["--flat", "-a", "2500"])
    p = config.into_params(cfg)
    assert p.earth_model == abi.EARTH_FLAT_DISTORTED and p.altitude.kind == abi.ALT_ABSOLUTE and p.altitude.value == 2500.0
    with pytest.raises(SystemExit):
        config.read_config(["--flat", "-R", "6371"])   # conflicting shapes (the reference panics)
    with pytest.raises(SystemExit):
        config.read_config(["-a", "1", "-e", "2"])


def test_custom_atmosphere_and_scope_errors(tmp_path):
    cfg = config.default_config()
    cfg["atmosphere"] = {
        "pressure": {"altitude": 0.0, "pressure": 101325.0},
        "temperature_fixed_point": {"altitude": 0.0, "temperature": 283.0},
        "first_temperature_function": {"Linear": {"gradient": 0.01}},
        "next_functions": [{"altitude": 300.0, "function": {"Linear": {"gradient": -0.0065}}}],
    }
    a = config.into_params(cfg).atmosphere
    assert a.n_functions == 2 and a.fn_gradient[0] == 0.01 and a.fn_start_altitude[1] == 300.0 and a.temperature == 283.0
    cfg["atmosphere"]["first_temperature_function"] = {"Spline": {"points": [[0, 288.0], [100, 287.0]], "boundary_condition": "Natural"}}
    a = config.into_params(cfg).atmosphere
    assert a.fn_kind[0] == abi.FUNCTION_SPLINE and a.fn_boundary[0] == abi.SPLINE_NATURAL and a.n_spline_points == 2
    cfg["atmosphere"]["first_temperature_function"] = {"Spline": {"points": [[0, 288.0]]}}
    with pytest.raises(config.ConfigError):
        config.into_params(cfg)
    # every EarthModel of the reference lowers (earth_model/mod.rs:19-28); the parameterless ones as the reference lowers them
    for shape, want in (("Wgs84", (abi.EARTH_ELLIPSOID, 6378137.0, 6356752.314245)), ({"Ellipsoid": {"a": 7.0e6, "b": 6.9e6}}, (abi.EARTH_ELLIPSOID, 7.0e6, 6.9e6)),
                        ("AzimuthalEquidistant", (abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 0.0)), ("SimpleObserverAe", (abi.EARTH_OBSERVER_AE, 6371000.0, 0.0)),
                        ({"ObserverAe": {"proj_radius": 6.4e6}}, (abi.EARTH_OBSERVER_AE, 6.4e6, 0.0)), ("SimpleSphere", (abi.EARTH_SPHERICAL, 6371000.0, 0.0))):
        c = config.default_config()
        c["earth_shape"] = shape
        q = config.into_params(c)
        assert (q.earth_model, q.radius, q.ellipsoid_b) == want
    c = config.default_config()
    c["earth_shape"] = "Geoid"
    with pytest.raises(config.ConfigError):
        config.into_params(c)
    c = config.default_config()
    c["output"]["generator"] = "Rectilinear"
    assert config.into_params(c).generator == abi.GENERATOR_RECTILINEAR and config.into_params(config.default_config()).generator == abi.GENERATOR_FAST
    c["output"]["generator"] = "InterpolatingRectilinear"
    assert config.into_params(c).generator == abi.GENERATOR_INTERPOLATING_RECTILINEAR
    c["output"]["generator"] = "Fisheye"
    with pytest.raises(config.ConfigError):
        config.into_params(c)
    c = config.default_config()
    c["output"]["width"] = 40000
    with pytest.raises(config.ConfigError):
        config.into_params(c)
    c = config.default_config()
    c["earth_shape"] = "SimpleSphere"
    assert config.into_params(c).radius == 6371000.0


def test_light_direction_formula():
    # zenith 0: straight up, whatever the azimuths
    for model in (abi.EARTH_SPHERICAL, abi.EARTH_FLAT_DISTORTED):
        north, east, up = config.world_directions(model, 45.0, 6.0)
        np.testing.assert_allclose(config.light_direction(model, 45.0, 6.0, 77.0, 0.0, 123.0), up, atol=1e-15)
        # zenith 90, light_dir 0, view north: light comes from behind the viewer (-front = south)
        np.testing.assert_allclose(config.light_direction(model, 45.0, 6.0, 0.0, 90.0, 0.0), -north, atol=1e-15)
        # light_dir 90: from the right (east when looking north)
        np.testing.assert_allclose(config.light_direction(model, 45.0, 6.0, 0.0, 90.0, 90.0), east, atol=1e-15)


def test_scenes_cover_the_baseline_configs():
    from atm_raytracer_b200 import scenes

    for name, (w, h, n_t) in {"c1": (640, 480, 2000), "c2": (1920, 1080, 4000), "c3_flat": (3840, 1080, 4000),
                              "c3_sph": (3840, 1080, 4000), "c4": (1920, 1080, 4000), "c5": (16384, 4096, 16000)}.items():
        cfg, grid = scenes.make_scene(name)
        p = config.into_params(cfg)
        assert (p.width, p.height) == (w, h)
        assert math.ceil(p.max_distance / p.simulation_step) == n_t
    cfg, grid = scenes.make_scene("c5")
    assert grid[2] * grid[3] == 64 and config.into_params(cfg).fov == 360.0
    cfg, _ = scenes.make_scene("c4")
    objs, tex = config.lower_objects(cfg)
    assert len(objs) == 6 and sum(o.kind == abi.OBJECT_BILLBOARD for o in objs) == 1 and tex[-1].shape == (32, 64, 4)
