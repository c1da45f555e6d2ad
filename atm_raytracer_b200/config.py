"""Host-side mirror of the reference's ``gen`` configuration surface.

Mirrors generator/params.rs: ``Config`` (YAML schema, :447-465, every key optional with the
reference's defaults), ``read_config`` (CLI flags override YAML, :694-777; ``--maxdist`` and
``--radius`` are km on the CLI but metres in YAML) and ``Config::into_params`` (:512-528), which is
lowered here to the flat ``atmrt_params`` POD of include/atmrt.h. Names follow the reference.
"""
import argparse
import math
import os

import numpy as np

from . import abi

DEGREE_DISTANCE = 10_000_000.0 / 90.0
EARTH_R = 6371000.0


class ConfigError(ValueError):
    pass


def _tagged(value, what):
    """serde externally-tagged enum: a bare string (unit variant) or a single-key map."""
    if isinstance(value, str):
        return value, None
    if isinstance(value, dict) and len(value) == 1:
        (k, v), = value.items()
        return k, v
    raise ConfigError(f"invalid {what}: {value!r}")


def default_config():
    """``Config::default()`` (params.rs:481-494)."""
    return {
        "scene": {"terrain_folder": "./terrain", "objects": [], "terrain_alpha": 1.0},
        "view": {
            "position": {"latitude": 0.0, "longitude": 0.0, "altitude": {"Relative": 1.0}},
            "frame": {"direction": 0.0, "tilt": 0.0, "fov": 30.0, "max_distance": 150_000.0},
            "coloring": {"Shading": {}},
            "fog_distance": None,
        },
        "atmosphere": None,  # AtmosphereDef::us_76()
        "earth_shape": {"Spherical": {"radius": 6_371_000.0}},
        "wavelength": 530e-9,
        "straight_rays": False,
        "simulation_step": 50.0,
        "output": {
            "file": "./output.png",
            "file_metadata": None,
            "width": 640,
            "height": 480,
            "ticks": [],
            "vertical_ticks": [],
            "show_eye_level": False,
            "show_flat_horizon": False,
            "generator": "Fast",
        },
    }


def _merge(base, over):
    out = dict(base)
    for k, v in (over or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and k not in ("altitude", "coloring", "earth_shape", "atmosphere"):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def parse_config(filename):
    """``parse_config`` (params.rs:678-692): YAML file -> Config dict with defaults filled in."""
    import yaml

    with open(filename) as f:
        doc = yaml.safe_load(f) or {}
    return _merge(default_config(), doc)


def subcommand_parser():
    """``subcommand_def`` (params.rs:531-676). ``-h`` is ``--height`` as in the reference."""
    p = argparse.ArgumentParser(prog="atm-raytracer gen", add_help=False, allow_abbrev=False)
    p.add_argument("--help", action="help")
    p.add_argument("-t", "--terrain")
    p.add_argument("-l", "--lat", dest="latitude")
    p.add_argument("-g", "--lon", dest="longitude")
    alt = p.add_mutually_exclusive_group()
    alt.add_argument("-a", "--alt", dest="altitude")
    alt.add_argument("-e", "--elev", dest="elevation")
    p.add_argument("-d", "--dir", dest="direction")
    p.add_argument("-f", "--fov")
    p.add_argument("-i", "--tilt")
    p.add_argument("-m", "--maxdist", dest="max_dist")
    p.add_argument("--step")
    shape = p.add_mutually_exclusive_group()
    shape.add_argument("-R", "--radius")
    shape.add_argument("--flat", action="store_true")
    p.add_argument("-s", "--straight", action="store_true")
    p.add_argument("--output")
    p.add_argument("--output-meta", dest="output_meta")
    p.add_argument("-w", "--width")
    p.add_argument("-h", "--height")
    p.add_argument("-c", "--config")
    return p


def read_config(argv):
    """``read_config`` (params.rs:694-777): YAML first, then CLI overrides."""
    m = subcommand_parser().parse_args(argv)
    cfg = parse_config(m.config) if m.config else default_config()
    if m.terrain is not None:
        cfg["scene"]["terrain_folder"] = m.terrain
    if m.output is not None:
        cfg["output"]["file"] = m.output
    if m.output_meta is not None:
        cfg["output"]["file_metadata"] = m.output_meta
    if m.width is not None:
        cfg["output"]["width"] = int(m.width)
    if m.height is not None:
        cfg["output"]["height"] = int(m.height)
    if m.latitude is not None:
        cfg["view"]["position"]["latitude"] = float(m.latitude)
    if m.longitude is not None:
        cfg["view"]["position"]["longitude"] = float(m.longitude)
    if m.altitude is not None:
        cfg["view"]["position"]["altitude"] = {"Absolute": float(m.altitude)}
    elif m.elevation is not None:
        cfg["view"]["position"]["altitude"] = {"Relative": float(m.elevation)}
    if m.direction is not None:
        cfg["view"]["frame"]["direction"] = float(m.direction)
    if m.fov is not None:
        cfg["view"]["frame"]["fov"] = float(m.fov)
    if m.tilt is not None:
        cfg["view"]["frame"]["tilt"] = float(m.tilt)
    if m.max_dist is not None:
        cfg["view"]["frame"]["max_distance"] = float(m.max_dist) * 1e3
    if m.step is not None:
        cfg["simulation_step"] = float(m.step)
    if m.flat:
        cfg["earth_shape"] = "FlatDistorted"
    elif m.radius is not None:
        cfg["earth_shape"] = {"Spherical": {"radius": float(m.radius) * 1e3}}
    if m.straight:
        cfg["straight_rays"] = True
    return cfg


def _altitude(node):
    kind, val = _tagged(node, "altitude")
    a = abi.Altitude()
    if kind == "Absolute":
        a.kind = abi.ALT_ABSOLUTE
    elif kind == "Relative":
        a.kind = abi.ALT_RELATIVE
    else:
        raise ConfigError(f"unknown altitude kind {kind}")
    a.value = float(val)
    return a


WGS84_A, WGS84_B = 6378137.0, 6356752.314245  # earth_model/mod.rs:15-16


def _earth_model(node):
    """``EarthModel`` (earth_model/mod.rs:19-28) -> (atmrt_earth_model, radius, ellipsoid_b); the parameterless
    variants are lowered as the reference lowers them (mod.rs:66-74, 127-143)."""
    kind, val = _tagged(node, "earth_shape")
    if kind == "Spherical":
        return abi.EARTH_SPHERICAL, float(val["radius"]), 0.0
    if kind == "SimpleSphere":
        return abi.EARTH_SPHERICAL, EARTH_R, 0.0
    if kind == "FlatDistorted":
        return abi.EARTH_FLAT_DISTORTED, 0.0, 0.0
    if kind == "Ellipsoid":
        return abi.EARTH_ELLIPSOID, float(val["a"]), float(val["b"])
    if kind == "Wgs84":
        return abi.EARTH_ELLIPSOID, WGS84_A, WGS84_B
    if kind == "AzimuthalEquidistant":
        return abi.EARTH_AZIMUTHAL_EQUIDISTANT, 0.0, 0.0
    if kind == "ObserverAe":
        return abi.EARTH_OBSERVER_AE, float(val["proj_radius"]), 0.0
    if kind == "SimpleObserverAe":
        return abi.EARTH_OBSERVER_AE, EARTH_R, 0.0
    raise ConfigError(f"unknown earth_shape {kind}")


def world_directions(model, lat, lon):
    """``EarthModel::world_directions`` (earth_model/mod.rs:31-57,155-172)."""
    lon_r = math.radians(lon)
    sinlon, coslon = math.sin(lon_r), math.cos(lon_r)
    if model in abi.FLAT_FAMILY:
        return (np.array([-coslon, -sinlon, 0.0]), np.array([-sinlon, coslon, 0.0]), np.array([0.0, 0.0, 1.0]))
    lat_r = math.radians(lat)
    sinlat, coslat = math.sin(lat_r), math.cos(lat_r)
    up = np.array([coslat * coslon, coslat * sinlon, sinlat])
    north = np.array([-sinlat * coslon, -sinlat * sinlon, coslat])
    east = np.array([-sinlon, coslon, 0.0])
    return north, east, up


def light_direction(model, lat, lon, direction, zenith_deg, light_dir_deg):
    """Light vector of ``ConfColoring::into_coloring`` (params.rs:243-259)."""
    zen, ld = math.radians(zenith_deg), math.radians(light_dir_deg)
    north, east, up = world_directions(model, lat, lon)
    az = math.radians(direction)
    front = north * math.cos(az) + east * math.sin(az)
    right = east * math.cos(az) - north * math.sin(az)
    v = -front * math.sin(zen) * math.cos(ld) + right * math.sin(zen) * math.sin(ld) + up * math.cos(zen)
    return v / math.sqrt(float(v @ v))


def atmosphere_def(node):
    """YAML ``atmosphere`` (README.md:281-323) -> atmrt_atmosphere_def. ``None`` = US-76."""
    if node is None:
        return abi.us_76()
    a = abi.AtmosphereDef()
    a.pressure_altitude = float(node["pressure"]["altitude"])
    a.pressure = float(node["pressure"]["pressure"])
    a.humidity = float(node.get("humidity", 0.0) or 0.0)
    fns = [(None, node["first_temperature_function"])]
    for nf in node.get("next_functions", []) or []:
        fns.append((float(nf["altitude"]), nf["function"]))
    if len(fns) > abi.MAX_ATM_FUNCTIONS:
        raise ConfigError("too many temperature functions")
    a.n_functions = len(fns)
    n_points = 0
    for i, (start, fn) in enumerate(fns):
        kind, val = _tagged(fn, "temperature function")
        a.fn_start_altitude[i] = 0.0 if start is None else start
        if kind == "Linear":
            a.fn_kind[i] = abi.FUNCTION_LINEAR
            a.fn_gradient[i] = float(val["gradient"])
        elif kind == "Spline":
            a.fn_kind[i] = abi.FUNCTION_SPLINE
            bc, bc_val = _tagged(val.get("boundary_condition", "Natural"), "boundary_condition")
            try:
                a.fn_boundary[i] = {"Natural": abi.SPLINE_NATURAL, "Derivatives": abi.SPLINE_DERIVATIVES,
                                    "SecondDerivatives": abi.SPLINE_SECOND_DERIVATIVES}[bc]
            except KeyError:
                raise ConfigError(f"unknown Spline boundary condition {bc!r}") from None
            if bc != "Natural":
                if bc_val is None or len(bc_val) != 2:
                    raise ConfigError(f"{bc} needs two numbers")
                a.fn_boundary_values[i][0], a.fn_boundary_values[i][1] = float(bc_val[0]), float(bc_val[1])
            points = val.get("points") or []
            if len(points) < 2:
                raise ConfigError("a Spline needs at least two points")
            if n_points + len(points) > abi.MAX_SPLINE_POINTS:
                raise ConfigError("too many Spline points")
            a.fn_first_point[i], a.fn_n_points[i] = n_points, len(points)
            for alt, temp in points:
                a.spline_points[n_points][0], a.spline_points[n_points][1] = float(alt), float(temp)
                n_points += 1
        else:
            raise ConfigError(f"unknown temperature function {kind!r}")
    a.n_spline_points = n_points
    tfp = node.get("temperature_fixed_point")
    if n_points == 0:
        if tfp is None:
            raise ConfigError("temperature_fixed_point is required when every function is Linear")
    if tfp is not None:  # README.md:318-323: with a Spline present it "shouldn't be present"; it is then ignored
        a.temperature_altitude = float(tfp["altitude"])
        a.temperature = float(tfp["temperature"])
    return a


def load_texture(path):
    """Decode a billboard texture to RGBA8 (object/mod.rs:60-66; `image::open` + get_pixel)."""
    from PIL import Image

    img = Image.open(path).convert("RGBA")
    return np.ascontiguousarray(np.asarray(img, dtype=np.uint8))


def lower_objects(cfg, cwd="."):
    """``ConfScene::into_scene`` (params.rs:90-107) -> (list[abi.Object], list[texture or None])."""
    objs, texs = [], []
    for node in cfg["scene"].get("objects", []) or []:
        o = abi.Object()
        pos = node["position"]
        o.latitude = float(pos.get("latitude", 0.0))
        o.longitude = float(pos.get("longitude", 0.0))
        o.altitude = _altitude(pos.get("altitude", {"Relative": 1.0}))
        kind, val = _tagged(node["shape"], "shape")
        if "color" not in node:  # `color` has no serde default (object/mod.rs:158-163): the reference rejects the config
            raise ConfigError("object needs a color")
        col = node["color"]
        o.color[0], o.color[1], o.color[2] = float(col["r"]), float(col["g"]), float(col["b"])
        o.color[3] = float(col.get("a", 1.0))
        tex = None
        if kind == "Cylinder":
            o.kind, o.r1, o.r2, o.height = abi.OBJECT_FRUSTUM, float(val["radius"]), float(val["radius"]), float(val["height"])
        elif kind == "Cone":
            o.kind, o.r1, o.r2, o.height = abi.OBJECT_FRUSTUM, float(val["radius"]), 0.0, float(val["height"])
        elif kind == "Frustum":
            o.kind, o.r1, o.r2, o.height = abi.OBJECT_FRUSTUM, float(val["r1"]), float(val["r2"]), float(val["height"])
        elif kind == "Billboard":
            o.kind, o.width, o.height = abi.OBJECT_BILLBOARD, float(val["width"]), float(val["height"])
            tex = val.get("texture")  # in-memory RGBA array (tests) ...
            if tex is None:
                tex = load_texture(os.path.join(cwd, val["texture_path"]))  # ... or a file, joined onto cwd
            tex = np.ascontiguousarray(tex, dtype=np.uint8)
            o.texture_height, o.texture_width = int(tex.shape[0]), int(tex.shape[1])
        else:
            raise ConfigError(f"unknown shape {kind}")
        objs.append(o)
        texs.append(tex)
    if len(objs) > abi.MAX_OBJECTS:
        raise ConfigError(f"at most {abi.MAX_OBJECTS} objects are supported")
    return objs, texs


def into_params(cfg, x0=None, x1=None):
    """``Config::into_params`` (params.rs:512-528) -> abi.Params."""
    p = abi.Params()
    view, frame, pos = cfg["view"], cfg["view"]["frame"], cfg["view"]["position"]
    p.latitude, p.longitude = float(pos["latitude"]), float(pos["longitude"])
    p.altitude = _altitude(pos["altitude"])
    p.direction, p.tilt = float(frame["direction"]), float(frame["tilt"])
    p.fov, p.max_distance = float(frame["fov"]), float(frame["max_distance"])
    p.earth_model, p.radius, p.ellipsoid_b = _earth_model(cfg["earth_shape"])
    p.straight_rays = 1 if cfg["straight_rays"] else 0
    p.wavelength = float(cfg["wavelength"])
    p.simulation_step = float(cfg["simulation_step"])
    p.atmosphere = atmosphere_def(cfg.get("atmosphere"))
    p.terrain_alpha = float(cfg["scene"].get("terrain_alpha", 1.0))
    kind, val = _tagged(view.get("coloring", {"Shading": {}}), "coloring")
    val = val or {}
    p.water_level = float(val.get("water_level", 0.0))
    p.simple_max_distance = p.max_distance
    if kind == "Simple":
        p.coloring = abi.COLORING_SIMPLE
    elif kind == "Shading":
        p.coloring = abi.COLORING_SHADING
        p.ambient_light = float(val.get("ambient_light", 0.4))
        pal = val.get("palette", "Improved")
        p.palette = {"Legacy": abi.PALETTE_LEGACY, "Improved": abi.PALETTE_IMPROVED}[pal]
        ld = light_direction(p.earth_model, p.latitude, p.longitude, p.direction,
                             float(val.get("light_zenith_angle", 45.0)), float(val.get("light_dir", 0.0)))
        p.light_dir[0], p.light_dir[1], p.light_dir[2] = float(ld[0]), float(ld[1]), float(ld[2])
    else:
        raise ConfigError(f"unknown coloring {kind}")
    fog = view.get("fog_distance")
    p.fog_enabled = 0 if fog is None else 1
    p.fog_distance = 0.0 if fog is None else float(fog)
    out = cfg["output"]
    p.width, p.height = int(out["width"]), int(out["height"])
    if not (0 < p.width <= 32767 and 0 < p.height <= 32767):
        raise ConfigError("width/height must fit the reference's i16 pixel centring (fast.rs:116,122)")
    gen = out.get("generator", "Fast")  # GeneratorDef, params.rs:387-392
    if gen == "Fast":
        p.generator = abi.GENERATOR_FAST
    elif gen == "Rectilinear":
        p.generator = abi.GENERATOR_RECTILINEAR
    elif gen == "InterpolatingRectilinear":
        p.generator = abi.GENERATOR_INTERPOLATING_RECTILINEAR
    else:
        raise ConfigError(f"unknown generator {gen!r}")
    p.x0 = 0 if x0 is None else int(x0)
    p.x1 = p.width if x1 is None else int(x1)
    return p
