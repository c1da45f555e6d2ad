"""The five BASELINE.json workloads as reference-style configs on synthetic DTED tiles.

Each scene is ``(cfg, tile_grid)`` where ``cfg`` is a Config dict in the reference's YAML schema
(see config.py) and ``tile_grid = (lat0, lon0, nlat_tiles, nlon_tiles)`` names the synthetic tiles
(synth.py). ``scale`` shrinks the image for parity tests without changing the geometry.
"""
import copy

import numpy as np

from . import config as cfgmod
from . import synth

NAMES = ("c1", "c2", "c3_flat", "c3_sph", "c4", "c5")


def _base():
    return copy.deepcopy(cfgmod.default_config())


def billboard_texture():
    """Seeded 64x32 RGBA texture with fully transparent and fully opaque texels."""
    rng = np.random.default_rng(synth.SEED + 7)
    tex = rng.integers(0, 256, size=(32, 64, 4), dtype=np.uint8)
    tex[:, :, 3] = 255
    tex[:8, :, 3] = 0  # transparent top band
    tex[8:16, ::2, 3] = 128  # half-transparent stripes
    tex[20:28, 20:44, :3] = (200, 30, 30)
    return tex


def scene_objects():
    """2 cylinders, 2 cones, 1 frustum, 1 billboard, 5-40 km out along the c2/c4 view axis."""
    return [
        {"position": {"latitude": 45.10, "longitude": 6.000, "altitude": {"Relative": 0.0}},
         "shape": {"Cylinder": {"radius": 60.0, "height": 900.0}}, "color": {"r": 0.9, "g": 0.1, "b": 0.1}},
        {"position": {"latitude": 45.16, "longitude": 6.012, "altitude": {"Absolute": 1500.0}},
         "shape": {"Cylinder": {"radius": 120.0, "height": 400.0}}, "color": {"r": 0.1, "g": 0.1, "b": 0.9, "a": 0.5}},
        {"position": {"latitude": 45.22, "longitude": 5.990, "altitude": {"Relative": 0.0}},
         "shape": {"Cone": {"radius": 250.0, "height": 1500.0}}, "color": {"r": 0.9, "g": 0.9, "b": 0.1}},
        {"position": {"latitude": 45.30, "longitude": 6.025, "altitude": {"Relative": 10.0}},
         "shape": {"Cone": {"radius": 400.0, "height": 2200.0}}, "color": {"r": 0.1, "g": 0.9, "b": 0.9, "a": 1.0}},
        {"position": {"latitude": 45.40, "longitude": 5.975, "altitude": {"Relative": 0.0}},
         "shape": {"Frustum": {"r1": 500.0, "r2": 200.0, "height": 2500.0}}, "color": {"r": 0.8, "g": 0.3, "b": 0.8}},
        {"position": {"latitude": 45.13, "longitude": 5.992, "altitude": {"Absolute": 1400.0}},
         "shape": {"Billboard": {"width": 600.0, "height": 600.0, "texture": billboard_texture()}},
         "color": {"r": 1.0, "g": 1.0, "b": 1.0}},
    ]


EARTH_VARIANTS = {
    "c3_wgs84": "Wgs84",
    "c3_ellipsoid": {"Ellipsoid": {"a": 6_400_000.0, "b": 6_300_000.0}},
    "c3_azeq": "AzimuthalEquidistant",
    "c3_obsae": {"ObserverAe": {"proj_radius": 6_371_000.0}},
    "c3_simple": "SimpleSphere",
    "c4_wgs84": "Wgs84",
    "c4_azeq": "AzimuthalEquidistant",
    "c4_obsae": "SimpleObserverAe",
}


def make_scene(name, scale=1.0):
    if name.startswith("rect_"):  # the same scene through the Rectilinear generator, looking slightly down
        c, grid = make_scene(name[5:], scale)
        c["output"]["generator"] = "Rectilinear"
        c["view"]["frame"]["tilt"] = -1.5
        return c, grid
    if name.startswith("interp_"):  # ... and through the InterpolatingRectilinear generator
        c, grid = make_scene(name[7:], scale)
        c["output"]["generator"] = "InterpolatingRectilinear"
        c["view"]["frame"]["tilt"] = -1.5
        return c, grid
    if name.startswith("spline_"):  # a surface inversion as a Spline temperature function between two Linear ones (README.md:296-316)
        c, grid = make_scene(name[7:], scale)
        c["atmosphere"] = {
            "pressure": {"altitude": 0.0, "pressure": 101325.0},
            "first_temperature_function": {"Linear": {"gradient": -0.0065}},
            "next_functions": [
                {"altitude": 1700.0, "function": {"Spline": {"boundary_condition": {"Derivatives": [-0.0065, -0.0065]},
                                                            "points": [[1700.0, 277.0], [1760.0, 280.5], [1840.0, 282.0], [2000.0, 280.0], [2500.0, 276.5]]}}},
                {"altitude": 2500.0, "function": {"Linear": {"gradient": -0.0065}}},
            ],
        }
        return c, grid
    c = _base()
    out, view = c["output"], c["view"]

    def size(w, h):
        out["width"] = max(8, int(round(w * scale)))
        out["height"] = max(8, int(round(h * scale)))

    if name == "c1":
        # straight rays, spherical R=6371 km, 640x480, fov 30, one tile, 100 km, step 50
        c["straight_rays"] = True
        c["earth_shape"] = {"Spherical": {"radius": 6_371_000.0}}
        size(640, 480)
        view["frame"].update(direction=0.0, tilt=0.0, fov=30.0, max_distance=100_000.0)
        view["position"] = {"latitude": 45.05, "longitude": 5.5, "altitude": {"Relative": 2.0}}
        c["simulation_step"] = 50.0
        grid = (45, 5, 1, 1)
    elif name in ("c2", "c4", "c3_flat", "c3_sph") or name in EARTH_VARIANTS:
        # refracted, US-76, 1920x1080, fov 10, 2x2 tiles, 200 km, step 50
        size(3840 if name.startswith("c3") else 1920, 1080)
        view["frame"].update(direction=0.0, tilt=0.0, fov=20.0 if name.startswith("c3") else 10.0, max_distance=200_000.0)
        view["position"] = {"latitude": 45.05, "longitude": 6.0, "altitude": {"Absolute": 1800.0}}
        c["simulation_step"] = 50.0
        grid = (45, 5, 2, 2)
        if name == "c2":
            out["file_metadata"] = "./output.dat"  # BASELINE config 2 is the one "with --output-meta"
        if name == "c3_flat":
            c["earth_shape"] = "FlatDistorted"
        if name == "c4" or name.startswith("c4_"):
            c["scene"]["objects"] = scene_objects()
            c["scene"]["terrain_alpha"] = 0.5
        if name in EARTH_VARIANTS:  # the remaining earth models, on c3's / c4's geometry
            c["earth_shape"] = EARTH_VARIANTS[name]
    elif name == "c5":
        # 360-degree 16384x4096 refracted panorama over 8x8 tiles, 400 km, step 25
        size(16384, 4096)
        view["frame"].update(direction=0.0, tilt=0.0, fov=360.0, max_distance=400_000.0)
        view["position"] = {"latitude": 45.0, "longitude": 9.0, "altitude": {"Absolute": 2500.0}}
        c["simulation_step"] = 25.0
        grid = (41, 5, 8, 8)
    else:
        raise KeyError(name)
    return c, grid


_TILE_CACHE = {}


def terrain_arrays(grid, level=1):
    """Synthetic tiles of ``grid`` as a list of (lat0, lon0, posts); cached per process."""
    lat0, lon0, nlat, nlon = grid
    out = []
    for i in range(nlat):
        for j in range(nlon):
            key = (lat0 + i, lon0 + j, level)
            if key not in _TILE_CACHE:
                _TILE_CACHE[key] = synth.make_tile(lat0 + i, lon0 + j, level)
            out.append((lat0 + i, lon0 + j, _TILE_CACHE[key]))
    return out
