"""ctypes binding of the CPU ORACLE (oracle/atmrt_oracle.cpp). TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
nothing under atm_raytracer_b200/ may import this package (tests/test_boundary.py enforces it).
The oracle is a C++ restatement of the reference's Fast-generator path; the reference itself (Rust)
cannot be built in this image, so there is no oracle/_ref. PARITY UNPINNED for the two external
crates (atm-refraction 0.6, dted 0.2) -- see the header of atmrt_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "atmrt_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _abi():
    from atm_raytracer_b200 import abi  # POD mirrors of include/atmrt.h only (no CUDA library needed)

    return abi


class Timing(C.Structure):
    _fields_ = [("s_terrain", C.c_double), ("s_paths", C.c_double), ("s_pixels", C.c_double), ("s_image", C.c_double),
                ("s_total", C.c_double), ("threads", C.c_int), ("_pad", C.c_int)]


_lib = None
BUILD = "-O2 -ffp-contract=off (portable: liboracle.so)"
NATIVE_LIB_PATH = os.path.join(_HERE, "liboracle_native.so")


def use_native():
    """Switch this process to the `baseline` build of the same source (-O3 -march=native, oracle/Makefile), compiled
    here and now for this host's cores -- what bench.py's CPU legs time. Falls back to the portable build (and says
    so in BUILD) when the compiler is not available. Must be called before the first use of lib()."""
    global _lib, BUILD
    if _lib is not None and BUILD.startswith("-O3"):
        return BUILD
    try:
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s", "baseline"], stdout=subprocess.DEVNULL)
        _lib = _bind(C.CDLL(NATIVE_LIB_PATH))
        BUILD = "-O3 -march=native -ffp-contract=off (liboracle_native.so, built on this host)"
    except Exception as e:  # noqa: BLE001
        BUILD = f"-O2 -ffp-contract=off (portable liboracle.so; native build failed: {type(e).__name__})"
    return BUILD


def _bind(L):
    L.oracle_air_index.restype = C.c_double
    L.oracle_air_index.argtypes = [C.c_double] * 4
    L.oracle_num_threads.restype = C.c_int
    L.oracle_set_num_threads.argtypes = [C.c_int]
    return L


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_air_index.restype = C.c_double
        _lib.oracle_air_index.argtypes = [C.c_double] * 4
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


META_DTYPE = np.dtype([("lat", "<f8"), ("lon", "<f8"), ("elevation", "<f8"), ("distance", "<f8")])
TRACE_DTYPE = np.dtype(
    [("lat", "<f8"), ("lon", "<f8"), ("distance", "<f8"), ("elevation", "<f8"), ("path_length", "<f8"),
     ("normal", "<f8", 3), ("color", "<f8", 4), ("is_terrain", "<i4"), ("step", "<i4")]
)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _tiles(tiles):
    """tiles: list of (abi.TileDesc, int16 posts[nlon][nlat])."""
    abi = _abi()
    n = len(tiles)
    descs = (abi.TileDesc * max(n, 1))()
    ptrs = (C.c_void_p * max(n, 1))()
    for i, (d, posts) in enumerate(tiles):
        assert posts.dtype == np.int16 and posts.flags["C_CONTIGUOUS"]
        descs[i] = d
        ptrs[i] = posts.ctypes.data
    return descs, ptrs, n


def _objects(objects, textures):
    abi = _abi()
    n = len(objects or [])
    arr = (abi.Object * max(n, 1))()
    ptrs = (C.c_void_p * max(n, 1))()
    keep = []
    for i, o in enumerate(objects or []):
        arr[i] = o
        t = None if textures is None else textures[i]
        if t is not None:
            t = np.ascontiguousarray(t, dtype=np.uint8)
            keep.append(t)
            ptrs[i] = t.ctypes.data
    return arr, ptrs, n, keep


def render(params, tiles, objects=None, textures=None, stride_x=1, stride_y=1, rgb=True, meta=True, steps=True,
           max_points=0, threads=None):
    """FastGenerator::generate + draw_image on the CPU. Returns dict(rgb, meta, steps, counts, points,
    stats, timing)."""
    abi = _abi()
    L = lib()
    if threads:
        L.oracle_set_num_threads(int(threads))
    descs, ptrs, n = _tiles(tiles)
    oarr, optrs, no, _keep = _objects(objects, textures)
    cols = (params.x1 - params.x0 + stride_x - 1) // stride_x
    rows = (params.height + stride_y - 1) // stride_y
    a_rgb = np.empty((rows, cols, 3), np.uint8) if rgb else None
    a_meta = np.empty((rows, cols), META_DTYPE) if meta else None
    a_steps = np.empty((rows, cols), np.int32) if steps else None
    a_counts = np.empty((rows, cols), np.int32)
    a_points = np.zeros((rows, cols, max(max_points, 1)), TRACE_DTYPE) if max_points > 0 else None
    st = abi.Stats()
    tm = Timing()
    rc = L.oracle_render(C.byref(params), descs, n, ptrs, oarr, no, optrs, int(stride_x), int(stride_y), _p(a_rgb), _p(a_meta),
                         _p(a_steps), _p(a_counts), _p(a_points), int(max_points), C.byref(st), C.byref(tm))
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    timing = {k: getattr(tm, k) for k, _ in Timing._fields_ if not k.startswith("_")}
    return {"rgb": a_rgb, "meta": a_meta, "steps": a_steps, "counts": a_counts, "points": a_points, "stats": st.as_dict(),
            "timing": timing}


def get_elev(tiles, lat, lon):
    descs, ptrs, n = _tiles(tiles)
    lat = np.ascontiguousarray(lat, np.float64)
    lon = np.ascontiguousarray(lon, np.float64)
    out = np.empty_like(lat)
    lib().oracle_get_elev(descs, n, ptrs, _p(lat), _p(lon), int(lat.size), _p(out))
    return out


def coords_at_dist(earth_model, radius, lat0, lon0, direction, dist, ellipsoid_b=0.0):
    dist = np.ascontiguousarray(dist, np.float64)
    lat, lon = np.empty_like(dist), np.empty_like(dist)
    lib().oracle_coords_at_dist(int(earth_model), C.c_double(radius), C.c_double(lat0), C.c_double(lon0), C.c_double(direction),
                                _p(dist), int(dist.size), _p(lat), _p(lon), C.c_double(ellipsoid_b))
    return lat, lon


def world_directions(earth_model, radius, lat, lon):
    out = np.empty(9)
    lib().oracle_world_directions(int(earth_model), C.c_double(radius), C.c_double(lat), C.c_double(lon), _p(out))
    return out[0:3], out[3:6], out[6:9]


def as_cartesian(earth_model, radius, lat, lon, elev, ellipsoid_b=0.0):
    out = np.empty(3)
    lib().oracle_as_cartesian(int(earth_model), C.c_double(radius), C.c_double(lat), C.c_double(lon), C.c_double(elev), _p(out),
                              C.c_double(ellipsoid_b))
    return out


def ray_params(params, x, y):
    """RectilinearGenerator::get_ray_params: (elevation, direction) of a pixel in degrees."""
    out = np.empty(2)
    lib().oracle_ray_params(C.byref(params), int(x), int(y), _p(out))
    return out[0], out[1]


def light_dir(earth_model, radius, lat, lon, direction, zenith_deg, light_dir_deg):
    out = np.empty(3)
    lib().oracle_light_dir(int(earth_model), C.c_double(radius), C.c_double(lat), C.c_double(lon), C.c_double(direction),
                           C.c_double(zenith_deg), C.c_double(light_dir_deg), _p(out))
    return out


def atmosphere(adef, wavelength, h):
    h = np.ascontiguousarray(h, np.float64)
    t, p, n = np.empty_like(h), np.empty_like(h), np.empty_like(h)
    rc = lib().oracle_atmosphere(C.byref(adef), C.c_double(wavelength), _p(h), int(h.size), _p(t), _p(p), _p(n))
    if rc != 0:
        raise RuntimeError(f"oracle_atmosphere failed: {rc}")
    return t, p, n


def air_index(lam, p, t, rh=0.0):
    return lib().oracle_air_index(lam, p, t, rh)


def ray_path(adef, wavelength, flat, radius, straight, start_h, ang_deg, step, nsteps):
    x, h = np.empty(nsteps), np.empty(nsteps)
    rc = lib().oracle_ray_path(C.byref(adef), C.c_double(wavelength), int(flat), C.c_double(radius), int(straight),
                               C.c_double(start_h), C.c_double(ang_deg), C.c_double(step), int(nsteps), _p(x), _p(h))
    if rc != 0:
        raise RuntimeError(f"oracle_ray_path failed: {rc}")
    return x, h


def path_cache(params, tiles, y):
    descs, ptrs, n = _tiles(tiles)
    cnt = C.c_int()
    L = lib()
    L.oracle_path_cache(C.byref(params), descs, n, ptrs, int(y), 0, None, None, None, C.byref(cnt))
    m = cnt.value
    dist, elev, plen = np.empty(m), np.empty(m), np.empty(m)
    L.oracle_path_cache(C.byref(params), descs, n, ptrs, int(y), m, _p(dist), _p(elev), _p(plen), C.byref(cnt))
    return {"dist": dist, "elev": elev, "path_length": plen}


def terrain_cache(params, tiles, x, objects=None):
    descs, ptrs, n = _tiles(tiles)
    oarr, _optrs, no, _keep = _objects(objects, None)
    cnt = C.c_int()
    L = lib()
    L.oracle_terrain_cache(C.byref(params), descs, n, ptrs, oarr, no, int(x), 0, None, None, None, None, None, C.byref(cnt))
    m = cnt.value
    lat, lon, elev = np.empty(m), np.empty(m), np.empty(m)
    normal = np.empty((m, 3))
    close = np.zeros(m, np.uint64)
    L.oracle_terrain_cache(C.byref(params), descs, n, ptrs, oarr, no, int(x), m, _p(lat), _p(lon), _p(elev), _p(normal), _p(close),
                           C.byref(cnt))
    return {"lat": lat, "lon": lon, "elev": elev, "normal": normal, "close": close}


def ray_angles(params):
    d, e = np.empty(params.width), np.empty(params.height)
    lib().oracle_ray_angles(C.byref(params), _p(d), _p(e))
    return d, e


def pixel_angles(params):
    """ResultPixel.elevation_angle / azimuth, [H][x1-x0] each (fast.rs:67-76, rectilinear.rs:78-116)."""
    shape = (params.height, params.x1 - params.x0)
    el, az = np.empty(shape), np.empty(shape)
    lib().oracle_pixel_angles(C.byref(params), _p(el), _p(az))
    return el, az


def _read_text(fn, *args):
    import tempfile

    with tempfile.NamedTemporaryFile(suffix=".tsv") as tmp:
        rc = fn(*args, os.fsencode(tmp.name))
        if rc != 0:
            raise RuntimeError(f"oracle dumper failed: {rc}")
        return open(tmp.name, "rb").read().decode()


def output_ray_paths(params, height=2.0, min_ang=-1.0, max_ang=1.0, step=0.1, ray_step=50.0, cutoff=10000.0, output_step=50.0):
    """The stdout of `output-ray-paths` (ray_path.rs) for a Params."""
    L = lib()
    L.oracle_output_ray_paths.argtypes = [C.c_void_p] + [C.c_double] * 7 + [C.c_char_p]
    return _read_text(L.oracle_output_ray_paths, C.byref(params), height, min_ang, max_ang, step, ray_step, cutoff, output_step)


def output_elev_profile(params, tiles, azim=0.0, step=50.0, cutoff=10000.0):
    """The stdout of `output-elev-profile` (elev_profile.rs), Terrain::from_folder's line included."""
    L = lib()
    descs, ptrs, n = _tiles(tiles)
    L.oracle_output_elev_profile.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_char_p]
    return _read_text(L.oracle_output_elev_profile, C.byref(params), descs, n, ptrs, azim, step, cutoff)


def output_atm(adef, min_alt=0.0, max_alt=1000.0, step=0.2, celsius=False):
    """The stdout of `output-atm` (atm_printer.rs)."""
    L = lib()
    L.oracle_output_atm.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_char_p]
    return _read_text(L.oracle_output_atm, C.byref(adef), min_alt, max_alt, step, int(bool(celsius)))


def check_collision(obj, texture, obj_alt_abs, earth_model, radius, p1, p2):
    p1 = np.ascontiguousarray(p1, np.float64)
    p2 = np.ascontiguousarray(p2, np.float64)
    props, normals, colors = np.empty(8), np.empty((8, 3)), np.empty((8, 4))
    tex = None if texture is None else np.ascontiguousarray(texture, np.uint8)
    n = lib().oracle_check_collision(C.byref(obj), _p(tex), C.c_double(obj_alt_abs), int(earth_model), C.c_double(radius), _p(p1),
                                     _p(p2), _p(props), _p(normals), _p(colors))
    n = min(n, 8)
    return props[:n], normals[:n], colors[:n]


def draw_pixel(params, points):
    """points: structured array of TRACE_DTYPE."""
    points = np.ascontiguousarray(points, TRACE_DTYPE)
    out = np.zeros(3, np.uint8)
    lib().oracle_draw_pixel(C.byref(params), _p(points), int(points.size), _p(out))
    return out


def read_dted(path, header_only=False):
    abi = _abi()
    d = abi.TileDesc()
    L = lib()
    L.oracle_read_dted.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t]
    rc = L.oracle_read_dted(os.fsencode(path), C.byref(d), None, 0)
    if rc != 0:
        raise RuntimeError(f"oracle_read_dted failed: {rc}")
    if header_only:
        return d, None
    posts = np.empty((d.nlon, d.nlat), np.int16)
    rc = L.oracle_read_dted(os.fsencode(path), C.byref(d), _p(posts), posts.size)
    if rc != 0:
        raise RuntimeError(f"oracle_read_dted failed: {rc}")
    return d, posts


def num_threads():
    return lib().oracle_num_threads()
