"""ctypes binding of ``libatmrt_host.so`` -- the C++ host helpers (DTED decode, PNG, `gen`)."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libatmrt_host.so")


class HostError(RuntimeError):
    pass


def _load():
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError(f"{HOST_LIB_PATH} is missing: run __graft_entry__.build()")
    lib = C.CDLL(HOST_LIB_PATH)
    lib.atmrt_host_read_dted.restype = C.c_int
    lib.atmrt_host_read_dted.argtypes = [C.c_char_p, C.POINTER(abi.TileDesc), C.c_void_p, C.c_size_t]
    for fn in (lib.atmrt_host_read_geotiff, lib.atmrt_host_read_tile):
        fn.restype = C.c_int
        fn.argtypes = [C.c_char_p, C.POINTER(abi.TileDesc), C.c_void_p, C.c_size_t]
    lib.atmrt_host_geotiff_coords_from_name.restype = C.c_int
    lib.atmrt_host_geotiff_coords_from_name.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.atmrt_host_gzip_write.restype = C.c_int
    lib.atmrt_host_gzip_write.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
    lib.atmrt_host_write_png.restype = C.c_int
    lib.atmrt_host_write_png.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.atmrt_host_read_png.restype = C.c_int
    lib.atmrt_host_read_png.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.atmrt_host_last_error.restype = C.c_char_p
    lib.atmrt_host_gen.restype = C.c_int
    lib.atmrt_host_gen.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    lib.atmrt_host_parse_config.restype = C.c_int
    lib.atmrt_host_parse_config.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(abi.Params), C.POINTER(abi.Object), C.c_int,
                                            C.POINTER(C.c_int), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.atmrt_host_draw_overlays.restype = C.c_int
    lib.atmrt_host_draw_overlays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Overlays)]
    lib.atmrt_host_gen_ticks.restype = C.c_int
    lib.atmrt_host_gen_ticks.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Overlays), C.c_int, C.POINTER(DrawTick), C.c_int,
                                         C.POINTER(C.c_int)]
    lib.atmrt_host_num_decimals.restype = C.c_int
    lib.atmrt_host_num_decimals.argtypes = [C.c_double]
    lib.atmrt_host_flat_horizon_elevation.restype = C.c_int
    lib.atmrt_host_flat_horizon_elevation.argtypes = [C.c_double, C.POINTER(C.c_double)]
    lib.atmrt_host_parse_overlays.restype = C.c_int
    lib.atmrt_host_parse_overlays.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(Tick), C.c_int, C.POINTER(C.c_int), C.POINTER(Tick), C.c_int,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return lib


class Tick(C.Structure):
    """atmrt_host_tick: Tick / VerticalTick of generator/params.rs:325-385."""

    _fields_ = [("multiple", C.c_int32), ("labelled", C.c_int32), ("size", C.c_uint32), ("reserved", C.c_uint32), ("angle", C.c_double),
                ("bias", C.c_double), ("step", C.c_double)]


class Overlays(C.Structure):
    """atmrt_host_overlays: what renderer::output_image reads of Params to draw over the picture."""

    _fields_ = [("ticks", C.POINTER(Tick)), ("vertical_ticks", C.POINTER(Tick)), ("nticks", C.c_int32), ("nvertical_ticks", C.c_int32),
                ("direction", C.c_double), ("fov", C.c_double), ("tilt", C.c_double), ("show_eye_level", C.c_int32),
                ("show_flat_horizon", C.c_int32), ("flat_horizon_elevation", C.c_double)]


class DrawTick(C.Structure):
    _fields_ = [("position", C.c_uint32), ("size", C.c_uint32), ("labelled", C.c_int32), ("label", C.c_char * 28)]


lib = _load()


def _check(rc):
    if rc != 0:
        raise HostError((lib.atmrt_host_last_error() or b"").decode())


def read_dted_header(path):
    d = abi.TileDesc()
    _check(lib.atmrt_host_read_dted(os.fsencode(path), C.byref(d), None, 0))
    return d


def read_dted(path):
    """(TileDesc, int16 posts [nlon][nlat]) decoded bit-exactly by the C++ host."""
    d = read_dted_header(path)
    posts = np.empty((d.nlon, d.nlat), dtype=np.int16)
    _check(lib.atmrt_host_read_dted(os.fsencode(path), C.byref(d), posts.ctypes.data_as(C.c_void_p), posts.size))
    return d, posts


def _read_posts(fn, path):
    d = abi.TileDesc()
    _check(fn(os.fsencode(path), C.byref(d), None, 0))
    posts = np.empty((d.nlon, d.nlat), dtype=np.int16)
    _check(fn(os.fsencode(path), C.byref(d), posts.ctypes.data_as(C.c_void_p), posts.size))
    return d, posts


def read_geotiff(path):
    """(TileDesc, int16 posts [3601][3601]) of a GeoTIFF tile (terrain/geotiff.rs): key and corner from the file name, posts
    [lon line][lat point] west->east, south->north, one arc-second apart."""
    return _read_posts(lib.atmrt_host_read_geotiff, path)


def read_tile(path):
    """TerrainDataInner::read_tile (terrain/mod.rs:23-31): a DTED tile if the header reads as one, else a GeoTIFF tile."""
    return _read_posts(lib.atmrt_host_read_tile, path)


def geotiff_coords_from_name(path):
    """GeoTiffWrapper::coords_from_name (terrain/geotiff.rs:16-31): (lat, lon) or None."""
    lat, lon = C.c_int(), C.c_int()
    if lib.atmrt_host_geotiff_coords_from_name(os.fsencode(path), C.byref(lat), C.byref(lon)) != 0:
        return None
    return lat.value, lon.value


def gzip_write(path, data, block_bytes=0, threads=0):
    """The metadata sidecar's writer on its own: ``data`` (bytes) as a gzip file of members compressed in parallel."""
    buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
    _check(lib.atmrt_host_gzip_write(os.fsencode(path), buf, len(data), int(block_bytes), int(threads)))


def write_png(path, pixels):
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w, ch = pixels.shape
    _check(lib.atmrt_host_write_png(os.fsencode(path), pixels.ctypes.data_as(C.c_void_p), w, h, ch))


def read_png(path):
    w, h = C.c_int(), C.c_int()
    _check(lib.atmrt_host_read_png(os.fsencode(path), None, 0, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 4), dtype=np.uint8)
    _check(lib.atmrt_host_read_png(os.fsencode(path), out.ctypes.data_as(C.c_void_p), out.size, C.byref(w), C.byref(h)))
    return out


def gen(argv):
    """Run the C++ `gen` subcommand in-process; returns its exit code."""
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    return lib.atmrt_host_gen(len(argv), arr)


def parse_config(argv, max_objects=64):
    """`read_config` + `Config::into_params` of the C++ host (YAML subset parser + CLI overrides) without
    touching the GPU: returns (Params, [Object], terrain_folder, output_file, metadata_file)."""
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    p = abi.Params()
    objs = (abi.Object * max_objects)()
    n = C.c_int()
    folder, out, meta = (C.create_string_buffer(4096) for _ in range(3))
    _check(lib.atmrt_host_parse_config(len(argv), arr, C.byref(p), objs, max_objects, C.byref(n), folder, 4096, out, 4096, meta, 4096))
    return p, list(objs[: n.value]), folder.value.decode(), out.value.decode(), meta.value.decode()


def _tick_array(ticks):
    """[{kind: Single|Multiple, angle | bias + step, size, labelled}] -> (Tick * n)"""
    arr = (Tick * max(len(ticks), 1))()
    for i, t in enumerate(ticks):
        arr[i].multiple = 1 if t["kind"] == "Multiple" else 0
        arr[i].labelled, arr[i].size = int(t["labelled"]), int(t["size"])
        arr[i].angle, arr[i].bias, arr[i].step = float(t.get("angle", 0.0)), float(t.get("bias", 0.0)), float(t.get("step", 0.0))
    return arr


def _overlays(ticks, vertical_ticks, frame, show_eye_level, flat_horizon_elev):
    ov = Overlays()
    keep = (_tick_array(ticks), _tick_array(vertical_ticks))
    ov.ticks, ov.vertical_ticks, ov.nticks, ov.nvertical_ticks = keep[0], keep[1], len(ticks), len(vertical_ticks)
    ov.direction, ov.fov, ov.tilt = frame["direction"], frame["fov"], frame["tilt"]
    ov.show_eye_level = int(show_eye_level)
    ov.show_flat_horizon = int(flat_horizon_elev is not None)
    ov.flat_horizon_elevation = 0.0 if flat_horizon_elev is None else flat_horizon_elev
    return ov, keep


def draw_overlays(img, el, az, ticks=(), vertical_ticks=(), frame=None, show_eye_level=False, flat_horizon_elev=None):
    """renderer::output_image's overlays (renderer/mod.rs:416-431) drawn into img[H][W][3] (uint8, C-contiguous, in place);
    el / az: ResultPixel.elevation_angle / .azimuth, [H][W] each. flat_horizon_elev: degrees, or None (not drawn)."""
    assert img.dtype == np.uint8 and img.flags.c_contiguous and img.ndim == 3 and img.shape[2] == 3
    h, w = img.shape[:2]
    el, az = (np.ascontiguousarray(a, dtype=np.float64) for a in (el, az))
    assert el.shape == (h, w) and az.shape == (h, w)
    ov, keep = _overlays(list(ticks), list(vertical_ticks), frame or dict(direction=0.0, fov=30.0, tilt=0.0), show_eye_level, flat_horizon_elev)
    _check(lib.atmrt_host_draw_overlays(img.ctypes.data_as(C.c_void_p), w, h, el.ctypes.data_as(C.c_void_p), az.ctypes.data_as(C.c_void_p), C.byref(ov)))
    del keep
    return img


def gen_ticks(el, az, ticks=(), vertical_ticks=(), frame=None, vertical=False):
    """gen_ticks (renderer/mod.rs:225-266): [(pixel position, size, labelled, label text)] by ascending position."""
    el, az = (np.ascontiguousarray(a, dtype=np.float64) for a in (el, az))
    h, w = el.shape
    ov, keep = _overlays(list(ticks), list(vertical_ticks), frame or dict(direction=0.0, fov=30.0, tilt=0.0), False, None)
    n = C.c_int()
    cap = 4096
    out = (DrawTick * cap)()
    _check(lib.atmrt_host_gen_ticks(w, h, el.ctypes.data_as(C.c_void_p), az.ctypes.data_as(C.c_void_p), C.byref(ov), int(vertical), out, cap, C.byref(n)))
    del keep
    return [(t.position, t.size, bool(t.labelled), t.label.decode()) for t in out[: min(n.value, cap)]]


def num_decimals(x):
    return lib.atmrt_host_num_decimals(float(x))


def flat_horizon_elevation(n_at_observer):
    out = C.c_double()
    _check(lib.atmrt_host_flat_horizon_elevation(float(n_at_observer), C.byref(out)))
    return out.value


def parse_overlays(argv, cap=64):
    """output.ticks / vertical_ticks / show_eye_level / show_flat_horizon as the C++ host's read_config lowers them."""
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    t, v = (Tick * cap)(), (Tick * cap)()
    nt, nv, eye, flat = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _check(lib.atmrt_host_parse_overlays(len(argv), arr, t, cap, C.byref(nt), v, cap, C.byref(nv), C.byref(eye), C.byref(flat)))

    def rows(a, n):
        return [dict(kind="Multiple" if x.multiple else "Single", angle=x.angle, bias=x.bias, step=x.step, size=x.size, labelled=bool(x.labelled))
                for x in a[:n]]

    return rows(t, nt.value), rows(v, nv.value), bool(eye.value), bool(flat.value)


EXECUTABLE = os.path.join(_HERE, "atm-raytracer")
