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
    return lib


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


EXECUTABLE = os.path.join(_HERE, "atm-raytracer")
