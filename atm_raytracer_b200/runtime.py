"""ctypes binding of ``libatmrt_cuda.so`` and the host-side mirror of the reference's generator seam.

``Terrain`` mirrors terrain/mod.rs (``Terrain::from_folder``), ``FastGenerator`` mirrors
generator/generators/fast.rs (``FastGenerator::new(&params, &terrain, start).generate()``) and
``output_image`` mirrors renderer/mod.rs:416-437. All arithmetic of the hot path runs in the CUDA
library; there is no CPU fallback -- importing this module without the built library raises.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libatmrt_cuda.so")
HOST_LIB_PATH = os.path.join(_HERE, "libatmrt_host.so")


class AtmrtError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"atmrt error {code}: {message}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The hot path has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    vp = C.c_void_p
    sig = {
        "atmrt_abi_version": (C.c_int, []),
        "atmrt_abi_sizes": (C.c_int, [P(C.c_size_t), C.c_int]),
        "atmrt_create": (C.c_int, [C.c_int, P(vp)]),
        "atmrt_destroy": (None, [vp]),
        "atmrt_last_error": (C.c_char_p, [vp]),
        "atmrt_terrain_packed_bytes": (C.c_int, [P(abi.TileDesc), C.c_int, P(C.c_size_t)]),
        "atmrt_pack_terrain": (C.c_int, [vp, P(abi.TileDesc), C.c_int, P(vp), vp]),
        "atmrt_bind_terrain": (C.c_int, [vp, P(abi.TileDesc), C.c_int, vp]),
        "atmrt_set_terrain": (C.c_int, [vp, P(abi.TileDesc), C.c_int, P(vp)]),
        "atmrt_get_elev": (C.c_int, [vp, vp, vp, C.c_int, vp]),
        "atmrt_read_tile": (C.c_int, [vp, C.c_int, vp]),
        "atmrt_set_params": (C.c_int, [vp, P(abi.Params)]),
        "atmrt_set_objects": (C.c_int, [vp, P(abi.Object), C.c_int, P(vp)]),
        "atmrt_render": (C.c_int, [vp, vp, vp, vp, P(abi.Stats)]),
        "atmrt_render_device": (C.c_int, [vp, vp, vp, vp, P(abi.Stats), vp]),
        "atmrt_stage_times": (C.c_int, [vp, P(abi.StageMs)]),
        "atmrt_kernel_times": (C.c_int, [vp, P(abi.KernelMs)]),
        "atmrt_kernel_name": (C.c_char_p, [C.c_int]),
        "atmrt_render_trace": (C.c_int, [vp, vp, vp, C.c_int]),
        "atmrt_pixel_angles": (C.c_int, [vp, vp, vp]),
        "atmrt_set_march_mode": (C.c_int, [vp, C.c_int]),
        "atmrt_set_sweep_bands": (C.c_int, [vp, C.c_int]),
        "atmrt_get_terrain_profile": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, P(C.c_int)]),
        "atmrt_get_path": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, P(C.c_int)]),
        "atmrt_atmosphere_probe": (C.c_int, [vp, vp, C.c_int, vp, vp, vp]),
        "atmrt_ray_paths": (C.c_int, [vp, C.c_double, vp, C.c_int, C.c_double, C.c_int, vp, vp]),
        "atmrt_elev_profile": (C.c_int, [vp, C.c_double, vp, C.c_int, vp, vp, vp]),
        "atmrt_refraction_probe": (C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp, P(C.c_int)]),
        "atmrt_set_path_mode": (C.c_int, [vp, C.c_int]),
        "atmrt_refraction_table": (C.c_int, [vp, C.c_double, vp, C.c_int, P(C.c_int), P(C.c_int), P(C.c_double), P(C.c_double), P(C.c_int), P(C.c_int)]),
        "atmrt_observer_altitude": (C.c_int, [vp, P(C.c_double)]),
        "atmrt_fp64_peak": (C.c_int, [vp, P(C.c_double), P(C.c_double)]),
        "atmrt_group_create": (C.c_int, [P(C.c_int), C.c_int, P(vp)]),
        "atmrt_group_destroy": (None, [vp]),
        "atmrt_group_last_error": (C.c_char_p, [vp]),
        "atmrt_group_size": (C.c_int, [vp]),
        "atmrt_group_context": (vp, [vp, C.c_int]),
        "atmrt_group_column_block": (C.c_int, [vp, C.c_int, C.c_int, P(C.c_int), P(C.c_int)]),
        "atmrt_group_set_terrain": (C.c_int, [vp, P(abi.TileDesc), C.c_int, P(vp)]),
        "atmrt_group_set_params": (C.c_int, [vp, P(abi.Params)]),
        "atmrt_group_set_objects": (C.c_int, [vp, P(abi.Object), C.c_int, P(vp)]),
        "atmrt_group_render": (C.c_int, [vp, vp, vp, vp, P(abi.Stats)]),
        "atmrt_group_pixel_angles": (C.c_int, [vp, vp, vp]),
        "atmrt_group_render_trace": (C.c_int, [vp, vp, vp, C.c_int]),
        "atmrt_group_render_tiles": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, vp, vp]),
        "atmrt_host_alloc": (vp, [C.c_size_t]),
        "atmrt_host_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()

META_DTYPE = np.dtype([("lat", "<f8"), ("lon", "<f8"), ("elevation", "<f8"), ("distance", "<f8")])
TRACE_DTYPE = np.dtype(
    [("lat", "<f8"), ("lon", "<f8"), ("distance", "<f8"), ("elevation", "<f8"), ("path_length", "<f8"),
     ("normal", "<f8", 3), ("color", "<f8", 4), ("is_terrain", "<i4"), ("step", "<i4")]
)
assert META_DTYPE.itemsize == C.sizeof(abi.Meta) and TRACE_DTYPE.itemsize == C.sizeof(abi.TracePoint)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def refraction_table(atmosphere, wavelength):
    """The ray-path stage's g(h) table for an abi.AtmosphereDef (host only, no GPU):
    (cells[ncoef, ncells], base, cell_height, cells_served, pieces)."""
    nc, nq, served, npieces = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    base, ch = C.c_double(), C.c_double()
    rc = lib.atmrt_refraction_table(C.byref(atmosphere), wavelength, None, 0, C.byref(nc), C.byref(nq), C.byref(base), C.byref(ch), C.byref(served), C.byref(npieces))
    if rc:
        raise AtmrtError(rc, "refraction_table: invalid atmosphere")
    cells = np.empty((nq.value, nc.value), dtype=np.float64)
    rc = lib.atmrt_refraction_table(C.byref(atmosphere), wavelength, _ptr(cells), cells.size, None, None, None, None, C.byref(served), None)
    if rc:
        raise AtmrtError(rc, "refraction_table: invalid atmosphere")
    return cells, base.value, ch.value, served.value, npieces.value


from .terrain import Terrain  # noqa: E402,F401  (re-exported: runtime.Terrain)


class Context:
    """One ``atmrt_ctx`` (one GPU)."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.atmrt_create(int(device), C.byref(h))
        if rc != 0:
            raise AtmrtError(rc, (lib.atmrt_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self._keep = []
        self.params = None

    def close(self):
        if getattr(self, "_h", None):
            lib.atmrt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise AtmrtError(rc, (lib.atmrt_last_error(self._h) or b"").decode())

    # ---- terrain ----
    def set_terrain(self, terrain):
        descs, ptrs, n = terrain.c_arrays()
        self._check(lib.atmrt_set_terrain(self._h, descs, n, ptrs))
        self._terrain = terrain

    def packed_bytes(self, terrain):
        descs, _, n = terrain.c_arrays()
        out = C.c_size_t()
        self._check(lib.atmrt_terrain_packed_bytes(descs, n, C.byref(out)))
        return out.value

    def pack_terrain(self, terrain, dev_ptr):
        descs, ptrs, n = terrain.c_arrays()
        self._check(lib.atmrt_pack_terrain(self._h, descs, n, ptrs, C.c_void_p(dev_ptr)))

    def bind_terrain(self, terrain, dev_ptr):
        descs, _, n = terrain.c_arrays()
        self._check(lib.atmrt_bind_terrain(self._h, descs, n, C.c_void_p(dev_ptr)))
        self._terrain = terrain

    def get_elev(self, lat, lon):
        lat = np.ascontiguousarray(lat, dtype=np.float64)
        lon = np.ascontiguousarray(lon, dtype=np.float64)
        out = np.empty_like(lat)
        self._check(lib.atmrt_get_elev(self._h, _ptr(lat), _ptr(lon), lat.size, _ptr(out)))
        return out

    def read_tile(self, index):
        d, posts = self._terrain.tiles[index]
        out = np.empty((d.nlon, d.nlat), dtype=np.int16)
        self._check(lib.atmrt_read_tile(self._h, index, _ptr(out)))
        return out

    # ---- scene ----
    def set_params(self, params):
        self._check(lib.atmrt_set_params(self._h, C.byref(params)))
        self.params = params

    def set_objects(self, objects, textures=None):
        n = len(objects)
        arr = (abi.Object * max(n, 1))()
        ptrs = (C.c_void_p * max(n, 1))()
        keep = []
        for i, o in enumerate(objects):
            arr[i] = o
            t = None if textures is None else textures[i]
            if t is not None:
                t = np.ascontiguousarray(t, dtype=np.uint8)
                keep.append(t)
                ptrs[i] = t.ctypes.data
        self._check(lib.atmrt_set_objects(self._h, arr, n, ptrs))

    def set_sweep_bands(self, bands):
        self._check(lib.atmrt_set_sweep_bands(self._h, int(bands)))

    def set_march_mode(self, mode):
        self._check(lib.atmrt_set_march_mode(self._h, int(mode)))

    # ---- render ----
    def shape(self):
        p = self.params
        return p.height, p.x1 - p.x0

    def render(self, rgb=True, meta=True, steps=True, out=None):
        """Host-buffer render (copies inside the call). Returns dict(rgb, meta, steps, stats).
        ``out`` may carry preallocated (e.g. pinned) arrays under the same keys."""
        if self.params is None:  # let the library report the call-order violation (ATMRT_ERR_STATE)
            self._check(lib.atmrt_render(self._h, None, None, None, None))
        h, w = self.shape()
        out = out or {}
        a_rgb = out.get("rgb") if rgb else None
        a_meta = out.get("meta") if meta else None
        a_steps = out.get("steps") if steps else None
        if rgb and a_rgb is None:
            a_rgb = np.empty((h, w, 3), dtype=np.uint8)
        if meta and a_meta is None:
            a_meta = np.empty((h, w), dtype=META_DTYPE)
        if steps and a_steps is None:
            a_steps = np.empty((h, w), dtype=np.int32)
        st = abi.Stats()
        self._check(lib.atmrt_render(self._h, _ptr(a_rgb), _ptr(a_meta), _ptr(a_steps), C.byref(st)))
        return {"rgb": a_rgb, "meta": a_meta, "steps": a_steps, "stats": st.as_dict()}

    def render_device(self, rgb_ptr=0, meta_ptr=0, steps_ptr=0, stream=0, want_stats=False):
        st = abi.Stats() if want_stats else None
        self._check(
            lib.atmrt_render_device(self._h, C.c_void_p(rgb_ptr or None), C.c_void_p(meta_ptr or None), C.c_void_p(steps_ptr or None),
                                    C.byref(st) if st is not None else None, C.c_void_p(stream or None))
        )
        return st.as_dict() if st is not None else None

    def stage_times(self):
        """Average device time per stage over the renders since the last call (synchronises)."""
        st = abi.StageMs()
        self._check(lib.atmrt_stage_times(self._h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in abi.StageMs._fields_ if not k.startswith("_")}

    def kernel_times(self):
        """Average device time of each hot kernel over the renders since the last stage_times() call:
        {kernel name: ms} for the kernels those renders launched (synchronises; does not reset the averages)."""
        k = abi.KernelMs()
        self._check(lib.atmrt_kernel_times(self._h, C.byref(k)))
        return {lib.atmrt_kernel_name(i).decode(): k.ms[i] for i in range(abi.KERNEL_COUNT) if k.renders_with[i] > 0}

    def render_trace(self, max_points=8):
        h, w = self.shape()
        pts = np.zeros((h, w, max(max_points, 1)), dtype=TRACE_DTYPE)
        cnt = np.zeros((h, w), dtype=np.int32)
        self._check(lib.atmrt_render_trace(self._h, _ptr(pts), _ptr(cnt), int(max_points)))
        return pts, cnt

    def pixel_angles(self):
        """ResultPixel.elevation_angle / azimuth of every pixel of the column block, [H][x1-x0] each (degrees)."""
        h, w = self.shape()
        el, az = np.empty((h, w)), np.empty((h, w))
        self._check(lib.atmrt_pixel_angles(self._h, _ptr(el), _ptr(az)))
        return el, az

    # ---- probes ----
    def terrain_profile(self, x):
        n = C.c_int()
        self._check(lib.atmrt_get_terrain_profile(self._h, x, 0, None, None, None, None, None, C.byref(n)))
        m = n.value
        lat, lon, elev = np.empty(m), np.empty(m), np.empty(m)
        normal = np.empty((m, 3))
        close = np.empty(m, dtype=np.uint64)
        self._check(lib.atmrt_get_terrain_profile(self._h, x, m, _ptr(lat), _ptr(lon), _ptr(elev), _ptr(normal), _ptr(close), C.byref(n)))
        return {"lat": lat, "lon": lon, "elev": elev, "normal": normal, "close": close}

    def path(self, y):
        n = C.c_int()
        self._check(lib.atmrt_get_path(self._h, y, 0, None, None, None, C.byref(n)))
        m = n.value
        dist, elev, plen = np.empty(m), np.empty(m), np.empty(m)
        self._check(lib.atmrt_get_path(self._h, y, m, _ptr(dist), _ptr(elev), _ptr(plen), C.byref(n)))
        return {"dist": dist, "elev": elev, "path_length": plen}

    def ray_paths(self, start_h, angles_deg, ray_step, nsteps):
        """The stepper behind `output-ray-paths`: (x[nsteps], h[n_angles][nsteps])."""
        a = np.ascontiguousarray(angles_deg, dtype=np.float64)
        x, h = np.empty(nsteps), np.empty((a.size, nsteps))
        self._check(lib.atmrt_ray_paths(self._h, float(start_h), _ptr(a), a.size, float(ray_step), int(nsteps), _ptr(x), _ptr(h)))
        return x, h

    def elev_profile(self, azimuth, dist):
        """The sampler behind `output-elev-profile`: (lat, lon, elev) at the given distances along `azimuth`."""
        d = np.ascontiguousarray(dist, dtype=np.float64)
        lat, lon, elev = np.empty_like(d), np.empty_like(d), np.empty_like(d)
        self._check(lib.atmrt_elev_profile(self._h, float(azimuth), _ptr(d), d.size, _ptr(lat), _ptr(lon), _ptr(elev)))
        return lat, lon, elev

    def atmosphere_probe(self, h):
        h = np.ascontiguousarray(h, dtype=np.float64)
        t, p, n = np.empty_like(h), np.empty_like(h), np.empty_like(h)
        self._check(lib.atmrt_atmosphere_probe(self._h, _ptr(h), h.size, _ptr(t), _ptr(p), _ptr(n)))
        return t, p, n

    def refraction_probe(self, h, with_pieces=True):
        """g(h) = dn/n of the ray-path stage: (without libm -- table and pieces --, NaN where unserved; through libm; cells served)."""
        h = np.ascontiguousarray(h, dtype=np.float64)
        gt, gl = np.empty_like(h), np.empty_like(h)
        served = C.c_int()
        self._check(lib.atmrt_refraction_probe(self._h, _ptr(h), h.size, int(bool(with_pieces)), _ptr(gt), _ptr(gl), C.byref(served)))
        return gt, gl, served.value

    def set_path_mode(self, mode):
        self._check(lib.atmrt_set_path_mode(self._h, int(mode)))

    def observer_altitude(self):
        v = C.c_double()
        self._check(lib.atmrt_observer_altitude(self._h, C.byref(v)))
        return v.value

    def fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(lib.atmrt_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return {"dfma_gflops": a.value, "dadd_ginstr": b.value}


def host_array(shape, dtype):
    """A numpy array in page-locked host memory visible to every GPU (atmrt_host_alloc); freed with the array."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.atmrt_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("atmrt_host_alloc failed")

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            lib.atmrt_host_free(self.ptr)

    buf = (C.c_char * max(n, 1)).from_address(p)
    buf._owner = _Owner(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Group:
    """One panorama over several GPUs of this box in ONE process (``atmrt_group``): column blocks, one context and one
    host thread per GPU; terrain uploaded in slices and all-gathered over the peer links; every GPU writes its block
    straight into the host's row-major image over its own PCIe link."""

    def __init__(self, n, devices=None):
        h = C.c_void_p()
        dev = None if devices is None else (C.c_int * n)(*devices)
        rc = lib.atmrt_group_create(dev, int(n), C.byref(h))
        if rc != 0:
            raise AtmrtError(rc, (lib.atmrt_group_last_error(None) or b"").decode())
        self._h = h
        self.n = n
        self.params = None

    def close(self):
        if getattr(self, "_h", None):
            lib.atmrt_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise AtmrtError(rc, (lib.atmrt_group_last_error(self._h) or b"").decode())

    def column_block(self, width, i):
        a, b = C.c_int(), C.c_int()
        self._check(lib.atmrt_group_column_block(self._h, int(width), int(i), C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_terrain(self, terrain):
        descs, ptrs, n = terrain.c_arrays()
        self._check(lib.atmrt_group_set_terrain(self._h, descs, n, ptrs))

    def set_params(self, params):
        self._check(lib.atmrt_group_set_params(self._h, C.byref(params)))
        self.params = params

    def set_objects(self, objects, textures=None):
        n = len(objects)
        arr = (abi.Object * max(n, 1))()
        ptrs = (C.c_void_p * max(n, 1))()
        keep = []
        for i, o in enumerate(objects):
            arr[i] = o
            t = None if textures is None else textures[i]
            if t is not None:
                t = np.ascontiguousarray(t, dtype=np.uint8)
                keep.append(t)
                ptrs[i] = t.ctypes.data
        self._check(lib.atmrt_group_set_objects(self._h, arr, n, ptrs))

    def pixel_angles(self):
        """ResultPixel.elevation_angle / azimuth of the whole image, [H][W] each."""
        p = self.params
        el, az = np.empty((p.height, p.width)), np.empty((p.height, p.width))
        self._check(lib.atmrt_group_pixel_angles(self._h, _ptr(el), _ptr(az)))
        return el, az

    def render_trace(self, max_points=8):
        """ResultPixel.trace_points of the whole image: (points[H][W][max_points], counts[H][W])."""
        p = self.params
        pts = np.zeros((p.height, p.width, max(max_points, 1)), dtype=TRACE_DTYPE)
        cnt = np.zeros((p.height, p.width), dtype=np.int32)
        self._check(lib.atmrt_group_render_trace(self._h, _ptr(pts), _ptr(cnt), int(max_points)))
        return pts, cnt

    def render(self, rgb=True, meta=True, steps=True, out=None, terrain=None):
        """The full image (all column blocks) in host memory; ``out`` may carry preallocated arrays (host_array). With
        ``terrain``: upload it and render in one call (atmrt_group_render_tiles: the ray paths are integrated while the tiles
        are on their way)."""
        p = self.params
        h, w = p.height, p.width
        out = out or {}
        a_rgb = out.get("rgb") if rgb else None
        a_meta = out.get("meta") if meta else None
        a_steps = out.get("steps") if steps else None
        if rgb and a_rgb is None:
            a_rgb = host_array((h, w, 3), np.uint8)
        if meta and a_meta is None:
            a_meta = host_array((h, w), META_DTYPE)
        if steps and a_steps is None:
            a_steps = host_array((h, w), np.int32)
        st = abi.Stats()
        if terrain is not None:
            descs, ptrs, n = terrain.c_arrays()
            self._check(lib.atmrt_group_render_tiles(self._h, descs, n, ptrs, _ptr(a_rgb), _ptr(a_meta), _ptr(a_steps), C.byref(st)))
        else:
            self._check(lib.atmrt_group_render(self._h, _ptr(a_rgb), _ptr(a_meta), _ptr(a_steps), C.byref(st)))
        return {"rgb": a_rgb, "meta": a_meta, "steps": a_steps, "stats": st.as_dict()}


class FastGenerator:
    """``FastGenerator`` (generator/generators/fast.rs:16-109): ``generate()`` returns the rendered
    column block (rgb, per-pixel metadata, ray-step counts, stats) instead of ``Vec<Vec<ResultPixel>>``
    because colouring and compositing are fused into the march kernel."""

    def __init__(self, params, terrain, objects=(), textures=None, device=0, context=None):
        self.ctx = context or Context(device)
        self.ctx.set_terrain(terrain)
        self.ctx.set_params(params)
        self.ctx.set_objects(list(objects), textures)

    def generate(self, **kw):
        return self.ctx.render(**kw)


def overlay_ticks(entries, single_key):
    """``Vec<Tick>`` / ``Vec<VerticalTick>`` of the YAML (params.rs:325-385) as the host's tick dictionaries."""
    from .config import ConfigError, _tagged

    out = []
    for e in entries or []:
        kind, body = _tagged(e, "tick")
        body = body or {}
        try:
            if kind == "Single":
                t = dict(kind="Single", angle=float(body[single_key]))
            elif kind == "Multiple":
                t = dict(kind="Multiple", bias=float(body["bias"]), step=float(body["step"]))
            else:
                raise ConfigError(f"unknown tick variant {kind} (Single, Multiple)")
            t["size"], t["labelled"] = int(body["size"]), bool(body["labelled"])
        except KeyError as k:  # serde has no defaults for these fields
            raise ConfigError(f"tick {kind} is missing field {k}") from None
        out.append(t)
    return out


def output_image(rgb, path, cfg=None, context=None):
    """``renderer::output_image`` (renderer/mod.rs:416-437): the overlays the configuration asks for -- ticks, the flat-earth
    horizon, the eye-level line -- drawn over ``rgb`` (in place) by the host library, then the PNG. ``cfg``: the dictionary of
    ``config.read_config``; ``context``: the Context (or Group context 0) that rendered ``rgb`` -- the overlays read its
    ``ResultPixel`` angles, the flat horizon its observer altitude and atmosphere. Without ``cfg`` only the PNG is written."""
    from . import host

    out = (cfg or {}).get("output", {})
    ticks, vticks = overlay_ticks(out.get("ticks"), "azimuth"), overlay_ticks(out.get("vertical_ticks"), "elevation")
    eye, flat = bool(out.get("show_eye_level")), bool(out.get("show_flat_horizon"))
    if ticks or vticks or eye or flat:
        if context is None:
            raise ValueError("output_image: the overlays need the context that rendered the image")
        from .config import _tagged

        el, az = context.pixel_angles()
        frame = cfg["view"]["frame"]
        shape, _ = _tagged(cfg.get("earth_shape", "SimpleSphere"), "earth_shape")
        flat_elev = None
        # show_flat_horizon && shape == Flat && !straight_rays (renderer/mod.rs:420-427; to_shape: earth_model/mod.rs:95-112)
        if flat and shape in ("FlatDistorted", "AzimuthalEquidistant", "ObserverAe", "SimpleObserverAe") and not cfg.get("straight_rays"):
            _, _, n = context.atmosphere_probe(np.array([context.observer_altitude()]))
            flat_elev = host.flat_horizon_elevation(float(n[0]))
        host.draw_overlays(rgb, el, az, ticks, vticks, dict(direction=float(frame["direction"]), fov=float(frame["fov"]), tilt=float(frame["tilt"])),
                           show_eye_level=eye, flat_horizon_elev=flat_elev)
    host.write_png(path, rgb)
