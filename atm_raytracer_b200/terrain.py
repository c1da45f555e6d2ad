"""Decoded DTED tiles on the host: the reference's ``Terrain`` store (terrain/mod.rs:55-126) as plain
descriptors + int16 arrays. No native library is loaded by importing this module (the CPU reference arm of
bench.py builds its workload from it without ever dlopening the CUDA library)."""
import ctypes as C
import os

import numpy as np

from . import abi


class Terrain:
    """Decoded DTED tiles on the host (``Terrain``, terrain/mod.rs:55-126).

    ``tiles`` is a list of ``(abi.TileDesc, int16 ndarray [nlon][nlat])``. Decoding is done by the
    C++ host library (``libatmrt_host.so``: atmrt_host_read_dted) or by the caller.
    """

    def __init__(self, tiles=None):
        self.tiles = list(tiles or [])

    @staticmethod
    def desc(lat0, lon0, posts, lat_interval=None, lon_interval=None):
        nlon, nlat = posts.shape
        d = abi.TileDesc()
        d.lat0, d.lon0 = int(lat0), int(lon0)  # `as i16` truncation of the header origin
        d.nlon, d.nlat = int(nlon), int(nlat)
        d.min_lat, d.min_lon = float(lat0), float(lon0)
        d.lat_interval = 3600.0 / (nlat - 1) if lat_interval is None else float(lat_interval)
        d.lon_interval = 3600.0 / (nlon - 1) if lon_interval is None else float(lon_interval)
        return d

    @classmethod
    def from_arrays(cls, items):
        """items: iterable of (lat0, lon0, posts[nlon][nlat])."""
        out = []
        for lat0, lon0, posts in items:
            posts = np.ascontiguousarray(posts, dtype=np.int16)
            out.append((cls.desc(lat0, lon0, posts), posts))
        return cls(out)

    @classmethod
    def from_folder(cls, folder):
        """``Terrain::from_folder`` (terrain/mod.rs:66-83): every entry must be a DTED or a GeoTIFF tile."""
        from . import host

        tiles = []
        names = sorted(os.listdir(folder))
        for name in names:
            tiles.append(host.read_tile(os.path.join(folder, name)))
        print(f"Detected {len(names)} terrain files")
        return cls(tiles)

    def c_arrays(self):
        n = len(self.tiles)
        descs = (abi.TileDesc * max(n, 1))()
        ptrs = (C.c_void_p * max(n, 1))()
        for i, (d, posts) in enumerate(self.tiles):
            descs[i] = d
            ptrs[i] = None if posts is None else posts.ctypes.data  # None: descriptor-only terrain (non-root ranks)
        return descs, ptrs, n

    def descriptors_only(self):
        """The same tile table without the posts (what a non-root rank needs to bind a broadcast copy)."""
        return Terrain([(d, None) for d, _ in self.tiles])

    @property
    def bytes(self):
        return sum(int(d.nlon) * int(d.nlat) * 2 for d, _ in self.tiles)
