"""Reader of the metadata sidecar `atm-raytracer gen --output-meta` writes (csrc/host/gen.cpp:write_metadata): what a consumer
of the reference's `.dat` gets from `AllData.result` (generator/mod.rs:20-24, generators/mod.rs:14-30) -- per pixel the
``ResultPixel`` angles and its trace points -- from this host's own documented layout (NOT the reference's bincode container).
No native library is needed to read it."""
import gzip

import numpy as np

META_DTYPE = np.dtype([("lat", "<f8"), ("lon", "<f8"), ("elevation", "<f8"), ("distance", "<f8")])
TRACE_DTYPE = np.dtype(
    [("lat", "<f8"), ("lon", "<f8"), ("distance", "<f8"), ("elevation", "<f8"), ("path_length", "<f8"),
     ("normal", "<f8", 3), ("color", "<f8", 4), ("is_terrain", "<i4"), ("step", "<i4")]
)
GENERATORS = ("Fast", "Rectilinear", "InterpolatingRectilinear")


class Sidecar:
    """width, height, generator; elevation_angle / azimuth [H][W] (degrees; ResultPixel.elevation_angle / .azimuth);
    first [H][W] of META_DTYPE (the first trace point of every pixel, NaN where the ray met nothing); version 3 only:
    counts [H][W] (true number of trace points), max_points, and the kept trace points of all pixels, row-major, in one array."""

    def trace_points(self, y, x):
        """ResultPixel.trace_points of pixel (y, x), front to back (the first max_points of them)."""
        if self.counts is None:  # version 2: opaque scenes hold at most the first point
            p = self.first[y, x]
            if np.isnan(p["distance"]):
                return np.empty(0, TRACE_DTYPE)
            out = np.zeros(1, TRACE_DTYPE)
            for k in ("lat", "lon", "elevation", "distance"):
                out[k] = p[k]
            for k in ("path_length", "normal", "color"):
                out[k] = np.nan  # not kept by version 2
            out["is_terrain"], out["step"] = 1, -1
            return out
        i = y * self.width + x
        return self.points[self._first_index[i]:self._first_index[i] + self._kept[i]]


def read_sidecar(path):
    raw = gzip.decompress(open(path, "rb").read())
    magic = raw[:11]
    if magic not in (b"ATMRTMETA2\n", b"ATMRTMETA3\n"):
        raise ValueError(f"{path}: not a metadata sidecar of this host (magic {magic!r})")
    s = Sidecar()
    s.version = int(magic[9:10])
    w, h, gen, _ = (int(v) for v in np.frombuffer(raw, "<i4", 4, 11))
    s.width, s.height, s.generator = w, h, GENERATORS[gen]
    off = 11 + 16
    if gen == 0:  # Fast: separable angles, one elevation per row and one azimuth per column
        el = np.frombuffer(raw, "<f8", h, off)
        az = np.frombuffer(raw, "<f8", w, off + 8 * h)
        off += 8 * (h + w)
        s.elevation_angle, s.azimuth = np.repeat(el[:, None], w, 1), np.repeat(az[None, :], h, 0)
    else:
        s.elevation_angle = np.frombuffer(raw, "<f8", h * w, off).reshape(h, w)
        s.azimuth = np.frombuffer(raw, "<f8", h * w, off + 8 * h * w).reshape(h, w)
        off += 16 * h * w
    s.first = np.frombuffer(raw, META_DTYPE, h * w, off).reshape(h, w)
    off += s.first.nbytes
    s.counts = s.points = None
    s.max_points = 1
    if s.version == 3:
        s.max_points = int(np.frombuffer(raw, "<i4", 1, off)[0])
        s.counts = np.frombuffer(raw, "<i4", h * w, off + 4).reshape(h, w)
        off += 4 + 4 * h * w
        s._kept = np.minimum(s.counts.ravel(), s.max_points).astype(np.int64)
        s._first_index = np.concatenate([[0], np.cumsum(s._kept)[:-1]])
        s.points = np.frombuffer(raw, TRACE_DTYPE, int(s._kept.sum()), off)
        off += s.points.nbytes
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} bytes behind the last record")
    return s
