"""Synthetic DTED tiles for tests and benchmarks (no real DEM data is available offline).

The file writer follows MIL-PRF-89020B (UHL 80 B + DSI 648 B + ACC 2700 B, then one 0xAA record per
longitude line with big-endian signed-magnitude posts and a 4-byte checksum), which is what the
reference reads through the external ``dted`` crate (terrain/mod.rs:24,85-98). The height field is
the seeded multi-octave function of SURVEY.md section 8(d): a function of absolute lat/lon, so tile
seams are continuous, with one below-sea-level basin (exercises signed magnitude) and one plateau.
"""
import os

import numpy as np

SEED = 20260101
LEVELS = {0: (30.0, 121), 1: (3.0, 1201), 2: (1.0, 3601)}  # arc-second spacing, posts per degree + 1


def _octaves():
    rng = np.random.default_rng(SEED)
    amps = np.array([800.0, 400.0, 200.0, 100.0, 50.0, 25.0])
    freq = np.array([0.9, 2.1, 4.7, 9.3, 19.0, 41.0])
    ang = rng.uniform(0.0, 2.0 * np.pi, size=6)
    phase = rng.uniform(0.0, 2.0 * np.pi, size=6)
    return amps, freq * np.cos(ang), freq * np.sin(ang), phase


_AMPS, _F, _G, _PH = _octaves()


def terrain_height(lat, lon):
    """Height in metres (float64, before rounding) at absolute lat/lon in degrees."""
    lat = np.asarray(lat, dtype=np.float64)
    lon = np.asarray(lon, dtype=np.float64)
    h = np.full(np.broadcast(lat, lon).shape, 400.0)
    for a, f, g, p in zip(_AMPS, _F, _G, _PH):
        h = h + a * np.sin(2.0 * np.pi * (f * lat + g * lon) + p)
    # a ridge running roughly east-west 0.6 degrees north of N45
    h = h + 900.0 * np.exp(-(((lat - 45.6) - 0.08 * np.sin(3.0 * lon)) / 0.05) ** 2)
    # a basin below sea level and a plateau
    basin = np.exp(-(((lat - 45.25) / 0.06) ** 2 + ((lon - 5.3) / 0.08) ** 2))
    h = h * (1.0 - basin) + (-40.0) * basin
    plateau = ((np.abs(lat - 45.8) < 0.04) & (np.abs(lon - 5.75) < 0.06))
    h = np.where(plateau, 1500.0, h)
    return h


def make_tile(lat0, lon0, level=1):
    """int16 posts [nlon][nlat] (west->east, south->north) of the synthetic field for the 1x1 degree
    tile with SW corner (lat0, lon0)."""
    _, n = LEVELS[level]
    lats = lat0 + np.arange(n, dtype=np.float64) / (n - 1)
    lons = lon0 + np.arange(n, dtype=np.float64) / (n - 1)
    h = terrain_height(lats[None, :], lons[:, None])
    return np.clip(np.rint(h), -50, 4000).astype(np.int16)


def _dms(value, deg_digits):
    hemi_pos, hemi_neg = ("N", "S") if deg_digits == 2 else ("E", "W")
    hemi = hemi_pos if value >= 0 else hemi_neg
    v = abs(value)
    total = int(round(v * 3600.0))
    deg, rem = divmod(total, 3600)
    mn, sc = divmod(rem, 60)
    return deg, mn, sc, hemi


def encode_dted(lat0, lon0, posts, lat_interval=None, lon_interval=None):
    """Serialise ``posts`` ([nlon][nlat] int16) as a DTED file image (bytes)."""
    posts = np.asarray(posts, dtype=np.int16)
    nlon, nlat = posts.shape
    if lat_interval is None:
        lat_interval = 3600.0 / (nlat - 1)
    if lon_interval is None:
        lon_interval = 3600.0 / (nlon - 1)
    d, m, s, hm = _dms(lon0, 3)
    lon_s = f"{d:03d}{m:02d}{s:02d}{hm}"
    d, m, s, hm = _dms(lat0, 2)
    lat_s = f"{d:03d}{m:02d}{s:02d}{hm}"  # UHL uses DDDMMSSH for both
    uhl = (
        "UHL1" + lon_s + lat_s + f"{int(round(lon_interval * 10)):04d}" + f"{int(round(lat_interval * 10)):04d}"
        + "NA  " + "U  " + " " * 12 + f"{nlon:04d}" + f"{nlat:04d}" + "0" + " " * 24
    )
    assert len(uhl) == 80, len(uhl)
    dsi = ("DSIU" + " " * 644)[:648]
    acc = ("ACC" + " " * 2697)[:2700]
    out = bytearray(uhl.encode("ascii") + dsi.encode("ascii") + acc.encode("ascii"))
    mag = np.abs(posts.astype(np.int32)).astype(np.uint16)
    enc = np.where(posts < 0, mag | 0x8000, mag).astype(">u2")  # signed magnitude, big endian
    for i in range(nlon):
        rec = bytearray([0xAA, (i >> 16) & 0xFF, (i >> 8) & 0xFF, i & 0xFF, (i >> 8) & 0xFF, i & 0xFF, 0, 0])
        rec += enc[i].tobytes()
        chk = sum(rec) & 0xFFFFFFFF
        rec += chk.to_bytes(4, "big")
        out += rec
    return bytes(out)


def write_dted(path, lat0, lon0, posts, **kw):
    with open(path, "wb") as f:
        f.write(encode_dted(lat0, lon0, posts, **kw))


def tile_name(lat0, lon0, level=1):
    ns = "n" if lat0 >= 0 else "s"
    ew = "e" if lon0 >= 0 else "w"
    return f"{ns}{abs(int(lat0)):02d}_{ew}{abs(int(lon0)):03d}.dt{level}"


def write_tile_grid(folder, lat0=45, lon0=5, nlat_tiles=1, nlon_tiles=1, level=1):
    """Write an nlat_tiles x nlon_tiles grid of synthetic tiles; returns the list of paths."""
    os.makedirs(folder, exist_ok=True)
    paths = []
    for i in range(nlat_tiles):
        for j in range(nlon_tiles):
            p = os.path.join(folder, tile_name(lat0 + i, lon0 + j, level))
            if not os.path.exists(p):
                write_dted(p, lat0 + i, lon0 + j, make_tile(lat0 + i, lon0 + j, level))
            paths.append(p)
    return paths


def tile_grid_arrays(lat0=45, lon0=5, nlat_tiles=1, nlon_tiles=1, level=1):
    """In-memory version: list of (lat0, lon0, posts) without touching the filesystem."""
    return [
        (lat0 + i, lon0 + j, make_tile(lat0 + i, lon0 + j, level))
        for i in range(nlat_tiles)
        for j in range(nlon_tiles)
    ]
