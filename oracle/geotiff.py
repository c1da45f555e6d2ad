"""TEST INFRASTRUCTURE -- CPU restatement of the reference's GeoTIFF tile (terrain/geotiff.rs:9-100). Only tests/ may import this.

`raster[row][column]` is the decoded picture, row 0 at the top (north). geotiff-rs 0.1 (external crate, not vendored with the
reference) supplies `get_pixel(lon, lat)`; PARITY UNPINNED on its row sense -- taken, as in csrc/host/geotiff.cpp, as the post
`lat` arc-seconds north of the south edge: raster[3600 - lat][lon]."""
import re

_RE = re.compile(r"(N|S)(\d+)(E|W)(\d+)", re.ASCII)


def coords_from_name(file_name):  # geotiff.rs:16-31
    m = _RE.search(file_name)
    if not m:
        return None
    lat, lon = int(m.group(2)), int(m.group(4))
    if lat > 32767 or lon > 32767:  # i16::from_str(..).ok()?
        return None
    return (-lat if m.group(1) == "S" else lat, -lon if m.group(3) == "W" else lon)


def get_pixel(raster, lon, lat):
    return raster[3600 - lat][lon]


def get_elev(raster, min_lat, min_lon, lat, lon):  # geotiff.rs:62-99, one point, Python floats (f64)
    if lat < min_lat or lat > min_lat + 1.0 or lon < min_lon or lon > min_lon + 1.0:
        return None
    lat = (lat - min_lat) * 3600.0
    lon = (lon - min_lon) * 3600.0
    lat_int, lon_int = int(lat), int(lon)
    lat_frac, lon_frac = lat - float(lat_int), lon - float(lon_int)
    if lat_int == 3600:
        lat_int -= 1
        lat_frac += 1.0
    if lon_int == 3600:
        lon_int -= 1
        lon_frac += 1.0
    e00 = float(get_pixel(raster, lon_int, lat_int))
    e01 = float(get_pixel(raster, lon_int, lat_int + 1))
    e10 = float(get_pixel(raster, lon_int + 1, lat_int))
    e11 = float(get_pixel(raster, lon_int + 1, lat_int + 1))
    return e00 * (1.0 - lon_frac) * (1.0 - lat_frac) + e01 * (1.0 - lon_frac) * lat_frac + e10 * lon_frac * (1.0 - lat_frac) + e11 * lon_frac * lat_frac
