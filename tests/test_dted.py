"""DTED decode (terrain/mod.rs:24,85-98 through the external `dted` 0.2 crate): the C++ host decoder
is bit-exact on files written by an independent MIL-PRF-89020B writer (synth.py), including negative
posts (signed magnitude), L1 geometry, header origins in all hemispheres; the oracle's own reader
agrees."""
import numpy as np
import pytest

from atm_raytracer_b200 import host, synth


@pytest.mark.parametrize("lat0,lon0", [(45, 5), (-12, -77), (0, 0), (-1, 179)])
def test_round_trip_bit_exact(tmp_path, oracle_lib, lat0, lon0):
    rng = np.random.default_rng(abs(lat0 * 1000 + lon0) + 1)
    posts = rng.integers(-500, 9000, size=(121, 121)).astype(np.int16)
    posts[3, 7] = -32767  # void marker is just another signed-magnitude value
    posts[0, 0] = -1
    path = tmp_path / synth.tile_name(lat0, lon0, 0)
    synth.write_dted(str(path), lat0, lon0, posts)
    d, got = host.read_dted(str(path))
    assert (d.lat0, d.lon0, d.nlon, d.nlat) == (lat0, lon0, 121, 121)
    assert (d.min_lat, d.min_lon) == (float(lat0), float(lon0))
    assert d.lat_interval == 30.0 and d.lon_interval == 30.0
    np.testing.assert_array_equal(got, posts)
    d2, got2 = oracle_lib.read_dted(str(path))
    np.testing.assert_array_equal(got2, posts)
    assert (d2.lat0, d2.lon0, d2.min_lat, d2.min_lon) == (d.lat0, d.lon0, d.min_lat, d.min_lon)


def test_level1_geometry_and_synthetic_field(tmp_path):
    posts = synth.make_tile(45, 5, 1)
    assert posts.shape == (1201, 1201) and posts.min() < 0 < posts.max()  # the basin is below sea level
    path = tmp_path / "n45_e005.dt1"
    synth.write_dted(str(path), 45, 5, posts)
    d, got = host.read_dted(str(path))
    assert d.lat_interval == 3.0 and d.nlat == 1201
    np.testing.assert_array_equal(got, posts)
    # seams: the east edge of (45,5) is the west edge of (45,6); north edge likewise
    np.testing.assert_array_equal(posts[-1, :], synth.make_tile(45, 6, 1)[0, :])
    np.testing.assert_array_equal(posts[:, -1], synth.make_tile(46, 5, 1)[:, 0])


def test_header_only_and_errors(tmp_path):
    path = tmp_path / "x.dt0"
    synth.write_dted(str(path), 10, 20, np.zeros((121, 121), np.int16))
    d = host.read_dted_header(str(path))
    assert (d.lat0, d.lon0) == (10, 20)
    bad = tmp_path / "bad.dt0"
    bad.write_bytes(b"not a dted file" * 10)
    with pytest.raises(host.HostError):
        host.read_dted(str(bad))
    with pytest.raises(host.HostError):
        host.read_dted(str(tmp_path / "missing.dt0"))
    trunc = tmp_path / "trunc.dt0"
    trunc.write_bytes(path.read_bytes()[:5000])
    with pytest.raises(host.HostError):
        host.read_dted(str(trunc))


def test_png_round_trip(tmp_path):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(37, 53, 3), dtype=np.uint8)
    p = tmp_path / "a.png"
    host.write_png(str(p), img)
    back = host.read_png(str(p))
    np.testing.assert_array_equal(back[..., :3], img)
    assert (back[..., 3] == 255).all()
    from PIL import Image  # independent decoder

    np.testing.assert_array_equal(np.asarray(Image.open(str(p)).convert("RGB")), img)


@pytest.mark.parametrize("shape", [(1, 1, 3), (7, 5, 4), (33, 1000, 3), (900, 1200, 3), (2600, 1700, 4), (5, 700001, 3)])
def test_png_written_in_parallel_bands_decodes_everywhere(tmp_path, shape):
    """write_png deflates the scanlines in bands on all host threads and writes one IDAT chunk per band (one zlib stream, the
    pigz construction): the host's own reader and an independent decoder (Pillow / libpng) must both return the pixels."""
    rng = np.random.default_rng(shape[1])
    h, w, ch = shape
    img = (np.add.outer(np.arange(h) * 3, np.arange(w) * 5)[..., None] // (1 + np.arange(ch)) + rng.integers(0, 9, (h, w, ch))).astype(np.uint8)
    path = str(tmp_path / "x.png")
    host.write_png(path, img)
    back = host.read_png(path)
    np.testing.assert_array_equal(back[..., :ch], img)
    if ch == 3:
        assert (back[..., 3] == 255).all()
    Image = pytest.importorskip("PIL.Image")
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        im.load()
        assert im.mode == ("RGB" if ch == 3 else "RGBA")
        np.testing.assert_array_equal(np.asarray(im), img)
    raw = open(path, "rb").read()
    assert raw.count(b"IDAT") >= max(1, (h * (w * ch + 1)) // (4 << 20))  # several bands for a large picture
