"""GeoTIFF terrain tiles (terrain/geotiff.rs): the host's TIFF reader (csrc/host/geotiff.cpp) against files
written by an independent encoder (Pillow / libtiff) and by a minimal writer below (tiles, big-endian), the lowering to a
DTED-like tile against the restatement of GeoTiffWrapper::get_elev in oracle/geotiff.py, Terrain::from_folder on mixed folders."""
import struct
import zlib

import numpy as np
import pytest

from atm_raytracer_b200 import host, synth, terrain
from oracle import geotiff as ref

N = 3601


def raster(seed, lo=-400, hi=4500):
    """A smooth-ish 3601 x 3601 i16 picture with structure in both directions (a transposed or flipped read cannot pass)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:N, 0:N]
    a = 800.0 + 600.0 * np.sin(x / 211.0) + 900.0 * np.cos(y / 173.0) + 0.3 * x - 0.2 * y + rng.integers(-40, 40, (N, N))
    return np.clip(a, lo, hi).astype(np.int16)


def write_pillow(path, a, compression, predictor=None):
    Image = pytest.importorskip("PIL.Image")  # libtiff through Pillow: the independent encoder

    im = Image.fromarray(a.view(np.uint16))  # mode I;16; SampleFormat 2 marks the samples as two's complement
    info = {339: 2}
    if predictor:
        info[317] = predictor
    im.save(path, compression=compression, tiffinfo=info)


def write_minimal(path, a, big=False, tile=None, deflate=False, sample_format=2, rows_per_strip=64):
    """Classic TIFF, one IFD behind the data: strips of rows_per_strip rows or tiles of tile x tile pixels."""
    e = ">" if big else "<"
    h, w = a.shape
    bits = a.dtype.itemsize * 8
    data = a.astype(a.dtype.newbyteorder(e))
    chunks = []
    if tile:
        for y in range(0, h, tile):
            for x in range(0, w, tile):
                t = np.zeros((tile, tile), data.dtype)
                part = data[y:y + tile, x:x + tile]
                t[:part.shape[0], :part.shape[1]] = part
                chunks.append(t.tobytes())
    else:
        for y in range(0, h, rows_per_strip):
            chunks.append(data[y:y + rows_per_strip].tobytes())
    if deflate:
        chunks = [zlib.compress(c, 6) for c in chunks]
    out = bytearray((b"MM" if big else b"II") + struct.pack(e + "HI", 42, 0))
    offsets = []
    for c in chunks:
        offsets.append(len(out))
        out += c
    counts = [len(c) for c in chunks]

    def longs(v):
        at = len(out)
        out.extend(struct.pack(e + f"{len(v)}I", *v))
        return at

    off_at, cnt_at = longs(offsets), longs(counts)
    n = len(chunks)
    ent = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, bits), (259, 3, 1, 8 if deflate else 1), (262, 3, 1, 1), (277, 3, 1, 1), (339, 3, 1, sample_format)]
    if tile:
        ent += [(322, 3, 1, tile), (323, 3, 1, tile), (324, 4, n, off_at if n > 1 else offsets[0]), (325, 4, n, cnt_at if n > 1 else counts[0])]
    else:
        ent += [(278, 3, 1, rows_per_strip), (273, 4, n, off_at if n > 1 else offsets[0]), (279, 4, n, cnt_at if n > 1 else counts[0])]
    ifd = len(out)
    out += struct.pack(e + "H", len(ent))
    for tag, typ, cnt, val in sorted(ent):
        out += struct.pack(e + "HHI", tag, typ, cnt)
        out += struct.pack(e + "HH", val, 0) if typ == 3 and cnt == 1 else struct.pack(e + "I", val)
    out += struct.pack(e + "I", 0)
    out[4:8] = struct.pack(e + "I", ifd)
    open(path, "wb").write(bytes(out))


def expect_posts(a):
    """posts[lon line][lat point], south -> north: raster row 3600 - lat, column lon (see geotiff.cpp)."""
    return np.ascontiguousarray(a[::-1, :].T[:N, :N])


@pytest.mark.parametrize("name,want", [("N45E006.tif", (45, 6)), ("ALPSMLC30_S012W077_DSM.tif", (-12, -77)), ("n45e006.tif", None),
                                       ("tile.tif", None), ("N99999E1.tif", None), ("xN1N45E6y", (45, 6)), ("N4E", None), ("S00W000", (0, 0)),
                                       ("N45E6_N46E7.tif", (45, 6)), ("N5E40000_N46E7.tif", None)])
def test_coords_from_name(name, want):
    assert ref.coords_from_name(name) == want
    assert host.geotiff_coords_from_name("/some/folder N9E9/" + name) == want


@pytest.fixture(scope="module")
def picture():
    return raster(3)


@pytest.mark.parametrize("compression,predictor", [(None, None), ("tiff_lzw", None), ("tiff_lzw", 2), ("tiff_adobe_deflate", None),
                                                   ("tiff_adobe_deflate", 2), ("packbits", None)])
def test_reads_what_libtiff_writes(tmp_path, picture, compression, predictor):
    path = tmp_path / "N46E007.tif"
    write_pillow(str(path), picture, compression, predictor)
    d, posts = host.read_geotiff(str(path))
    assert (d.lat0, d.lon0, d.nlon, d.nlat, d.min_lat, d.min_lon, d.lat_interval, d.lon_interval) == (46, 7, N, N, 46.0, 7.0, 1.0, 1.0)
    np.testing.assert_array_equal(posts, expect_posts(picture))


@pytest.mark.parametrize("kw", [dict(big=True), dict(tile=256), dict(tile=512, deflate=True, big=True), dict(deflate=True, rows_per_strip=1000),
                                dict(rows_per_strip=N)], ids=["big_endian", "tiles", "tiles_deflate_big", "strips_deflate", "one_strip"])
def test_reads_minimal_writer_variants(tmp_path, picture, kw):
    path = tmp_path / "S01W001.tif"
    write_minimal(str(path), picture, **kw)
    d, posts = host.read_geotiff(str(path))
    assert (d.lat0, d.lon0, d.min_lat, d.min_lon) == (-1, -1, -1.0, -1.0)
    np.testing.assert_array_equal(posts, expect_posts(picture))


def test_other_integer_types_and_larger_pictures(tmp_path):
    a = raster(4, 0, 250)
    write_minimal(str(tmp_path / "N00E000_u8.tif"), a.astype(np.uint8), sample_format=1)
    np.testing.assert_array_equal(host.read_geotiff(str(tmp_path / "N00E000_u8.tif"))[1], expect_posts(a))
    b = raster(5)
    write_minimal(str(tmp_path / "N00E000_i32.tif"), b.astype(np.int32), deflate=True, rows_per_strip=512)
    np.testing.assert_array_equal(host.read_geotiff(str(tmp_path / "N00E000_i32.tif"))[1], expect_posts(b))
    c = np.pad(raster(6), ((0, 7), (0, 11)))  # more than 3601 x 3601: the reference never addresses the rest
    write_minimal(str(tmp_path / "N00E000_big.tif"), c, tile=256)
    np.testing.assert_array_equal(host.read_geotiff(str(tmp_path / "N00E000_big.tif"))[1], np.ascontiguousarray(c[N - 1::-1, :N].T))


def test_refusals(tmp_path, picture):
    small = tmp_path / "N10E010.tif"
    write_minimal(str(small), picture[:3600, :3600])
    with pytest.raises(host.HostError, match="3601 x 3601"):
        host.read_geotiff(str(small))
    f32 = tmp_path / "N11E010.tif"
    write_minimal(str(f32), picture.astype(np.float32), sample_format=3)
    with pytest.raises(host.HostError, match="floating-point"):
        host.read_geotiff(str(f32))
    wide = tmp_path / "N12E010.tif"
    big = picture.astype(np.int32)
    big[100, 200] = 40000
    write_minimal(str(wide), big)
    with pytest.raises(host.HostError, match="does not fit"):
        host.read_geotiff(str(wide))
    u16 = tmp_path / "N13E010.tif"
    write_minimal(str(u16), picture.view(np.uint16), sample_format=1)  # negative heights read as unsigned exceed i16
    with pytest.raises(host.HostError, match="does not fit"):
        host.read_geotiff(str(u16))
    junk = tmp_path / "N14E010.tif"
    junk.write_bytes(b"II\x2b\x00" + bytes(64))
    with pytest.raises(host.HostError, match="BigTIFF"):
        host.read_geotiff(str(junk))
    noname = tmp_path / "tile.tif"
    write_minimal(str(noname), picture)
    with pytest.raises(host.HostError, match="file name"):
        host.read_geotiff(str(noname))
    with pytest.raises(host.HostError, match="Could not buffer terrain file"):
        host.read_tile(str(noname))


def test_lowered_tile_samples_like_the_wrapper(tmp_path, picture, oracle_lib):
    """GeoTiffWrapper::get_elev (geotiff.rs:62-99) == DtedData::get_elev on the lowered tile (3601 x 3601 posts, 1 arc-second),
    bit for bit: the oracle's DTED sampler on the descriptor + posts the host produced, against the wrapper restated."""
    path = tmp_path / "N46E007.tif"
    write_pillow(str(path), picture, "tiff_adobe_deflate", 2)
    d, posts = host.read_tile(str(path))
    t = terrain.Terrain([(d, posts)])
    rng = np.random.default_rng(9)
    lat = np.concatenate([46.0 + rng.random(4000), [46.0, 47.0, 46.0, 47.0, 46.5, 46.0 + 3599.5 / 3600, 47.0 - 1e-13, 46.0 + 1e-13]])
    lon = np.concatenate([7.0 + rng.random(4000), [7.0, 8.0, 8.0, 7.0, 8.0, 7.0 + 3600 / 3600, 8.0 - 1e-13, 7.0 + 1e-13]])
    got = oracle_lib.get_elev(t.tiles, lat, lon)
    rows = picture.tolist()
    # Terrain::get_elev (terrain/mod.rs:120-126) keys on (floor(lat), floor(lon)): the tile's own north and east edges belong to
    # its neighbours' keys, so the wrapper's "3600" edge case is reachable only just below them
    want = np.array([ref.get_elev(rows, 46.0, 7.0, float(a), float(b)) if (np.floor(a), np.floor(b)) == (46.0, 7.0) else np.nan for a, b in zip(lat, lon)])
    assert np.isnan(want).sum() == 5 and np.array_equal(got, want, equal_nan=True)


def test_from_folder_takes_dted_and_geotiff(tmp_path, picture, capsys):
    folder = tmp_path / "terrain"
    folder.mkdir()
    synth.write_tile_grid(str(folder), 45, 5, 1, 2, level=0)
    write_pillow(str(folder / "N46E005_dsm.tif"), picture, "tiff_lzw", 2)
    t = terrain.Terrain.from_folder(str(folder))
    assert "Detected 3 terrain files" in capsys.readouterr().out
    keys = sorted((d.lat0, d.lon0, d.nlat) for d, _ in t.tiles)
    assert keys[-1] == (46, 5, N) and [k[:2] for k in keys[:2]] == [(45, 5), (45, 6)]
    (folder / "notes.txt").write_text("not a tile")
    with pytest.raises(host.HostError, match="Could not buffer terrain file"):
        terrain.Terrain.from_folder(str(folder))


@pytest.mark.gpu
def test_device_samples_a_geotiff_tile_like_the_wrapper(tmp_path, picture, ctx):
    path = tmp_path / "S03W071.tif"
    write_pillow(str(path), picture, "tiff_lzw", 2)
    t = terrain.Terrain([host.read_tile(str(path))])
    ctx.set_terrain(t)
    rng = np.random.default_rng(10)
    lat = np.concatenate([-3.0 + rng.random(20000), [-3.0, -2.0, -2.0, -3.0]])
    lon = np.concatenate([-71.0 + rng.random(20000), [-71.0, -70.0, -71.0, -70.0]])
    got = ctx.get_elev(lat, lon)
    rows = picture.tolist()
    want = np.array([ref.get_elev(rows, -3.0, -71.0, float(a), float(b)) if (np.floor(a), np.floor(b)) == (-3.0, -71.0) else np.nan for a, b in zip(lat, lon)])
    assert np.isnan(want).sum() == 3 and np.array_equal(got, want, equal_nan=True)


def test_damaged_files_are_refused_not_crashed_on(tmp_path, picture):
    """Truncated and bit-flipped TIFFs (LZW strips, deflate tiles): the reader answers with an error or with a tile, never
    with a crash -- every offset, count and code of the file is checked before it is used."""
    good = {}
    write_pillow(str(tmp_path / "N20E020.tif"), picture, "tiff_lzw", 2)
    good["lzw"] = (tmp_path / "N20E020.tif").read_bytes()
    write_minimal(str(tmp_path / "N20E020.tif"), picture, tile=512, deflate=True)
    good["tiles"] = (tmp_path / "N20E020.tif").read_bytes()
    rng = np.random.default_rng(12)
    path = tmp_path / "N21E021.tif"
    outcomes = set()
    for kind, data in good.items():
        n = len(data)
        for trial in range(10):
            b = bytearray(data)
            if trial < 3:
                b = b[: int(n * rng.random())]  # truncated
            elif trial < 7:
                for pos in rng.integers(0, n, 40):  # noise anywhere (mostly in the compressed data)
                    b[pos] ^= 1 << int(rng.integers(0, 8))
            else:
                ifd = struct.unpack("<I", data[4:8])[0]  # noise in the directory: tags, types, counts, offsets
                for pos in rng.integers(ifd, min(ifd + 200, n), 6):
                    b[pos] = int(rng.integers(0, 256))
            path.write_bytes(bytes(b))
            try:
                d, posts = host.read_geotiff(str(path))
                assert posts.shape == (N, N)
                outcomes.add("read")
            except host.HostError:
                outcomes.add("refused")
    assert "refused" in outcomes


@pytest.mark.gpu
def test_gen_renders_from_a_folder_with_a_geotiff_tile(tmp_path, picture, ctx):
    """`atm-raytracer gen` on a folder holding a GeoTIFF tile and DTED tiles == the library driven through the Python mirror on
    Terrain.from_folder of the same folder (Terrain::from_folder / buffer_file, terrain/mod.rs:66-118)."""
    import subprocess

    from atm_raytracer_b200 import config, runtime

    folder = tmp_path / "terrain"
    folder.mkdir()
    synth.write_tile_grid(str(folder), 45, 5, 1, 2, level=0)           # DTED: (45, 5), (45, 6)
    write_pillow(str(folder / "srtm_N46E005.tif"), picture, "tiff_lzw", 2)  # GeoTIFF: (46, 5), 3601 x 3601
    png = tmp_path / "out.png"
    argv = ["-t", str(folder), "--output", str(png), "-l", "45.9", "-g", "5.5", "-a", "3000", "-d", "10", "-f", "50", "-m", "60", "--step", "100",
            "-w", "320", "-h", "200", "-i", "-8"]
    r = subprocess.run([host.EXECUTABLE, "gen"] + argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Detected 3 terrain files" in r.stdout
    img = host.read_png(str(png))[..., :3]
    t = terrain.Terrain.from_folder(str(folder))
    want = runtime.FastGenerator(config.into_params(config.read_config(argv)), t, [], [], context=ctx).generate()
    np.testing.assert_array_equal(img, want["rgb"])
    north = np.isfinite(want["meta"]["lat"]) & (want["meta"]["lat"] > 46.0)
    assert north.mean() > 0.05  # the picture does look onto the GeoTIFF tile
