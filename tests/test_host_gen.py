"""The thin C++ host of `gen` (csrc/host/gen.cpp; generator/mod.rs:47-99, generator/params.rs:447-777):
its hand-written YAML-subset parser + CLI overrides must lower to the same flat `atmrt_params` as the
Python mirror (PyYAML + argparse), and the executable must render the same image as the library."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config, host, synth
from test_config import YAML

BLOCK_YAML = """
# block style, comments, quoted strings, list at the key's indentation
scene:
  terrain_folder: "./my terrain"   # trailing comment
  objects:
  - position:
      latitude: 45.1
      longitude: 6.0
      altitude:
        Relative: 0.0
    shape:
      Cylinder:
        radius: 60.0
        height: 900.0
    color:
      r: 0.9
      g: 0.1
      b: 0.1
  - position: {latitude: 45.2, longitude: 6.1, altitude: {Absolute: 1500.0}}
    shape:
      Frustum: {r1: 500.0, r2: 200.0, height: 2500.0}
    color: {r: 1.0, g: 1.0, b: 1.0, a: 0.25}
view:
  position:
    latitude: 45.05
    longitude: 6.0
  frame:
    fov: 20
    max_distance: 1.2e5
  coloring:
    Simple:
      water_level: 2.0
atmosphere:
  pressure: {altitude: 0.0, pressure: 101325.0}
  temperature_fixed_point: {altitude: 0.0, temperature: 283.0}
  first_temperature_function:
    Linear: {gradient: 0.01}
  next_functions:
    - altitude: 300.0
      function:
        Linear:
          gradient: -0.0065
earth_shape:
  Spherical:
    radius: 6400000.0
wavelength: 600e-9
output:
  width: 320
  height: 200
"""


def _fields(s, skip=()):
    out = {}
    for name, _ in s._fields_:
        if name.startswith("_") or name in skip:
            continue
        v = getattr(s, name)
        if isinstance(v, C.Structure):
            out[name] = _fields(v)
        elif isinstance(v, C.Array):
            out[name] = [list(x) if isinstance(x, C.Array) else x for x in v]
        else:
            out[name] = v
    return out


def _same_params(a, b):
    fa, fb = _fields(a), _fields(b)
    la, lb = fa.pop("light_dir"), fb.pop("light_dir")
    np.testing.assert_allclose(la, lb, atol=1e-15)  # host libm on both sides, different expression grouping
    # unused tail of the fixed-size atmosphere arrays
    n = fa["atmosphere"]["n_functions"]
    for f in (fa, fb):
        for k in ("fn_start_altitude", "fn_gradient"):
            f["atmosphere"][k] = f["atmosphere"][k][:n]
        f["atmosphere"]["fn_start_altitude"][0] = 0.0
        if f["earth_model"] == abi.EARTH_FLAT_DISTORTED:
            f["radius"] = 0.0  # unused by the flat model
    assert fa == fb


# The README's own Spline example (README.md:296-316: nested block lists), and the compact spellings of the same
SPLINE_YAML = """
atmosphere:
    pressure:
        altitude: 0.0
        pressure: 101325
    first_temperature_function:
        Linear:
            gradient: -0.0065
    next_functions:
        - altitude: 100.0
          function:
            Spline:
                boundary_condition:
                    Derivatives:
                        - -0.0065
                        - 0.0
                points:
                    -
                        - 100.0
                        - 288.0
                    -
                        - 110.0
                        - 285.0
                    - - 120.0
                      - 291.0
        - altitude: 400.0
          function:
            Spline:
                boundary_condition: Natural
                points: [[400.0, 290.0], [600.0, 288.0], [1000.0, 286.0]]
        - altitude: 1000.0
          function: {Spline: {boundary_condition: {SecondDerivatives: [1.0e-6, 0.0]}, points: [[1000.0, 286.0], [3000.0, 270.0]]}}
"""


@pytest.mark.parametrize("text", [YAML, BLOCK_YAML, SPLINE_YAML], ids=["flow", "block", "spline"])
def test_cpp_yaml_parser_matches_pyyaml(tmp_path, text):
    f = tmp_path / "c.yaml"
    f.write_text(text)
    argv = ["-c", str(f)]
    p_cpp, objs_cpp, folder, out, meta = host.parse_config(argv)
    cfg = config.read_config(argv)
    _same_params(p_cpp, config.into_params(cfg))
    objs_py, _ = config.lower_objects(cfg)
    assert len(objs_cpp) == len(objs_py)
    for a, b in zip(objs_cpp, objs_py):
        assert _fields(a, skip=("texture_width", "texture_height")) == _fields(b, skip=("texture_width", "texture_height"))
    assert folder == cfg["scene"]["terrain_folder"] and out == cfg["output"]["file"]
    assert meta == (cfg["output"].get("file_metadata") or "")


@pytest.mark.parametrize("shape", ["Wgs84", "SimpleSphere", "AzimuthalEquidistant", "SimpleObserverAe", "FlatDistorted",
                                   "{Ellipsoid: {a: 6400000.0, b: 6300000.0}}", "{ObserverAe: {proj_radius: 6380000.0}}",
                                   "{Spherical: {radius: 7000000.0}}"])
def test_cpp_parser_lowers_every_earth_model_like_the_mirror(tmp_path, shape):
    f = tmp_path / "c.yaml"
    f.write_text(f"earth_shape: {shape}\nview:\n  coloring:\n    Shading: {{water_level: 0.0, ambient_light: 0.4, light_zenith_angle: 45.0, light_dir: 0.0}}\n")
    p_cpp, *_ = host.parse_config(["-c", str(f)])
    _same_params(p_cpp, config.into_params(config.read_config(["-c", str(f)])))


def test_cpp_cli_overrides_match_the_mirror(tmp_path):
    f = tmp_path / "c.yaml"
    f.write_text(YAML)
    argv = ["-c", str(f), "-m", "100", "-R", "7000", "-w", "320", "-h", "200", "-e", "12", "-s", "--step", "50", "-d", "-30", "-l", "46",
            "-g", "7", "-f", "45", "-i", "2.5", "--output", "x.png", "--output-meta", "x.dat", "-t", "/data/terrain"]
    p_cpp, _, folder, out, meta = host.parse_config(argv)
    _same_params(p_cpp, config.into_params(config.read_config(argv)))
    assert (folder, out, meta) == ("/data/terrain", "x.png", "x.dat")
    assert p_cpp.max_distance == 100e3 and p_cpp.radius == 7000e3 and p_cpp.altitude.kind == abi.ALT_RELATIVE  # km on the CLI
    p_def, objs, folder, out, meta = host.parse_config([])
    _same_params(p_def, config.into_params(config.default_config()))
    assert (objs, folder, out, meta) == ([], "./terrain", "./output.png", "")
    p_flat, *_ = host.parse_config(["--flat", "-a", "2500"])
    assert p_flat.earth_model == abi.EARTH_FLAT_DISTORTED and p_flat.altitude.kind == abi.ALT_ABSOLUTE and p_flat.altitude.value == 2500.0
    for bad in (["--flat", "-R", "6371"], ["-a", "1", "-e", "2"], ["--bogus"], ["-w"]):
        with pytest.raises(host.HostError):
            host.parse_config(bad)


def test_cpp_scope_errors(tmp_path):
    for body in ("earth_shape: Geoid\n", "output: {generator: Fisheye}\n",
                 "atmosphere:\n  pressure: {altitude: 0, pressure: 101325}\n  first_temperature_function:\n    Spline: {points: [[0, 288]]}\n"):
        f = tmp_path / "bad.yaml"
        f.write_text(body)
        with pytest.raises(host.HostError):
            host.parse_config(["-c", str(f)])


def test_executable_reports_errors_like_the_reference(tmp_path):
    r = subprocess.run([host.EXECUTABLE, "gen", "-t", str(tmp_path / "missing")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("ERROR: ")      # main.rs:36-38
    r = subprocess.run([host.EXECUTABLE, "view", "x.dat"], capture_output=True, text=True)
    assert r.returncode == 2


@pytest.mark.gpu
def test_gen_executable_end_to_end(tmp_path, ctx):
    """`atm-raytracer gen` on DTED files on disk == the library driven through the Python mirror."""
    from atm_raytracer_b200 import runtime

    folder = tmp_path / "terrain"
    folder.mkdir()
    synth.write_tile_grid(str(folder), 45, 5, 1, 2, level=0)
    conf = tmp_path / "c.yaml"
    conf.write_text("view:\n  position: {latitude: 45.4, longitude: 5.9, altitude: {Relative: 30.0}}\n"
                    "  frame: {direction: 80.0, tilt: -2.0, fov: 40.0, max_distance: 60000.0}\n  fog_distance: 50000.0\n"
                    "scene:\n  terrain_alpha: 0.75\noutput: {width: 200, height: 120}\n")
    png, dat = tmp_path / "out.png", tmp_path / "out.dat"
    argv = ["-c", str(conf), "-t", str(folder), "--output", str(png), "--output-meta", str(dat), "--step", "100"]
    r = subprocess.run([host.EXECUTABLE, "gen"] + argv, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for stamp in ("Using terrain data directory", "Detected 2 terrain files", "Generating terrain cache...", "Generating path cache...",
                  "Calculating pixels...", "Outputting image...", "Outputting metadata...", "Done."):
        assert stamp in r.stdout  # the reference's stdout protocol (generator/mod.rs:62-96)
    img = host.read_png(str(png))[..., :3]
    assert img.shape == (120, 200, 3)
    raw = gzip.decompress(dat.read_bytes())
    # the host's own sidecar (NOT the reference's bincode container, and it says so): translucent terrain -> version 3,
    # with every trace point of every pixel behind the version-2 content
    assert raw.startswith(b"ATMRTMETA3\n") and "not the reference's bincode container" in r.stderr
    w, h, generator, _ = np.frombuffer(raw, "<i4", 4, 11)
    assert (w, h, generator) == (200, 120, 0)
    off = 11 + 16
    el = np.frombuffer(raw, "<f8", h, off)         # ResultPixel.elevation_angle, one per row (Fast generator)
    az = np.frombuffer(raw, "<f8", w, off + 8 * h)  # ResultPixel.azimuth, one per column, wrapped into [0, 360)
    off += 8 * (h + w)
    meta = np.frombuffer(raw, runtime.META_DTYPE, h * w, off).reshape(h, w)
    off += meta.nbytes
    max_points = int(np.frombuffer(raw, "<i4", 1, off)[0])
    counts = np.frombuffer(raw, "<i4", h * w, off + 4).reshape(h, w)
    kept = np.minimum(counts, max_points)
    lists = np.frombuffer(raw, runtime.TRACE_DTYPE, int(kept.sum()), off + 4 + 4 * h * w)
    assert max_points == 16 and len(raw) == off + 4 + 4 * h * w + lists.nbytes
    assert el[h // 2] == -2.0 and (np.diff(el) < 0).all() and az[w // 2] == 80.0 and (np.diff(az) > 0).all()
    # the package's reader of the sidecar sees the same content
    from atm_raytracer_b200 import sidecar

    sc = sidecar.read_sidecar(str(dat))
    assert (sc.version, sc.width, sc.height, sc.generator, sc.max_points) == (3, 200, 120, "Fast", 16)
    np.testing.assert_array_equal(sc.elevation_angle[:, 7], el)
    np.testing.assert_array_equal(sc.azimuth[5], az)
    np.testing.assert_array_equal(sc.first, meta)
    np.testing.assert_array_equal(sc.counts, counts)
    yy, xx = np.unravel_index(int(np.argmax(counts)), counts.shape)
    tp = sc.trace_points(yy, xx)
    assert len(tp) == min(counts[yy, xx], 16) and tp["distance"][0] == meta["distance"][yy, xx]
    # the same render through the Python mirror of the host
    cfg = config.read_config(argv)
    terrain = runtime.Terrain.from_folder(str(folder))
    gen = runtime.FastGenerator(config.into_params(cfg), terrain, [], [], context=ctx)
    want = gen.generate()
    np.testing.assert_array_equal(img, want["rgb"])
    for k in ("lat", "lon", "elevation", "distance"):
        np.testing.assert_array_equal(meta[k], want["meta"][k])
    assert np.isfinite(meta["distance"]).mean() > 0.2
    # ResultPixel.trace_points: the lists of the sidecar are the library's own trace render, pixel by pixel
    pts, cnt = ctx.render_trace(max_points=16)
    np.testing.assert_array_equal(counts, cnt)
    assert counts.max() >= 2  # translucent terrain: rays go on behind the first surface
    first = np.concatenate([[0], np.cumsum(kept.ravel())[:-1]]).reshape(h, w)
    hit = counts > 0
    np.testing.assert_array_equal(lists["distance"][first[hit]], meta["distance"][hit])
    ys, xs = np.nonzero(counts >= 2)
    for y, x in list(zip(ys, xs))[:200]:
        got = lists[first[y, x]:first[y, x] + kept[y, x]]
        for f in ("lat", "lon", "distance", "elevation", "path_length", "is_terrain", "step"):
            np.testing.assert_array_equal(got[f], pts[f][y, x, :kept[y, x]])


@pytest.mark.gpu
@pytest.mark.parametrize("gpus", [1, 2])
def test_gen_executable_draws_the_overlays(tmp_path, ctx, oracle_lib, gpus):
    """`gen` with output.ticks / vertical_ticks / show_eye_level / show_flat_horizon (renderer::output_image,
    renderer/mod.rs:416-431) == the plain render + the oracle's restatement of the overlays on the oracle's ResultPixel angles,
    the flat-earth horizon at acos(1 / n) of the oracle's atmosphere at the observer's altitude."""
    import torch
    from atm_raytracer_b200 import runtime
    from oracle import overlays as ref

    if gpus > torch.cuda.device_count():
        pytest.skip("needs two GPUs")
    folder = tmp_path / "terrain"
    folder.mkdir()
    synth.write_tile_grid(str(folder), 45, 5, 1, 2, level=0)
    conf = tmp_path / "c.yaml"
    conf.write_text("view:\n  position: {latitude: 45.4, longitude: 5.9, altitude: {Absolute: 1200.0}}\n"
                    "  frame: {direction: 80.0, tilt: -1.0, fov: 40.0, max_distance: 60000.0}\n"
                    "earth_shape: FlatDistorted\n"
                    "output:\n  width: 240\n  height: 140\n  show_eye_level: true\n  show_flat_horizon: true\n"
                    "  ticks:\n    - Multiple: {bias: 0, step: 10, size: 10, labelled: true}\n    - Multiple: {bias: 0, step: 2, size: 5, labelled: false}\n"
                    "    - Single: {azimuth: 75.5, size: 15, labelled: true}\n"
                    "  vertical_ticks:\n    - Multiple: {bias: 0, step: 5, size: 8, labelled: true}\n    - Multiple: {bias: 0, step: 1, size: 4, labelled: false}\n")
    png = tmp_path / "out.png"
    argv = ["-c", str(conf), "-t", str(folder), "--output", str(png), "--step", "100"]
    r = subprocess.run([host.EXECUTABLE, "gen"] + argv + ["--gpus", str(gpus)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "built-in bitmap face" in r.stderr and "ignored" not in r.stderr
    img = host.read_png(str(png))[..., :3]

    cfg = config.read_config(argv)
    params = config.into_params(cfg)
    terrain = runtime.Terrain.from_folder(str(folder))
    want = runtime.FastGenerator(params, terrain, [], [], context=ctx).generate()["rgb"].copy()
    plain = want.copy()
    el, az = oracle_lib.pixel_angles(params)
    _, _, n = oracle_lib.atmosphere(params.atmosphere, params.wavelength, np.array([1200.0]))
    ticks, vticks, eye, flat = host.parse_overlays(argv)
    assert eye and flat and len(ticks) == 3 and len(vticks) == 2
    labels = ref.output_overlays(want, el.tolist(), az.tolist(), ticks, vticks, dict(direction=80.0, fov=40.0, tilt=-1.0), show_eye_level=True,
                                 flat_horizon_elev=ref.flat_horizon_elevation(float(n[0])))
    boxes = np.zeros(img.shape[:2], bool)
    for x, y, text in labels:
        boxes[max(y, 0):max(y + 15, 0), max(x, 0):max(x + 8 * len(text) + 2, 0)] = True
    assert len(labels) >= 6 and boxes.mean() < 0.2
    np.testing.assert_array_equal(img[~boxes], want[~boxes])
    drawn = (want != plain).any(axis=2)
    assert (want[drawn] == [255, 128, 255]).all(axis=1).sum() >= 200  # the eye-level line crosses the picture
    assert (want[drawn] == [0, 128, 255]).all(axis=1).sum() >= 200    # so does the flat-earth horizon, 1.3 degrees above it
    assert ((img != want).any(axis=2) & boxes).any()                   # and the labels left glyphs
    # the Python mirror of output_image draws the same picture from the same configuration
    mirror = tmp_path / "mirror.png"
    runtime.output_image(plain.copy(), str(mirror), cfg, ctx)
    np.testing.assert_array_equal(host.read_png(str(mirror))[..., :3], img)


def test_example_configuration_parses_the_same_in_both_hosts():
    """examples/panorama.yaml: every key of the schema, through the C++ YAML-subset parser and through PyYAML."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "panorama.yaml")
    p_cpp, objs, folder, out, meta = host.parse_config(["-c", path])
    cfg = config.read_config(["-c", path])
    _same_params(p_cpp, config.into_params(cfg))
    assert (len(objs), folder, out, meta) == (2, "./terrain", "./panorama.png", "")
    from atm_raytracer_b200 import runtime

    ticks, vticks, eye, flat = host.parse_overlays(["-c", path])
    assert (len(ticks), len(vticks), eye, flat) == (3, 1, True, False)
    assert [t["size"] for t in runtime.overlay_ticks(cfg["output"]["ticks"], "azimuth")] == [12, 5, 18]


@pytest.mark.parametrize("size,block,threads", [(0, 0, 0), (1, 0, 0), (70_000, 65536, 3), (65536 * 4, 65536, 2), (9_000_001, 1 << 20, 0), (40_000_000, 0, 0),
                                                (3_000_000, 65536, 1)])
def test_parallel_gzip_reads_back_as_one_stream(tmp_path, size, block, threads):
    """The sidecar's writer (csrc/host/pgzip.cpp): a series of gzip members compressed on several threads is, to any gzip reader,
    the stream of the bytes that went in -- sizes around the block edges, more blocks than one wave holds, an empty input."""
    rng = np.random.default_rng(size % 1000)
    data = (np.cumsum(rng.integers(-3, 4, size, dtype=np.int64)) & 0xff).astype(np.uint8).tobytes()  # compressible, not trivial
    path = tmp_path / "x.gz"
    host.gzip_write(str(path), data, block, threads)
    raw = path.read_bytes()
    assert gzip.decompress(raw) == data
    assert raw[:2] == b"\x1f\x8b" and (size < 1000 or len(raw) < size)
    r = subprocess.run(["gzip", "-dc", str(path)], capture_output=True)
    assert r.returncode == 0 and r.stdout == data


def test_sidecar_reader_on_hand_made_files(tmp_path):
    """atm_raytracer_b200/sidecar.py against the layout write_metadata documents (versions 2 and 3), without a GPU."""
    from atm_raytracer_b200 import sidecar

    h, w = 3, 4
    el, az = np.linspace(1, -1, h), np.linspace(10, 13, w)
    first = np.zeros((h, w), sidecar.META_DTYPE)
    first["distance"] = np.arange(12).reshape(h, w)
    first["distance"][0, 0] = np.nan
    body = np.array([w, h, 0, 0], "<i4").tobytes() + el.tobytes() + az.tobytes() + first.tobytes()
    p2 = tmp_path / "v2.meta"
    p2.write_bytes(gzip.compress(b"ATMRTMETA2\n" + body))
    s = sidecar.read_sidecar(str(p2))
    assert (s.version, s.generator, s.counts) == (2, "Fast", None) and s.elevation_angle.shape == (h, w) and s.azimuth[2, 3] == 13.0
    assert len(s.trace_points(0, 0)) == 0 and s.trace_points(1, 1)["distance"][0] == 5.0
    counts = np.array([[0, 1, 2, 3], [0, 0, 0, 0], [5, 0, 0, 1]], "<i4")
    kept = np.minimum(counts, 2)
    pts = np.zeros(int(kept.sum()), sidecar.TRACE_DTYPE)
    pts["distance"] = np.arange(len(pts))
    p3 = tmp_path / "v3.meta"
    p3.write_bytes(gzip.compress(b"ATMRTMETA3\n" + body + np.array([2], "<i4").tobytes() + counts.tobytes() + pts.tobytes()))
    s = sidecar.read_sidecar(str(p3))
    assert (s.version, s.max_points) == (3, 2)
    assert [len(s.trace_points(y, x)) for y in range(h) for x in range(w)] == kept.ravel().tolist()
    assert s.trace_points(2, 0)["distance"].tolist() == [5.0, 6.0]
    per = np.array([w, h, 1, 0], "<i4").tobytes() + np.zeros((h, w)).tobytes() + np.ones((h, w)).tobytes() + first.tobytes()
    p4 = tmp_path / "rect.meta"
    p4.write_bytes(gzip.compress(b"ATMRTMETA2\n" + per))
    assert sidecar.read_sidecar(str(p4)).generator == "Rectilinear" and sidecar.read_sidecar(str(p4)).azimuth.sum() == h * w
    bad = tmp_path / "bad.meta"
    bad.write_bytes(gzip.compress(b"ATMRTMETA2\n" + body + b"x"))
    with pytest.raises(ValueError):
        sidecar.read_sidecar(str(bad))
