"""The reference's three text dumpers (ray_path.rs, elev_profile.rs, atm_printer.rs): the oracle's
restatement pins the text layout (Rust's `{}` of an f64, tabs, row and column structure) without a GPU; on the GPU the
`atm-raytracer` executable's subcommands must print the same layout with the same numbers -- byte for byte where the
arithmetic is exact (distances, altitudes, temperatures), within the parity tolerances elsewhere."""
import subprocess

import numpy as np
import pytest

from atm_raytracer_b200 import abi, config, host, synth
from test_noise_floor import PATH_ATOL

CONF = """view:
  position: {latitude: 45.4, longitude: 5.9, altitude: {Relative: 30.0}}
earth_shape: {Spherical: {radius: 6371000.0}}
"""


def test_rust_display_formatting_through_output_atm(oracle_lib):
    """`{}` of an f64: shortest round-trip digits, no exponent, no trailing '.0' (the accumulated altitudes show it)."""
    text = oracle_lib.output_atm(abi.us_76(), 0.0, 1.0, 0.2)
    lines = text.splitlines()
    assert lines[0] == "0 288.15 101325 0"
    alts = [ln.split(" ")[0] for ln in lines]
    assert alts == ["0", "0.2", "0.4", "0.6000000000000001", "0.8", "1"]  # alt += 0.2, printed like Rust prints it
    assert all(len(ln.split(" ")) == 4 for ln in lines)
    t = [float(ln.split(" ")[1]) for ln in lines]
    assert t[1] == 288.15 + (-0.0065) * (0.2 - 0.0)
    # celsius: T - 273.15 (atm_printer.rs:41); 11 km and above: the isothermal function of US-76
    c = oracle_lib.output_atm(abi.us_76(), 11000.0, 11001.0, 0.5, celsius=True).splitlines()
    assert [ln.split(" ")[0] for ln in c] == ["11000", "11000.5", "11001"]
    assert abs(float(c[0].split(" ")[1]) - (216.65 - 273.15)) < 1e-9 and abs(float(c[0].split(" ")[2]) - 22632.0) < 2.0


def test_output_ray_paths_layout(oracle_lib):
    p = abi.Params()
    p.earth_model, p.radius, p.wavelength, p.atmosphere = abi.EARTH_SPHERICAL, 6371000.0, 530e-9, abi.us_76()
    text = oracle_lib.output_ray_paths(p, height=2.0, min_ang=-0.2, max_ang=0.2, step=0.1, ray_step=50.0, cutoff=2000.0, output_step=100.0)
    rows = [ln.split("\t") for ln in text.splitlines()]
    # ang = -0.2; ang += 0.1 while ang <= 0.2 gives -0.2, -0.1, ~0, 0.1, 0.2 (the last one survives the rounding here)
    nang = len(rows[0]) - 2
    assert nang in (4, 5) and all(len(r) == nang + 2 and r[-1] == "" for r in rows)  # every field is followed by a tab
    assert rows[0][0] == "0" and all(v == "2" for v in rows[0][1:-1])
    x = np.array([float(r[0]) for r in rows])
    assert len(x) == 21 and np.allclose(np.diff(x), 100.0, atol=1e-6)  # every second 50 m state is printed
    h = np.array([[float(v) for v in r[1:-1]] for r in rows])
    assert (np.diff(h[-1]) > 0).all()  # a higher elevation angle ends higher
    flat = np.tan(np.radians(-0.2 + 0.1 * np.arange(nang))) * 2000.0 + 2.0
    assert np.allclose(h[-1], flat + 2000.0**2 / (2 * 6371000.0) * (1 - 0.17), atol=0.05)  # earth curvature minus refraction (k ~ 0.17)


def test_output_elev_profile_layout(oracle_lib):
    from atm_raytracer_b200 import runtime

    p = abi.Params()
    p.earth_model, p.radius, p.latitude, p.longitude = abi.EARTH_SPHERICAL, 6371000.0, 45.5, 5.5
    posts = synth.make_tile(45, 5, 0)
    terrain = runtime.Terrain.from_arrays([(45, 5, posts)])
    text = oracle_lib.output_elev_profile(p, terrain.tiles, azim=90.0, step=250.0, cutoff=1000.0)
    lines = text.splitlines()
    assert lines[0] == "Detected 1 terrain files"  # Terrain::from_folder prints to the same stdout (terrain/mod.rs:80)
    assert [ln.split("\t")[0] for ln in lines[1:]] == ["0", "250", "500", "750", "1000"]
    e0 = float(lines[1].split("\t")[1])
    # x = 0 is the observer's position after a round trip through the walker's asin / atan2
    assert abs(e0 - oracle_lib.get_elev(terrain.tiles, np.array([45.5]), np.array([5.5]))[0]) < 1e-6


def test_executable_dumpers_fail_loudly(tmp_path):
    r = subprocess.run([host.EXECUTABLE, "output-atm"], capture_output=True, text=True)
    assert r.returncode == 1 and "please provide an input file" in r.stderr
    conf = tmp_path / "c.yaml"
    conf.write_text(CONF)
    r = subprocess.run([host.EXECUTABLE, "output-ray-paths", str(conf), "-s", "0"], capture_output=True, text=True)
    assert r.returncode == 1 and "step must be positive" in r.stderr  # assert!(step > 0.0), ray_path.rs:53
    r = subprocess.run([host.EXECUTABLE, "output-elev-profile", str(conf), "--bogus", "1"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("ERROR: ")


def _no_gpu():
    import torch

    return not torch.cuda.is_available()


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_executable_dumpers_have_no_cpu_fallback(tmp_path):
    conf = tmp_path / "c.yaml"
    conf.write_text(CONF)
    r = subprocess.run([host.EXECUTABLE, "output-atm", str(conf)], capture_output=True, text=True)
    assert r.returncode == 1 and "atmrt_create" in r.stderr and r.stdout == ""


# ---------------------------------------------------------------------------------------------
# GPU: the executable against the oracle's text
# ---------------------------------------------------------------------------------------------
def _run(args):
    r = subprocess.run([host.EXECUTABLE] + [str(a) for a in args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


@pytest.mark.gpu
def test_output_atm_matches_oracle_text(tmp_path, oracle_lib):
    conf = tmp_path / "c.yaml"
    conf.write_text(CONF + "atmosphere:\n  pressure: {altitude: 0.0, pressure: 100800.0}\n  temperature_fixed_point: {altitude: 0.0, temperature: 291.0}\n"
                    "  humidity: 0.6\n  first_temperature_function: {Linear: {gradient: -0.004}}\n"
                    "  next_functions:\n    - {altitude: 2050.0, function: {Linear: {gradient: 0.05}}}\n    - {altitude: 2090.0, function: {Linear: {gradient: -0.0065}}}\n")
    p = config.into_params(config.read_config(["-c", str(conf)]))
    for extra, kw in ((["-a", "1900", "-b", "2300", "-s", "12.5"], dict(min_alt=1900.0, max_alt=2300.0, step=12.5)),
                      (["-c"], dict(celsius=True)), (["--min-alt", "-100", "--max-alt", "10", "--step", "0.3", "--celsius"], dict(min_alt=-100.0, max_alt=10.0, step=0.3, celsius=True))):
        got = _run(["output-atm", conf] + extra).stdout
        want = oracle_lib.output_atm(p.atmosphere, **kw)
        g, w = [ln.split(" ") for ln in got.splitlines()], [ln.split(" ") for ln in want.splitlines()]
        assert len(g) == len(w) > 10 and all(len(r) == 4 for r in g)
        assert [r[0] for r in g] == [r[0] for r in w]  # altitudes: the same running sum, the same text
        assert [r[1] for r in g] == [r[1] for r in w]  # temperatures: a multiply-add, bit-exact (test_atmosphere_probe)
        assert [r[3] for r in g] == [r[3] for r in w] and g[0][3] == "0.6"
        np.testing.assert_allclose([float(r[2]) for r in g], [float(r[2]) for r in w], rtol=1e-14)  # pow / exp: libm ulps


@pytest.mark.gpu
def test_output_ray_paths_matches_oracle_text(tmp_path, oracle_lib):
    for shape, extra in (("{Spherical: {radius: 6371000.0}}", []), ("FlatDistorted", ["-h", "150", "-a", "-0.5", "-b", "0.75", "-s", "0.25", "-r", "25", "-c", "30000", "-o", "500"])):
        conf = tmp_path / "c.yaml"
        conf.write_text(f"earth_shape: {shape}\nwavelength: 600e-9\n")
        p = config.into_params(config.read_config(["-c", str(conf)]))
        kw = dict(height=150.0, min_ang=-0.5, max_ang=0.75, step=0.25, ray_step=25.0, cutoff=30000.0, output_step=500.0) if extra else {}
        r = _run(["output-ray-paths", conf] + extra)
        assert r.stderr.startswith("Elevation angle ")  # eprintln!, ray_path.rs:67
        want = oracle_lib.output_ray_paths(p, **kw)
        g, w = [ln.split("\t") for ln in r.stdout.splitlines()], [ln.split("\t") for ln in want.splitlines()]
        assert len(g) == len(w) > 50 and [len(x) for x in g] == [len(x) for x in w]
        assert [x[0] for x in g] == [x[0] for x in w]  # RayState::x: the same running sum on both sides, the same text
        assert g[0] == w[0]                            # the observer's height
        gh, wh = np.array([[float(v) for v in x[1:-1]] for x in g]), np.array([[float(v) for v in x[1:-1]] for x in w])
        np.testing.assert_allclose(gh, wh, rtol=1e-9, atol=PATH_ATOL)  # the noise floor of the reference's own dn/dh


@pytest.mark.gpu
def test_output_elev_profile_matches_oracle_text(tmp_path, oracle_lib):
    from atm_raytracer_b200 import runtime

    folder = tmp_path / "terrain"
    folder.mkdir()
    synth.write_tile_grid(str(folder), 45, 5, 1, 2, level=0)
    for shape in ("{Spherical: {radius: 6371000.0}}", "FlatDistorted", "Wgs84"):
        conf = tmp_path / "c.yaml"
        conf.write_text(f"scene:\n  terrain_folder: {folder}\nview:\n  position: {{latitude: 45.4, longitude: 5.9}}\nearth_shape: {shape}\n")
        p = config.into_params(config.read_config(["-c", str(conf)]))
        terrain = runtime.Terrain.from_folder(str(folder))
        got = _run(["output-elev-profile", conf, "-a", "77.5", "-s", "130", "-c", "90000"]).stdout
        want = oracle_lib.output_elev_profile(p, terrain.tiles, azim=77.5, step=130.0, cutoff=90000.0)
        g, w = got.splitlines(), want.splitlines()
        assert g[0] == w[0] == "Detected 2 terrain files" and len(g) == len(w) == 2 + int(90000 / 130)
        gx, wx = [ln.split("\t") for ln in g[1:]], [ln.split("\t") for ln in w[1:]]
        assert [x[0] for x in gx] == [x[0] for x in wx]  # distances: the same running sum, the same text
        ge, we = np.array([float(x[1]) for x in gx]), np.array([float(x[1]) for x in wx])
        np.testing.assert_allclose(ge, we, rtol=0, atol=1e-7)  # bilinear of i16 posts at coordinates that differ by libm ulps
        assert (ge != 0.0).mean() > 0.5 and (ge == we).mean() > 0.2
