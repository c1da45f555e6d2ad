"""The overlays of renderer::output_image (renderer/mod.rs:22-365, 416-431): the host's
C++ (csrc/host/overlay.cpp, through the C ABI of libatmrt_host.so) against the Python restatement in oracle/overlays.py, and
both against the reference's own known answers for this path (`test_decimals`, renderer/mod.rs:438-459). CPU only."""
import os

import numpy as np
import pytest

from atm_raytracer_b200 import host
from oracle import overlays as ref

# renderer/mod.rs:443-458 -- the reference's own test vector, verbatim values
REFERENCE_TEST_DECIMALS = [(0.0, 0), (1.0, 0), (15.0, 0), (183.0, 0), (0.1, 1), (0.3, 1), (0.9, 1), (1.8, 1), (12.6, 1), (133.5, 1),
                           (0.25, 2), (33.99, 2), (33.01, 2), (133.01002, 5)]


@pytest.mark.parametrize("x,want", REFERENCE_TEST_DECIMALS)
def test_num_decimals_reference_vector(x, want):
    assert ref.num_decimals(x) == want
    assert host.num_decimals(x) == want


def test_num_decimals_agree_on_random_inputs():
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-400, 400, 300), np.round(rng.uniform(-400, 400, 300), 3), np.round(rng.uniform(0, 90, 300), 1), [1e-12, 1 / 3]])
    for x in xs:
        assert host.num_decimals(x) == ref.num_decimals(float(x)), x


def fast_angles(width, height, direction, fov, tilt):
    """ResultPixel angles of the Fast generator (fast.rs:67-76, 114-125): one azimuth per column wrapped once into [0, 360),
    one elevation per row."""
    x = np.arange(width) - width // 2
    y = np.arange(height) - height // 2
    az = direction + x * fov / width
    az = np.where(az < 0.0, az + 360.0, np.where(az >= 360.0, az - 360.0, az))
    el = tilt - y * fov / width
    return np.repeat(el[:, None], width, 1), np.repeat(az[None, :], height, 0)


def rectilinear_angles(width, height, direction, fov, tilt):
    """Angles that differ in every pixel (as the Rectilinear generator's do): find_elev then finds a curved line."""
    f = (width / 2) / np.tan(np.radians(fov / 2))
    x = (np.arange(width) - width / 2)[None, :]
    y = (np.arange(height) - height / 2)[:, None]
    t = np.radians(tilt)
    vx, vy, vz = x + 0 * y, f * np.cos(t) + y * np.sin(t), f * np.sin(t) - y * np.cos(t)
    el = np.degrees(np.arctan2(vz, np.hypot(vx, vy)))
    az = direction + np.degrees(np.arctan2(vx, vy))
    return el, az


TICKS = [dict(kind="Multiple", bias=0.0, step=10.0, size=10, labelled=True), dict(kind="Multiple", bias=0.0, step=2.0, size=5, labelled=False),
         dict(kind="Single", angle=45.0, size=15, labelled=True), dict(kind="Single", angle=200.0, size=15, labelled=True)]
VTICKS = [dict(kind="Multiple", bias=0.5, step=2.5, size=8, labelled=True), dict(kind="Single", angle=0.0, size=20, labelled=False),
          dict(kind="Single", angle=80.0, size=20, labelled=True)]

SCENES = {
    "readme": dict(shape=(120, 160), frame=dict(direction=40.0, fov=30.0, tilt=0.0), ticks=TICKS, vticks=VTICKS, angles=fast_angles),
    "wrap_north": dict(shape=(90, 200), frame=dict(direction=355.0, fov=40.0, tilt=-2.0), ticks=TICKS, vticks=VTICKS, angles=fast_angles),
    "wrap_negative": dict(shape=(64, 128), frame=dict(direction=3.0, fov=20.0, tilt=5.0),
                          ticks=[dict(kind="Multiple", bias=0.25, step=1.25, size=6, labelled=True)],
                          vticks=[dict(kind="Multiple", bias=0.0, step=0.5, size=4, labelled=True)], angles=fast_angles),
    "rectilinear": dict(shape=(100, 150), frame=dict(direction=120.0, fov=60.0, tilt=10.0), ticks=TICKS[:2], vticks=VTICKS[:2], angles=rectilinear_angles),
    "odd_sizes": dict(shape=(33, 47), frame=dict(direction=180.0, fov=90.0, tilt=-30.0),
                      ticks=[dict(kind="Multiple", bias=0.0, step=15.0, size=40, labelled=False)],
                      vticks=[dict(kind="Multiple", bias=0.0, step=15.0, size=60, labelled=False)], angles=fast_angles),
}


def label_boxes(shape, labels):
    """Pixels a label may touch: the layout box draw_text_mut gets (15 px tall, 8 px per character and one to spare)."""
    mask = np.zeros(shape, dtype=bool)
    for x, y, text in labels:
        y0, y1 = max(y, 0), max(min(y + 15, shape[0]), 0)
        x0, x1 = max(x, 0), max(min(x + 8 * len(text) + 2, shape[1]), 0)
        mask[y0:y1, x0:x1] = True
    return mask


@pytest.mark.parametrize("name", sorted(SCENES))
def test_ticks_match_the_restatement(name):
    s = SCENES[name]
    h, w = s["shape"]
    el, az = s["angles"](w, h, **s["frame"])
    hor, ver = ref.gen_ticks(s["ticks"], s["vticks"], s["frame"], el.tolist(), az.tolist())
    got_h = host.gen_ticks(el, az, s["ticks"], s["vticks"], s["frame"], vertical=False)
    got_v = host.gen_ticks(el, az, s["ticks"], s["vticks"], s["frame"], vertical=True)
    assert got_h == [(x, t["size"], t["labelled"], t["angle"]) for x, t in sorted(hor.items())]
    assert got_v == [(y, t["size"], t["labelled"], t["angle"]) for y, t in sorted(ver.items())]
    assert got_h or got_v


@pytest.mark.parametrize("name", sorted(SCENES))
def test_overlay_pixels_match_the_restatement(name):
    s = SCENES[name]
    h, w = s["shape"]
    el, az = s["angles"](w, h, **s["frame"])
    rng = np.random.default_rng(11)
    base = rng.integers(0, 200, (h, w, 3), dtype=np.uint8)
    flat = 0.35 if name in ("readme", "rectilinear") else None
    want = base.copy()
    labels = ref.output_overlays(want, el.tolist(), az.tolist(), s["ticks"], s["vticks"], s["frame"], show_eye_level=True, flat_horizon_elev=flat)
    got = host.draw_overlays(base.copy(), el, az, s["ticks"], s["vticks"], s["frame"], show_eye_level=True, flat_horizon_elev=flat)
    boxes = label_boxes((h, w), labels)
    # every line (ticks, eye level, flat horizon) pixel for pixel; label glyphs are the one thing that may differ, inside their boxes
    # (a line drawn AFTER the labels -- eye level, flat horizon -- is the reference's there too, the restatement draws no glyphs)
    assert np.array_equal(got[~boxes], want[~boxes])
    changed = (got != want).any(axis=2)
    assert not changed[~boxes].any()
    if labels:
        assert changed.any(), "labelled ticks drew no glyph"
        assert (got[changed] == 255).all(), "glyphs are white"
    else:
        assert np.array_equal(got, want)
    assert (want != base).any()


def test_larger_tick_wins_and_labels_carry_the_common_precision():
    el, az = fast_angles(160, 120, 40.0, 30.0, 0.0)
    ticks = [dict(kind="Multiple", bias=0.0, step=2.5, size=5, labelled=True), dict(kind="Multiple", bias=0.0, step=10.0, size=10, labelled=True)]
    got = host.gen_ticks(el, az, ticks, [], dict(direction=40.0, fov=30.0, tilt=0.0))
    by_label = {t[3]: t for t in got}
    assert by_label["30.0"][1] == 10 and by_label["40.0"][1] == 10 and by_label["32.5"][1] == 5  # one decimal everywhere (step 2.5)
    assert all(t[2] for t in got)


def test_line_rasteriser_known_answers():
    """imageproc's BresenhamLineIter, worked by hand: a tick (vertical, end point included), a 45 degree line, a shallow one."""
    img = np.zeros((8, 8, 3), np.uint8)
    el = np.repeat(np.linspace(3.5, -3.5, 8)[:, None], 8, 1)
    az = np.repeat((100.0 + np.arange(8))[None, :], 8, 0)
    host.draw_overlays(img, el, az, [dict(kind="Single", angle=103.0, size=4, labelled=False)], [], dict(direction=104.0, fov=8.0, tilt=0.0))
    ys, xs = np.nonzero(img[:, :, 0])
    assert sorted(zip(xs.tolist(), ys.tolist())) == [(3, y) for y in range(5)]  # (3, 0) .. (3, 4): size + 1 pixels
    a = np.zeros((6, 6, 3), np.uint8)
    ref.draw_line_segment(a, (0, 0), (5, 5), (1, 1, 1))
    assert np.array_equal(np.nonzero(a[:, :, 0]), (np.arange(6), np.arange(6)))
    b = np.zeros((4, 8, 3), np.uint8)
    ref.draw_line_segment(b, (0, 0), (7, 2), (1, 1, 1))
    assert [int(np.nonzero(b[:, x, 0])[0][0]) for x in range(8)] == [0, 0, 1, 1, 1, 1, 2, 2]  # error 3.5 -> 1.5 -> -0.5 (step) -> 6.5 .. 0.5 -> -1.5 (step)


def test_eye_level_off_the_picture_draws_nothing():
    el, az = fast_angles(64, 48, 90.0, 20.0, 30.0)  # the picture looks 30 degrees up: elevation 0 is far below the last row
    img = np.zeros((48, 64, 3), np.uint8)
    host.draw_overlays(img, el, az, [], [], dict(direction=90.0, fov=20.0, tilt=30.0), show_eye_level=True)
    assert not img.any()
    want = np.zeros_like(img)
    ref.output_overlays(want, el.tolist(), az.tolist(), [], [], dict(direction=90.0, fov=20.0, tilt=30.0), show_eye_level=True)
    assert not want.any()


def test_flat_horizon_elevation():
    n = 1.000277
    assert host.flat_horizon_elevation(n) == pytest.approx(ref.flat_horizon_elevation(n), rel=0, abs=1e-15)
    assert host.flat_horizon_elevation(n) == pytest.approx(np.degrees(np.sqrt(2 * (n - 1))), rel=1e-3)  # acos(1/n) ~ sqrt(2 (n - 1))


OVERLAY_YAML = """
view:
  frame: {direction: 40, fov: 30}
output:
  width: 160
  height: 120
  ticks:
    - Multiple:
        bias: 0
        step: 10
        size: 10
        labelled: true
    - Multiple: {bias: 0.5, step: 2, size: 5, labelled: false}
    - Single:
        azimuth: 45
        size: 15
        labelled: true
  vertical_ticks:
    - Single: {elevation: -1.5, size: 7, labelled: true}
  show_eye_level: true
  show_flat_horizon: false
"""


def test_yaml_overlay_keys(tmp_path):
    cfg = tmp_path / "o.yaml"
    cfg.write_text(OVERLAY_YAML)
    ticks, vticks, eye, flat = host.parse_overlays(["-c", str(cfg)])
    assert [(t["kind"], t["size"], t["labelled"]) for t in ticks] == [("Multiple", 10, True), ("Multiple", 5, False), ("Single", 15, True)]
    assert (ticks[0]["bias"], ticks[0]["step"], ticks[1]["bias"], ticks[1]["step"], ticks[2]["angle"]) == (0.0, 10.0, 0.5, 2.0, 45.0)
    assert [(t["kind"], t["angle"], t["size"], t["labelled"]) for t in vticks] == [("Single", -1.5, 7, True)]
    assert eye and not flat
    assert host.parse_overlays([]) == ([], [], False, False)  # Output::default (params.rs:432-444)
    # the Python mirror lowers the same YAML to the same ticks
    from atm_raytracer_b200 import config, runtime

    out = config.read_config(["-c", str(cfg)])["output"]
    strip = lambda ts: [{k: v for k, v in t.items() if not (k == "angle" and t["kind"] == "Multiple") and not (k in ("bias", "step") and t["kind"] == "Single")} for t in ts]
    assert runtime.overlay_ticks(out["ticks"], "azimuth") == strip(ticks)
    assert runtime.overlay_ticks(out["vertical_ticks"], "elevation") == strip(vticks)
    with pytest.raises(config.ConfigError):
        runtime.overlay_ticks([{"Single": {"azimuth": 1.0, "size": 3}}], "azimuth")


@pytest.mark.parametrize("bad", ["- Single: {azimuth: 1, size: 3}", "- Triple: {azimuth: 1, size: 3, labelled: true}",
                                 "- Multiple: {bias: 0, step: 1, size: -2, labelled: true}", "- Single: {elevation: 1, size: 3, labelled: true}"])
def test_yaml_tick_errors(tmp_path, bad):
    cfg = tmp_path / "bad.yaml"
    cfg.write_text("output:\n  ticks:\n    " + bad + "\n")
    with pytest.raises(host.HostError):
        host.parse_overlays(["-c", str(cfg)])


def test_zero_or_negative_step_terminates():
    """The reference never leaves `while current_az < max_az` with a step <= 0; this host draws nothing instead of hanging."""
    el, az = fast_angles(64, 48, 90.0, 20.0, 0.0)
    assert host.gen_ticks(el, az, [dict(kind="Multiple", bias=0.0, step=0.0, size=3, labelled=False)], [], dict(direction=90.0, fov=20.0, tilt=0.0)) == []
    assert host.gen_ticks(el, az, [dict(kind="Multiple", bias=0.0, step=-1.0, size=3, labelled=False)], [], dict(direction=90.0, fov=20.0, tilt=0.0)) == []
