"""TEST INFRASTRUCTURE -- CPU restatement of the overlays `renderer::output_image` draws over the picture
(renderer/mod.rs:22-365, 416-431). Pure-Python loops, for small pictures:
only tests/ may import this; the product (csrc/host/overlay.cpp) must never.

`pixels` are two nested sequences el[y][x], az[y][x] (ResultPixel.elevation_angle / .azimuth). The line rasteriser
restates imageproc's `draw_line_segment_mut` (BresenhamLineIter; external crate, not vendored with the reference;
f32 arithmetic through numpy.float32). Labels: the reference rasterises DejaVuSans.ttf through rusttype; neither this
restatement nor the product does, so a label here is DATA (anchor and text), not pixels.

Pinned on the reference's own known answers for this path: `test_decimals` (renderer/mod.rs:438-459), held in
tests/test_overlays.py.
"""
import math

import numpy as np

F = np.float32


def diff_azimuth(az1, az2):  # renderer/mod.rs:28-37
    diff = az1 - az2
    if diff < -180.0:
        return diff + 360.0
    if diff > 180.0:
        return diff - 360.0
    return diff


def _first_min(keys):  # Iterator::min_by: the first of equal minima
    best = 0
    for i, k in enumerate(keys):
        if k < keys[best]:
            best = i
    return best


def azimuth_to_x(azimuth, az):  # renderer/mod.rs:39-59
    row = az[0]
    cand = _first_min([abs(diff_azimuth(azimuth, a)) for a in row])
    nb = 1 if cand == 0 else cand - 1
    per_pixel = abs(diff_azimuth(row[cand], row[nb]))
    return cand if abs(diff_azimuth(row[cand], azimuth)) < per_pixel * 1.5 else None


def elevation_to_y(elevation, el):  # renderer/mod.rs:61-81
    col = [r[0] for r in el]
    cand = _first_min([abs(elevation - e) for e in col])
    nb = 1 if cand == 0 else cand - 1
    per_pixel = abs(col[cand] - col[nb])
    return cand if abs(col[cand] - elevation) < per_pixel * 1.5 else None


def rust_round(x):  # f64::round: half away from zero
    return math.copysign(math.floor(abs(x) + 0.5), x)


def num_decimals(x):  # renderer/mod.rs:206-214
    for i in range(10):
        mul_x = x * 10.0 ** i
        if abs(rust_round(mul_x) - mul_x) < 0.001:
            return i
    return 10


def tick_angle(t):  # TickLike::angle, params.rs:347-352 / 379-384
    return t["step"] if t["kind"] == "Multiple" else t["angle"]


def round_decimals(ticks):  # renderer/mod.rs:216-223
    return max([num_decimals(tick_angle(t)) for t in ticks if t["labelled"]], default=0)


def fmt(v, decimals):  # format!("{:.1$}", v, decimals)
    return "%.*f" % (decimals, v)


def into_draw_ticks(t, frame, az, decimals):  # renderer/mod.rs:83-140
    if t["kind"] == "Single":
        x = azimuth_to_x(t["angle"], az)
        return [] if x is None else [(x, dict(size=t["size"], labelled=t["labelled"], angle=fmt(t["angle"], decimals)))]
    min_az = frame["direction"] - frame["fov"] / 2.0
    max_az = frame["direction"] + frame["fov"] / 2.0
    cur = math.ceil((min_az - t["bias"]) / t["step"]) * t["step"] + t["bias"]
    out = []
    while cur < max_az:
        azimuth = cur + 360.0 if cur < 0.0 else cur - 360.0 if cur >= 360.0 else cur
        x = azimuth_to_x(cur, az)
        if x is not None:
            out.append((x, dict(size=t["size"], labelled=t["labelled"], angle=fmt(azimuth, decimals))))
        cur += t["step"]
    return out


def into_draw_ticks_vertical(t, frame, el, width, height, decimals):  # renderer/mod.rs:142-199
    if t["kind"] == "Single":
        y = elevation_to_y(t["angle"], el)
        return [] if y is None else [(y, dict(size=t["size"], labelled=t["labelled"], angle=fmt(t["angle"], decimals)))]
    aspect = float(height) / float(width)
    min_elev = frame["tilt"] - frame["fov"] * aspect / 2.0
    max_elev = frame["tilt"] + frame["fov"] * aspect / 2.0
    cur = math.ceil((min_elev - t["bias"]) / t["step"]) * t["step"] + t["bias"]
    out = []
    while cur < max_elev:
        elevation = -180.0 - cur if cur < -90.0 else 180.0 - cur if cur > 90.0 else cur
        y = elevation_to_y(elevation, el)
        if y is not None:
            out.append((y, dict(size=t["size"], labelled=t["labelled"], angle=fmt(elevation, decimals))))
        cur += t["step"]
    return out


def gen_ticks(ticks, vertical_ticks, frame, el, az):  # renderer/mod.rs:225-266
    height, width = len(el), len(el[0])
    horizontal, vertical = {}, {}
    hd, vd = round_decimals(ticks), round_decimals(vertical_ticks)
    for t in ticks:
        for x, d in into_draw_ticks(t, frame, az, hd):
            if x not in horizontal or horizontal[x]["size"] < d["size"]:
                horizontal[x] = d
    for t in vertical_ticks:
        for y, d in into_draw_ticks_vertical(t, frame, el, width, height, vd):
            if y not in vertical or vertical[y]["size"] < d["size"]:
                vertical[y] = d
    return horizontal, vertical


def draw_line_segment(img, start, end, color):
    """imageproc::drawing::draw_line_segment_mut: BresenhamLineIter over f32 end points, end point included, pixels off
    the canvas skipped."""
    h, w = img.shape[:2]
    x0, y0, x1, y1 = F(start[0]), F(start[1]), F(end[0]), F(end[1])
    steep = abs(F(y1 - y0)) > abs(F(x1 - x0))
    if steep:
        x0, y0, x1, y1 = y0, x0, y1, x1
    if x0 > x1:
        x0, x1, y0, y1 = x1, x0, y1, y0
    dx, dy = F(x1 - x0), abs(F(y1 - y0))
    error = F(dx / F(2.0))
    y_step = 1 if y0 < y1 else -1
    x, y, end_x = int(x0), int(y0), int(x1)
    while x <= end_x:
        px, py = (y, x) if steep else (x, y)
        if 0 <= px < w and 0 <= py < h:
            img[py, px] = color
        x += 1
        error = F(error - dy)
        if error < 0:
            y += y_step
            error = F(error + dx)


def find_elev(el, column, elev):  # renderer/mod.rs:328-346
    closest, idx = math.inf, 0
    for y, row in enumerate(el):
        if abs(row[column] - elev) < abs(closest - elev):
            closest, idx = row[column], y
    nb = 1 if idx == 0 else idx - 1
    return idx if abs(closest - elev) < abs(el[nb][column] - closest) * 1.5 else None


def draw_const_elev(img, el, elev, color):  # renderer/mod.rs:348-367
    width = len(el[0])
    y_old = find_elev(el, 0, elev)
    for x in range(1, width):
        y_new = find_elev(el, x, elev)
        if y_old is not None and y_new is not None:
            draw_line_segment(img, (x - 1, y_old), (x, y_new), color)
        y_old = y_new


def flat_horizon_elevation(n_at_observer):  # renderer/mod.rs:424-425
    return math.degrees(math.acos(1.0 / n_at_observer))


def output_overlays(img, el, az, ticks, vertical_ticks, frame, show_eye_level=False, flat_horizon_elev=None):
    """renderer/mod.rs:416-431 between draw_image and img.save, on img[H][W][3] (numpy uint8, in place). Returns the labels
    draw_ticks would hand to draw_text_mut: [(x, y, text)], anchors as in renderer/mod.rs:291-299, 310-318."""
    horizontal, vertical = gen_ticks(ticks, vertical_ticks, frame, el, az)
    labels = []
    white = (255, 255, 255)
    for x, t in horizontal.items():
        draw_line_segment(img, (x, 0.0), (x, t["size"]), white)
        if t["labelled"]:
            labels.append((x - 8, t["size"] + 5, t["angle"]))
    for y, t in vertical.items():
        draw_line_segment(img, (0.0, y), (t["size"], y), white)
        if t["labelled"]:
            labels.append((t["size"] + 5, y - 7, t["angle"]))
    if flat_horizon_elev is not None:
        draw_const_elev(img, el, flat_horizon_elev, (0, 128, 255))
    if show_eye_level:
        draw_const_elev(img, el, 0.0, (255, 128, 255))
    return labels
