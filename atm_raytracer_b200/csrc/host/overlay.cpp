// overlay.cpp -- the overlays renderer::output_image draws over the finished picture (renderer/mod.rs:22-365, 416-431;
// SURVEY section 8 f4): azimuth / elevation ticks with their labels, the eye-level line and the flat-earth "horizon" line.
// Host work in the reference too (it runs after draw_image on the ResultPixel angles); nothing here touches the GPU.
// The line rasteriser is imageproc's draw_line_segment_mut (BresenhamLineIter; the crate is not vendored under
// /root/reference -- restated from its published source). Labels: the reference rasterises DejaVuSans.ttf through
// rusttype at 15 px with anti-aliasing; this host carries no TrueType rasteriser, it draws the same strings at the same
// anchor with a built-in 7x10 bitmap face of the same advance (8 px per digit), so label PIXELS differ from the reference's
// while tick positions, label texts and every line are the reference's.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "atmrt_host.h"

namespace atmrt_host {
int fail(int code, const std::string& msg);
}

namespace {

struct Image {
    uint8_t* rgb;
    int w, h;
    void put(int x, int y, const uint8_t c[3]) const {
        if (x >= 0 && x < w && y >= 0 && y < h) memcpy(rgb + ((size_t)y * w + x) * 3, c, 3);
    }
};

// imageproc::drawing::draw_line_segment_mut: Bresenham on f32 end points truncated to i32, end point included,
// pixels outside the canvas skipped.
void draw_line_segment(const Image& img, float x0, float y0, float x1, float y1, const uint8_t c[3]) {
    const bool steep = std::fabs(y1 - y0) > std::fabs(x1 - x0);
    if (steep) std::swap(x0, y0), std::swap(x1, y1);
    if (x0 > x1) std::swap(x0, x1), std::swap(y0, y1);
    const float dx = x1 - x0, dy = std::fabs(y1 - y0);
    float error = dx / 2.0f;
    const int end_x = (int)x1, y_step = y0 < y1 ? 1 : -1;
    int y = (int)y0;
    for (int x = (int)x0; x <= end_x; ++x) {
        if (steep) img.put(y, x, c); else img.put(x, y, c);
        error -= dy;
        if (error < 0.0f) y += y_step, error += dx;
    }
}

// 7x10 bitmap face for the characters format!("{:.N}") of an angle can hold.
struct Glyph { char ch; int advance; const char* rows[10]; };
const Glyph GLYPHS[] = {
    {'0', 8, {"..###..", ".#...#.", "#.....#", "#.....#", "#.....#", "#.....#", "#.....#", "#.....#", ".#...#.", "..###.."}},
    {'1', 8, {"...#...", "..##...", ".#.#...", "...#...", "...#...", "...#...", "...#...", "...#...", "...#...", ".#####."}},
    {'2', 8, {".####..", "#....#.", ".....#.", ".....#.", "....#..", "...#...", "..#....", ".#.....", "#......", "######."}},
    {'3', 8, {".####..", "#....#.", ".....#.", ".....#.", "..###..", ".....#.", "......#", "......#", "#....#.", ".####.."}},
    {'4', 8, {"....##.", "...#.#.", "..#..#.", ".#...#.", "#....#.", "#....#.", "#######", ".....#.", ".....#.", ".....#."}},
    {'5', 8, {"######.", "#......", "#......", "#####..", ".....#.", "......#", "......#", "......#", "#....#.", ".####.."}},
    {'6', 8, {"..###..", ".#.....", "#......", "#.###..", "##...#.", "#.....#", "#.....#", "#.....#", ".#...#.", "..###.."}},
    {'7', 8, {"#######", "......#", ".....#.", ".....#.", "....#..", "....#..", "...#...", "...#...", "..#....", "..#...."}},
    {'8', 8, {"..###..", ".#...#.", "#.....#", ".#...#.", "..###..", ".#...#.", "#.....#", "#.....#", ".#...#.", "..###.."}},
    {'9', 8, {"..###..", ".#...#.", "#.....#", "#.....#", "#.....#", ".#...##", "..###.#", "......#", ".....#.", ".###..."}},
    {'.', 4, {"...", "...", "...", "...", "...", "...", "...", "...", ".##", ".##"}},
    {'-', 5, {"....", "....", "....", "....", "....", "....", "####", "....", "....", "...."}},
};

// draw_text_mut(img, white, x, y, Scale 15, font, text): (x, y) is the top-left corner of the layout box; the digits of
// DejaVu Sans at that scale stand on a base line 12 px below it and are 9.4 px tall: rows y + 2 .. y + 11 here.
void draw_text(const Image& img, int x, int y, const std::string& text, const uint8_t c[3]) {
    int pen = x + 1;
    for (char ch : text) {
        const Glyph* g = nullptr;
        for (const Glyph& cand : GLYPHS)
            if (cand.ch == ch) g = &cand;
        if (!g) {  // NaN / inf: nothing to draw, keep the advance
            pen += 8;
            continue;
        }
        for (int r = 0; r < 10; ++r)
            for (int q = 0; g->rows[r][q]; ++q)
                if (g->rows[r][q] == '#') img.put(pen + q, y + 2 + r, c);
        pen += g->advance;
    }
}

// renderer/mod.rs:28-37
double diff_azimuth(double az1, double az2) {
    const double diff = az1 - az2;
    if (diff < -180.0) return diff + 360.0;
    if (diff > 180.0) return diff - 360.0;
    return diff;
}

// The reference reads pixels[0][x].azimuth and pixels[y][0].elevation_angle; with the flat [H][W] planes of
// atmrt_pixel_angles that is az[x] and el[y * W].
struct Angles {
    const double* el;
    const double* az;
    int w, h;
    double azimuth0(int x) const { return az[x]; }
    double elevation0(int y) const { return el[(size_t)y * w]; }
};

// Iterator::min_by keeps the FIRST of equal minima; a NaN key makes partial_cmp().unwrap() panic in the reference
// (the generators never produce one).
// renderer/mod.rs:39-59
bool azimuth_to_x(double azimuth, const Angles& a, uint32_t* out) {
    if (a.w < 2) return false;  // the reference indexes the neighbour unconditionally: a one-column picture panics there
    int cand = 0;
    double best = std::fabs(diff_azimuth(azimuth, a.azimuth0(0)));
    for (int x = 1; x < a.w; ++x) {
        const double d = std::fabs(diff_azimuth(azimuth, a.azimuth0(x)));
        if (d < best) best = d, cand = x;
    }
    const int nb = cand == 0 ? 1 : cand - 1;
    const double per_pixel = std::fabs(diff_azimuth(a.azimuth0(cand), a.azimuth0(nb)));
    if (!(std::fabs(diff_azimuth(a.azimuth0(cand), azimuth)) < per_pixel * 1.5)) return false;
    *out = (uint32_t)cand;
    return true;
}

// renderer/mod.rs:61-81
bool elevation_to_y(double elevation, const Angles& a, uint32_t* out) {
    if (a.h < 2) return false;
    int cand = 0;
    double best = std::fabs(elevation - a.elevation0(0));
    for (int y = 1; y < a.h; ++y) {
        const double d = std::fabs(elevation - a.elevation0(y));
        if (d < best) best = d, cand = y;
    }
    const int nb = cand == 0 ? 1 : cand - 1;
    const double per_pixel = std::fabs(a.elevation0(cand) - a.elevation0(nb));
    if (!(std::fabs(a.elevation0(cand) - elevation) < per_pixel * 1.5)) return false;
    *out = (uint32_t)cand;
    return true;
}

// renderer/mod.rs:206-214; f64::powi(10.0, i) is exact for i < 10, f64::round is C's round()
int num_decimals(double x) {
    for (int i = 0; i < 10; ++i) {
        double p = 1.0;
        for (int k = 0; k < i; ++k) p *= 10.0;
        const double mul_x = x * p;
        if (std::fabs(std::round(mul_x) - mul_x) < 0.001) return i;
    }
    return 10;
}

// TickLike::angle (params.rs:347-352, 379-384): the azimuth / elevation of a Single, the STEP of a Multiple
double tick_angle(const atmrt_host_tick& t) { return t.multiple ? t.step : t.angle; }

// renderer/mod.rs:216-223
int round_decimals(const atmrt_host_tick* ticks, int n) {
    int best = 0;
    for (int i = 0; i < n; ++i)
        if (ticks[i].labelled) best = std::max(best, num_decimals(tick_angle(ticks[i])));
    return best;
}

struct DrawTick {
    uint32_t size;
    std::string angle;
    bool labelled;
};

std::string fixed(double v, int decimals) {  // format!("{:.1$}", v, decimals): exact decimal expansion, ties to even -- as glibc's %.*f
    char buf[512];
    snprintf(buf, sizeof buf, "%.*f", decimals, v);
    return buf;
}

// renderer/mod.rs:83-140
void into_draw_ticks(const atmrt_host_tick& t, const atmrt_host_overlays& ov, const Angles& a, int decimals,
                     std::vector<std::pair<uint32_t, DrawTick>>* out) {
    uint32_t x;
    if (!t.multiple) {
        if (azimuth_to_x(t.angle, a, &x)) out->push_back({x, DrawTick{t.size, fixed(t.angle, decimals), t.labelled != 0}});
        return;
    }
    const double min_az = ov.direction - ov.fov / 2.0, max_az = ov.direction + ov.fov / 2.0;
    double current_az = std::ceil((min_az - t.bias) / t.step) * t.step + t.bias;
    if (!(t.step > 0.0)) return;  // the reference loops for ever on a step <= 0; a NaN step draws nothing there either
    while (current_az < max_az) {
        const double azimuth = current_az < 0.0 ? current_az + 360.0 : current_az >= 360.0 ? current_az - 360.0 : current_az;
        if (azimuth_to_x(current_az, a, &x)) out->push_back({x, DrawTick{t.size, fixed(azimuth, decimals), t.labelled != 0}});
        current_az += t.step;
    }
}

// renderer/mod.rs:142-199
void into_draw_ticks_vertical(const atmrt_host_tick& t, const atmrt_host_overlays& ov, const Angles& a, int decimals,
                              std::vector<std::pair<uint32_t, DrawTick>>* out) {
    uint32_t y;
    if (!t.multiple) {
        if (elevation_to_y(t.angle, a, &y)) out->push_back({y, DrawTick{t.size, fixed(t.angle, decimals), t.labelled != 0}});
        return;
    }
    const double aspect = (double)a.h / (double)a.w;
    const double min_elev = ov.tilt - ov.fov * aspect / 2.0, max_elev = ov.tilt + ov.fov * aspect / 2.0;
    double current_elev = std::ceil((min_elev - t.bias) / t.step) * t.step + t.bias;
    if (!(t.step > 0.0)) return;
    while (current_elev < max_elev) {
        const double elevation = current_elev < -90.0 ? -180.0 - current_elev : current_elev > 90.0 ? 180.0 - current_elev : current_elev;
        if (elevation_to_y(elevation, a, &y)) out->push_back({y, DrawTick{t.size, fixed(elevation, decimals), t.labelled != 0}});
        current_elev += t.step;
    }
}

// renderer/mod.rs:225-266: one tick per pixel position, a LARGER one replaces the one that is there
void merge_ticks(std::map<uint32_t, DrawTick>* into, const std::vector<std::pair<uint32_t, DrawTick>>& ticks) {
    for (const auto& pt : ticks) {
        auto it = into->find(pt.first);
        if (it == into->end())
            into->emplace(pt.first, pt.second);
        else if (it->second.size < pt.second.size)
            it->second = pt.second;
    }
}

void gen_ticks(const atmrt_host_overlays& ov, const Angles& a, std::map<uint32_t, DrawTick>* horizontal, std::map<uint32_t, DrawTick>* vertical) {
    const int hd = round_decimals(ov.ticks, ov.nticks), vd = round_decimals(ov.vertical_ticks, ov.nvertical_ticks);
    for (int i = 0; i < ov.nticks; ++i) {
        std::vector<std::pair<uint32_t, DrawTick>> t;
        into_draw_ticks(ov.ticks[i], ov, a, hd, &t);
        merge_ticks(horizontal, t);
    }
    for (int i = 0; i < ov.nvertical_ticks; ++i) {
        std::vector<std::pair<uint32_t, DrawTick>> t;
        into_draw_ticks_vertical(ov.vertical_ticks[i], ov, a, vd, &t);
        merge_ticks(vertical, t);
    }
}

// renderer/mod.rs:268-326 (the reference walks two HashMaps in arbitrary order: every pixel it writes is white)
void draw_ticks(const Image& img, const atmrt_host_overlays& ov, const Angles& a) {
    const uint8_t white[3] = {255, 255, 255};
    std::map<uint32_t, DrawTick> horizontal, vertical;
    gen_ticks(ov, a, &horizontal, &vertical);
    for (const auto& xt : horizontal) {
        draw_line_segment(img, (float)xt.first, 0.0f, (float)xt.first, (float)xt.second.size, white);
        if (xt.second.labelled) draw_text(img, (int)xt.first - 8, (int)xt.second.size + 5, xt.second.angle, white);
    }
    for (const auto& yt : vertical) {
        draw_line_segment(img, 0.0f, (float)yt.first, (float)yt.second.size, (float)yt.first, white);
        if (yt.second.labelled) draw_text(img, (int)yt.second.size + 5, (int)yt.first - 7, yt.second.angle, white);
    }
}

// find_elev for every column at once (renderer/mod.rs:328-346): rows in the reference's order, strict "<" keeps the first
// closest row; walking the [H][W] plane row by row instead of column by column reads it once, in memory order.
void find_elev_all(const Angles& a, double elev, std::vector<int64_t>* found) {
    const size_t w = (size_t)a.w;
    std::vector<double> closest(w, std::numeric_limits<double>::infinity());
    std::vector<int32_t> idx(w, 0);
    for (int y = 0; y < a.h; ++y) {
        const double* row = a.el + (size_t)y * w;
        for (size_t x = 0; x < w; ++x)
            if (std::fabs(row[x] - elev) < std::fabs(closest[x] - elev)) closest[x] = row[x], idx[x] = y;
    }
    found->assign(w, -1);
    if (a.h < 2) return;  // the reference indexes the neighbouring row unconditionally
    for (size_t x = 0; x < w; ++x) {
        const int nb = idx[x] == 0 ? 1 : idx[x] - 1;
        const double nb_elev = a.el[(size_t)nb * w + x];
        if (std::fabs(closest[x] - elev) < std::fabs(nb_elev - closest[x]) * 1.5) (*found)[x] = idx[x];
    }
}

// renderer/mod.rs:348-367
void draw_const_elev(const Image& img, const Angles& a, double elev, const uint8_t color[3]) {
    std::vector<int64_t> y;
    find_elev_all(a, elev, &y);
    for (int x = 1; x < a.w; ++x)
        if (y[(size_t)x - 1] >= 0 && y[(size_t)x] >= 0) draw_line_segment(img, (float)(x - 1), (float)y[(size_t)x - 1], (float)x, (float)y[(size_t)x], color);
}

}  // namespace

// gen_ticks (renderer/mod.rs:225-266) as data: the ticks draw_ticks would draw, by ascending pixel position.
extern "C" int atmrt_host_gen_ticks(int width, int height, const double* elevation_angle, const double* azimuth, const atmrt_host_overlays* ov,
                                    int vertical, atmrt_host_draw_tick* out, int capacity, int* n) {
    if (!ov || !n || width < 1 || height < 1 || !elevation_angle || !azimuth || (capacity > 0 && !out))
        return atmrt_host::fail(ATMRT_ERR_INVALID, "gen_ticks: bad argument");
    const Angles a{elevation_angle, azimuth, width, height};
    std::map<uint32_t, DrawTick> horizontal, vert;
    gen_ticks(*ov, a, &horizontal, &vert);
    const std::map<uint32_t, DrawTick>& m = vertical ? vert : horizontal;
    *n = (int)m.size();
    int i = 0;
    for (const auto& pt : m) {
        if (i >= capacity) break;
        out[i].position = pt.first, out[i].size = pt.second.size, out[i].labelled = pt.second.labelled;
        snprintf(out[i].label, sizeof out[i].label, "%s", pt.second.angle.c_str());
        ++i;
    }
    return 0;
}

extern "C" int atmrt_host_num_decimals(double x) { return num_decimals(x); }

extern "C" int atmrt_host_flat_horizon_elevation(double n_at_observer, double* elevation_deg) {
    if (!elevation_deg) return atmrt_host::fail(ATMRT_ERR_INVALID, "flat_horizon_elevation: NULL argument");
    *elevation_deg = std::acos(1.0 / n_at_observer) * (180.0 / M_PI);  // (1.0 / n).acos().to_degrees(), renderer/mod.rs:424-425
    return 0;
}

// renderer/mod.rs:416-431, between draw_image and img.save
extern "C" int atmrt_host_draw_overlays(uint8_t* rgb, int width, int height, const double* elevation_angle, const double* azimuth,
                                        const atmrt_host_overlays* ov) {
    if (!rgb || !ov || width < 1 || height < 1) return atmrt_host::fail(ATMRT_ERR_INVALID, "draw_overlays: bad argument");
    if ((ov->nticks > 0 && !ov->ticks) || (ov->nvertical_ticks > 0 && !ov->vertical_ticks) || ov->nticks < 0 || ov->nvertical_ticks < 0)
        return atmrt_host::fail(ATMRT_ERR_INVALID, "draw_overlays: tick list missing");
    const bool any = ov->nticks || ov->nvertical_ticks || ov->show_flat_horizon || ov->show_eye_level;
    if (!any) return 0;
    if (!elevation_angle || !azimuth) return atmrt_host::fail(ATMRT_ERR_INVALID, "draw_overlays: the ResultPixel angles are required");
    const Image img{rgb, width, height};
    const Angles a{elevation_angle, azimuth, width, height};
    draw_ticks(img, *ov, a);
    if (ov->show_flat_horizon) {
        const uint8_t c[3] = {0, 128, 255};
        draw_const_elev(img, a, ov->flat_horizon_elevation, c);
    }
    if (ov->show_eye_level) {
        const uint8_t c[3] = {255, 128, 255};
        draw_const_elev(img, a, 0.0, c);
    }
    return 0;
}
