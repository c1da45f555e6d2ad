// device_shade.cuh -- scene objects, colouring and front-to-back compositing on the device.
//
// Restates object/frustum.rs, object/billboard.rs, object/mod.rs:89-118 (texture fetch),
// coloring/shading.rs, coloring/simple.rs and renderer/mod.rs:367-414 (fog/add/draw_image), including
// the u8 re-quantisation after every compositing term.
#pragma once

#include "device_math.cuh"

namespace atmrt {

struct Color4 {
    double r, g, b, a;
};

struct DevObject {
    int kind;
    int tex_w, tex_h;
    int _pad;
    double lat, lon, elev;  // position with the altitude resolved (Altitude::abs)
    double r1, r2, width, height;
    double close_r;  // max(r1, r2) for a frustum, width for a billboard (is_close)
    Color4 color;
    V3 pos;  // as_cartesian(position)
    V3 up;   // world_directions(position).2
    const uint8_t* tex;  // RGBA8, row 0 = top
};

struct DevShade {
    int coloring, palette;
    int fog_enabled, _pad;
    double water_level, ambient_light;
    double light[3];
    double simple_max_distance;
    double fog_distance;
    double terrain_alpha;
    unsigned char def_color[4];  // fog colour if fog is enabled, else sky colour
    int _pad2;
};

// x / 255.0 and the palette's x / (thr_b - thr_a): IEEE divisions by a constant, evaluated as div_by (device_math.cuh:
// Markstein's correctly rounded quotient from RN(1/b), 3 instructions) -- the same bits as the division the reference
// performs, without its slow-path bookkeeping.
__device__ __forceinline__ double over_255(double x) { return div_by(x, 255.0, 1.0 / 255.0); }

struct Collision {
    double prop;
    V3 normal;
    Color4 color;
};

// Image::get_pixel, object/mod.rs:89-118
__device__ inline void texture_get_pixel(const DevObject& o, double x, double y, unsigned char out[4]) {
    double w = (double)o.tex_w, h = (double)o.tex_h;
    x = x * w - 0.5;
    double x1 = floor(x);
    x1 = x1 < 0.0 ? 0.0 : (x1 > w - 2.0 ? w - 2.0 : x1);
    double x2 = x1 + 1.0;
    int ix1 = (int)x1, ix2 = (int)x2;
    y = (1.0 - y) * h - 0.5;
    double y1 = floor(y);
    y1 = y1 < 0.0 ? 0.0 : (y1 > h - 2.0 ? h - 2.0 : y1);
    double y2 = y1 + 1.0;
    int iy1 = (int)y1, iy2 = (int)y2;
    double px = x - x1, py = y - y1;
    for (int c = 0; c < 4; ++c) {
        double p00 = over_255((double)o.tex[((size_t)iy1 * o.tex_w + ix1) * 4 + c]);
        double p01 = over_255((double)o.tex[((size_t)iy2 * o.tex_w + ix1) * 4 + c]);
        double p10 = over_255((double)o.tex[((size_t)iy1 * o.tex_w + ix2) * 4 + c]);
        double p11 = over_255((double)o.tex[((size_t)iy2 * o.tex_w + ix2) * 4 + c]);
        double v = p00 * (1.0 - px) * (1.0 - py) + p01 * (1.0 - px) * py + p10 * px * (1.0 - py) + p11 * px * py;
        out[c] = as_u8(v * 255.0);
    }
}

__device__ inline void sort_collisions(Collision* c, int n) {  // stable insertion sort by prop
    for (int i = 1; i < n; ++i) {
        Collision key = c[i];
        int j = i - 1;
        while (j >= 0 && key.prop < c[j].prop) {
            c[j + 1] = c[j];
            --j;
        }
        c[j + 1] = key;
    }
}

// Frustum::check_collision, frustum.rs:18-101. Writes <= 4 collisions sorted by prop.
__device__ inline int frustum_collision(const DevObject& o, V3 pos1, V3 pos2, Collision* results) {
    int nres = 0;
    V3 p1 = pos1 - o.pos;
    double p1sq = dot(p1, p1);
    V3 v = o.up;
    V3 w = pos2 - pos1;
    double wsq = dot(w, w), p1v = dot(p1, v), p1w = dot(p1, w), wv = dot(w, v);
    double aa = (o.r2 - o.r1) / o.height;
    double aa1 = 1.0 + aa * aa;
    double a = wsq - wv * wv * (1.0 + aa * aa);
    double b = 2.0 * (p1w - wv * (p1v * aa1 + aa * o.r1));
    double c = p1sq - p1v * p1v * aa1 - o.r1 * o.r1 - 2.0 * aa * o.r1 * p1v;
    double delta = b * b - 4.0 * a * c;
    if (delta >= 0.0) {
        double x1 = (-b - sqrt(delta)) / 2.0 / a;
        double x2 = (-b + sqrt(delta)) / 2.0 / a;
        if (a < 0.0) {
            double t = x1;
            x1 = x2;
            x2 = t;
        }
        double tmp[2];
        int nt = 0;
        if (in_range(0.0, 1.0, x1)) tmp[nt++] = x1;
        if (in_range(0.0, 1.0, x2)) tmp[nt++] = x2;
        for (int i = 0; i < nt; ++i) {
            double x = tmp[i];
            V3 intersection = p1 + w * x;
            double h = dot(intersection, v);
            if (!in_range(0.0, o.height, h)) continue;
            V3 outward = intersection - h * v;
            double o_len = sqrt(dot(outward, outward));
            outward = outward / o_len;
            double ang = atan2(o.r1 - o.r2, o.height);
            V3 normal = outward * cos(ang) + v * sin(ang);
            results[nres++] = {x, normal, o.color};
        }
    }
    for (int i = 0; i < 2; ++i) {
        double hh = i == 0 ? 0.0 : o.height;
        double rr = i == 0 ? o.r1 : o.r2;
        V3 nn = i == 0 ? -v : v;
        double x = (hh - p1v) / wv;
        V3 out = p1 + w * x - hh * v;
        double d = dot(out, out);
        if (d < rr * rr && in_range(0.0, 1.0, x)) results[nres++] = {x, nn, o.color};
    }
    sort_collisions(results, nres);
    return nres;
}

// Billboard::check_collision, billboard.rs:17-66. Writes <= 1 collision.
__device__ inline int billboard_collision(const DevObject& o, V3 pos1, V3 pos2, Collision* results) {
    V3 ray = pos2 - pos1;
    V3 up = o.up;
    V3 right = cross(ray, up);
    double right_len = sqrt(dot(right, right));
    right = right / right_len;
    V3 front = cross(right, up);
    V3 p1 = pos1 - o.pos;
    double prop = -dot(p1, front) / dot(ray, front);
    if (!in_range(0.0, 1.0, prop)) return 0;
    V3 intersection = p1 + ray * prop;
    double y = dot(intersection, up), x = dot(intersection, right);
    if (!in_range(0.0, o.height, y) || !in_range(-o.width / 2.0, o.width / 2.0, x)) return 0;
    x = (x + o.width / 2.0) / o.width;
    y = y / o.height;
    unsigned char px[4];
    texture_get_pixel(o, x, y, px);
    results[0] = {prop, front, Color4{over_255((double)px[0]), over_255((double)px[1]), over_255((double)px[2]), over_255((double)px[3])}};
    return 1;
}

// ---- colouring -------------------------------------------------------------------------------
__device__ inline V3 elev_to_color(int palette, double elev) {  // shading.rs:30-83
    const double thr1 = 300.0, thr3 = 1800.0, thr4 = 3000.0;
    double thr2;
    V3 c0, c1, c2, c3;
    if (palette == ATMRT_PALETTE_LEGACY) {
        thr2 = 1200.0;
        c0 = {0.0, 1.0, 0.0};
        c1 = {0.6, 1.0, 0.0};
        c2 = {0.5, 0.5, 0.5};
        c3 = {1.0, 1.0, 1.0};
    } else {
        thr2 = 1000.0;
        c0 = {0.4, 0.8, 0.3};
        c1 = {0.77, 0.84, 0.4};
        c2 = {0.41, 0.52, 0.4};
        c3 = {0.85, 0.92, 0.95};
    }
    const bool legacy = palette == ATMRT_PALETTE_LEGACY;
    if (elev < thr1) return c0;
    if (elev < thr2) {  // (elev - thr1) / (thr2 - thr1)
        double prop = legacy ? div_by(elev - thr1, 900.0, 1.0 / 900.0) : div_by(elev - thr1, 700.0, 1.0 / 700.0);
        return c1 * prop + c0 * (1.0 - prop);
    }
    if (elev < thr3) {  // (elev - thr2) / (thr3 - thr2)
        double prop = legacy ? div_by(elev - thr2, 600.0, 1.0 / 600.0) : div_by(elev - thr2, 800.0, 1.0 / 800.0);
        return c2 * prop + c1 * (1.0 - prop);
    }
    if (elev < thr4) {  // (elev - thr3) / (thr4 - thr3)
        double prop = div_by(elev - thr3, 1200.0, 1.0 / 1200.0);
        return c3 * prop + c2 * (1.0 - prop);
    }
    return c3;
}

struct Rgb8 {
    unsigned char c[3];
};

__device__ inline Rgb8 hsv(double h, double s, double v) {  // simple.rs:55-87
    double c = v * s;
    h = fmod(h, 360.0) < 0.0 ? fmod(h, 360.0) + 360.0 : fmod(h, 360.0);
    double x = c * (1.0 - fabs(fmod(h / 60.0, 2.0) - 1.0));
    double m = v - c;
    double rp = 0, gp = 0, bp = 0;
    if (in_range(0.0, 60.0, h)) {
        rp = c, gp = x, bp = 0.0;
    } else if (in_range(60.0, 120.0, h)) {
        rp = x, gp = c, bp = 0.0;
    } else if (in_range(120.0, 180.0, h)) {
        rp = 0.0, gp = c, bp = x;
    } else if (in_range(180.0, 240.0, h)) {
        rp = 0.0, gp = x, bp = c;
    } else if (in_range(240.0, 300.0, h)) {
        rp = x, gp = 0.0, bp = c;
    } else if (in_range(300.0, 360.0, h)) {
        rp = c, gp = 0.0, bp = x;
    }
    return {{as_u8((rp + m) * 255.0), as_u8((gp + m) * 255.0), as_u8((bp + m) * 255.0)}};
}

// ColoringMethod::color_for_pixel for one trace point.
__device__ inline Rgb8 color_for_pixel(const DevShade& p, bool is_terrain, double elevation, double distance, V3 normal, Color4 color) {
    if (p.coloring == ATMRT_COLORING_SIMPLE) {  // simple.rs:22-44
        double dist_ratio = distance / p.simple_max_distance;
        if (elevation <= p.water_level) {
            double mul = 1.0 - dist_ratio * 0.6;
            return {{0, as_u8(128.0 * mul), as_u8(255.0 * mul)}};
        }
        double elev_ratio = elevation / 4500.0;
        double h = 120.0 - 240.0 * (elev_ratio < 0.0 ? -pow(-elev_ratio, 0.65) : pow(elev_ratio, 0.65));
        double v = (elev_ratio > 0.7 ? 2.1 - elev_ratio * 2.0 : 0.9 - elev_ratio / 0.7 * 0.2) * (1.0 - dist_ratio * 0.6);
        double s = 1.0 - dist_ratio * 0.9;
        return hsv(h, s, v);
    }
    // shading.rs:108-132
    V3 light{p.light[0], p.light[1], p.light[2]};
    double light_dot = dot(light, normal);
    light_dot = light_dot >= 0.0 ? light_dot : 0.0;
    double brightness = p.ambient_light + (1.0 - p.ambient_light) * light_dot * light_dot;
    V3 base;
    if (!is_terrain)
        base = {color.r, color.g, color.b};
    else if (elevation <= p.water_level)
        base = p.palette == ATMRT_PALETTE_LEGACY ? V3{0.0, 0.5, 1.0} : V3{0.23, 0.41, 0.55};
    else
        base = elev_to_color(p.palette, elevation);
    V3 c = base * brightness;
    return {{as_u8(c.x * 255.0), as_u8(c.y * 255.0), as_u8(c.z * 255.0)}};
}

__device__ inline Rgb8 apply_fog(double fog_dist, double pixel_dist, Rgb8 color) {  // renderer/mod.rs:367-376
    double fog_coeff = 1.0 - exp(-pixel_dist / fog_dist);
    Rgb8 out;
    for (int i = 0; i < 3; ++i) out.c[i] = as_u8((double)color.c[i] * (1.0 - fog_coeff) + 160.0 * fog_coeff);
    return out;
}

__device__ inline Rgb8 add_rgb(Rgb8 rgb1, Rgb8 rgb2, double a) {  // renderer/mod.rs:378-383
    Rgb8 out;
    for (int i = 0; i < 3; ++i) {
        double c1 = over_255((double)rgb1.c[i]), c2 = over_255((double)rgb2.c[i]);
        out.c[i] = as_u8((c1 + c2 * a) * 255.0);
    }
    return out;
}

}  // namespace atmrt
