// atmrt_oracle.cpp -- CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement (C++17, f64, OpenMP) of the reference's Fast-generator panorama path, written
// in the reference's own structure and operation order so that the CUDA product can be checked
// against it. Nothing in the product (atm_raytracer_b200/, include/) links, imports or calls this
// file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// PARITY STATUS -- read before trusting:
//   * The reference (fizyk20/atm-raytracer 0.13.0, Rust) cannot be built here (no cargo/rustc, no
//     network) and ships no tests, fixtures or golden vectors for this path.
//   * In-tree reference arithmetic (earth models, generators, objects, colouring, compositing) is
//     restated line by line; each function cites the file:line it follows.
//   * Two external crates carry arithmetic whose source is NOT under /root/reference:
//       atm-refraction "0.6" (Cargo.toml:8)  -- ray ODE + RK4 stepper, Ciddor n(h), US-76 atmosphere
//       dted           "0.2" (Cargo.toml:11) -- DTED parse + bilinear get_elev
//     They are restated from their published algorithms (Ciddor 1996; US Standard Atmosphere 1976;
//     MIL-PRF-89020B; Fermat-principle ray ODE) and from how the reference calls them.
//     ==> PARITY UNPINNED for those two pieces: they are anchored on analytic known answers
//     (tests/test_oracle_known_answers.py), not on outputs of the real crates.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off -fopenmp). -ffp-contract=off because Rust
// never contracts a*b+c into an FMA.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/atmrt.h"
#include "../include/atmrt_fmt.h"

namespace {

constexpr double PI = 3.14159265358979323846;
// f64::to_radians / to_degrees are one multiply by a precomputed constant.
inline double to_radians(double d) { return d * (PI / 180.0); }
inline double to_degrees(double r) { return r * (180.0 / PI); }

struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
// nalgebra dot for Vector3: x*x' + y*y' + z*z' accumulated left to right.
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// nalgebra normalize: divide every component by sqrt(dot) (no rsqrt).
inline V3 normalized(V3 a) {
    double n = std::sqrt(dot(a, a));
    return a / n;
}

// Rust `as` casts saturate and map NaN to 0.
inline uint8_t as_u8(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}
inline int16_t as_i16(double v) {
    if (!(v == v)) return 0;
    if (v <= -32768.0) return INT16_MIN;
    if (v >= 32767.0) return INT16_MAX;
    return (int16_t)v;
}
inline size_t as_usize(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 1.8446744073709552e19) return SIZE_MAX;
    return (size_t)v;
}
// (lo..hi).contains(&v): lo <= v < hi; NaN fails.
inline bool in_range(double lo, double hi, double v) { return lo <= v && v < hi; }

// ------------------------------------------------------------------------------------------
// Earth models: utils/earth_model/mod.rs, directional_calc.rs
// ------------------------------------------------------------------------------------------
constexpr double DEGREE_DISTANCE = 10000000.0 / 90.0;  // mod.rs:12

struct Coords {
    double lat, lon, elev;
};

struct EarthModel {
    int kind;       // atmrt_earth_model
    double radius;  // Spherical: radius; Ellipsoid: a; ObserverAe: proj_radius
    double b = 0.0; // Ellipsoid: b
};

// the family whose world is the azimuthal-equidistant plane (mod.rs:37-51, 80-91, 106-110)
inline bool flat_family(const EarthModel& m) {
    return m.kind == ATMRT_EARTH_FLAT_DISTORTED || m.kind == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT || m.kind == ATMRT_EARTH_OBSERVER_AE;
}

struct Dirs {
    V3 north, east, up;
};

// spherical_directions, mod.rs:155-172
Dirs spherical_directions(double lat, double lon) {
    double lat_rad = to_radians(lat), lon_rad = to_radians(lon);
    double sinlon = std::sin(lon_rad), coslon = std::cos(lon_rad);
    double sinlat = std::sin(lat_rad), coslat = std::cos(lat_rad);
    Dirs d;
    d.up = {coslat * coslon, coslat * sinlon, sinlat};
    d.north = {-sinlat * coslon, -sinlat * sinlon, coslat};
    d.east = {-sinlon, coslon, 0.0};
    return d;
}

// EarthModel::world_directions, mod.rs:31-57
Dirs world_directions(const EarthModel& m, double lat, double lon) {
    if (flat_family(m)) {
        double lon_rad = to_radians(lon);
        double sinlon = std::sin(lon_rad), coslon = std::cos(lon_rad);
        Dirs d;
        d.north = {-coslon, -sinlon, 0.0};
        d.east = {-sinlon, coslon, 0.0};
        d.up = {0.0, 0.0, 1.0};
        return d;
    }
    return spherical_directions(lat, lon);
}

// spherical_to_cartesian, mod.rs:148-153
V3 spherical_to_cartesian(double r, double lat, double lon) {
    double x = r * std::cos(to_radians(lat)) * std::cos(to_radians(lon));
    double y = r * std::cos(to_radians(lat)) * std::sin(to_radians(lon));
    double z = r * std::sin(to_radians(lat));
    return {x, y, z};
}

// EarthModel::as_cartesian, mod.rs:59-93
V3 as_cartesian(const EarthModel& m, const Coords& c) {
    if (flat_family(m)) {
        double z = c.elev;
        double r = (90.0 - c.lat) * DEGREE_DISTANCE;
        double x = r * std::cos(to_radians(c.lon));
        double y = r * std::sin(to_radians(c.lon));
        return {x, y, z};
    }
    if (m.kind == ATMRT_EARTH_ELLIPSOID) {  // mod.rs:70-79
        double a = m.radius, b = m.b;
        double e2 = 1.0 - (b * b) / (a * a);
        double lat = to_radians(c.lat), lon = to_radians(c.lon);
        double sl = std::sin(lat);
        double n = a / std::sqrt(1.0 - e2 * (sl * sl));
        double x = (n + c.elev) * std::cos(lat) * std::cos(lon);
        double y = (n + c.elev) * std::cos(lat) * std::sin(lon);
        double z = (n * (1.0 - e2) + c.elev) * std::sin(lat);
        return {x, y, z};
    }
    return spherical_to_cartesian(m.radius + c.elev, c.lat, c.lon);
}

// EarthModel::to_shape, mod.rs:95-112: the ray physics sees a plane or a sphere
inline bool shape_flat(const EarthModel& m) { return flat_family(m); }
inline double shape_radius(const EarthModel& m) { return m.kind == ATMRT_EARTH_ELLIPSOID ? (2.0 * m.radius + m.b) / 3.0 : m.radius; }

// DirectionalCalc (directional_calc.rs): SphericalCalc (:50-86), FlDsCalc (:30-48), AzEqCalc (:9-28),
// EllipsoidCalc (:88-185, Vincenty's direct formula after NGS "inverse.pdf")
enum { CALC_SPHERICAL = 0, CALC_FLDS = 1, CALC_AZEQ = 2, CALC_ELLIPSOID = 3 };
struct DirCalc {
    int kind;
    // spherical
    double radius;
    V3 pos, dir;  // AzEq: pos = start in the plane, dir = dir_v
    // flat distorted
    double start_lat, start_lon, dir_deg;
    // ellipsoid
    double eb, f, red_lat, lon, az1, alfa, sig1, cap_a, cap_b, cap_c;
};

DirCalc coords_at_dist_calc(const EarthModel& m, double lat, double lon, double dir_deg) {
    DirCalc c{};
    if (m.kind == ATMRT_EARTH_FLAT_DISTORTED) {
        c.kind = CALC_FLDS;
        c.start_lat = lat;
        c.start_lon = lon;
        c.dir_deg = dir_deg;
        return c;
    }
    if (m.kind == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT) {  // mod.rs:116-125
        c.kind = CALC_AZEQ;
        c.pos = as_cartesian(m, Coords{lat, lon, 0.0});
        Dirs d = world_directions(m, lat, lon);
        c.dir = d.north * std::cos(to_radians(dir_deg)) + d.east * std::sin(to_radians(dir_deg));
        return c;
    }
    if (m.kind == ATMRT_EARTH_ELLIPSOID) {  // EllipsoidCalc::new, directional_calc.rs:104-135
        c.kind = CALC_ELLIPSOID;
        double a = m.radius, b = m.b;
        double la = to_radians(lat), lo = to_radians(lon), az1 = to_radians(dir_deg);
        double f = (a - b) / a;
        double red_lat = std::atan((1.0 - f) * std::tan(la));
        double sig1 = std::atan(std::tan(red_lat) / std::cos(az1));
        double alfa = std::asin(std::cos(red_lat) * std::sin(az1));
        double ca = std::cos(alfa);
        double u2 = ca * ca * (a * a - b * b) / (b * b);
        c.cap_a = 1.0 + u2 / 256.0 * (64.0 + u2 * (-12.0 + 5.0 * u2));
        c.cap_b = u2 / 512.0 * (128.0 + u2 * (-64.0 + 37.0 * u2));
        c.cap_c = f / 16.0 * (ca * ca) * (4.0 + f * (4.0 - 3.0 * (ca * ca)));
        c.eb = b, c.f = f, c.red_lat = red_lat, c.lon = lo, c.az1 = az1, c.alfa = alfa, c.sig1 = sig1;
        return c;
    }
    // SphericalCalc::new, directional_calc.rs:56-69 (Spherical{radius}, ObserverAe{proj_radius}: mod.rs:127-131)
    c.kind = CALC_SPHERICAL;
    Dirs d = spherical_directions(lat, lon);
    double dir_rad = to_radians(dir_deg);
    double sindir = std::sin(dir_rad), cosdir = std::cos(dir_rad);
    c.radius = m.radius;
    c.pos = d.up;
    c.dir = d.north * cosdir + d.east * sindir;
    return c;
}

void coords_at_dist(const DirCalc& c, double dist, double* lat, double* lon) {
    if (c.kind == CALC_FLDS) {
        // FlDsCalc::coords_at_dist, directional_calc.rs:41-47
        double d_lat = std::cos(to_radians(c.dir_deg)) * dist / DEGREE_DISTANCE;
        double d_lon = std::sin(to_radians(c.dir_deg)) * dist / DEGREE_DISTANCE / std::cos(to_radians(c.start_lat));
        *lat = c.start_lat + d_lat;
        *lon = c.start_lon + d_lon;
        return;
    }
    if (c.kind == CALC_AZEQ) {
        // AzEqCalc::coords_at_dist, directional_calc.rs:20-27
        V3 pos2 = c.pos + c.dir * dist;
        *lon = to_degrees(std::atan2(pos2.y, pos2.x));
        double r = std::sqrt(pos2.x * pos2.x + pos2.y * pos2.y);
        *lat = 90.0 - r / DEGREE_DISTANCE;
        return;
    }
    if (c.kind == CALC_ELLIPSOID) {
        // EllipsoidCalc::coords_at_dist, directional_calc.rs:139-184
        const double EPSILON = 1e-10;
        double sig = dist / c.eb / c.cap_a;
        for (;;) {
            double sigm = 2.0 * c.sig1 + sig;
            double csm = std::cos(sigm);
            double dsig = c.cap_b * std::sin(sig) * (csm + c.cap_b / 4.0 * std::cos(sig) * (-1.0 + 2.0 * (csm * csm)));
            double new_sig = dist / c.eb / c.cap_a + dsig;
            double ds = std::fabs(new_sig - sig);
            sig = new_sig;
            if (ds < EPSILON) break;
            if (!(ds == ds)) break;  // NaN: the reference would spin; never reached with finite inputs
        }
        double sigm = 2.0 * c.sig1 + sig;
        double sr = std::sin(c.red_lat), cr = std::cos(c.red_lat), ss = std::sin(sig), cs = std::cos(sig);
        double caz = std::cos(c.az1), saz = std::sin(c.az1), sal = std::sin(c.alfa);
        double t = sr * ss - cr * cs * caz;
        double lat2 = std::atan((sr * cs + cr * ss * caz) / ((1.0 - c.f) * std::sqrt(sal * sal + t * t)));
        double lambda = std::atan(ss * saz / (cr * cs - sr * ss * caz));
        double csm = std::cos(sigm);
        double dl = lambda - (1.0 - c.cap_c) * c.f * sal * (sig + c.cap_c * ss * (csm + c.cap_c * cs * (-1.0 + 2.0 * (csm * csm))));
        *lat = to_degrees(lat2);
        *lon = to_degrees(c.lon + dl);
        return;
    }
    // SphericalCalc::coords_at_dist, directional_calc.rs:71-86
    double ang = dist / c.radius;
    double sinang = std::sin(ang), cosang = std::cos(ang);
    V3 fpos = c.pos * cosang + c.dir * sinang;
    *lat = to_degrees(std::asin(fpos.z));
    *lon = to_degrees(std::atan2(fpos.y, fpos.x));
}

// ------------------------------------------------------------------------------------------
// Terrain: terrain/mod.rs + external dted 0.2 (restated; PARITY UNPINNED -- SURVEY Appendix A.2)
// ------------------------------------------------------------------------------------------
struct Tile {
    atmrt_tile_desc d;
    const int16_t* posts;  // [lon][lat]
    double max_lat, max_lon;
};

struct Terrain {
    std::map<std::pair<int16_t, int16_t>, Tile> tiles;
};

// DtedData::get_elev (external; bilinear form witnessed by terrain/geotiff.rs:61-100)
bool tile_get_elev(const Tile& t, double lat, double lon, double* out) {
    double min_lat = t.d.min_lat, min_lon = t.d.min_lon;
    if (lat < min_lat || lat > t.max_lat || lon < min_lon || lon > t.max_lon) return false;
    double plat = (lat - min_lat) * 3600.0 / t.d.lat_interval;
    double plon = (lon - min_lon) * 3600.0 / t.d.lon_interval;
    size_t lat_int = as_usize(plat), lon_int = as_usize(plon);
    double lat_frac = plat - (double)lat_int, lon_frac = plon - (double)lon_int;
    if (lat_int == (size_t)(t.d.nlat - 1)) {  // max-edge fix-up (geotiff.rs:79-87)
        lat_int -= 1;
        lat_frac += 1.0;
    }
    if (lon_int == (size_t)(t.d.nlon - 1)) {
        lon_int -= 1;
        lon_frac += 1.0;
    }
    const int16_t* p = t.posts;
    size_t nlat = (size_t)t.d.nlat;
    double e00 = (double)p[lon_int * nlat + lat_int];
    double e01 = (double)p[lon_int * nlat + lat_int + 1];
    double e10 = (double)p[(lon_int + 1) * nlat + lat_int];
    double e11 = (double)p[(lon_int + 1) * nlat + lat_int + 1];
    *out = e00 * (1.0 - lon_frac) * (1.0 - lat_frac) + e01 * (1.0 - lon_frac) * lat_frac +
           e10 * lon_frac * (1.0 - lat_frac) + e11 * lon_frac * lat_frac;
    return true;
}

// Terrain::get_elev, terrain/mod.rs:120-126
bool terrain_get_elev(const Terrain& t, double latitude, double longitude, double* out) {
    int16_t lat = as_i16(std::floor(latitude));
    int16_t lon = as_i16(std::floor(longitude));
    auto it = t.tiles.find({lat, lon});
    if (it == t.tiles.end()) return false;
    return tile_get_elev(it->second, latitude, longitude, out);
}
inline double elev_or_zero(const Terrain& t, double lat, double lon) {  // .unwrap_or(0.0)
    double e;
    return terrain_get_elev(t, lat, lon, &e) ? e : 0.0;
}

Terrain make_terrain(const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts) {
    Terrain t;
    for (int i = 0; i < ntiles; ++i) {
        Tile tile;
        tile.d = tiles[i];
        tile.posts = posts[i];
        tile.max_lat = tiles[i].min_lat + (double)(tiles[i].nlat - 1) * tiles[i].lat_interval / 3600.0;
        tile.max_lon = tiles[i].min_lon + (double)(tiles[i].nlon - 1) * tiles[i].lon_interval / 3600.0;
        t.tiles[{(int16_t)tiles[i].lat0, (int16_t)tiles[i].lon0}] = tile;  // later insert wins (HashMap::insert)
    }
    return t;
}

// Altitude::abs, params.rs:23-30
double altitude_abs(const atmrt_altitude& a, const Terrain& t, double lat, double lon) {
    if (a.kind == ATMRT_ALT_ABSOLUTE) return a.value;
    return elev_or_zero(t, lat, lon) + a.value;
}

// ------------------------------------------------------------------------------------------
// Atmosphere + refractive index: external atm-refraction 0.6 (restated; PARITY UNPINNED --
// SURVEY Appendix A.1). US-76 hydrostatic layers; Ciddor (1996) refractive index.
// ------------------------------------------------------------------------------------------
constexpr double ATM_G = 9.80665;      // m/s^2
constexpr double ATM_M = 0.0289644;    // kg/mol
constexpr double ATM_R = 8.31432;      // J/(mol K): R* of the US Standard Atmosphere 1976, which reproduces its published p(h) table

struct AtmLayer {
    double start;   // lower boundary (-inf for layer 0)
    double h_ref, t_ref, p_ref;
    double gradient;
    // One segment of a Spline temperature function: T = k[0] + k[1] x + k[2] x^2 + k[3] x^3, x = h - origin.
    bool spline_segment;
    double origin, k[4];
};

constexpr int ORACLE_MAX_LAYERS = 32;  // Linear functions + Spline segments
struct Atmosphere {
    int n;
    AtmLayer layer[ORACLE_MAX_LAYERS];
    double humidity;
};

int atm_layer_index(const Atmosphere& a, double h) {
    int idx = 0;
    for (int i = 1; i < a.n; ++i)
        if (h >= a.layer[i].start) idx = i;
    return idx;
}

double segment_temperature(const AtmLayer& l, double h) {
    const double x = h - l.origin;
    return std::fma(std::fma(std::fma(l.k[3], x, l.k[2]), x, l.k[1]), x, l.k[0]);
}

double layer_temperature(const AtmLayer& l, double h) {
    if (l.spline_segment) return segment_temperature(l, h);
    return l.t_ref + l.gradient * (h - l.h_ref);
}

// Integral of dh / T from the segment's reference altitude to h with the 8-point Gauss-Legendre rule (include/atmrt.h:
// the documented quadrature of the hydrostatic integral inside a Spline function), nodes visited in ascending order.
double cubic_inverse_integral(const AtmLayer& l, double h) {
    static const double node[8] = {-0.9602898564975362316835609, -0.7966664774136267395915539, -0.5255324099163289858177390, -0.1834346424956498049394761,
                                   0.1834346424956498049394761,  0.5255324099163289858177390,  0.7966664774136267395915539,  0.9602898564975362316835609};
    static const double weight[8] = {0.1012285362903762591525314, 0.2223810344533744705443560, 0.3137066458778872873379622, 0.3626837833783619829651504,
                                     0.3626837833783619829651504, 0.3137066458778872873379622, 0.2223810344533744705443560, 0.1012285362903762591525314};
    const double half = 0.5 * (h - l.h_ref), mid = 0.5 * (h + l.h_ref);
    double sum = 0.0;
    for (int i = 0; i < 8; ++i) sum += weight[i] / segment_temperature(l, mid + half * node[i]);
    return sum * half;
}

double layer_pressure(const AtmLayer& l, double h) {
    if (l.spline_segment) return l.p_ref * std::exp(-ATM_G * ATM_M / ATM_R * cubic_inverse_integral(l, h));
    if (l.gradient != 0.0) {
        double t = layer_temperature(l, h);
        return l.p_ref * std::pow(t / l.t_ref, -ATM_G * ATM_M / (ATM_R * l.gradient));
    }
    return l.p_ref * std::exp(-ATM_G * ATM_M * (h - l.h_ref) / (ATM_R * l.t_ref));
}

// Atmosphere::from_def with at least one Spline function (README.md:296-323). The spline's second derivatives come
// from the classical forward-elimination / back-substitution recurrence (decomposition factors `u`, as in the textbook
// routine), the cubic of segment i is then
//   T = y_i + x ((y_{i+1} - y_i) / d - d (2 z_i + z_{i+1}) / 6) + x^2 z_i / 2 + x^3 (z_{i+1} - z_i) / (6 d),  d = x_{i+1} - x_i.
bool atmosphere_with_splines(const atmrt_atmosphere_def& def, Atmosphere* out) {
    const double inf = std::numeric_limits<double>::infinity();
    Atmosphere a{};
    a.humidity = def.humidity;
    int n = 0;
    for (int f = 0; f < def.n_functions; ++f) {
        if (f >= 2 && !(def.fn_start_altitude[f] > def.fn_start_altitude[f - 1])) return false;
        const double from = f == 0 ? -inf : def.fn_start_altitude[f];
        const double to = f + 1 < def.n_functions ? def.fn_start_altitude[f + 1] : inf;
        if (def.fn_kind[f] == ATMRT_FUNCTION_LINEAR) {
            if (n == ORACLE_MAX_LAYERS) return false;
            a.layer[n].start = from;
            a.layer[n].gradient = def.fn_gradient[f];
            ++n;
            continue;
        }
        if (def.fn_kind[f] != ATMRT_FUNCTION_SPLINE) return false;
        const int m = def.fn_n_points[f], p0 = def.fn_first_point[f];
        if (m < 2 || p0 < 0 || p0 + m > def.n_spline_points || def.n_spline_points > ATMRT_MAX_SPLINE_POINTS) return false;
        const double (*pt)[2] = def.spline_points + p0;
        for (int i = 1; i < m; ++i)
            if (!(pt[i][0] > pt[i - 1][0])) return false;
        std::vector<double> z(m, 0.0), u(m, 0.0);
        const int bc = def.fn_boundary[f];
        const double bc0 = def.fn_boundary_values[f][0], bc1 = def.fn_boundary_values[f][1];
        // row 0: z_0 = -1/2 z_1 + u_0 (Derivatives) or z_0 = value (Natural: 0)
        if (bc == ATMRT_SPLINE_DERIVATIVES) {
            z[0] = -0.5;
            u[0] = (3.0 / (pt[1][0] - pt[0][0])) * ((pt[1][1] - pt[0][1]) / (pt[1][0] - pt[0][0]) - bc0);
        } else if (bc == ATMRT_SPLINE_SECOND_DERIVATIVES) {
            u[0] = bc0;
        } else if (bc != ATMRT_SPLINE_NATURAL) {
            return false;
        }
        for (int i = 1; i + 1 < m; ++i) {
            const double sig = (pt[i][0] - pt[i - 1][0]) / (pt[i + 1][0] - pt[i - 1][0]);
            const double piv = sig * z[i - 1] + 2.0;
            z[i] = (sig - 1.0) / piv;
            const double dd = (pt[i + 1][1] - pt[i][1]) / (pt[i + 1][0] - pt[i][0]) - (pt[i][1] - pt[i - 1][1]) / (pt[i][0] - pt[i - 1][0]);
            u[i] = (6.0 * dd / (pt[i + 1][0] - pt[i - 1][0]) - sig * u[i - 1]) / piv;
        }
        double qn = 0.0, un = 0.0;  // last row: z_{m-1} = (un - qn u_{m-2}) / (qn z_{m-2} + 1)
        if (bc == ATMRT_SPLINE_DERIVATIVES) {
            const double d = pt[m - 1][0] - pt[m - 2][0];
            qn = 0.5;
            un = (3.0 / d) * (bc1 - (pt[m - 1][1] - pt[m - 2][1]) / d);
        } else if (bc == ATMRT_SPLINE_SECOND_DERIVATIVES) {
            un = bc1;
        }
        z[m - 1] = (un - qn * u[m - 2]) / (qn * z[m - 2] + 1.0);
        for (int i = m - 2; i >= 0; --i) z[i] = z[i] * z[i + 1] + u[i];
        for (int i = 0; i + 1 < m; ++i) {
            const double lo = i == 0 ? -inf : pt[i][0], hi = i + 2 == m ? inf : pt[i + 1][0];
            if (hi <= from || lo >= to) continue;
            if (n == ORACLE_MAX_LAYERS) return false;
            const double d = pt[i + 1][0] - pt[i][0];
            AtmLayer& l = a.layer[n++];
            l.start = lo > from ? lo : from;
            l.spline_segment = true;
            l.origin = pt[i][0];
            l.k[0] = pt[i][1];
            l.k[1] = (pt[i + 1][1] - pt[i][1]) / d - d * (2.0 * z[i] + z[i + 1]) / 6.0;
            l.k[2] = z[i] / 2.0;
            l.k[3] = (z[i + 1] - z[i]) / (6.0 * d);
        }
    }
    a.n = n;
    // Linear layers take their temperature from the nearest Spline by continuity: sweep up, then down.
    // A Linear layer is written as T = t_fix + gradient (h - h_fix); keep (h_fix, t_fix) in (h_ref, t_ref) for now.
    std::vector<bool> have(n);
    for (int i = 0; i < n; ++i) have[i] = a.layer[i].spline_segment;
    for (int i = 1; i < n; ++i)
        if (!have[i] && have[i - 1]) {
            a.layer[i].h_ref = a.layer[i].start;
            a.layer[i].t_ref = layer_temperature(a.layer[i - 1], a.layer[i].start);
            have[i] = true;
        }
    for (int i = n - 2; i >= 0; --i)
        if (!have[i] && have[i + 1]) {
            a.layer[i].h_ref = a.layer[i + 1].start;
            a.layer[i].t_ref = layer_temperature(a.layer[i + 1], a.layer[i + 1].start);
            have[i] = true;
        }
    // Pressure reference points, outwards from the pressure fixed point (as in the Linear-only lowering below).
    int jp = 0;
    for (int i = 1; i < n; ++i)
        if (def.pressure_altitude >= a.layer[i].start) jp = i;
    auto move_reference = [&](AtmLayer& l, double h) {  // after this, (h_ref, t_ref) is the pressure reference point
        const double t = layer_temperature(l, h);
        l.h_ref = h;
        l.t_ref = t;
    };
    move_reference(a.layer[jp], def.pressure_altitude);
    a.layer[jp].p_ref = def.pressure;
    for (int i = jp + 1; i < n; ++i) {
        move_reference(a.layer[i], a.layer[i].start);
        a.layer[i].p_ref = layer_pressure(a.layer[i - 1], a.layer[i].start);
    }
    for (int i = jp - 1; i >= 0; --i) {
        move_reference(a.layer[i], a.layer[i + 1].start);
        a.layer[i].p_ref = layer_pressure(a.layer[i + 1], a.layer[i + 1].start);
    }
    *out = a;
    return true;
}

// Atmosphere::from_def.
bool atmosphere_from_def(const atmrt_atmosphere_def& def, Atmosphere* out) {
    int n = def.n_functions;
    if (n < 1 || n > ATMRT_MAX_ATM_FUNCTIONS) return false;
    for (int i = 0; i < n; ++i)
        if (def.fn_kind[i] != ATMRT_FUNCTION_LINEAR) return atmosphere_with_splines(def, out);
    Atmosphere a{};
    a.n = n;
    a.humidity = def.humidity;
    for (int i = 0; i < n; ++i) {
        a.layer[i].start = i == 0 ? -std::numeric_limits<double>::infinity() : def.fn_start_altitude[i];
        a.layer[i].gradient = def.fn_gradient[i];
        if (i >= 2 && !(def.fn_start_altitude[i] > def.fn_start_altitude[i - 1])) return false;
    }
    // temperature: anchor the layer holding the fixed point, then walk outwards by continuity.
    auto find = [&](double h) {
        int idx = 0;
        for (int i = 1; i < n; ++i)
            if (h >= a.layer[i].start) idx = i;
        return idx;
    };
    int jt = find(def.temperature_altitude);
    std::vector<double> t_at_start(n, 0.0);  // temperature at layer[i].start, i>=1
    auto temp_in = [&](int i, double h_anchor, double t_anchor, double h) {
        return t_anchor + a.layer[i].gradient * (h - h_anchor);
    };
    // temperatures at boundaries above the fixed point
    {
        double h_anchor = def.temperature_altitude, t_anchor = def.temperature;
        for (int i = jt; i + 1 < n; ++i) {
            double tb = temp_in(i, h_anchor, t_anchor, a.layer[i + 1].start);
            t_at_start[i + 1] = tb;
            h_anchor = a.layer[i + 1].start;
            t_anchor = tb;
        }
        h_anchor = def.temperature_altitude;
        t_anchor = def.temperature;
        for (int i = jt; i >= 1; --i) {
            double tb = temp_in(i, h_anchor, t_anchor, a.layer[i].start);
            t_at_start[i] = tb;
            h_anchor = a.layer[i].start;
            t_anchor = tb;
        }
    }
    // Temperature anywhere (used once, to evaluate T at the pressure fixed point).
    auto temp_at = [&](double h) {
        int i = find(h);
        if (i == jt) return temp_in(i, def.temperature_altitude, def.temperature, h);
        if (i > jt) return temp_in(i, a.layer[i].start, t_at_start[i], h);
        return temp_in(i, a.layer[i + 1].start, t_at_start[i + 1], h);
    };
    // Reference point of every layer (altitude with known T and p): the pressure fixed point in
    // its own layer, the lower boundary for layers above it, the upper boundary for layers below.
    int jp = find(def.pressure_altitude);
    a.layer[jp].h_ref = def.pressure_altitude;
    a.layer[jp].t_ref = temp_at(def.pressure_altitude);
    a.layer[jp].p_ref = def.pressure;
    for (int i = jp + 1; i < n; ++i) {
        a.layer[i].h_ref = a.layer[i].start;
        a.layer[i].t_ref = t_at_start[i];
        a.layer[i].p_ref = layer_pressure(a.layer[i - 1], a.layer[i].start);
    }
    for (int i = jp - 1; i >= 0; --i) {
        a.layer[i].h_ref = a.layer[i + 1].start;
        a.layer[i].t_ref = t_at_start[i + 1];
        a.layer[i].p_ref = layer_pressure(a.layer[i + 1], a.layer[i + 1].start);
    }
    *out = a;
    return true;
}

double atm_temperature(const Atmosphere& a, double h) { return layer_temperature(a.layer[atm_layer_index(a, h)], h); }
double atm_pressure(const Atmosphere& a, double h) { return layer_pressure(a.layer[atm_layer_index(a, h)], h); }

// Ciddor (1996) refractive index of air; lambda in metres, p in Pa, t in K, rh in 0..1, 450 ppm CO2.
double air_index(double lambda, double p, double t_kelvin, double rh) {
    const double w0 = 295.235, w1 = 2.6422, w2 = -0.032380, w3 = 0.004028;
    const double k0 = 238.0185, k1 = 5792105.0, k2 = 57.362, k3 = 167917.0;
    const double a0 = 1.58123e-6, a1 = -2.9331e-8, a2 = 1.1043e-10;
    const double b0 = 5.707e-6, b1 = -2.051e-8;
    const double c0 = 1.9898e-4, c1 = -2.376e-6;
    const double d = 1.83e-11, e = -0.765e-8;
    const double p_r1 = 101325.0, t_r1 = 288.15;
    const double z_a = 0.9995922115;
    const double rho_vs = 0.00985938;
    const double gas_r = 8.314510, m_v = 0.018015;
    const double x_c = 450.0;
    const double alpha = 1.00062, beta = 3.14e-8, gamma = 5.6e-7;
    const double sa = 1.2378847e-5, sb = -1.9121316e-2, sc = 33.93711047, sd = -6.3431645e3;

    double lambda_um = lambda * 1.0e6;
    double s = 1.0 / (lambda_um * lambda_um);
    double r_as = 1.0e-8 * (k1 / (k0 - s) + k3 / (k2 - s));
    double r_vs = 1.022e-8 * (w0 + w1 * s + w2 * s * s + w3 * s * s * s);
    double m_a = 0.0289635 + 1.2011e-8 * (x_c - 400.0);
    double r_axs = r_as * (1.0 + 5.34e-7 * (x_c - 450.0));

    double t_c = t_kelvin - 273.15;
    double x_v = 0.0;
    if (rh != 0.0) {
        double svp = std::exp(sa * t_kelvin * t_kelvin + sb * t_kelvin + sc + sd / t_kelvin);
        double f = alpha + beta * p + gamma * t_c * t_c;
        x_v = rh * f * svp / p;
    }
    double pt = p / t_kelvin;
    double z_m = 1.0 - pt * (a0 + a1 * t_c + a2 * t_c * t_c + (b0 + b1 * t_c) * x_v + (c0 + c1 * t_c) * x_v * x_v) +
                 pt * pt * (d + e * x_v * x_v);
    double rho_axs = p_r1 * m_a / (z_a * gas_r * t_r1);
    double rho_v = x_v * p * m_v / (z_m * gas_r * t_kelvin);
    double rho_a = (1.0 - x_v) * p * m_a / (z_m * gas_r * t_kelvin);
    return 1.0 + (rho_a / rho_axs) * r_axs + (rho_v / rho_vs) * r_vs;
}

struct Environment {
    int flat;  // EarthShape::Flat
    double radius;
    Atmosphere atm;
    double wavelength;
};

// Environment::n(h)
double env_n(const Environment& env, double h) {
    int i = atm_layer_index(env.atm, h);
    double t = layer_temperature(env.atm.layer[i], h);
    double p = layer_pressure(env.atm.layer[i], h);
    return air_index(env.wavelength, p, t, env.atm.humidity);
}
// Environment::dn(h): central difference, epsilon 0.01 m [recalled]
double env_dn(const Environment& env, double h) {
    const double eps = 0.01;
    double n1 = env_n(env, h - eps);
    double n2 = env_n(env, h + eps);
    return (n2 - n1) / (2.0 * eps);
}

struct RayState {
    double x, h, dh;
};

// PathStepper restated: classical RK4 on the Fermat ray ODE (spherical: r(phi); flat: h(x)),
// or the closed-form straight line.
struct Stepper {
    const Environment* env;
    bool straight;
    double step;
    // spherical RK4 state
    double r, dr, phi;
    // flat RK4 state
    double h, dh, x;
    // straight-line parameters
    double h0, ang, xs;

    Stepper(const Environment* e, double start_h, double ang_rad, bool straight_)
        : env(e), straight(straight_), step(1.0) {
        h0 = start_h;
        ang = ang_rad;
        xs = 0.0;
        r = e->radius + start_h;
        dr = r * std::tan(ang_rad);
        phi = 0.0;
        h = start_h;
        dh = std::tan(ang_rad);
        x = 0.0;
    }
    void set_step_size(double s) { step = s; }

    void deriv_sph(double r_, double dr_, double* o_r, double* o_dr) const {
        double hh = r_ - env->radius;
        double n = env_n(*env, hh);
        double dn = env_dn(*env, hh);
        *o_r = dr_;
        *o_dr = dr_ * dr_ * dn / n + r_ * r_ * dn / n + 2.0 * dr_ * dr_ / r_ + r_;
    }
    void deriv_flat(double h_, double dh_, double* o_h, double* o_dh) const {
        double n = env_n(*env, h_);
        double dn = env_dn(*env, h_);
        *o_h = dh_;
        *o_dh = dn / n * (1.0 + dh_ * dh_);
    }

    RayState next() {
        if (straight) {
            xs += step;
            if (env->flat) {
                return {xs, h0 + xs * std::tan(ang), std::tan(ang)};
            }
            double r0 = env->radius + h0;
            double ph = xs / env->radius;
            double rr = r0 * std::cos(ang) / std::cos(ph + ang);
            return {xs, rr - env->radius, std::tan(ph + ang)};
        }
        if (env->flat) {
            double s = step;
            double k1a, k1b, k2a, k2b, k3a, k3b, k4a, k4b;
            deriv_flat(h, dh, &k1a, &k1b);
            deriv_flat(h + 0.5 * s * k1a, dh + 0.5 * s * k1b, &k2a, &k2b);
            deriv_flat(h + 0.5 * s * k2a, dh + 0.5 * s * k2b, &k3a, &k3b);
            deriv_flat(h + s * k3a, dh + s * k3b, &k4a, &k4b);
            h = h + (k1a + 2.0 * k2a + 2.0 * k3a + k4a) * s / 6.0;
            dh = dh + (k1b + 2.0 * k2b + 2.0 * k3b + k4b) * s / 6.0;
            x += s;
            return {x, h, dh};
        }
        double s = step / env->radius;
        double k1a, k1b, k2a, k2b, k3a, k3b, k4a, k4b;
        deriv_sph(r, dr, &k1a, &k1b);
        deriv_sph(r + 0.5 * s * k1a, dr + 0.5 * s * k1b, &k2a, &k2b);
        deriv_sph(r + 0.5 * s * k2a, dr + 0.5 * s * k2b, &k3a, &k3b);
        deriv_sph(r + s * k3a, dr + s * k3b, &k4a, &k4b);
        r = r + (k1a + 2.0 * k2a + 2.0 * k3a + k4a) * s / 6.0;
        dr = dr + (k1b + 2.0 * k2b + 2.0 * k3b + k4b) * s / 6.0;
        phi += s;
        return {phi * env->radius, r - env->radius, dr / r};
    }
};

// ------------------------------------------------------------------------------------------
// Scene objects: object/mod.rs, frustum.rs, billboard.rs
// ------------------------------------------------------------------------------------------
struct Color {
    double r, g, b, a;
};

struct Object {
    int kind;
    Coords position;  // altitude resolved
    double r1, r2, width, height;
    Color color;
    int tex_w, tex_h;
    const uint8_t* tex;  // RGBA8 row-major, row 0 = top
};

struct Collision {
    double prop;
    V3 normal;
    Color color;
};

// Image::get_pixel, object/mod.rs:89-118
void texture_get_pixel(const Object& o, double x, double y, uint8_t out[4]) {
    double w = (double)o.tex_w, h = (double)o.tex_h;
    x = x * w - 0.5;
    double x1 = std::floor(x);
    x1 = x1 < 0.0 ? 0.0 : (x1 > w - 2.0 ? w - 2.0 : x1);  // clamp(0, w-2)
    double x2 = x1 + 1.0;
    uint32_t ix1 = (uint32_t)x1, ix2 = (uint32_t)x2;
    y = (1.0 - y) * h - 0.5;
    double y1 = std::floor(y);
    y1 = y1 < 0.0 ? 0.0 : (y1 > h - 2.0 ? h - 2.0 : y1);
    double y2 = y1 + 1.0;
    uint32_t iy1 = (uint32_t)y1, iy2 = (uint32_t)y2;
    double px = x - x1, py = y - y1;
    auto pix = [&](uint32_t ix, uint32_t iy, int c) { return (double)o.tex[((size_t)iy * o.tex_w + ix) * 4 + c] / 255.0; };
    for (int c = 0; c < 4; ++c) {
        double v = pix(ix1, iy1, c) * (1.0 - px) * (1.0 - py) + pix(ix1, iy2, c) * (1.0 - px) * py +
                   pix(ix2, iy1, c) * px * (1.0 - py) + pix(ix2, iy2, c) * px * py;
        out[c] = as_u8(v * 255.0);
    }
}

// Frustum::check_collision, frustum.rs:18-101
void frustum_collision(const Object& o, const EarthModel& m, const Coords& point1, const Coords& point2,
                       std::vector<Collision>* results) {
    V3 pos1 = as_cartesian(m, point1), pos2 = as_cartesian(m, point2), obj_pos = as_cartesian(m, o.position);
    V3 p1 = pos1 - obj_pos;
    double p1sq = dot(p1, p1);
    V3 v = world_directions(m, o.position.lat, o.position.lon).up;
    V3 w = pos2 - pos1;
    double wsq = dot(w, w), p1v = dot(p1, v), p1w = dot(p1, w), wv = dot(w, v);
    double aa = (o.r2 - o.r1) / o.height;
    double aa1 = 1.0 + aa * aa;
    double a = wsq - wv * wv * (1.0 + aa * aa);
    double b = 2.0 * (p1w - wv * (p1v * aa1 + aa * o.r1));
    double c = p1sq - p1v * p1v * aa1 - o.r1 * o.r1 - 2.0 * aa * o.r1 * p1v;
    double delta = b * b - 4.0 * a * c;
    if (delta >= 0.0) {
        double x1 = (-b - std::sqrt(delta)) / 2.0 / a;
        double x2 = (-b + std::sqrt(delta)) / 2.0 / a;
        if (a < 0.0) std::swap(x1, x2);
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
            double o_len = std::sqrt(dot(outward, outward));
            outward = outward / o_len;
            double ang = std::atan2(o.r1 - o.r2, o.height);
            V3 normal = outward * std::cos(ang) + v * std::sin(ang);
            results->push_back({x, normal, o.color});
        }
    }
    const double hs[2] = {0.0, o.height}, rs[2] = {o.r1, o.r2};
    const V3 ns[2] = {-v, v};
    for (int i = 0; i < 2; ++i) {
        double x = (hs[i] - p1v) / wv;
        V3 out = p1 + w * x - hs[i] * v;
        double d = dot(out, out);
        if (d < rs[i] * rs[i] && in_range(0.0, 1.0, x)) results->push_back({x, ns[i], o.color});
    }
    std::stable_sort(results->begin(), results->end(), [](const Collision& l, const Collision& r) { return l.prop < r.prop; });
}

// Billboard::check_collision, billboard.rs:17-66
void billboard_collision(const Object& o, const EarthModel& m, const Coords& point1, const Coords& point2,
                         std::vector<Collision>* results) {
    V3 pos1 = as_cartesian(m, point1), pos2 = as_cartesian(m, point2), obj_pos = as_cartesian(m, o.position);
    V3 ray = pos2 - pos1;
    V3 up = world_directions(m, o.position.lat, o.position.lon).up;
    V3 right = cross(ray, up);
    double right_len = std::sqrt(dot(right, right));
    right = right / right_len;
    V3 front = cross(right, up);
    V3 p1 = pos1 - obj_pos;
    double prop = -dot(p1, front) / dot(ray, front);
    if (!in_range(0.0, 1.0, prop)) return;
    V3 intersection = p1 + ray * prop;
    double y = dot(intersection, up), x = dot(intersection, right);
    if (!in_range(0.0, o.height, y) || !in_range(-o.width / 2.0, o.width / 2.0, x)) return;
    x = (x + o.width / 2.0) / o.width;
    y = y / o.height;
    uint8_t px[4];
    texture_get_pixel(o, x, y, px);
    Color c{(double)px[0] / 255.0, (double)px[1] / 255.0, (double)px[2] / 255.0, (double)px[3] / 255.0};
    results->push_back({prop, front, c});
}

void check_collision(const Object& o, const EarthModel& m, const Coords& p1, const Coords& p2, std::vector<Collision>* out) {
    out->clear();
    if (o.kind == ATMRT_OBJECT_FRUSTUM)
        frustum_collision(o, m, p1, p2, out);
    else
        billboard_collision(o, m, p1, p2, out);
}

// Object::is_close, frustum.rs:103-114 / billboard.rs:68-78
bool is_close(const Object& o, const EarthModel& m, double sim_step, double lat, double lon) {
    V3 obj_pos = as_cartesian(m, o.position);
    V3 pos = as_cartesian(m, Coords{lat, lon, o.position.elev});
    V3 dist_v = pos - obj_pos;
    double r = o.kind == ATMRT_OBJECT_FRUSTUM ? std::fmax(o.r1, o.r2) : o.width;
    return dot(dist_v, dist_v) < 2.0 * (r + sim_step) * (r + sim_step);
}

// ------------------------------------------------------------------------------------------
// Generator: generators/utils.rs, generators/fast.rs
// ------------------------------------------------------------------------------------------
struct Scene {
    EarthModel model;
    Environment env;
    Terrain terrain;
    std::vector<Object> objects;
    atmrt_params p;
    double observer_alt;
};

// find_normal, utils.rs:15-40
V3 find_normal(const Scene& s, double lat, double lon) {
    const double DIFF = 15.0;
    DirCalc ns_calc = coords_at_dist_calc(s.model, lat, lon, 0.0);
    DirCalc ew_calc = coords_at_dist_calc(s.model, lat, lon, 90.0);
    double n_lat, n_lon, s_lat, s_lon, e_lat, e_lon, w_lat, w_lon;
    coords_at_dist(ns_calc, DIFF, &n_lat, &n_lon);
    coords_at_dist(ns_calc, -DIFF, &s_lat, &s_lon);
    coords_at_dist(ew_calc, DIFF, &e_lat, &e_lon);
    coords_at_dist(ew_calc, -DIFF, &w_lat, &w_lon);
    Dirs d = world_directions(s.model, lat, lon);
    double diff_ew = elev_or_zero(s.terrain, e_lat, e_lon) - elev_or_zero(s.terrain, w_lat, w_lon);
    double diff_ns = elev_or_zero(s.terrain, n_lat, n_lon) - elev_or_zero(s.terrain, s_lat, s_lon);
    V3 vec_ns = 2.0 * DIFF * d.north + diff_ns * d.up;
    V3 vec_ew = 2.0 * DIFF * d.east + diff_ew * d.up;
    return normalized(cross(vec_ew, vec_ns));
}

// calc_dist, utils.rs:42-53
double calc_dist(const Environment& env, RayState o, RayState n) {
    double dx = n.x - o.x, dh = n.h - o.h;
    if (env.flat) return std::sqrt(dx * dx + dh * dh);
    double avg_h = (n.h + o.h) / 2.0;
    dx = dx / env.radius * (avg_h + env.radius);
    return std::sqrt(dx * dx + dh * dh);
}

struct PathElem {
    double dist, elev, path_length;
};

struct TerrainData {
    double lat, lon, elev;
    V3 normal;
    uint64_t objects_close;  // bit i = object i (Vec<usize> in the reference)
};

// TerrainData::from_lat_lon, utils.rs:72-88
TerrainData terrain_data_from(const Scene& s, double lat, double lon) {
    TerrainData t;
    t.normal = find_normal(s, lat, lon);
    t.objects_close = 0;
    for (size_t i = 0; i < s.objects.size(); ++i)
        if (is_close(s.objects[i], s.model, s.p.simulation_step, lat, lon)) t.objects_close |= (1ull << i);
    t.lat = lat;
    t.lon = lon;
    t.elev = elev_or_zero(s.terrain, lat, lon);
    return t;
}

// gen_path_cache, utils.rs:136-174
std::vector<PathElem> gen_path_cache(const Scene& s, double ray_elev) {
    double alt = altitude_abs(s.p.altitude, s.terrain, s.p.latitude, s.p.longitude);
    Stepper ray(&s.env, alt, to_radians(ray_elev), s.p.straight_rays != 0);
    ray.set_step_size(s.p.simulation_step);
    std::vector<PathElem> path;
    path.push_back({0.0, alt, 0.0});
    RayState ray_state{0.0, alt, 0.0};
    double path_length = 0.0;
    for (;;) {
        RayState nw = ray.next();
        path_length += calc_dist(s.env, ray_state, nw);
        path.push_back({nw.x, nw.h, path_length});
        if (ray_state.x > s.p.max_distance || ray_state.h < -1000.0) break;
        ray_state = nw;
    }
    return path;
}

// gen_terrain_cache, utils.rs:176-199
std::vector<TerrainData> gen_terrain_cache(const Scene& s, double dir) {
    std::vector<TerrainData> result;
    double distance = 0.0;
    DirCalc calc = coords_at_dist_calc(s.model, s.p.latitude, s.p.longitude, dir);
    while (distance < s.p.max_distance) {
        double lat, lon;
        coords_at_dist(calc, distance, &lat, &lon);
        result.push_back(terrain_data_from(s, lat, lon));
        distance += s.p.simulation_step;
    }
    return result;
}

// get_ray_elev / get_ray_dir, fast.rs:111-125 (pixel centring goes through i16)
double get_ray_elev(const atmrt_params& p, int y) {
    double width = (double)p.width, height = (double)p.height;
    double aspect = width / height;
    double yy = (double)(int16_t)((int16_t)y - (int16_t)p.height / 2) / height;
    return p.tilt - yy * p.fov / aspect;
}
double get_ray_dir(const atmrt_params& p, int x) {
    double width = (double)p.width;
    double xx = (double)(int16_t)((int16_t)x - (int16_t)p.width / 2) / width;
    return p.direction + xx * p.fov;
}

struct TracePoint {
    double lat, lon, distance, elevation, path_length;
    V3 normal;
    bool is_terrain;
    Color color;  // Rgba; for terrain {0,0,0,alpha}
    int step;
    double alpha() const { return color.a; }
};

struct TracingState {
    double lat, lon, elev;
    V3 normal;
    uint64_t close;
    double ray_elev, dist, path_len;
};

// TracingState::interpolate, utils.rs:108-125: a + (b - a) * prop
TracingState interpolate(const TracingState& a, const TracingState& b, double prop) {
    TracingState r;
    r.lat = a.lat + (b.lat - a.lat) * prop;
    r.lon = a.lon + (b.lon - a.lon) * prop;
    r.elev = a.elev + (b.elev - a.elev) * prop;
    r.normal = a.normal + (b.normal - a.normal) * prop;
    r.close = 0;
    r.ray_elev = a.ray_elev + (b.ray_elev - a.ray_elev) * prop;
    r.dist = a.dist + (b.dist - a.dist) * prop;
    r.path_len = a.path_len + (b.path_len - a.path_len) * prop;
    return r;
}

// The stream get_single_pixel consumes: the Fast generator zips two caches (fast.rs:56-64), the Rectilinear
// generator walks its own PathIterator per pixel (rectilinear.rs:112-186).
struct ZipSource {
    const std::vector<TerrainData>& terr;
    const std::vector<PathElem>& path;
    bool has(size_t k) const { return k < terr.size() && k < path.size(); }
};

// get_single_pixel, utils.rs:201-289. Returns iterations consumed.
template <class Source>
int get_single_pixel(const Scene& s, Source& src, std::vector<TracePoint>* result, uint64_t* step_overflows) {
    auto& terr = src.terr;
    auto& path = src.path;
    auto make = [&](size_t k, bool first) {
        TracingState t;
        t.lat = terr[k].lat;
        t.lon = terr[k].lon;
        t.elev = terr[k].elev;
        t.normal = terr[k].normal;
        t.close = terr[k].objects_close;
        t.ray_elev = path[k].elev;
        t.dist = first ? 0.0 : path[k].dist;
        t.path_len = first ? 0.0 : path[k].path_length;
        return t;
    };
    double terrain_alpha = s.p.terrain_alpha;
    std::vector<Collision> coll;
    std::vector<std::pair<double, TracePoint>> step_result;
    int consumed = 0;
    for (size_t k = 1; src.has(k); ++k) {
        consumed = (int)k;
        // old_tracing_state is always (terr[k-1], path[k-1]) -- with dist/path_len forced to 0.0 for
        // k-1 == 0 (utils.rs:207-208) -- so it is rebuilt on demand instead of being cloned per step.
        double diff1 = path[k - 1].elev - terr[k - 1].elev;
        double diff2 = path[k].elev - terr[k].elev;
        bool terrain_hit = diff1 * diff2 < 0.0;
        uint64_t mask = terr[k].objects_close | terr[k - 1].objects_close;
        if (!terrain_hit && mask == 0) continue;
        bool finish = false;
        step_result.clear();
        TracingState old_state = make(k - 1, k - 1 == 0);
        TracingState new_state = make(k, false);
        if (terrain_hit) {
            double prop = diff1 / (diff1 - diff2);
            TracingState it = interpolate(old_state, new_state, prop);
            TracePoint tp{it.lat, it.lon, it.dist, it.elev, it.path_len, it.normal, true, Color{0, 0, 0, terrain_alpha}, (int)k};
            step_result.push_back({prop, tp});
            if (terrain_alpha == 1.0) finish = true;
        }
        if (mask != 0) {
            // The reference iterates a HashSet (random order); we iterate by object index. Order only
            // matters for exactly equal `prop` values (stable sort below).
            Coords c1{old_state.lat, old_state.lon, old_state.ray_elev};
            Coords c2{new_state.lat, new_state.lon, new_state.ray_elev};
            for (size_t oi = 0; oi < s.objects.size(); ++oi) {
                if (!((mask >> oi) & 1)) continue;
                check_collision(s.objects[oi], s.model, c1, c2, &coll);
                for (const Collision& c : coll) {
                    if (c.color.a == 0.0) continue;
                    TracingState it = interpolate(old_state, new_state, c.prop);
                    TracePoint tp{it.lat, it.lon, it.dist, it.ray_elev, it.path_len, c.normal, false, c.color, (int)k};
                    step_result.push_back({c.prop, tp});
                    if (c.color.a == 1.0) {
                        finish = true;
                        break;
                    }
                }
            }
        }
        std::stable_sort(step_result.begin(), step_result.end(),
                         [](const std::pair<double, TracePoint>& l, const std::pair<double, TracePoint>& r) { return l.first < r.first; });
        if (step_result.size() > ATMRT_MAX_STEP_POINTS && step_overflows) ++*step_overflows;
        for (auto& pr : step_result) result->push_back(pr.second);
        if (finish) break;
    }
    return consumed;
}

// ------------------------------------------------------------------------------------------
// Rectilinear generator: generators/rectilinear.rs
// ------------------------------------------------------------------------------------------
struct RayParams {
    double elevation, direction;  // radians
};

// RectilinearGenerator::get_ray_params, rectilinear.rs:80-105. nalgebra's Matrix4::from_euler_angles(roll,
// pitch, yaw) is Rz(yaw) Ry(pitch) Rx(roll); its product with a vector accumulates column by column.
RayParams get_ray_params(const atmrt_params& p, int x, int y) {
    double width = (double)p.width;
    double xf = (double)(int16_t)((int16_t)x - (int16_t)p.width / 2);
    double yf = (double)(int16_t)((int16_t)y - (int16_t)p.height / 2);
    double z = width / 2.0 / std::tan(to_radians(p.fov) / 2.0);
    double roll = 0.0, pitch = -to_radians(p.tilt), yaw = to_radians(p.direction);
    double sr = std::sin(roll), cr = std::cos(roll), sp = std::sin(pitch), cp = std::cos(pitch), sy = std::sin(yaw), cy = std::cos(yaw);
    double m[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                      {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                      {-sp, cp * sr, cp * cr}};
    double v[3] = {z, xf, -yf};  // [forward, right, up]
    double d[3];
    for (int i = 0; i < 3; ++i) d[i] = (m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2];
    double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; ++i) d[i] = d[i] / n;
    return {std::asin(d[2]), std::atan2(d[1], d[0])};
}

// PathIterator (rectilinear.rs:112-186) as a lazily materialised stream: point k is the stepper's state k
// with the terrain under it; the stream ends at the first state past max_distance or below -1000 m.
struct PathIteratorSource {
    const Scene& s;
    Stepper ray;
    DirCalc dist_calc;
    RayState ray_state;
    double path_length = 0.0;
    bool exhausted = false;
    std::vector<TerrainData> terr;
    std::vector<PathElem> path;

    PathIteratorSource(const Scene& scene, RayParams rp)
        : s(scene),
          ray(&scene.env, altitude_abs(scene.p.altitude, scene.terrain, scene.p.latitude, scene.p.longitude), rp.elevation, scene.p.straight_rays != 0),
          dist_calc(coords_at_dist_calc(scene.model, scene.p.latitude, scene.p.longitude, to_degrees(rp.direction))) {
        ray.set_step_size(scene.p.simulation_step);
        ray_state = {0.0, altitude_abs(scene.p.altitude, scene.terrain, scene.p.latitude, scene.p.longitude), 0.0};
    }
    bool next() {  // Iterator::next, rectilinear.rs:173-185
        PathElem elem{ray_state.x, ray_state.h, path_length};
        if (elem.dist > s.p.max_distance || elem.elev < -1000.0) return false;
        double lat, lon;
        coords_at_dist(dist_calc, elem.dist, &lat, &lon);
        terr.push_back(terrain_data_from(s, lat, lon));
        path.push_back(elem);
        RayState nw = ray.next();
        path_length += calc_dist(s.env, ray_state, nw);
        ray_state = nw;
        return true;
    }
    bool has(size_t k) {
        while (!exhausted && terr.size() <= k) exhausted = !next();
        return k < terr.size();
    }
};

// ------------------------------------------------------------------------------------------
// InterpolatingRectilinear generator: generators/interpolating_rectilinear.rs. A regular grid in (elevation,
// direction) -- steps = 1.5 x the smallest angular distance between neighbouring pixels of the rectilinear image
// (gen_fov_data, :432-521) -- whose points are Fast-generator pixels (one path cache per elevation index, one
// terrain cache per direction index, get_single_pixel on the pair, :86-117); an image pixel blends the trace points
// of the four grid points around its own ray (interpolate, :394-419).
// ------------------------------------------------------------------------------------------
struct FovData {
    double min_elev_step, min_dir_step;
};

// gen_fov_data, interpolating_rectilinear.rs:432-521 (over the WHOLE image, whatever column block is rendered)
FovData gen_fov_data(const atmrt_params& p, const std::vector<RayParams>& table) {
    const int W = p.width, H = p.height;
    const double two_pi = 360.0 * (PI / 180.0), min_diff = p.fov * (PI / 180.0) / (double)W / 3.0, scale = 1.5;
    double min_elev = std::numeric_limits<double>::infinity(), min_dir = std::numeric_limits<double>::infinity();
    for (int x = 0; x < W; ++x) {
        double m = two_pi, last = table[x].elevation;
        for (int y = 1; y < H; ++y) {
            const double next = table[(size_t)y * W + x].elevation;
            double diff = std::fabs(next - last);
            if (diff < min_diff) diff = min_diff;
            if (diff < m) m = diff;
            last = next;
        }
        min_elev = std::fmin(min_elev, m);
    }
    for (int y = 0; y < H; ++y) {
        double m = two_pi, last = table[(size_t)y * W].direction;
        for (int x = 1; x < W; ++x) {
            const double next = table[(size_t)y * W + x].direction;
            double diff = std::fabs(next - last);
            if (diff > two_pi) diff -= two_pi;
            if (diff < min_diff) diff = min_diff;
            if (diff < m) m = diff;
            last = next;
        }
        min_dir = std::fmin(min_dir, m);
    }
    return {min_elev * scale, min_dir * scale};
}

// PixelColor::same_class / TracePoint::interpolate (generators/mod.rs:32-80): self * (1 - coeff) + other * coeff
TracePoint blend(const TracePoint& a, const TracePoint& b, double coeff) {
    auto mix = [&](double u, double v) { return u * (1.0 - coeff) + v * coeff; };
    TracePoint r;
    r.lat = mix(a.lat, b.lat), r.lon = mix(a.lon, b.lon), r.distance = mix(a.distance, b.distance);
    r.elevation = mix(a.elevation, b.elevation), r.path_length = mix(a.path_length, b.path_length);
    r.normal = a.normal * (1.0 - coeff) + b.normal * coeff;
    r.is_terrain = a.is_terrain || b.is_terrain;  // a mixed pair is Terrain (it does not occur: groups hold one class)
    if (a.is_terrain && b.is_terrain) r.color = Color{0, 0, 0, mix(a.color.a, b.color.a)};
    else if (!a.is_terrain && !b.is_terrain) r.color = Color{mix(a.color.r, b.color.r), mix(a.color.g, b.color.g), mix(a.color.b, b.color.b), mix(a.color.a, b.color.a)};
    else r.color = Color{0, 0, 0, a.is_terrain ? a.color.a : b.color.a};
    r.step = a.step;
    return r;
}

// interpolate_trace_points, :274-345: slot i = SEQUENCE[i] = (elev + i / 2, dir + i % 2). False: the group yields nothing.
bool blend_group(const TracePoint* e[4], double re, double rd, TracePoint* out) {
    const int have = (e[0] ? 1 : 0) | (e[1] ? 2 : 0) | (e[2] ? 4 : 0) | (e[3] ? 8 : 0);
    auto two_adjacent = [&](const TracePoint& a, const TracePoint& b, double r_elev, double r_dir) {
        if (r_elev >= 0.5) return false;
        *out = blend(a, b, r_dir);
        return true;
    };
    auto two_diagonal = [&](const TracePoint& a, const TracePoint& b, double r_elev, double r_dir) {
        if ((r_elev >= 0.5 && r_dir < 0.5) || (r_elev < 0.5 && r_dir >= 0.5)) return false;
        const double coeff = r_elev * r_dir / (r_elev * r_dir + (1.0 - r_elev) * (1.0 - r_dir));
        *out = blend(a, b, coeff);
        return true;
    };
    auto three = [&](const TracePoint& a, const TracePoint& b, const TracePoint& c, double r_elev, double r_dir) {
        if (r_elev >= 0.5 && r_dir >= 0.5) return false;
        const double sum = 1.0 - r_elev + r_elev * (1.0 - r_dir);
        *out = blend(blend(a, b, r_dir), c, r_elev * (1.0 - r_dir) / sum);
        return true;
    };
    auto single = [&](const TracePoint& a, bool cond) {
        if (!cond) return false;
        *out = a;
        return true;
    };
    switch (have) {
        case 0: return false;
        case 1: return single(*e[0], re < 0.5 && rd < 0.5);
        case 2: return single(*e[1], re < 0.5 && rd >= 0.5);
        case 4: return single(*e[2], re >= 0.5 && rd < 0.5);
        case 8: return single(*e[3], re >= 0.5 && rd >= 0.5);
        case 1 | 2: return two_adjacent(*e[0], *e[1], re, rd);
        case 1 | 4: return two_adjacent(*e[0], *e[2], rd, re);
        case 1 | 8: return two_diagonal(*e[0], *e[3], re, rd);
        case 2 | 4: return two_diagonal(*e[1], *e[2], re, 1.0 - rd);
        case 2 | 8: return two_adjacent(*e[1], *e[3], 1.0 - rd, re);
        case 4 | 8: return two_adjacent(*e[2], *e[3], 1.0 - re, rd);
        case 1 | 2 | 4: return three(*e[0], *e[1], *e[2], re, rd);
        case 1 | 2 | 8: return three(*e[1], *e[0], *e[3], re, 1.0 - rd);
        case 1 | 4 | 8: return three(*e[0], *e[3], *e[2], 1.0 - re, rd);
        case 2 | 4 | 8: return three(*e[3], *e[2], *e[1], 1.0 - re, 1.0 - rd);
        default: *out = blend(blend(*e[0], *e[1], rd), blend(*e[2], *e[3], rd), re); return true;
    }
}

// interpolate, :394-419 with collect_trace_points, :213-243: a trace point joins the FIRST group that holds a point of its
// class closer than one simulation step in distance, else it opens a group; a later point of the same grid pixel replaces
// an earlier one in the group's slot (match_sequence, :245-266).
std::vector<TracePoint> blend_pixels(const std::vector<TracePoint>* px[4], double re, double rd, double step_size) {
    struct Member {
        int slot;
        const TracePoint* tp;
    };
    std::vector<std::vector<Member>> groups;
    for (int slot = 0; slot < 4; ++slot)
        for (const TracePoint& tp : *px[slot]) {
            size_t g = 0;
            for (; g < groups.size(); ++g) {
                bool close = false;
                for (const Member& m : groups[g])
                    if (std::fabs(tp.distance - m.tp->distance) < step_size && tp.is_terrain == m.tp->is_terrain) {
                        close = true;
                        break;
                    }
                if (close) break;
            }
            if (g == groups.size()) groups.emplace_back();
            groups[g].push_back({slot, &tp});
        }
    std::vector<TracePoint> out;
    for (const auto& grp : groups) {
        const TracePoint* e[4] = {nullptr, nullptr, nullptr, nullptr};
        for (const Member& m : grp) e[m.slot] = m.tp;
        TracePoint tp;
        if (blend_group(e, re, rd, &tp)) out.push_back(tp);
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// Colouring: coloring/shading.rs, coloring/simple.rs ; compositing: renderer/mod.rs:367-414
// ------------------------------------------------------------------------------------------
struct Rgb8 {
    uint8_t c[3];
};

V3 palette_sky(int palette) { return palette == ATMRT_PALETTE_LEGACY ? V3{0.11, 0.11, 0.11} : V3{0.23, 0.41, 0.55}; }
V3 palette_water(int palette) { return palette == ATMRT_PALETTE_LEGACY ? V3{0.0, 0.5, 1.0} : V3{0.23, 0.41, 0.55}; }

// ColorPalette::elev_to_color, shading.rs:30-83
V3 elev_to_color(int palette, double elev) {
    double thr1 = 300.0, thr2, thr3 = 1800.0, thr4 = 3000.0;
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
    if (elev < thr1) return c0;
    if (elev < thr2) {
        double prop = (elev - thr1) / (thr2 - thr1);
        return c1 * prop + c0 * (1.0 - prop);
    }
    if (elev < thr3) {
        double prop = (elev - thr2) / (thr3 - thr2);
        return c2 * prop + c1 * (1.0 - prop);
    }
    if (elev < thr4) {
        double prop = (elev - thr3) / (thr4 - thr3);
        return c3 * prop + c2 * (1.0 - prop);
    }
    return c3;
}

// hsv, simple.rs:55-87
Rgb8 hsv(double h, double s, double v) {
    double c = v * s;
    h = std::fmod(h, 360.0) < 0.0 ? std::fmod(h, 360.0) + 360.0 : std::fmod(h, 360.0);
    double x = c * (1.0 - std::fabs(std::fmod(h / 60.0, 2.0) - 1.0));
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
    }  // else: unreachable!() in the reference (h == 360 after rounding, NaN)
    return {{as_u8((rp + m) * 255.0), as_u8((gp + m) * 255.0), as_u8((bp + m) * 255.0)}};
}

Rgb8 color_for_pixel(const atmrt_params& p, const TracePoint& tp) {
    if (p.coloring == ATMRT_COLORING_SIMPLE) {
        // SimpleColors::color_for_pixel, simple.rs:22-44
        double dist_ratio = tp.distance / p.simple_max_distance;
        if (tp.elevation <= p.water_level) {
            double mul = 1.0 - dist_ratio * 0.6;
            return {{0, as_u8(128.0 * mul), as_u8(255.0 * mul)}};
        }
        double elev_ratio = tp.elevation / 4500.0;
        double h = 120.0 - 240.0 * (elev_ratio < 0.0 ? -std::pow(-elev_ratio, 0.65) : std::pow(elev_ratio, 0.65));
        double v = (elev_ratio > 0.7 ? 2.1 - elev_ratio * 2.0 : 0.9 - elev_ratio / 0.7 * 0.2) * (1.0 - dist_ratio * 0.6);
        double s = 1.0 - dist_ratio * 0.9;
        return hsv(h, s, v);
    }
    // Shading::color_for_pixel, shading.rs:115-132
    V3 light{p.light_dir[0], p.light_dir[1], p.light_dir[2]};
    double light_dot = dot(light, tp.normal);
    light_dot = light_dot >= 0.0 ? light_dot : 0.0;
    double brightness = p.ambient_light + (1.0 - p.ambient_light) * light_dot * light_dot;
    V3 base;
    if (!tp.is_terrain)
        base = {tp.color.r, tp.color.g, tp.color.b};
    else if (tp.elevation <= p.water_level)
        base = palette_water(p.palette);
    else
        base = elev_to_color(p.palette, tp.elevation);
    V3 c = base * brightness;
    return {{as_u8(c.x * 255.0), as_u8(c.y * 255.0), as_u8(c.z * 255.0)}};
}

Rgb8 sky_color(const atmrt_params& p) {
    if (p.coloring == ATMRT_COLORING_SIMPLE) return {{28, 28, 28}};
    V3 c = palette_sky(p.palette);
    return {{as_u8(c.x * 255.0), as_u8(c.y * 255.0), as_u8(c.z * 255.0)}};
}

// fog, renderer/mod.rs:367-376
Rgb8 fog(double fog_dist, double pixel_dist, Rgb8 color) {
    double fog_coeff = 1.0 - std::exp(-pixel_dist / fog_dist);
    Rgb8 out;
    for (int i = 0; i < 3; ++i) out.c[i] = as_u8((double)color.c[i] * (1.0 - fog_coeff) + 160.0 * fog_coeff);
    return out;
}
// add, renderer/mod.rs:378-383 (utils/mod.rs:16-29)
Rgb8 add(Rgb8 rgb1, Rgb8 rgb2, double a) {
    Rgb8 out;
    for (int i = 0; i < 3; ++i) {
        double c1 = (double)rgb1.c[i] / 255.0, c2 = (double)rgb2.c[i] / 255.0;
        out.c[i] = as_u8((c1 + c2 * a) * 255.0);
    }
    return out;
}

// draw_image body for one pixel, renderer/mod.rs:395-411
Rgb8 draw_pixel(const atmrt_params& p, const std::vector<TracePoint>& tps) {
    Rgb8 def_color = p.fog_enabled ? Rgb8{{160, 160, 160}} : sky_color(p);
    Rgb8 result{{0, 0, 0}};
    double accum_neg_alpha = 1.0;
    for (const TracePoint& tp : tps) {
        Rgb8 color1 = color_for_pixel(p, tp);
        Rgb8 color2 = p.fog_enabled ? fog(p.fog_distance, tp.path_length, color1) : color1;
        result = add(result, color2, accum_neg_alpha * tp.alpha());
        accum_neg_alpha *= 1.0 - tp.alpha();
    }
    return add(result, def_color, accum_neg_alpha);
}

bool build_scene(const atmrt_params* p, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts,
                 const atmrt_object* objects, int nobjects, const uint8_t* const* textures, Scene* s) {
    s->p = *p;
    s->model = {p->earth_model, p->radius, p->ellipsoid_b};
    s->env.flat = shape_flat(s->model);  // EarthModel::to_shape, mod.rs:95-112
    s->env.radius = shape_radius(s->model);
    s->env.wavelength = p->wavelength;
    if (!atmosphere_from_def(p->atmosphere, &s->env.atm)) return false;
    s->terrain = make_terrain(tiles, ntiles, posts);
    s->objects.clear();
    for (int i = 0; i < nobjects; ++i) {
        const atmrt_object& o = objects[i];
        Object ob{};
        ob.kind = o.kind;
        ob.position = {o.latitude, o.longitude, altitude_abs(o.altitude, s->terrain, o.latitude, o.longitude)};
        ob.r1 = o.r1;
        ob.r2 = o.r2;
        ob.width = o.width;
        ob.height = o.height;
        ob.color = {o.color[0], o.color[1], o.color[2], o.color[3]};
        ob.tex_w = o.texture_width;
        ob.tex_h = o.texture_height;
        ob.tex = textures ? textures[i] : nullptr;
        s->objects.push_back(ob);
    }
    s->observer_alt = altitude_abs(p->altitude, s->terrain, p->latitude, p->longitude);
    return true;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ==========================================================================================
// C API (ctypes). Prefix oracle_ so it can never be mistaken for the product ABI.
// ==========================================================================================
extern "C" {

struct oracle_timing {
    double s_terrain, s_paths, s_pixels, s_image, s_total;
    int threads;
    int _pad;
};

// FastGenerator::generate (fast.rs:22-98) + draw_image (renderer/mod.rs:385-414) for the column
// block [x0,x1), optionally on a row/column sub-sample (every `stride_x`-th column and
// `stride_y`-th row, used for the bounded CPU baseline). Buffers are [rows][cols] of the sampled
// grid: rows = ceil(H/stride_y), cols = ceil((x1-x0)/stride_x).
// InterpolatingRectilinearGenerator::generate (interpolating_rectilinear.rs:121-170) + draw_image, for the pixels
// (x0 + c * stride_x, r * stride_y). The reference fills its three caches lazily; here the grid points the pixels
// need are listed first and evaluated in parallel. `steps` is 0 for every pixel: a blended pixel has no march of its own.
static int render_interpolating(const Scene& s, int stride_x, int stride_y, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, int32_t* counts,
                                atmrt_trace_point* points, int max_points, atmrt_stats* stats, oracle_timing* timing) {
    const atmrt_params& p = s.p;
    const int W = p.width, H = p.height, x0 = p.x0, x1 = p.x1;
    const int cols = (x1 - x0 + stride_x - 1) / stride_x, rows = (H + stride_y - 1) / stride_y;
    const double t0 = now_s();
    std::vector<RayParams> table((size_t)W * H);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) table[(size_t)y * W + x] = get_ray_params(p, x, y);
    const FovData fov = gen_fov_data(p, table);
    // FovData::cache_coords, :185-204
    struct Corner {
        int elev_index, dir_index;
        double rem_elev, rem_dir;
    };
    std::vector<Corner> corner((size_t)rows * cols);
    std::vector<int> elev_ids, dir_ids;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const RayParams rp = table[(size_t)(r * stride_y) * W + (x0 + c * stride_x)];
            const double ef = rp.elevation / fov.min_elev_step, df = rp.direction / fov.min_dir_step;
            Corner k;
            k.elev_index = (int)std::floor(ef), k.dir_index = (int)std::floor(df);
            k.rem_elev = ef - (double)k.elev_index, k.rem_dir = df - (double)k.dir_index;
            corner[(size_t)r * cols + c] = k;
            elev_ids.push_back(k.elev_index), elev_ids.push_back(k.elev_index + 1);
            dir_ids.push_back(k.dir_index), dir_ids.push_back(k.dir_index + 1);
        }
    auto unique = [](std::vector<int>& v) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
    };
    unique(elev_ids), unique(dir_ids);
    auto pos = [](const std::vector<int>& v, int id) { return (size_t)(std::lower_bound(v.begin(), v.end(), id) - v.begin()); };
    // Cache::get_path_cache / get_terrain_cache, :46-84: index * step, to degrees
    std::vector<std::vector<PathElem>> paths(elev_ids.size());
    std::vector<std::vector<TerrainData>> terrains(dir_ids.size());
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < dir_ids.size(); ++i) terrains[i] = gen_terrain_cache(s, to_degrees((double)dir_ids[i] * fov.min_dir_step));
    const double t1 = now_s();
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t i = 0; i < elev_ids.size(); ++i) paths[i] = gen_path_cache(s, to_degrees((double)elev_ids[i] * fov.min_elev_step));
    const double t2 = now_s();
    // Cache::get_pixel, :86-117: the grid points in use
    const size_t ne = elev_ids.size(), nd = dir_ids.size();
    std::vector<char> used(ne * nd, 0);
    for (const Corner& k : corner)
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) used[pos(elev_ids, k.elev_index + i) * nd + pos(dir_ids, k.dir_index + j)] = 1;
    std::vector<std::vector<TracePoint>> grid(ne * nd);
    uint64_t ray_steps = 0, overflows = 0, path_steps = 0;
    for (const auto& pc : paths) path_steps += pc.size() - 1;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ray_steps, overflows)
    for (size_t g = 0; g < ne * nd; ++g) {
        if (!used[g]) continue;
        ZipSource zip{terrains[g % nd], paths[g / nd]};
        uint64_t ov = 0;
        ray_steps += (uint64_t)get_single_pixel(s, zip, &grid[g], &ov);
        overflows += ov;
    }
    uint64_t ntp = 0, nhit = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ntp, nhit)
    for (size_t idx = 0; idx < (size_t)rows * cols; ++idx) {
        const Corner& k = corner[idx];
        const std::vector<TracePoint>* px[4];
        for (int q = 0; q < 4; ++q) px[q] = &grid[pos(elev_ids, k.elev_index + q / 2) * nd + pos(dir_ids, k.dir_index + q % 2)];
        const std::vector<TracePoint> tps = blend_pixels(px, k.rem_elev, k.rem_dir, p.simulation_step);
        ntp += tps.size();
        nhit += tps.empty() ? 0 : 1;
        if (steps) steps[idx] = 0;
        if (counts) counts[idx] = (int32_t)tps.size();
        if (meta) {
            const double nan = std::numeric_limits<double>::quiet_NaN();
            if (tps.empty()) meta[idx] = {nan, nan, nan, nan};
            else meta[idx] = {tps[0].lat, tps[0].lon, tps[0].elevation, tps[0].distance};
        }
        if (points && max_points > 0)
            for (size_t i = 0; i < tps.size() && i < (size_t)max_points; ++i) {
                const TracePoint& t = tps[i];
                atmrt_trace_point& o = points[idx * max_points + i];
                o.lat = t.lat, o.lon = t.lon, o.distance = t.distance, o.elevation = t.elevation, o.path_length = t.path_length;
                o.normal[0] = t.normal.x, o.normal[1] = t.normal.y, o.normal[2] = t.normal.z;
                o.color[0] = t.color.r, o.color[1] = t.color.g, o.color[2] = t.color.b, o.color[3] = t.color.a;
                o.is_terrain = t.is_terrain ? 1 : 0;
                o.step = t.step;
            }
        if (rgb) {
            const Rgb8 c = draw_pixel(p, tps);
            rgb[idx * 3 + 0] = c.c[0], rgb[idx * 3 + 1] = c.c[1], rgb[idx * 3 + 2] = c.c[2];
        }
    }
    const double t3 = now_s();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->ray_steps = ray_steps, stats->trace_points = ntp, stats->pixels_hit = nhit, stats->step_overflows = overflows;
        stats->n_terrain = terrains.empty() ? 0 : (int)terrains[0].size();
        stats->terrain_samples = (uint64_t)terrains.size() * (uint64_t)stats->n_terrain;
        stats->path_steps = path_steps;
        int mx = 0;
        for (const auto& pc : paths) mx = std::max(mx, (int)pc.size());
        stats->n_path_max = mx;
        stats->ms_terrain = (float)((t1 - t0) * 1e3), stats->ms_paths = (float)((t2 - t1) * 1e3), stats->ms_march = (float)((t3 - t2) * 1e3);
        stats->ms_total = (float)((t3 - t0) * 1e3);
    }
    if (timing) {
        timing->s_terrain = t1 - t0, timing->s_paths = t2 - t1, timing->s_pixels = t3 - t2, timing->s_image = 0.0, timing->s_total = t3 - t0;
#ifdef _OPENMP
        timing->threads = omp_get_max_threads();
#else
        timing->threads = 1;
#endif
    }
    return 0;
}

int oracle_render(const atmrt_params* p, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts,
                  const atmrt_object* objects, int nobjects, const uint8_t* const* textures, int stride_x,
                  int stride_y, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, int32_t* counts,
                  atmrt_trace_point* points, int max_points, atmrt_stats* stats, oracle_timing* timing) {
    Scene s;
    if (!build_scene(p, tiles, ntiles, posts, objects, nobjects, textures, &s)) return ATMRT_ERR_INVALID;
    if (stride_x < 1) stride_x = 1;
    if (stride_y < 1) stride_y = 1;
    const int x0 = p->x0, x1 = p->x1, H = p->height;
    const int cols = (x1 - x0 + stride_x - 1) / stride_x, rows = (H + stride_y - 1) / stride_y;
    if (p->generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR)
        return render_interpolating(s, stride_x, stride_y, rgb, meta, steps, counts, points, max_points, stats, timing);
    const bool rectilinear = p->generator == ATMRT_GENERATOR_RECTILINEAR;
    double t0 = now_s();
    std::vector<std::vector<TerrainData>> terrain_cache(cols);
    if (!rectilinear) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int c = 0; c < cols; ++c) terrain_cache[c] = gen_terrain_cache(s, get_ray_dir(*p, x0 + c * stride_x));
    }
    double t1 = now_s();
    std::vector<std::vector<PathElem>> path_cache(rows);
    if (!rectilinear) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int r = 0; r < rows; ++r) path_cache[r] = gen_path_cache(s, get_ray_elev(*p, r * stride_y));
    }
    double t2 = now_s();
    uint64_t ray_steps = 0, ntp = 0, nhit = 0, overflows = 0, path_steps = 0;
    for (int r = 0; r < rows && !rectilinear; ++r) path_steps += path_cache[r].size() - 1;
#pragma omp parallel for schedule(dynamic, 16) collapse(2) reduction(+ : ray_steps, ntp, nhit, overflows, path_steps)
    for (int r = 0; r < rows; ++r) {
        for (int c = 0; c < cols; ++c) {
            std::vector<TracePoint> tps;
            uint64_t ov = 0;
            int consumed;
            if (rectilinear) {  // RectilinearGenerator::gen_pixel, rectilinear.rs:107-124
                PathIteratorSource it(s, get_ray_params(*p, x0 + c * stride_x, r * stride_y));
                consumed = get_single_pixel(s, it, &tps, &ov);
                path_steps += it.path.size();
            } else {
                ZipSource zip{terrain_cache[c], path_cache[r]};
                consumed = get_single_pixel(s, zip, &tps, &ov);
            }
            size_t idx = (size_t)r * cols + c;
            ray_steps += (uint64_t)consumed;
            ntp += tps.size();
            nhit += tps.empty() ? 0 : 1;
            overflows += ov;
            if (steps) steps[idx] = consumed;
            if (counts) counts[idx] = (int32_t)tps.size();
            if (meta) {
                if (tps.empty()) {
                    double nan = std::numeric_limits<double>::quiet_NaN();
                    meta[idx] = {nan, nan, nan, nan};
                } else {
                    meta[idx] = {tps[0].lat, tps[0].lon, tps[0].elevation, tps[0].distance};
                }
            }
            if (points && max_points > 0) {
                for (size_t i = 0; i < tps.size() && i < (size_t)max_points; ++i) {
                    const TracePoint& t = tps[i];
                    atmrt_trace_point& o = points[idx * max_points + i];
                    o.lat = t.lat, o.lon = t.lon, o.distance = t.distance, o.elevation = t.elevation;
                    o.path_length = t.path_length;
                    o.normal[0] = t.normal.x, o.normal[1] = t.normal.y, o.normal[2] = t.normal.z;
                    o.color[0] = t.color.r, o.color[1] = t.color.g, o.color[2] = t.color.b, o.color[3] = t.color.a;
                    o.is_terrain = t.is_terrain ? 1 : 0;
                    o.step = t.step;
                }
            }
            if (rgb) {
                Rgb8 px = draw_pixel(*p, tps);
                rgb[idx * 3 + 0] = px.c[0], rgb[idx * 3 + 1] = px.c[1], rgb[idx * 3 + 2] = px.c[2];
            }
        }
    }
    double t3 = now_s();
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->ray_steps = ray_steps;
        stats->trace_points = ntp;
        stats->pixels_hit = nhit;
        stats->step_overflows = overflows;
        stats->n_terrain = cols && !rectilinear ? (int)terrain_cache[0].size() : 0;
        stats->terrain_samples = (uint64_t)cols * (uint64_t)stats->n_terrain;
        stats->path_steps = path_steps;
        int mx = 0;
        for (int r = 0; r < rows; ++r) mx = std::max(mx, (int)path_cache[r].size());
        stats->n_path_max = mx;
        stats->ms_terrain = (float)((t1 - t0) * 1e3);
        stats->ms_paths = (float)((t2 - t1) * 1e3);
        stats->ms_march = (float)((t3 - t2) * 1e3);
        stats->ms_total = (float)((t3 - t0) * 1e3);
    }
    if (timing) {
        timing->s_terrain = t1 - t0;
        timing->s_paths = t2 - t1;
        timing->s_pixels = t3 - t2;
        timing->s_image = 0.0;  // draw_image is fused into the pixel loop here
        timing->s_total = t3 - t0;
#ifdef _OPENMP
        timing->threads = omp_get_max_threads();
#else
        timing->threads = 1;
#endif
    }
    return 0;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---- probes ------------------------------------------------------------------------------
int oracle_get_elev(const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, const double* lat,
                    const double* lon, int n, double* elev) {
    Terrain t = make_terrain(tiles, ntiles, posts);
    for (int i = 0; i < n; ++i) {
        double e;
        elev[i] = terrain_get_elev(t, lat[i], lon[i], &e) ? e : std::numeric_limits<double>::quiet_NaN();
    }
    return 0;
}

int oracle_coords_at_dist(int earth_model, double radius, double lat0, double lon0, double dir, const double* dist,
                          int n, double* lat, double* lon, double ellipsoid_b) {
    EarthModel m{earth_model, radius, ellipsoid_b};
    DirCalc c = coords_at_dist_calc(m, lat0, lon0, dir);
    for (int i = 0; i < n; ++i) coords_at_dist(c, dist[i], &lat[i], &lon[i]);
    return 0;
}

int oracle_world_directions(int earth_model, double radius, double lat, double lon, double* out9) {
    Dirs d = world_directions(EarthModel{earth_model, radius}, lat, lon);
    out9[0] = d.north.x, out9[1] = d.north.y, out9[2] = d.north.z;
    out9[3] = d.east.x, out9[4] = d.east.y, out9[5] = d.east.z;
    out9[6] = d.up.x, out9[7] = d.up.y, out9[8] = d.up.z;
    return 0;
}

int oracle_as_cartesian(int earth_model, double radius, double lat, double lon, double elev, double* out3, double ellipsoid_b) {
    V3 v = as_cartesian(EarthModel{earth_model, radius, ellipsoid_b}, Coords{lat, lon, elev});
    out3[0] = v.x, out3[1] = v.y, out3[2] = v.z;
    return 0;
}

// ConfColoring::into_coloring light vector, params.rs:243-259
int oracle_light_dir(int earth_model, double radius, double lat, double lon, double direction, double zenith_deg,
                     double light_dir_deg, double* out3) {
    EarthModel m{earth_model, radius};
    double zen = to_radians(zenith_deg), ld = to_radians(light_dir_deg);
    Dirs d = world_directions(m, lat, lon);
    double front_az = to_radians(direction);
    V3 dir_front = d.north * std::cos(front_az) + d.east * std::sin(front_az);
    V3 dir_right = d.east * std::cos(front_az) - d.north * std::sin(front_az);
    V3 l = normalized(-dir_front * std::sin(zen) * std::cos(ld) + dir_right * std::sin(zen) * std::sin(ld) + d.up * std::cos(zen));
    out3[0] = l.x, out3[1] = l.y, out3[2] = l.z;
    return 0;
}

// atm_printer.rs:35-46 + Environment::n
int oracle_atmosphere(const atmrt_atmosphere_def* def, double wavelength, const double* h, int n, double* temperature,
                      double* pressure, double* index) {
    Environment env{};
    if (!atmosphere_from_def(*def, &env.atm)) return ATMRT_ERR_INVALID;
    env.wavelength = wavelength;
    for (int i = 0; i < n; ++i) {
        if (temperature) temperature[i] = atm_temperature(env.atm, h[i]);
        if (pressure) pressure[i] = atm_pressure(env.atm, h[i]);
        if (index) index[i] = env_n(env, h[i]);
    }
    return 0;
}

double oracle_air_index(double lambda, double p, double t, double rh) { return air_index(lambda, p, t, rh); }

// ray_path.rs:65-91 core: the raw stepper states for one ray (x, h) for nsteps steps.
int oracle_ray_path(const atmrt_atmosphere_def* def, double wavelength, int flat, double radius, int straight,
                    double start_h, double ang_deg, double step, int nsteps, double* x, double* h) {
    Environment env{};
    if (!atmosphere_from_def(*def, &env.atm)) return ATMRT_ERR_INVALID;
    env.wavelength = wavelength;
    env.flat = flat;
    env.radius = radius;
    Stepper st(&env, start_h, to_radians(ang_deg), straight != 0);
    st.set_step_size(step);
    for (int i = 0; i < nsteps; ++i) {
        RayState s = st.next();
        x[i] = s.x;
        h[i] = s.h;
    }
    return 0;
}

// RectilinearGenerator::get_ray_params (rectilinear.rs:80-105): elevation and direction of a pixel, degrees
int oracle_ray_params(const atmrt_params* p, int x, int y, double* out2) {
    RayParams r = get_ray_params(*p, x, y);
    out2[0] = to_degrees(r.elevation), out2[1] = to_degrees(r.direction);
    return 0;
}

// gen_path_cache for one row (utils.rs:136-174)
int oracle_path_cache(const atmrt_params* p, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts,
                      int y, int capacity, double* dist, double* elev, double* path_length, int* n) {
    Scene s;
    if (!build_scene(p, tiles, ntiles, posts, nullptr, 0, nullptr, &s)) return ATMRT_ERR_INVALID;
    std::vector<PathElem> path = gen_path_cache(s, get_ray_elev(*p, y));
    *n = (int)path.size();
    for (int i = 0; i < (int)path.size() && i < capacity; ++i) {
        dist[i] = path[i].dist, elev[i] = path[i].elev, path_length[i] = path[i].path_length;
    }
    return 0;
}

// gen_terrain_cache for one column (utils.rs:176-199)
int oracle_terrain_cache(const atmrt_params* p, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts,
                         const atmrt_object* objects, int nobjects, int x, int capacity, double* lat, double* lon,
                         double* elev, double* normal, uint64_t* close, int* n) {
    Scene s;
    if (!build_scene(p, tiles, ntiles, posts, objects, nobjects, nullptr, &s)) return ATMRT_ERR_INVALID;
    std::vector<TerrainData> t = gen_terrain_cache(s, get_ray_dir(*p, x));
    *n = (int)t.size();
    for (int i = 0; i < (int)t.size() && i < capacity; ++i) {
        lat[i] = t[i].lat, lon[i] = t[i].lon, elev[i] = t[i].elev;
        normal[3 * i] = t[i].normal.x, normal[3 * i + 1] = t[i].normal.y, normal[3 * i + 2] = t[i].normal.z;
        if (close) close[i] = t[i].objects_close;
    }
    return 0;
}

int oracle_ray_angles(const atmrt_params* p, double* dir /*[W]*/, double* elev /*[H]*/) {
    for (int x = 0; x < p->width; ++x) dir[x] = get_ray_dir(*p, x);
    for (int y = 0; y < p->height; ++y) elev[y] = get_ray_elev(*p, y);
    return 0;
}

// ResultPixel.elevation_angle / azimuth of every pixel of the column block, [H][x1-x0] each.
// Fast generator (fast.rs:67-76): elevation_angle = get_ray_elev(y); azimuth = get_ray_dir(x) wrapped ONCE into
// [0, 360) (`if azimuth < 0.0 { += 360.0 } else if azimuth >= 360.0 { -= 360.0 }`).
// Rectilinear generator (rectilinear.rs:78-116): the pixel's own (elevation, direction).to_degrees(), not wrapped.
int oracle_pixel_angles(const atmrt_params* p, double* elevation_angle, double* azimuth) {
    const int wl = p->x1 - p->x0;
    FovData fov{0.0, 0.0};
    if (p->generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR) {
        std::vector<RayParams> table((size_t)p->width * p->height);
        for (int y = 0; y < p->height; ++y)
            for (int x = 0; x < p->width; ++x) table[(size_t)y * p->width + x] = get_ray_params(*p, x, y);
        fov = gen_fov_data(*p, table);
    }
    for (int y = 0; y < p->height; ++y) {
        for (int c = 0; c < wl; ++c) {
            double el, az;
            if (p->generator == ATMRT_GENERATOR_RECTILINEAR) {
                const RayParams r = get_ray_params(*p, p->x0 + c, y);
                el = to_degrees(r.elevation), az = to_degrees(r.direction);
            } else if (p->generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR) {
                // interpolate(), interpolating_rectilinear.rs:405-417: the four grid points' angles (Cache::get_pixel, :97-107:
                // azimuth wrapped once) weighted with the remainders
                const RayParams r = get_ray_params(*p, p->x0 + c, y);
                const double ef = r.elevation / fov.min_elev_step, df = r.direction / fov.min_dir_step;
                const int ei = (int)std::floor(ef), di = (int)std::floor(df);
                const double re = ef - (double)ei, rd = df - (double)di;
                double e4[4], a4[4];
                for (int q = 0; q < 4; ++q) {
                    e4[q] = to_degrees((double)(ei + q / 2) * fov.min_elev_step);
                    double a = to_degrees((double)(di + q % 2) * fov.min_dir_step);
                    if (a < 0.0) a += 360.0;
                    else if (a >= 360.0) a -= 360.0;
                    a4[q] = a;
                }
                el = e4[0] * (1.0 - re) * (1.0 - rd) + e4[1] * (1.0 - re) * rd + e4[2] * re * (1.0 - rd) + e4[3] * re * rd;
                az = a4[0] * (1.0 - re) * (1.0 - rd) + a4[1] * (1.0 - re) * rd + a4[2] * re * (1.0 - rd) + a4[3] * re * rd;
            } else {
                el = get_ray_elev(*p, y);
                az = get_ray_dir(*p, p->x0 + c);
                if (az < 0.0) az += 360.0;
                else if (az >= 360.0) az -= 360.0;
            }
            if (elevation_angle) elevation_angle[(size_t)y * wl + c] = el;
            if (azimuth) azimuth[(size_t)y * wl + c] = az;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// The three text dumpers, restated (the reference's only text windows onto the external crates: a real binary's
// stdout can be diffed against these files). `{}` of an f64 is atmrt_fmt_f64 (include/atmrt_fmt.h). Each writes the
// text the reference prints to stdout into `path` and returns 0.
// ---------------------------------------------------------------------------------------------
// ray_path.rs:6-106 (`output-ray-paths`): rays are always refracted (cast_ray_stepper(height, ang, false), :71).
int oracle_output_ray_paths(const atmrt_params* p, double height, double min_ang, double max_ang, double step, double ray_step, double cutoff,
                            double output_step, const char* path) {
    if (!(step > 0.0)) return ATMRT_ERR_INVALID;  // assert!(step > 0.0), :53
    Environment env{};
    if (!atmosphere_from_def(p->atmosphere, &env.atm)) return ATMRT_ERR_INVALID;
    const EarthModel model{p->earth_model, p->radius, p->ellipsoid_b};
    env.flat = shape_flat(model), env.radius = shape_radius(model), env.wavelength = p->wavelength;
    std::vector<std::vector<double>> rays;
    std::vector<double> xs{0.0};
    for (double ang = min_ang; ang <= max_ang; ang += step) {  // :65-94
        Stepper stepper(&env, height, to_radians(ang), false);
        stepper.set_step_size(ray_step);
        std::vector<double> ray{height};
        for (;;) {
            const RayState st = stepper.next();
            if (std::floor((st.x - ray_step / 2.0) / output_step) != std::floor((st.x + ray_step / 2.0) / output_step)) {
                if (ang == min_ang) xs.push_back(st.x);
                ray.push_back(st.h);
            }
            if (st.x >= cutoff) break;
        }
        rays.push_back(ray);
    }
    FILE* f = fopen(path, "wb");
    if (!f) return ATMRT_ERR_IO;
    for (size_t i = 0; i < xs.size(); ++i) {  // :97-103
        fprintf(f, "%s\t", atmrt_fmt_f64(xs[i]).c_str());
        for (const auto& ray : rays) fprintf(f, "%s\t", atmrt_fmt_f64(ray[i]).c_str());
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

// elev_profile.rs:9-67 (`output-elev-profile`); Terrain::from_folder's "Detected N terrain files" line (terrain/mod.rs:80)
// goes to the same stdout first.
int oracle_output_elev_profile(const atmrt_params* p, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, double azim,
                               double step, double cutoff, const char* path) {
    if (!(step > 0.0)) return ATMRT_ERR_INVALID;
    const Terrain terrain = make_terrain(tiles, ntiles, posts);
    const EarthModel model{p->earth_model, p->radius, p->ellipsoid_b};
    const DirCalc calc = coords_at_dist_calc(model, p->latitude, p->longitude, azim);
    FILE* f = fopen(path, "wb");
    if (!f) return ATMRT_ERR_IO;
    fprintf(f, "Detected %d terrain files\n", ntiles);
    for (double x = 0.0; x <= cutoff; x += step) {  // :53-64
        double lat, lon, elev = 0.0;
        coords_at_dist(calc, x, &lat, &lon);
        if (!terrain_get_elev(terrain, lat, lon, &elev)) elev = 0.0;
        fprintf(f, "%s\t%s\n", atmrt_fmt_f64(x).c_str(), atmrt_fmt_f64(elev).c_str());
    }
    fclose(f);
    return 0;
}

// atm_printer.rs:6-49 (`output-atm`)
int oracle_output_atm(const atmrt_atmosphere_def* def, double min_alt, double max_alt, double step, int celsius, const char* path) {
    if (!(step > 0.0)) return ATMRT_ERR_INVALID;
    Atmosphere atm;
    if (!atmosphere_from_def(*def, &atm)) return ATMRT_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return ATMRT_ERR_IO;
    for (double alt = min_alt; alt <= max_alt; alt += step)  // :35-46
        fprintf(f, "%s %s %s %s\n", atmrt_fmt_f64(alt).c_str(), atmrt_fmt_f64(atm_temperature(atm, alt) - (celsius ? 273.15 : 0.0)).c_str(),
                atmrt_fmt_f64(atm_pressure(atm, alt)).c_str(), atmrt_fmt_f64(atm.humidity).c_str());
    fclose(f);
    return 0;
}

// Object::check_collision for one segment; returns number of collisions (<= 8 written)
int oracle_check_collision(const atmrt_object* obj, const uint8_t* texture, double obj_alt_abs, int earth_model,
                           double radius, const double* p1 /*lat,lon,elev*/, const double* p2, double* props,
                           double* normals /*[8][3]*/, double* colors /*[8][4]*/) {
    Object o{};
    o.kind = obj->kind;
    o.position = {obj->latitude, obj->longitude, obj_alt_abs};
    o.r1 = obj->r1, o.r2 = obj->r2, o.width = obj->width, o.height = obj->height;
    o.color = {obj->color[0], obj->color[1], obj->color[2], obj->color[3]};
    o.tex_w = obj->texture_width, o.tex_h = obj->texture_height, o.tex = texture;
    std::vector<Collision> out;
    check_collision(o, EarthModel{earth_model, radius}, Coords{p1[0], p1[1], p1[2]}, Coords{p2[0], p2[1], p2[2]}, &out);
    int n = 0;
    for (const Collision& c : out) {
        if (n >= 8) break;
        props[n] = c.prop;
        normals[3 * n] = c.normal.x, normals[3 * n + 1] = c.normal.y, normals[3 * n + 2] = c.normal.z;
        colors[4 * n] = c.color.r, colors[4 * n + 1] = c.color.g, colors[4 * n + 2] = c.color.b, colors[4 * n + 3] = c.color.a;
        ++n;
    }
    return (int)out.size();
}

// draw_image for one explicit trace-point list (renderer/mod.rs:395-411)
int oracle_draw_pixel(const atmrt_params* p, const atmrt_trace_point* pts, int n, uint8_t* rgb) {
    std::vector<TracePoint> tps;
    for (int i = 0; i < n; ++i) {
        TracePoint t{pts[i].lat, pts[i].lon, pts[i].distance, pts[i].elevation, pts[i].path_length,
                     V3{pts[i].normal[0], pts[i].normal[1], pts[i].normal[2]}, pts[i].is_terrain != 0,
                     Color{pts[i].color[0], pts[i].color[1], pts[i].color[2], pts[i].color[3]}, pts[i].step};
        tps.push_back(t);
    }
    Rgb8 px = draw_pixel(*p, tps);
    rgb[0] = px.c[0], rgb[1] = px.c[1], rgb[2] = px.c[2];
    return 0;
}

// DTED file reader (external dted 0.2 restated from MIL-PRF-89020B; PARITY UNPINNED).
// Pass posts == NULL to read only the header. Returns 0 or ATMRT_ERR_IO / ATMRT_ERR_INVALID.
static double parse_dms(const unsigned char* s, int deg_digits) {
    auto num = [&](int off, int n) {
        int v = 0;
        for (int i = 0; i < n; ++i) v = v * 10 + (s[off + i] - '0');
        return v;
    };
    int deg = num(0, deg_digits), mn = num(deg_digits, 2), sc = num(deg_digits + 2, 2);
    char hemi = (char)s[deg_digits + 4];
    double v = (double)deg + (double)mn / 60.0 + (double)sc / 3600.0;
    return (hemi == 'W' || hemi == 'S') ? -v : v;
}

int oracle_read_dted(const char* path, atmrt_tile_desc* desc, int16_t* posts, size_t capacity) {
    FILE* f = fopen(path, "rb");
    if (!f) return ATMRT_ERR_IO;
    unsigned char uhl[80];
    if (fread(uhl, 1, 80, f) != 80 || memcmp(uhl, "UHL1", 4) != 0) {
        fclose(f);
        return ATMRT_ERR_INVALID;
    }
    auto num = [&](int off, int n) {
        int v = 0;
        for (int i = 0; i < n; ++i) v = v * 10 + (uhl[off + i] - '0');
        return v;
    };
    desc->min_lon = parse_dms(uhl + 4, 3);
    desc->min_lat = parse_dms(uhl + 12, 3);
    desc->lon_interval = (double)num(20, 4) / 10.0;
    desc->lat_interval = (double)num(24, 4) / 10.0;
    desc->nlon = num(47, 4);
    desc->nlat = num(51, 4);
    desc->lat0 = (int32_t)as_i16(desc->min_lat);
    desc->lon0 = (int32_t)as_i16(desc->min_lon);
    if (!posts) {
        fclose(f);
        return 0;
    }
    size_t need = (size_t)desc->nlon * desc->nlat;
    if (capacity < need) {
        fclose(f);
        return ATMRT_ERR_INVALID;
    }
    const long data_off = 80 + 648 + 2700;
    size_t rec = 12 + 2 * (size_t)desc->nlat;
    std::vector<unsigned char> buf(rec);
    fseek(f, data_off, SEEK_SET);
    for (int i = 0; i < desc->nlon; ++i) {
        if (fread(buf.data(), 1, rec, f) != rec || buf[0] != 0xAA) {
            fclose(f);
            return ATMRT_ERR_INVALID;
        }
        for (int j = 0; j < desc->nlat; ++j) {
            unsigned v = ((unsigned)buf[8 + 2 * j] << 8) | buf[9 + 2 * j];
            int mag = (int)(v & 0x7fff);
            posts[(size_t)i * desc->nlat + j] = (int16_t)((v & 0x8000) ? -mag : mag);  // signed magnitude
        }
    }
    fclose(f);
    return 0;
}

}  // extern "C"
