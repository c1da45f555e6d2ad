// device_math.cuh -- f64 device functions for the earth models and the terrain store.
//
// Every function cites the reference code it restates (paths relative to /root/reference/src).
// The library is compiled with --fmad=false: Rust never contracts a*b+c, and the CPU oracle is
// built with -ffp-contract=off, so the only CPU/GPU differences left are libm ulps.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/atmrt.h"

namespace atmrt {

constexpr double PI = 3.14159265358979323846;
constexpr double DEGREE_DISTANCE = 10000000.0 / 90.0;  // utils/earth_model/mod.rs:12
// sin/cos of f64::to_radians(90.0) as glibc (and therefore the reference on Linux) returns them.
constexpr double SIN_90 = 1.0;
constexpr double COS_90 = 6.123233995736766e-17;
constexpr double NORMAL_DIFF = 15.0;  // generators/utils.rs:16

__host__ __device__ __forceinline__ double to_radians(double d) { return d * (PI / 180.0); }
__host__ __device__ __forceinline__ double to_degrees(double r) { return r * (180.0 / PI); }

struct V3 {
    double x, y, z;
};
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__host__ __device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
__host__ __device__ __forceinline__ V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__host__ __device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Rust `as` casts saturate and send NaN to 0; CUDA conversions of
// out-of-range values are not defined that way, so clamp explicitly.
__device__ __forceinline__ uint8_t as_u8(double v) {
    if (!(v == v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)(int)v;
}
__device__ __forceinline__ int as_i16(double v) {
    if (!(v == v)) return 0;
    if (v <= -32768.0) return -32768;
    if (v >= 32767.0) return 32767;
    return (int)v;
}
__device__ __forceinline__ bool in_range(double lo, double hi, double v) { return lo <= v && v < hi; }

// ---------------------------------------------------------------------------------------------
// Terrain: packed, micro-tiled DTED grids in HBM
// ---------------------------------------------------------------------------------------------
// Each decoded tile ([lon][lat] i16, external crate dted 0.2) is stored as 8x8-post micro-tiles of
// 128 B (one L2 line): a bilinear tap's 2x2 footprint falls inside one line 77% of the time and a
// warp sampling neighbouring azimuths/distances touches a handful of lines instead of one 2.4 KB
// longitude line per post.
constexpr int MT = 8;  // micro-tile edge in posts

struct DevTile {
    double min_lat, min_lon, max_lat, max_lon;
    double lat_interval, lon_interval;  // arc-seconds
    double inv_lat_interval, inv_lon_interval;  // RN(1 / interval), for the exact division below
    int nlat, nlon;
    int mt_lat;           // micro-tiles along latitude
    int _pad;
    long long post_offset;  // first post of this tile in the packed i16 array
};

struct DevTerrain {
    const DevTile* tiles;
    const int* lookup;  // [(klat-lat_min)*nlon_tiles + (klon-lon_min)] -> tile index or -1
    const int16_t* posts;
    int lat_min, lon_min, nlat_tiles, nlon_tiles;
    int ntiles;
    // A REGULAR terrain -- every tile covers exactly its one-degree cell [lat0, lat0 + 1] x [lon0, lon0 + 1] with the
    // same grid (what a folder of DTED files of one level is): the descriptor of a tile is then known without loading
    // it (only its offset in the packed posts differs) and a sample is always inside the tile its floor selects.
    int regular;
    int r_nlat, r_nlon, r_mt_lat;
    double r_lat_interval, r_lon_interval, r_inv_lat_interval, r_inv_lon_interval;
};

__device__ __forceinline__ long long post_index(const DevTile& t, int ilon, int ilat) {
    return t.post_offset + ((long long)((ilon >> 3) * t.mt_lat + (ilat >> 3)) << 6) + ((ilon & 7) << 3) + (ilat & 7);
}

// a / b correctly rounded from y = RN(1/b) (Markstein: q0 = RN(a y), r = a - b q0 exactly by FMA,
// q = RN(q0 + r y)); a finite, b a normal number whose significand is not all ones. Same bits as the
// IEEE division the reference performs, without the division's slow-path bookkeeping.
__device__ __forceinline__ double div_by(double a, double b, double y) {
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

// The bilinear form of DtedData::get_elev (external; witnessed by terrain/geotiff.rs:61-100) on the 2 x 2 posts whose
// lower-left one is (lon_int, lat_int) of the tile whose posts start at `off`. The four indices of the micro-tiled
// layout follow from the first: +1 / +8 inside an 8 x 8 micro-tile, a step into the next micro-tile at its edge.
__device__ __forceinline__ double bilinear_posts(const int16_t* __restrict__ p, long long off, int mt_lat, int lon_int, int lat_int, double lon_frac, double lat_frac) {
    const long long i00 = off + ((long long)((lon_int >> 3) * mt_lat + (lat_int >> 3)) << 6) + ((lon_int & 7) << 3) + (lat_int & 7);
    const int dlat = (lat_int & 7) == 7 ? 57 : 1;                 // next micro-tile along latitude: + 64 - 7
    const int dlon = (lon_int & 7) == 7 ? mt_lat * 64 - 56 : 8;   // next micro-tile along longitude: + 64 mt_lat - 7 * 8
    const double e00 = (double)__ldg(p + i00), e01 = (double)__ldg(p + i00 + dlat);
    const double e10 = (double)__ldg(p + i00 + dlon), e11 = (double)__ldg(p + i00 + dlon + dlat);
    return e00 * (1.0 - lon_frac) * (1.0 - lat_frac) + e01 * (1.0 - lon_frac) * lat_frac + e10 * lon_frac * (1.0 - lat_frac) + e11 * lon_frac * lat_frac;
}

// DtedData::get_elev
__device__ __forceinline__ bool tile_get_elev(const DevTerrain& T, const DevTile& t, double lat, double lon, double* out) {
    if (lat < t.min_lat || lat > t.max_lat || lon < t.min_lon || lon > t.max_lon) return false;
    double plat = div_by((lat - t.min_lat) * 3600.0, t.lat_interval, t.inv_lat_interval);
    double plon = div_by((lon - t.min_lon) * 3600.0, t.lon_interval, t.inv_lon_interval);
    int lat_int = (int)plat, lon_int = (int)plon;  // in [0, n-1] after the bounds check
    double lat_frac = plat - (double)lat_int, lon_frac = plon - (double)lon_int;
    if (lat_int == t.nlat - 1) {
        lat_int -= 1;
        lat_frac += 1.0;
    }
    if (lon_int == t.nlon - 1) {
        lon_int -= 1;
        lon_frac += 1.0;
    }
    *out = bilinear_posts(T.posts, t.post_offset, t.mt_lat, lon_int, lat_int, lon_frac, lat_frac);
    return true;
}

// Terrain::get_elev, terrain/mod.rs:120-126
__device__ __forceinline__ bool terrain_get_elev(const DevTerrain& T, double latitude, double longitude, double* out) {
    const double flat = floor(latitude), flon = floor(longitude);
    int klat = as_i16(flat) - T.lat_min;
    int klon = as_i16(flon) - T.lon_min;
    if (klat < 0 || klat >= T.nlat_tiles || klon < 0 || klon >= T.nlon_tiles) return false;
    int ti = __ldg(T.lookup + klat * T.nlon_tiles + klon);
    if (ti < 0) return false;
    if (T.regular && flat == (double)(klat + T.lat_min) && flon == (double)(klon + T.lon_min)) {  // (not a saturated or NaN key)
        // tile_get_elev with the descriptor of a regular tile: min = the floor (its key), max = min + 1, so the bounds
        // hold by construction; the same operations on the same values otherwise
        const double plat = div_by((latitude - flat) * 3600.0, T.r_lat_interval, T.r_inv_lat_interval);
        const double plon = div_by((longitude - flon) * 3600.0, T.r_lon_interval, T.r_inv_lon_interval);
        int lat_int = (int)plat, lon_int = (int)plon;
        double lat_frac = plat - (double)lat_int, lon_frac = plon - (double)lon_int;
        if (lat_int == T.r_nlat - 1) {
            lat_int -= 1;
            lat_frac += 1.0;
        }
        if (lon_int == T.r_nlon - 1) {
            lon_int -= 1;
            lon_frac += 1.0;
        }
        *out = bilinear_posts(T.posts, __ldg(&T.tiles[ti].post_offset), T.r_mt_lat, lon_int, lat_int, lon_frac, lat_frac);
        return true;
    }
    return tile_get_elev(T, T.tiles[ti], latitude, longitude, out);
}
__device__ __forceinline__ double elev_or_zero(const DevTerrain& T, double lat, double lon) {  // .unwrap_or(0.0)
    double e;
    return terrain_get_elev(T, lat, lon, &e) ? e : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Earth models
// ---------------------------------------------------------------------------------------------
struct Dirs {
    V3 north, east, up;
};

// spherical_directions (mod.rs:155-172) from already evaluated sin/cos.
__device__ __forceinline__ Dirs spherical_directions_sc(double sinlat, double coslat, double sinlon, double coslon) {
    Dirs d;
    d.up = {coslat * coslon, coslat * sinlon, sinlat};
    d.north = {-sinlat * coslon, -sinlat * sinlon, coslat};
    d.east = {-sinlon, coslon, 0.0};
    return d;
}

// The geometry of an EarthModel (earth_model/mod.rs:19-28) as the kernels need it.
enum { WALK_SPHERICAL = 0, WALK_FLDS = 1, WALK_AZEQ = 2, WALK_ELLIPSOID = 3 };
struct DevEarth {
    int model;      // atmrt_earth_model
    int flat_dirs;  // world_directions / as_cartesian of the azimuthal-equidistant plane (mod.rs:37-51, 80-91)
    int walker;     // coords_at_dist_calc (mod.rs:114-145): which DirectionalCalc walks the azimuth
    int _pad;
    double radius;  // Spherical: radius; Ellipsoid: a; ObserverAe: proj_radius
    double b, f, e2;  // Ellipsoid: b, (a - b) / a, 1 - b^2 / a^2
};

// EarthModel::world_directions, mod.rs:31-57
__device__ __forceinline__ Dirs world_directions(const DevEarth& E, double lat, double lon) {
    double sinlon, coslon;
    sincos(to_radians(lon), &sinlon, &coslon);
    if (E.flat_dirs) {
        Dirs d;
        d.north = {-coslon, -sinlon, 0.0};
        d.east = {-sinlon, coslon, 0.0};
        d.up = {0.0, 0.0, 1.0};
        return d;
    }
    double sinlat, coslat;
    sincos(to_radians(lat), &sinlat, &coslat);
    return spherical_directions_sc(sinlat, coslat, sinlon, coslon);
}

// EarthModel::as_cartesian (mod.rs:59-93) from already evaluated sin/cos of lat/lon.
__device__ __forceinline__ V3 as_cartesian_sc(const DevEarth& E, double lat, double elev, double sinlat, double coslat, double sinlon,
                                               double coslon) {
    if (E.flat_dirs) {
        double r = (90.0 - lat) * DEGREE_DISTANCE;
        return {r * coslon, r * sinlon, elev};
    }
    if (E.model == ATMRT_EARTH_ELLIPSOID) {  // mod.rs:70-79
        const double n = E.radius / sqrt(1.0 - E.e2 * (sinlat * sinlat));
        return {(n + elev) * coslat * coslon, (n + elev) * coslat * sinlon, (n * (1.0 - E.e2) + elev) * sinlat};
    }
    double r = E.radius + elev;  // spherical_to_cartesian, mod.rs:148-153
    return {r * coslat * coslon, r * coslat * sinlon, r * sinlat};
}
__device__ __forceinline__ V3 as_cartesian(const DevEarth& E, double lat, double lon, double elev) {
    double sinlat = 0.0, coslat = 1.0, sinlon, coslon;
    sincos(to_radians(lon), &sinlon, &coslon);
    if (!E.flat_dirs) sincos(to_radians(lat), &sinlat, &coslat);
    return as_cartesian_sc(E, lat, elev, sinlat, coslat, sinlon, coslon);
}

// EllipsoidCalc (directional_calc.rs:88-185): Vincenty's direct formula, after NGS "inverse.pdf".
struct EllipsoidCalc {
    double red_lat, lon, cos_az1, sin_az1, sin_alfa, sig1, cap_a, cap_b, cap_c;
};
__device__ __forceinline__ EllipsoidCalc ellipsoid_calc(const DevEarth& E, double lat_deg, double lon_deg, double dir_deg) {  // ::new, :104-135
    const double a = E.radius, b = E.b, f = E.f;
    const double lat = to_radians(lat_deg), az1 = to_radians(dir_deg);
    EllipsoidCalc c;
    c.lon = to_radians(lon_deg);
    c.red_lat = atan((1.0 - f) * tan(lat));
    sincos(az1, &c.sin_az1, &c.cos_az1);
    c.sig1 = atan(tan(c.red_lat) / c.cos_az1);
    const double alfa = asin(cos(c.red_lat) * c.sin_az1);
    double ca;
    sincos(alfa, &c.sin_alfa, &ca);
    const double u2 = ca * ca * (a * a - b * b) / (b * b);
    c.cap_a = 1.0 + u2 / 256.0 * (64.0 + u2 * (-12.0 + 5.0 * u2));
    c.cap_b = u2 / 512.0 * (128.0 + u2 * (-64.0 + 37.0 * u2));
    c.cap_c = f / 16.0 * (ca * ca) * (4.0 + f * (4.0 - 3.0 * (ca * ca)));
    return c;
}
__device__ __forceinline__ void ellipsoid_walk(const DevEarth& E, const EllipsoidCalc& c, double dist, double* lat, double* lon) {  // :139-184
    const double s0 = dist / E.b / c.cap_a;
    double sig = s0;
    for (int it = 0; it < 64; ++it) {  // the reference loops until |d sigma| < 1e-10 (2-4 rounds); the cap only guards NaN input
        const double sigm = 2.0 * c.sig1 + sig;
        double ss, cs;
        sincos(sig, &ss, &cs);
        const double csm = cos(sigm);
        const double dsig = c.cap_b * ss * (csm + c.cap_b / 4.0 * cs * (-1.0 + 2.0 * (csm * csm)));
        const double new_sig = s0 + dsig;
        const double ds = fabs(new_sig - sig);
        sig = new_sig;
        if (ds < 1e-10) break;
    }
    const double sigm = 2.0 * c.sig1 + sig;
    double sr, cr, ss, cs;
    sincos(c.red_lat, &sr, &cr);
    sincos(sig, &ss, &cs);
    const double t = sr * ss - cr * cs * c.cos_az1;
    const double lat2 = atan((sr * cs + cr * ss * c.cos_az1) / ((1.0 - E.f) * sqrt(c.sin_alfa * c.sin_alfa + t * t)));
    const double lambda = atan(ss * c.sin_az1 / (cr * cs - sr * ss * c.cos_az1));
    const double csm = cos(sigm);
    const double dl = lambda - (1.0 - c.cap_c) * E.f * c.sin_alfa * (sig + c.cap_c * ss * (csm + c.cap_c * cs * (-1.0 + 2.0 * (csm * csm))));
    *lat = to_degrees(lat2);
    *lon = to_degrees(c.lon + dl);
}

// AzEqCalc::coords_at_dist (directional_calc.rs:20-27): a straight line in the azimuthal-equidistant plane.
__device__ __forceinline__ void azeq_walk(V3 pos, V3 dir_v, double dist, double* lat, double* lon) {
    const V3 pos2 = pos + dir_v * dist;
    *lon = to_degrees(atan2(pos2.y, pos2.x));
    const double r = sqrt(pos2.x * pos2.x + pos2.y * pos2.y);
    *lat = 90.0 - r / DEGREE_DISTANCE;
}

// SphericalCalc::coords_at_dist (directional_calc.rs:71-86) with sin/cos of dist/radius given.
__device__ __forceinline__ void spherical_walk(V3 pos, V3 dir, double sinang, double cosang, double* lat, double* lon) {
    V3 fpos = pos * cosang + dir * sinang;
    *lat = to_degrees(asin(fpos.z));
    *lon = to_degrees(atan2(fpos.y, fpos.x));
}

}  // namespace atmrt
