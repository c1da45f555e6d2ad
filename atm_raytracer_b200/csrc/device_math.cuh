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

// DtedData::get_elev (external; bilinear form witnessed by terrain/geotiff.rs:61-100)
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
    const int16_t* p = T.posts;
    double e00 = (double)__ldg(p + post_index(t, lon_int, lat_int));
    double e01 = (double)__ldg(p + post_index(t, lon_int, lat_int + 1));
    double e10 = (double)__ldg(p + post_index(t, lon_int + 1, lat_int));
    double e11 = (double)__ldg(p + post_index(t, lon_int + 1, lat_int + 1));
    *out = e00 * (1.0 - lon_frac) * (1.0 - lat_frac) + e01 * (1.0 - lon_frac) * lat_frac +
           e10 * lon_frac * (1.0 - lat_frac) + e11 * lon_frac * lat_frac;
    return true;
}

// Terrain::get_elev, terrain/mod.rs:120-126
__device__ __forceinline__ bool terrain_get_elev(const DevTerrain& T, double latitude, double longitude, double* out) {
    int klat = as_i16(floor(latitude)) - T.lat_min;
    int klon = as_i16(floor(longitude)) - T.lon_min;
    if (klat < 0 || klat >= T.nlat_tiles || klon < 0 || klon >= T.nlon_tiles) return false;
    int ti = __ldg(T.lookup + klat * T.nlon_tiles + klon);
    if (ti < 0) return false;
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

// EarthModel::world_directions, mod.rs:31-57
__device__ __forceinline__ Dirs world_directions(int model, double lat, double lon) {
    double sinlon, coslon;
    sincos(to_radians(lon), &sinlon, &coslon);
    if (model == ATMRT_EARTH_FLAT_DISTORTED) {
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
__device__ __forceinline__ V3 as_cartesian_sc(int model, double radius, double lat, double elev, double sinlat,
                                               double coslat, double sinlon, double coslon) {
    if (model == ATMRT_EARTH_FLAT_DISTORTED) {
        double r = (90.0 - lat) * DEGREE_DISTANCE;
        return {r * coslon, r * sinlon, elev};
    }
    double r = radius + elev;  // spherical_to_cartesian, mod.rs:148-153
    return {r * coslat * coslon, r * coslat * sinlon, r * sinlat};
}
__device__ __forceinline__ V3 as_cartesian(int model, double radius, double lat, double lon, double elev) {
    double sinlat = 0.0, coslat = 1.0, sinlon, coslon;
    sincos(to_radians(lon), &sinlon, &coslon);
    if (model != ATMRT_EARTH_FLAT_DISTORTED) sincos(to_radians(lat), &sinlat, &coslat);
    return as_cartesian_sc(model, radius, lat, elev, sinlat, coslat, sinlon, coslon);
}

// SphericalCalc::coords_at_dist (directional_calc.rs:71-86) with sin/cos of dist/radius given.
__device__ __forceinline__ void spherical_walk(V3 pos, V3 dir, double sinang, double cosang, double* lat, double* lon) {
    V3 fpos = pos * cosang + dir * sinang;
    *lat = to_degrees(asin(fpos.z));
    *lon = to_degrees(atan2(fpos.y, fpos.x));
}

}  // namespace atmrt
