// kernels.cuh -- the CUDA kernels of the panorama ray march (sm_100a, f64).
//
//   Stage A  k_terrain_profile : gen_terrain_cache  (generators/utils.rs:176-199, 72-88, 15-40)
//   Stage B  k_ray_paths       : gen_path_cache     (generators/utils.rs:136-174)
//            k_pyramid         : min/max (and objects_close OR) pyramids over both caches
//   Stage C  k_march           : get_single_pixel   (generators/utils.rs:201-289) fused with
//                                Object::check_collision, ColoringMethod::color_for_pixel,
//                                renderer::draw_image compositing and the per-pixel metadata.
#pragma once

#include "device_atm.cuh"
#include "device_paths.cuh"
#include "device_math.cuh"
#include "device_shade.cuh"

namespace atmrt {

// Path cache layout: groups of PATH_ROWS adjacent rows, step-major inside a group: [row / 4][k][row % 4].
// A warp of the march (32 adjacent rows, one step) reads 8 whole 32-byte sectors; a warp of the horizon
// sweep (one row group, 32 consecutive steps) reads 1 KB of contiguous memory; the ray-path kernel's
// rows store next to each other.
constexpr int PATH_ROWS = 4;
__host__ __device__ __forceinline__ size_t path_index(int n_t, int k, int y) {
    return ((size_t)(y / PATH_ROWS) * (size_t)n_t + (size_t)k) * PATH_ROWS + (size_t)(y % PATH_ROWS);
}

constexpr int CHUNK = 32;  // march steps per level-1 chunk == warp width
constexpr unsigned FULL = 0xffffffffu;

// Everything a kernel needs about the scene, passed by value as a __grid_constant__ argument.
struct DevScene {
    double lat0, lon0, direction, tilt, fov, max_distance, step;
    double radius;  // of the ray physics (EarthModel::to_shape, mod.rs:95-112); the geometry's own is earth.radius
    DevEarth earth;
    double sin_diff, cos_diff;  // sin/cos(NORMAL_DIFF / earth.radius), spherical find_normal
    double tan_diff, versin_diff, diff_deg;  // tan(delta), 1 - cos(delta), delta in degrees
    atmrt_altitude altitude;
    int earth_model, straight, flat;
    int width, height, x0, x1;
    int generator;  // atmrt_generator
    // explicit ray angles in degrees instead of the image's (the grid of the InterpolatingRectilinear generator): one
    // elevation per row, one direction per column; null for the image's own
    const double* row_elev_deg;
    const double* col_dir_deg;
    int n_x;     // entries of path_x / path_dxr (n_t + 2)
    int n_sc_off;  // offset of walk_sc behind dist_k
    int n_t;     // terrain samples per column (N_t)
    int path_k_far;  // first path element with dist > max_distance (n_t if none): the row-independent half of utils.rs:167
    int n_pad;   // row stride of the [column][k] and [row][k] caches
    int n1;      // level-1 chunks  = ceil(n_t / 32)
    int n1_pad;  // row stride of the level-1 pyramids
    int n2;      // level-2 chunks  = ceil(n1 / 32)
    int h_pad;   // row stride of the step-major [k][row] path cache and of the node-major path pyramids
    int nobjects;
    int anchor_shift;  // log2 of the samples per walk anchor (k_terrain_profile), 0: no anchors
    int n_anchor, _pad;  // anchors per column
    DevAtmosphere atm;
    DevShade shade;
};

struct WalkAnchor;

struct DevBuffers {
    const double* dist_k;  // [n_t] accumulated `distance += step` (utils.rs:191-196), host-computed
    double* colcalc;       // [wl][8]: spherical {dir.xyz, pos.xyz}; flat {cos_az, sin_az, cos_lat0}
    // Stage A cache, [wl][n_pad]
    DevTerrain terrain;  // the packed terrain (the march samples it for the deferred normals)
    const double2* walk_sc;  // [n_t]: (sin, cos)(dist_k / R) of the spherical walker, host libm like the reference; nullptr otherwise
    WalkAnchor* walk_anchor;  // [wl][n_anchor]: see anchored_lat_lon; nullptr: every sample through asin / atan2
    double *t_lat, *t_lon, *t_elev;
    unsigned long long* t_close;
    // Stage B cache, step-major [n_t][h_pad]
    const double *path_x, *path_dxr;  // [n_t] each: PathElem::dist (row-independent) and calc_dist's dx / R per step
    double *p_elev, *p_len;
    int* p_n;  // [h] elements per row (capped at n_t)
    // pyramids
    double *tmin1, *tmax1, *tmin2, *tmax2, *tmin3, *tmax3;  // [wl][n1_pad], [wl][n2], [wl]
    unsigned long long *close1, *close2, *close3;
    double *rmin1, *rmax1, *rmin2, *rmax2, *rmin3, *rmax3;  // [n1][h_pad], [n2][h_pad], [h]
    double* obs_alt;  // [1]
    DevObject* objects;
    unsigned long long* counters;  // see Counter
    unsigned* sweep_flags;   // [1]: bit 0 = the path cache of this render is not monotone in the row (no sweep)
    unsigned char* sweep_col;  // [wl]: 1 = column needs the general march
    int* sweep_hit;          // [wl][h_pad]: first-hit step per pixel (0 = none)
    const double* atm_cells;  // g(h) table of the ray-path stage (device_paths.cuh): [ATM_FIELDS][ATM_CELLS]
    const DevGPiece* atm_pieces;  // its cells that hold the start of a temperature function, piecewise
    int n_atm_pieces;
    const double* atm_bnd;           // sorted altitudes where g is not smooth (k_ray_paths_macro), +inf padded
    const unsigned char* atm_first;  // per table cell: index of the first of them at or above the cell's lower edge
};

enum Counter { CNT_RAY_STEPS = 0, CNT_TRACE_POINTS, CNT_PIXELS_HIT, CNT_OVERFLOWS, CNT_PATH_STEPS, CNT_COUNT };

// ---------------------------------------------------------------------------------------------
// Terrain packing
// ---------------------------------------------------------------------------------------------
__global__ void k_retile(const int16_t* __restrict__ raw, int16_t* __restrict__ posts, DevTile t) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)t.nlon * t.nlat;
    if (i >= n) return;
    int ilon = (int)(i / t.nlat), ilat = (int)(i % t.nlat);
    posts[post_index(t, ilon, ilat)] = raw[i];
}
__global__ void k_untile(const int16_t* __restrict__ posts, int16_t* __restrict__ raw, DevTile t) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)t.nlon * t.nlat;
    if (i >= n) return;
    int ilon = (int)(i / t.nlat), ilat = (int)(i % t.nlat);
    raw[i] = posts[post_index(t, ilon, ilat)];
}

__global__ void k_get_elev(DevTerrain T, const double* lat, const double* lon, int n, double* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double e;
    out[i] = terrain_get_elev(T, lat[i], lon[i], &e) ? e : __longlong_as_double(0x7ff8000000000000LL);
}

__global__ void k_atm_probe(const __grid_constant__ DevScene S, const double* h, int n, double* t, double* p, double* idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevAtmLayer& l = S.atm.layer[atm_layer_index(S.atm, h[i])];
    double tt = layer_temperature(l, h[i]);
    double pp = layer_pressure(l, h[i], tt);
    t[i] = tt;
    p[i] = pp;
    idx[i] = air_index(S.atm, pp, tt);
}

// Probe of the ray-path stage's atmosphere function g(h) = dn/n (device_paths.cuh): through the table (NaN
// where the table does not serve the altitude) and through the libm path.
__global__ void __launch_bounds__(128) k_refraction_probe(const __grid_constant__ DevScene S, const double* __restrict__ table, const double* h,
                                                          int n, double* g_tab, double* g_ref, const DevGPiece* pieces, int npieces) {
    __shared__ double tab_smem[ATM_FIELDS * ATM_CELLS];
    for (int i = threadIdx.x; i < ATM_FIELDS * ATM_CELLS; i += blockDim.x) tab_smem[i] = table[i];
    __syncthreads();
    const unsigned tab = (unsigned)__cvta_generic_to_shared(tab_smem);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    g_tab[i] = pieces ? g_fallback(tab, pieces, npieces, S.atm, h[i]) : g_table(tab, h[i] - ATM_BASE);
    g_ref[i] = g_libm(S.atm, h[i]);
}

// ---------------------------------------------------------------------------------------------
// Probes behind the reference's text dumpers.
// ---------------------------------------------------------------------------------------------
// ray_path.rs:65-91: env.cast_ray_stepper(height, ang.to_radians(), false), set_step_size(ray_step), nsteps x next().
// One thread per ray. libm_only: PathStepper::next op for op (device_atm.cuh: stepper_next); else the ray-path stage's
// step with g(h) from the table, its pieces, then libm (device_paths.cuh: rk4_step<FLAT, 1>).
template <bool FLAT>
__global__ void __launch_bounds__(128) k_ray_path_probe(const __grid_constant__ DevScene S, DevBuffers B, double start_h, const double* __restrict__ ang_deg,
                                                        int n, double step, int nsteps, int libm_only, double* __restrict__ h_out) {
    __shared__ double tab_smem[ATM_FIELDS * ATM_CELLS];
    for (int i = threadIdx.x; i < ATM_FIELDS * ATM_CELLS; i += blockDim.x) tab_smem[i] = B.atm_cells[i];
    __syncthreads();
    const GSource gs{(unsigned)__cvta_generic_to_shared(tab_smem), B.atm_pieces, B.n_atm_pieces};
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Stepper st;
    stepper_init(st, FLAT, S.radius, start_h, to_radians(ang_deg[i]));
    const double d = FLAT ? step : step / S.radius, hd = 0.5 * d, d6 = d / 6.0;
    double* out = h_out + (size_t)i * nsteps;
    for (int k = 0; k < nsteps; ++k) {
        if (libm_only) {
            out[k] = stepper_next(st, S.atm, FLAT, 0, S.radius, step).h;
        } else {
            double a_new, b_new;
            rk4_step<FLAT, 1>(S.atm, gs, S.radius, d, hd, d6, st.a, st.b, &a_new, &b_new);
            st.a = a_new, st.b = b_new;
            out[k] = FLAT ? st.a : st.a - S.radius;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Scene preparation: Altitude::abs for the observer (params.rs:23-30) and the objects
// (object/mod.rs:165-186), object frames.
// ---------------------------------------------------------------------------------------------
__global__ void k_prepare_scene(const __grid_constant__ DevScene S, DevTerrain T, DevBuffers B, const atmrt_object* objs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        double alt = S.altitude.value;
        if (S.altitude.kind == ATMRT_ALT_RELATIVE) alt = elev_or_zero(T, S.lat0, S.lon0) + S.altitude.value;
        *B.obs_alt = alt;
    }
    if (i >= 1 && i <= S.nobjects) {
        const atmrt_object& o = objs[i - 1];
        DevObject& d = B.objects[i - 1];
        double alt = o.altitude.value;
        if (o.altitude.kind == ATMRT_ALT_RELATIVE) alt = elev_or_zero(T, o.latitude, o.longitude) + o.altitude.value;
        d.lat = o.latitude;
        d.lon = o.longitude;
        d.elev = alt;
        d.pos = as_cartesian(S.earth, o.latitude, o.longitude, alt);
        d.up = world_directions(S.earth, o.latitude, o.longitude).up;
        // kind, sizes, colour and texture pointer are filled by the host
    }
}

// get_ray_dir / get_ray_elev, generators/fast.rs:111-125 (pixel centring goes through i16)
__device__ __forceinline__ double get_ray_dir(const DevScene& S, int x) {
    if (S.col_dir_deg) return S.col_dir_deg[x];
    double width = (double)S.width;
    double xx = (double)(short)((short)x - (short)S.width / 2) / width;
    return S.direction + xx * S.fov;
}
__device__ __forceinline__ double get_ray_elev(const DevScene& S, int y) {
    if (S.row_elev_deg) return S.row_elev_deg[y];
    double width = (double)S.width, height = (double)S.height;
    double aspect = width / height;
    double yy = (double)(short)((short)y - (short)S.height / 2) / height;
    return S.tilt - yy * S.fov / aspect;
}

// coords_at_dist_calc for every column (mod.rs:114-145): SphericalCalc::new (directional_calc.rs:56-69; also
// ObserverAe with its proj_radius), FlDsCalc::new (:35-38), AzEqCalc::new (:15-17), EllipsoidCalc::new (:104-135).
// Eight doubles per column.
__device__ __forceinline__ void direction_calc(const DevScene& S, double dir, double* c) {
    if (S.earth.walker == WALK_ELLIPSOID) {
        const EllipsoidCalc e = ellipsoid_calc(S.earth, S.lat0, S.lon0, dir);
        c[0] = e.cos_az1, c[1] = e.sin_az1, c[2] = e.sin_alfa, c[3] = e.sig1, c[4] = e.cap_a, c[5] = e.cap_b, c[6] = e.cap_c, c[7] = e.red_lat;
        return;
    }
    double sindir, cosdir;
    sincos(to_radians(dir), &sindir, &cosdir);
    if (S.earth.walker == WALK_FLDS) {
        c[0] = cosdir;
        c[1] = sindir;
        c[2] = cos(to_radians(S.lat0));
        return;
    }
    if (S.earth.walker == WALK_AZEQ) {  // mod.rs:116-125
        const V3 pos = as_cartesian(S.earth, S.lat0, S.lon0, 0.0);
        const Dirs d = world_directions(S.earth, S.lat0, S.lon0);
        const V3 dv = d.north * cosdir + d.east * sindir;
        c[0] = dv.x, c[1] = dv.y, c[2] = dv.z;
        c[3] = pos.x, c[4] = pos.y, c[5] = pos.z;
        return;
    }
    double sinlat, coslat, sinlon, coslon;
    sincos(to_radians(S.lat0), &sinlat, &coslat);
    sincos(to_radians(S.lon0), &sinlon, &coslon);
    Dirs d = spherical_directions_sc(sinlat, coslat, sinlon, coslon);
    V3 dv = d.north * cosdir + d.east * sindir;
    c[0] = dv.x, c[1] = dv.y, c[2] = dv.z;
    c[3] = d.up.x, c[4] = d.up.y, c[5] = d.up.z;
}

__global__ void k_column_setup(const __grid_constant__ DevScene S, DevBuffers B) {
    int xl = blockIdx.x * blockDim.x + threadIdx.x;
    if (xl >= S.x1 - S.x0) return;
    direction_calc(S, get_ray_dir(S, S.x0 + xl), B.colcalc + (size_t)xl * 8);
}

__device__ __forceinline__ EllipsoidCalc column_ellipsoid_calc(const DevScene& S, const double* __restrict__ c) {
    EllipsoidCalc e;
    e.cos_az1 = c[0], e.sin_az1 = c[1], e.sin_alfa = c[2], e.sig1 = c[3], e.cap_a = c[4], e.cap_b = c[5], e.cap_c = c[6], e.red_lat = c[7];
    e.lon = to_radians(S.lon0);
    return e;
}

// The walker of a kernel instantiation: W >= 0 fixes it at compile time (the hot kernels are instantiated for
// the spherical walker, so that the common case carries none of the other models' code), W < 0 reads the scene.
template <int W>
__device__ __forceinline__ int walker_of(const DevScene& S) {
    return W >= 0 ? W : S.earth.walker;
}

// find_normal, utils.rs:15-40: central differences +-15 m north/south and east/west on the terrain, walked
// with the model's own DirectionalCalc from (lat, lon) at azimuths 0 and 90 degrees.
// A pure function of (lat, lon); sin/cos of both are passed in because the callers have them.
template <int W = -1>
__device__ __forceinline__ V3 find_normal(const DevScene& S, const DevTerrain& T, double lat, double lon, double sinlat, double coslat,
                                          double sinlon, double coslon) {
    const int walker = walker_of<W>(S);
    double n_lat, n_lon, s_lat, s_lon, e_lat, e_lon, w_lat, w_lon;
    Dirs D;  // world_directions(lat, lon)
    if (S.earth.flat_dirs) {
        D.north = {-coslon, -sinlon, 0.0};
        D.east = {-sinlon, coslon, 0.0};
        D.up = {0.0, 0.0, 1.0};
    } else {
        D = spherical_directions_sc(sinlat, coslat, sinlon, coslon);
    }
    if (walker == WALK_FLDS) {
        // FlDsCalc::new((lat, lon), 0.0 / 90.0).coords_at_dist(+-DIFF)
        n_lat = lat + 1.0 * NORMAL_DIFF / DEGREE_DISTANCE;
        n_lon = lon + 0.0 * NORMAL_DIFF / DEGREE_DISTANCE / coslat;
        s_lat = lat + 1.0 * -NORMAL_DIFF / DEGREE_DISTANCE;
        s_lon = lon + 0.0 * -NORMAL_DIFF / DEGREE_DISTANCE / coslat;
        e_lat = lat + COS_90 * NORMAL_DIFF / DEGREE_DISTANCE;
        e_lon = lon + SIN_90 * NORMAL_DIFF / DEGREE_DISTANCE / coslat;
        w_lat = lat + COS_90 * -NORMAL_DIFF / DEGREE_DISTANCE;
        w_lon = lon + SIN_90 * -NORMAL_DIFF / DEGREE_DISTANCE / coslat;
    } else if (walker == WALK_AZEQ) {
        // AzEqCalc::new(north cos(az) + east sin(az), as_cartesian(lat, lon, 0)).coords_at_dist(+-DIFF), mod.rs:116-125
        const double r = (90.0 - lat) * DEGREE_DISTANCE;
        const V3 pos{r * coslon, r * sinlon, 0.0};
        const V3 dir_ns = D.north * 1.0 + D.east * 0.0;
        const V3 dir_ew = D.north * COS_90 + D.east * SIN_90;
        azeq_walk(pos, dir_ns, NORMAL_DIFF, &n_lat, &n_lon);
        azeq_walk(pos, dir_ns, -NORMAL_DIFF, &s_lat, &s_lon);
        azeq_walk(pos, dir_ew, NORMAL_DIFF, &e_lat, &e_lon);
        azeq_walk(pos, dir_ew, -NORMAL_DIFF, &w_lat, &w_lon);
    } else if (walker == WALK_ELLIPSOID) {
        // EllipsoidCalc::new(a, b, (lat, lon), 0.0 / 90.0).coords_at_dist(+-DIFF)
        const EllipsoidCalc ns = ellipsoid_calc(S.earth, lat, lon, 0.0), ew = ellipsoid_calc(S.earth, lat, lon, 90.0);
        ellipsoid_walk(S.earth, ns, NORMAL_DIFF, &n_lat, &n_lon);
        ellipsoid_walk(S.earth, ns, -NORMAL_DIFF, &s_lat, &s_lon);
        ellipsoid_walk(S.earth, ew, NORMAL_DIFF, &e_lat, &e_lon);
        ellipsoid_walk(S.earth, ew, -NORMAL_DIFF, &w_lat, &w_lon);
    } else {
        // SphericalCalc::new(radius, (lat, lon), 0.0 / 90.0).coords_at_dist(+-DIFF). With delta = DIFF/radius
        // (2.4e-6 rad) the great-circle walk has closed forms that agree with asin(fpos.z), atan2(fpos.y,
        // fpos.x) to ~1e-15 degrees -- below the 1-2 ulp (1e-14 degrees) those libm calls carry themselves:
        //   north/south: fpos.z = sin(lat +- delta), fpos.xy parallel to (cos lon, sin lon)  => (lat +- delta, lon)
        //   east/west:   fpos.z = sin(lat) cos(delta)       => lat - tan(lat) (1 - cos delta)   (next term 1e-23)
        //                lon +- atan(tan(delta) / cos(lat)) => q - q^3/3, q = tan(delta)/cos(lat) (next term 1e-28)
        // (the cos(90 deg) = 6e-17 north component of the east direction moves the point by 1e-22 rad).
        // Near the poles (|lat| > 85 deg) the expansions lose accuracy: use the walk itself.
        if (fabs(lat) <= 85.0) {
            const double icos = 1.0 / coslat;
            const double q = S.tan_diff * icos;
            const double dlon = to_degrees(q - q * q * q * (1.0 / 3.0));
            const double lat_ew = lat - to_degrees(sinlat * icos * S.versin_diff);
            n_lat = lat + S.diff_deg, n_lon = lon;
            s_lat = lat - S.diff_deg, s_lon = lon;
            e_lat = lat_ew, e_lon = lon + dlon;
            w_lat = lat_ew, w_lon = lon - dlon;
        } else {
            const Dirs G = spherical_directions_sc(sinlat, coslat, sinlon, coslon);  // SphericalCalc's own frame (ObserverAe: not D)
            V3 dir_ns = G.north * 1.0 + G.east * 0.0;
            V3 dir_ew = G.north * COS_90 + G.east * SIN_90;
            spherical_walk(G.up, dir_ns, S.sin_diff, S.cos_diff, &n_lat, &n_lon);
            spherical_walk(G.up, dir_ns, -S.sin_diff, S.cos_diff, &s_lat, &s_lon);
            spherical_walk(G.up, dir_ew, S.sin_diff, S.cos_diff, &e_lat, &e_lon);
            spherical_walk(G.up, dir_ew, -S.sin_diff, S.cos_diff, &w_lat, &w_lon);
        }
    }
    double diff_ew = elev_or_zero(T, e_lat, e_lon) - elev_or_zero(T, w_lat, w_lon);
    double diff_ns = elev_or_zero(T, n_lat, n_lon) - elev_or_zero(T, s_lat, s_lon);
    V3 vec_ns = (2.0 * NORMAL_DIFF) * D.north + diff_ns * D.up;
    V3 vec_ew = (2.0 * NORMAL_DIFF) * D.east + diff_ew * D.up;
    V3 normal = cross(vec_ew, vec_ns);
    return normal / sqrt(dot(normal, normal));
}
// ---------------------------------------------------------------------------------------------
// Stage A: terrain profile. One thread per (column, sample); lanes run along the ray so the
// [column][k] stores are coalesced and the bilinear taps of a warp walk along one azimuth.
//
// The reference computes find_normal for every sample (utils.rs:72-88) although only the two samples
// that bracket a hit are ever read (utils.rs:108-125 through :233). The normal is a pure function of the
// sample's coordinates, so it is deferred to the hit (sample_normal below, called from process_step):
// same arithmetic, same result, 2 evaluations per hit instead of one per sample.
// ---------------------------------------------------------------------------------------------
struct SampleTrig {
    double sinlat, coslat, sinlon, coslon;
};

// Sine and cosine of a sample's latitude and longitude as the reference's world_directions / as_cartesian
// need them (sin/cos of the degree values).
// Spherical: fpos is the unit vector of (lat, lon): its components ARE sin(lat), cos(lat) cos(lon),
// cos(lat) sin(lon) to the rounding of the walk (|fpos| = 1 +- 2e-16), so they are recovered without four
// more libm calls. Near the poles (cos lat < 0.05) the division loses bits: evaluate them from the angles
// as the reference does.
template <int W = -1>
__device__ __forceinline__ SampleTrig sample_trig(const DevScene& S, V3 fpos, double lat, double lon) {
    SampleTrig t;
    const double c2 = fpos.x * fpos.x + fpos.y * fpos.y;
    if (walker_of<W>(S) == WALK_SPHERICAL && c2 > 0.0025) {
        t.sinlat = fpos.z;
        t.coslat = sqrt(c2);
        const double ic = 1.0 / t.coslat;
        t.sinlon = fpos.y * ic, t.coslon = fpos.x * ic;
    } else {
        sincos(to_radians(lat), &t.sinlat, &t.coslat);
        sincos(to_radians(lon), &t.sinlon, &t.coslon);
    }
    return t;
}

// SphericalCalc::coords_at_dist (directional_calc.rs:71-86) up to the unit vector of the sample.
// `sc`: sin and cos of d / R when the caller has them (the Fast generator's distances are the same for every
// column, so the host tabulates sin/cos(dist_k / R) once: DevBuffers::walk_sc), else nullptr.
__device__ __forceinline__ V3 walk_fpos(const DevScene& S, const double* __restrict__ cc, double d, const double2* sc = nullptr) {
    double sinang, cosang;
    if (sc) {
        const double2 v = *sc;
        sinang = v.x, cosang = v.y;
    } else {
        sincos(d / S.earth.radius, &sinang, &cosang);
    }
    return V3{cc[3], cc[4], cc[5]} * cosang + V3{cc[0], cc[1], cc[2]} * sinang;
}

// DirectionalCalc::coords_at_dist of the walker lowered into cc (direction_calc). fpos: the unit vector of the
// sample for the spherical walker (sample_trig reads the sample's sines and cosines off it), else untouched.
template <int W = -1>
__device__ __forceinline__ void walk_coords(const DevScene& S, const double* __restrict__ cc, double d, double* lat, double* lon, V3* fpos,
                                            const double2* sc = nullptr) {
    const int walker = walker_of<W>(S);
    if (walker == WALK_FLDS) {  // FlDsCalc::coords_at_dist, directional_calc.rs:41-47
        double d_lat = cc[0] * d / DEGREE_DISTANCE;
        double d_lon = cc[1] * d / DEGREE_DISTANCE / cc[2];
        *lat = S.lat0 + d_lat;
        *lon = S.lon0 + d_lon;
    } else if (walker == WALK_AZEQ) {  // AzEqCalc::coords_at_dist, directional_calc.rs:20-27
        azeq_walk(V3{cc[3], cc[4], cc[5]}, V3{cc[0], cc[1], cc[2]}, d, lat, lon);
    } else if (walker == WALK_ELLIPSOID) {  // EllipsoidCalc::coords_at_dist, directional_calc.rs:139-184
        ellipsoid_walk(S.earth, column_ellipsoid_calc(S, cc), d, lat, lon);
    } else {  // SphericalCalc::coords_at_dist, directional_calc.rs:71-86
        *fpos = walk_fpos(S, cc, d, sc);
        *lat = to_degrees(asin(fpos->z));
        *lon = to_degrees(atan2(fpos->y, fpos->x));
    }
}

// ---------------------------------------------------------------------------------------------
// Latitude and longitude of a great-circle sample from an ANCHOR sample of the same column.
// SphericalCalc::coords_at_dist (directional_calc.rs:71-86) ends in lat = asin(z), lon = atan2(y, x) of the unit
// vector f = (x, y, z) -- two of the slowest f64 libm calls, 40 % of stage A's instructions. Neighbouring samples
// of a column differ by 4e-6 rad, so with (lat_a, lon_a) of an anchor sample a few kilometres away (libm, once per
// 512 or 1024 samples: k_walk_anchors) the sample's angles follow from exact identities with SMALL arguments:
//     sin(lat - lat_a) = z cos(lat_a) - c sin(lat_a),                 c = sqrt(x^2 + y^2) = cos(lat)
//     tan(lon - lon_a) = (y cos(lon_a) - x sin(lon_a)) / (x cos(lon_a) + y sin(lon_a))
// and asin / atan of an argument below 1e-2 are four / five terms of their series (truncation < 1e-22 rad). The
// rounding errors are absolute ones of ~1e-16 rad on values of order 1 -- what the 1-2 ulp of the libm calls amount to
// (CUDA's and glibc's differ from each other by as much). Returns false (the caller uses asin / atan2) near the
// poles, across the antimeridian and when an argument is not small.
// ---------------------------------------------------------------------------------------------
struct WalkAnchor {
    double lat, lon;  // radians
    double sinlat, coslat, sinlon, coslon;
};

__device__ __forceinline__ bool anchored_lat_lon(const WalkAnchor& a, V3 f, double* lat_deg, double* lon_deg) {
    const double c = sqrt_nr(fma(f.x, f.x, f.y * f.y));
    const double u = fma(f.z, a.coslat, -(c * a.sinlat));    // sin(lat - lat_a)
    const double p = fma(f.y, a.coslon, -(f.x * a.sinlon));  // c sin(lon - lon_a)
    const double q = fma(f.x, a.coslon, f.y * a.sinlon);     // c cos(lon - lon_a)
    const double t = p * rcp_nr(q);
    const double u2 = u * u, t2 = t * t;
    // asin u = u + u^3/6 + 3 u^5/40 + 5 u^7/112 + ...;  atan t = t - t^3/3 + t^5/5 - t^7/7 + t^9/9 - ...
    const double dlat = fma(u * u2, fma(u2, fma(u2, 5.0 / 112.0, 3.0 / 40.0), 1.0 / 6.0), u);
    const double dlon = fma(-(t * t2), fma(t2, fma(t2, fma(t2, -1.0 / 9.0, 1.0 / 7.0), -1.0 / 5.0), 1.0 / 3.0), t);
    const double lon = a.lon + dlon;
    *lat_deg = to_degrees(a.lat + dlat);
    *lon_deg = to_degrees(lon);
    return c > 0.05 && q > 0.0 && fabs(t) < 1.0e-2 && fabs(u) < 1.0e-2 && fabs(lon) < 3.1;  // (false for NaN)
}

// TerrainData::objects_close (utils.rs:76-82): Object::is_close, frustum.rs:103-114 / billboard.rs:68-78
__device__ __forceinline__ unsigned long long objects_close(const DevScene& S, const DevBuffers& B, V3 fpos, double lat, double lon) {
    const SampleTrig t = sample_trig(S, fpos, lat, lon);
    unsigned long long mask = 0;
    for (int i = 0; i < S.nobjects; ++i) {
        const DevObject& o = B.objects[i];
        V3 pos = as_cartesian_sc(S.earth, lat, o.elev, t.sinlat, t.coslat, t.sinlon, t.coslon);
        V3 dist_v = pos - o.pos;
        if (dot(dist_v, dist_v) < 2.0 * (o.close_r + S.step) * (o.close_r + S.step)) mask |= 1ull << i;
    }
    return mask;
}

// The anchors of every column: sample (b << shift) + (1 << shift) / 2 of block b, through libm. One thread per anchor.
__global__ void __launch_bounds__(128) k_walk_anchors(const __grid_constant__ DevScene S, DevBuffers B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, xl = blockIdx.y;
    if (b >= S.n_anchor) return;
    const int ka = min((b << S.anchor_shift) + (1 << S.anchor_shift) / 2, S.n_t - 1);
    const V3 f = walk_fpos(S, B.colcalc + (size_t)xl * 8, B.dist_k[ka], B.walk_sc + ka);
    WalkAnchor a;
    a.lat = asin(f.z), a.lon = atan2(f.y, f.x);
    sincos(a.lat, &a.sinlat, &a.coslat);
    sincos(a.lon, &a.sinlon, &a.coslon);
    B.walk_anchor[(size_t)xl * S.n_anchor + b] = a;
}

template <int W>
__global__ void __launch_bounds__(128) k_terrain_profile(const __grid_constant__ DevScene S, DevTerrain T, DevBuffers B, int col0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int xl = col0 + blockIdx.y;  // the render is issued in column chunks (atmrt_lib.cu:launch_render)
    if (k >= S.n_t) return;
    const double d = B.dist_k[k];
    const double* cc = B.colcalc + (size_t)xl * 8;
    double lat, lon;
    V3 fpos{0.0, 0.0, 0.0};
    bool done = false;
    if (W == WALK_SPHERICAL && B.walk_anchor) {  // asin / atan2 of the walk's unit vector through the column's anchors
        fpos = walk_fpos(S, cc, d, B.walk_sc + k);
        done = anchored_lat_lon(B.walk_anchor[(size_t)xl * S.n_anchor + (k >> S.anchor_shift)], fpos, &lat, &lon);
    }
    if (!done) walk_coords<W>(S, cc, d, &lat, &lon, &fpos, B.walk_sc ? B.walk_sc + k : nullptr);
    double elev = elev_or_zero(T, lat, lon);
    size_t idx = (size_t)xl * S.n_pad + k;
    B.t_lat[idx] = lat;
    B.t_lon[idx] = lon;
    B.t_elev[idx] = elev;

    if (S.nobjects > 0) B.t_close[idx] = objects_close(S, B, fpos, lat, lon);
}

// TerrainData::normal of sample k of column xl (find_normal at the sample's coordinates, utils.rs:84):
// the coordinates come from the cache, the unit vector of the walk is recomputed (one sincos).
template <int W = -1>
__device__ __forceinline__ V3 sample_normal(const DevScene& S, const DevTerrain& T, const DevBuffers& B, int xl, int k, double lat, double lon) {
    V3 fpos{0.0, 0.0, 0.0};
    if (walker_of<W>(S) == WALK_SPHERICAL) fpos = walk_fpos(S, B.colcalc + (size_t)xl * 8, B.dist_k[k], B.walk_sc ? B.walk_sc + k : nullptr);
    const SampleTrig t = sample_trig<W>(S, fpos, lat, lon);
    return find_normal<W>(S, T, lat, lon, t.sinlat, t.coslat, t.sinlon, t.coslon);
}

// The normals of one column (probe: atmrt_get_terrain_profile), [k][3].
__global__ void __launch_bounds__(128) k_profile_normals(const __grid_constant__ DevScene S, DevTerrain T, DevBuffers B, int xl, double* out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= S.n_t) return;
    const size_t idx = (size_t)xl * S.n_pad + k;
    const V3 n = sample_normal(S, T, B, xl, k, B.t_lat[idx], B.t_lon[idx]);
    out[3 * k] = n.x, out[3 * k + 1] = n.y, out[3 * k + 2] = n.z;
}

// elev_profile.rs:43-60: params.model.coords_at_dist_calc((lat0, lon0), azimuth).coords_at_dist(x) and
// terrain.get_elev(..).unwrap_or(0.0) at n distances -- the walker of stage A without its anchors or tables.
__global__ void __launch_bounds__(128) k_elev_profile(const __grid_constant__ DevScene S, DevTerrain T, double azimuth, const double* __restrict__ dist, int n,
                                                      double* __restrict__ lat, double* __restrict__ lon, double* __restrict__ elev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double cc[8];
    direction_calc(S, azimuth, cc);
    double la, lo;
    V3 fpos{0.0, 0.0, 0.0};
    walk_coords(S, cc, dist[i], &la, &lo, &fpos);
    if (lat) lat[i] = la;
    if (lon) lon[i] = lo;
    elev[i] = elev_or_zero(T, la, lo);
}

// ---------------------------------------------------------------------------------------------
// Stage B: ray paths (gen_path_cache, utils.rs:136-174). The physics offers one serial RK4 chain per
// image row, so the stage is bound by the dependent-issue latency of ONE chain, not by FP64
// throughput; see device_paths.cuh for how the critical path of a step is shortened (g(h) = dn/n from
// a per-atmosphere polynomial table in shared memory, overlapped lookups, reciprocals). One lane per
// row, one warp per block so that the chains spread over all SMs. The cache is written tiled by groups
// of four rows ([row/4][k][row%4]): four adjacent lanes store one 32-byte sector per plane.
// ---------------------------------------------------------------------------------------------
template <bool FLAT, bool LIBM>
__global__ void __launch_bounds__(32) k_ray_paths(const __grid_constant__ DevScene S, DevBuffers B) {
    __shared__ double tab_smem[ATM_FIELDS * ATM_CELLS];
    const int lane = threadIdx.x;
#pragma unroll 8
    for (int i = lane; i < ATM_FIELDS * ATM_CELLS; i += 32) tab_smem[i] = B.atm_cells[i];
    __syncwarp();
    const GSource gs{(unsigned)__cvta_generic_to_shared(tab_smem), B.atm_pieces, B.n_atm_pieces};
    // All 32 lanes stay in the loop (rows past the image shadow the last row and write nothing).
    const int y_raw = blockIdx.x * 32 + lane;
    const int y = min(y_raw, S.height - 1);
    const bool writer = y_raw < S.height;
    const double alt = *B.obs_alt;
    const double radius = S.radius;
    const double d = FLAT ? S.step : S.step / radius;
    const double hd = 0.5 * d, d6 = d / 6.0;
    const double shift = FLAT ? ATM_BASE : radius + ATM_BASE;
    const int k_far = S.path_k_far;
    const size_t row0 = path_index(S.n_t, 0, y);
    double* o_elev = B.p_elev + row0;
    double* o_len = B.p_len + row0;
    const double* __restrict__ dxr = B.path_dxr;

    // Outputs of step i-1 (`cur_h`, reached from `prev_h`): calc_dist (utils.rs:42-53; its dx / R is the
    // row-independent path_dxr[i-1]), the two cache entries (PathElem::dist is the row-independent
    // path_x[i-1]), and the termination test of utils.rs:167-170 on the state before it (x > max_distance
    // is i - 2 >= k_far). Branch-free so that it is scheduled into the stalls of the integration chain; for
    // i == 1 it (re)writes the initial element (alt, 0).
#define ATMRT_PATH_OUTPUTS()                                                                                     \
    {                                                                                                            \
        double dx = dx_next;                                                                                     \
        dx_next = dxr[min(i, S.n_t - 1)]; /* for the next iteration: the load leaves the critical path */       \
        const double dh = cur_h - prev_h;                                                                        \
        if (!FLAT) dx = dx * ((cur_h + prev_h) * 0.5 + radius);                                                  \
        const double seg = sqrt_nr(dx * dx + dh * dh);                                                           \
        path_length = i >= 2 ? path_length + seg : 0.0;                                                          \
        const bool emit = writer && !done;                                                                       \
        stg_if(o_elev, cur_h, emit);                                                                             \
        stg_if(o_len, path_length, emit);                                                                        \
        o_elev += PATH_ROWS, o_len += PATH_ROWS;                                                                 \
        n = emit ? i : n;                                                                                        \
        done = done || (i >= 2 && (i - 2 >= k_far || prev_h < -1000.0));                                         \
    }

    // stepper state: spherical (r, dr/dphi) or flat (h, dh/dx). The loop is software-pipelined: iteration i
    // integrates step i (the latency chain) while the outputs of step i-1 are produced in its shadow.
    double a = FLAT ? alt : radius + alt;
    double b = FLAT ? tan(to_radians(get_ray_elev(S, y))) : a * tan(to_radians(get_ray_elev(S, y)));
    // bases of the lookups (device_paths.cuh): E serves stage 1, M stages 2-3, N stage 4 and the next step's stage 1
    PathBase bE = path_base<FLAT>(gs.tab, shift, a), bM = path_base<FLAT>(gs.tab, shift, fma(hd, b, a)), bN = path_base<FLAT>(gs.tab, shift, fma(d, b, a));
    const double d15 = 1.5 * d, d2 = 2.0 * d;
    double dx_next = dxr[0];
    double prev_h = alt;  // state i-2
    double cur_h = alt;   // state i-1
    double path_length = 0.0;
    int n = 1;
    bool done = false;  // this row's cache is complete
    int i = 1;
#pragma unroll 1
    for (; i < S.n_t; ++i) {
        // A NaN state (the ray climbed above the altitude where the last temperature function reaches
        // 0 K, e.g. 178 km for US-76) stays NaN: h = NaN, path_length = NaN, exactly what the arithmetic
        // produces (NaN in, NaN out). When every row of the warp is complete or NaN the integration stops
        // (tested on the incoming state, acted upon at the end of the iteration).
        const bool idle = __all_sync(FULL, done || a != a);
        double a_new, b_new;
        // one basic block: the outputs of step i-1 are scheduled into the stalls of the integration chain
        bool ok;
        if (LIBM) {
            ok = rk4_step<FLAT, 2>(S.atm, gs, radius, d, hd, d6, a, b, &a_new, &b_new);
        } else {
            ok = rk4_step_shared<FLAT>(d, hd, d6, a, b, bE, bM, bN, &a_new, &b_new) || done || a != a || b != b;
            // the bases of step i+1, extrapolated from the state at the start of step i: independent of the chain above
            bE = bN;
            bM = path_base<FLAT>(gs.tab, shift, fma(d15, b, a));
            bN = path_base<FLAT>(gs.tab, shift, fma(d2, b, a));
        }
        ATMRT_PATH_OUTPUTS()
        if (!ok) rk4_step<FLAT, 1>(S.atm, gs, radius, d, hd, d6, a, b, &a_new, &b_new);  // rare: an altitude the table does not serve
        // a complete row stops moving: its state would otherwise leave the table (below -2 km) and drag the
        // warp through the libm path on every step
        a = done ? a : a_new, b = done ? b : b_new;
        prev_h = cur_h;
        cur_h = FLAT ? a : a - radius;
        if (idle || __all_sync(FULL, done)) {
            ++i;
            break;
        }
    }
    // Here cur_h = state i-1 and the outputs of steps < i-1 are written; then the last state's outputs
    // (rows that are NaN stay NaN).
#pragma unroll 1
    for (; i <= S.n_t; ++i) {
        ATMRT_PATH_OUTPUTS()
        if (__all_sync(FULL, done)) break;
        prev_h = cur_h;
    }
#undef ATMRT_PATH_OUTPUTS
    if (writer) {
        B.p_n[y] = n;
        atomicAdd(B.counters + CNT_PATH_STEPS, (unsigned long long)(n - 1));
    }
}

// ---------------------------------------------------------------------------------------------
// Stage B with macro steps. The ray equation is smooth wherever g(h) is -- everywhere except at the starts
// of the temperature functions, where g jumps -- and there classical RK4 with a 25 m step is converged far
// below f64 resolution of the path: integrating the same equation with ONE RK4 step of 16 x 25 m lands on the
// same state to 4e-11 m for near-horizontal rays and 2e-9 m at 44 degrees (extended-precision measurement in
// DESIGN.md section 4.B), and the fifteen states in between follow from the cubic Hermite interpolant of the
// two ends (error < 2e-11 m). At a start of a temperature function the reference's result DOES depend on how its 25 m steps
// straddle the jump (by millimetres), so there the kernel takes the reference's own single steps: a macro
// step is taken only when no start lies in the altitude span it covers (exact test against the sorted
// starts), single steps otherwise. The chain is 16x shorter where it matters, and the sixteen states of a
// macro step are sixteen independent outputs: a warp is 2 rows x 16 sub-lanes (lane = 2 j + row), every
// sub-lane integrates the row's macro step redundantly (identical values, no exchange), evaluates ITS
// state from the interpolant, its calc_dist segment, takes part in a prefix sum for path_length and stores
// its cache entry.
// ---------------------------------------------------------------------------------------------
constexpr int MACRO_THREADS = 128;  // 4 warps share one copy of the table
constexpr double MACRO_MAX_METRES = 800.0;  // longest macro step: 16 x 50 m (truncation error < 3e-8 m, section 4.B); longer simulation steps get fewer per macro step
constexpr int ATM_MAX_BND = ATMRT_MAX_ATM_FUNCTIONS + 4;

template <bool FLAT, int MACRO>  // MACRO steps per macro step (16, 8, 4 or 2), 32 / MACRO rows per warp; rows [row0, row1)
__global__ void __launch_bounds__(MACRO_THREADS) k_ray_paths_macro(const __grid_constant__ DevScene S, DevBuffers B, int row0, int row1) {
    __shared__ double tab_smem[ATM_FIELDS * ATM_CELLS];
    __shared__ double s_bnd[ATM_MAX_BND];          // sorted altitudes where g is not smooth (+inf padded)
    __shared__ unsigned char s_first[ATM_CELLS];   // per cell: index of the first of them at or above the cell's lower edge
#pragma unroll 4
    for (int i = threadIdx.x; i < ATM_FIELDS * ATM_CELLS; i += MACRO_THREADS) tab_smem[i] = B.atm_cells[i];
    for (int i = threadIdx.x; i < ATM_CELLS; i += MACRO_THREADS) s_first[i] = B.atm_first[i];
    if (threadIdx.x < ATM_MAX_BND) s_bnd[threadIdx.x] = B.atm_bnd[threadIdx.x];
    __syncthreads();
    constexpr int MACRO_ROWS = 32 / MACRO;
    const GSource gs{(unsigned)__cvta_generic_to_shared(tab_smem), B.atm_pieces, B.n_atm_pieces};
    const int lane = threadIdx.x & 31, rr = lane % MACRO_ROWS, j = lane / MACRO_ROWS;
    const int y_raw = row0 + (blockIdx.x * (MACRO_THREADS / 32) + (threadIdx.x >> 5)) * MACRO_ROWS + rr;
    const int y = min(y_raw, row1 - 1);
    const bool writer = y_raw < row1;
    const double alt = *B.obs_alt;
    const double radius = S.radius, off = FLAT ? 0.0 : radius;
    const double d = FLAT ? S.step : S.step / radius;
    const double hd = 0.5 * d, d6 = d / 6.0;
    const double D = (double)MACRO * d, hD = 0.5 * D, D6 = D / 6.0;
    const double shift = FLAT ? ATM_BASE : radius + ATM_BASE;
    const double d15 = 1.5 * d, d2 = 2.0 * d;
    const int n_t = S.n_t, k_far = S.path_k_far;
    double* const o_elev = B.p_elev + path_index(n_t, 0, y);
    double* const o_len = B.p_len + path_index(n_t, 0, y);
    const double* __restrict__ dxr = B.path_dxr;
    // the cubic Hermite basis at this sub-lane's state, s = (j + 1) / MACRO
    const double sj = (double)(j + 1) * (1.0 / MACRO);
    const double h01 = sj * sj * (3.0 - 2.0 * sj), h10 = sj * (sj - 1.0) * (sj - 1.0), h11 = sj * sj * (sj - 1.0);
    unsigned rowbits = 0u;  // this row's sub-lanes in a ballot
    for (int t = 0; t < MACRO; ++t) rowbits |= 1u << (t * MACRO_ROWS + rr);

    double a = FLAT ? alt : radius + alt;
    double b = FLAT ? tan(to_radians(get_ray_elev(S, y))) : a * tan(to_radians(get_ray_elev(S, y)));
    PathBase bE{0.0, 0.0, 0.0, 0.0}, bM = bE, bN = bE;
    bool bases_valid = false;
    // element 0: (alt, 0)
    if (writer && j == 0) o_elev[0] = alt, o_len[0] = 0.0;
    double h_e = alt;           // altitude of element e
    double path_length = 0.0;   // of element e
    bool done_c = false;        // an element <= e - 1 is past max_distance or below -1000 m (utils.rs:167-170)
    bool trig_e = alt < -1000.0 || 0 >= k_far;  // ... element e is
    int n = 1;
    int e = 0;
    double dx_q = dxr[min(j + 1, n_t - 1)];
#pragma unroll 1
    while (e < n_t - 1) {
        if (__all_sync(FULL, done_c)) break;
        const bool still = done_c || a != a;  // complete (it stops moving) or NaN (NaN in, NaN out): nothing to integrate
        const bool idle = __all_sync(FULL, still);
        // may this step be a macro step? no start of a temperature function in the altitudes it spans
        bool macro = e + MACRO <= n_t - 1;
        if (macro && !idle) {
            const double a_end = fma(D, b, a);
            const double lo = fmin(a, a_end) - off - 1.0, hi = fmax(a, a_end) - off + 1.0;
            const int cell = min(max((int)floor((lo - (ATM_BASE - 0.5 * ATM_CELL)) * (1.0 / ATM_CELL)), 0), ATM_CELLS - 1);
            int t = s_first[cell];
            while (s_bnd[t] < lo) ++t;
            const bool unsafe = !still && !(s_bnd[t] > hi);  // (NaN spans are not safe either)
            macro = !__any_sync(FULL, unsafe);
        }
        double a1 = a, b1 = b;
        if (macro && !idle) {
            const bool ok = rk4_step<FLAT, 0>(S.atm, gs, radius, D, hD, D6, a, b, &a1, &b1) || still;
            macro = __all_sync(FULL, ok);  // a cell the table does not serve: the reference's single steps handle it
        }
        int m;
        bool valid;
        double a_q;  // this sub-lane's state
        if (macro) {
            m = MACRO;
            valid = true;
            a_q = j == MACRO - 1 ? a1 : a + fma(h01, a1 - a, D * fma(h10, b, h11 * b1));
            if (still) a_q = a, a1 = a, b1 = b;
            bases_valid = false;
        } else {
            m = 1;
            valid = j == 0;
            if (!bases_valid) {
                bE = path_base<FLAT>(gs.tab, shift, a), bM = path_base<FLAT>(gs.tab, shift, fma(hd, b, a)), bN = path_base<FLAT>(gs.tab, shift, fma(d, b, a));
                bases_valid = true;
            }
            const bool ok = rk4_step_shared<FLAT>(d, hd, d6, a, b, bE, bM, bN, &a1, &b1) || still;
            bE = bN;
            bM = path_base<FLAT>(gs.tab, shift, fma(d15, b, a));
            bN = path_base<FLAT>(gs.tab, shift, fma(d2, b, a));
            if (!ok) rk4_step<FLAT, 1>(S.atm, gs, radius, d, hd, d6, a, b, &a1, &b1);  // rare: an altitude the table does not serve
            if (still) a1 = a, b1 = b;
            a_q = a1;
        }
        // this sub-lane's element q = e + j + 1: calc_dist from the element before it, path_length by prefix sum
        const int q = min(e + j + 1, n_t - 1);
        const double h_q = FLAT ? a_q : a_q - radius;
        const double h_up = __shfl_up_sync(FULL, h_q, MACRO_ROWS);
        const double h_p = j == 0 ? h_e : h_up;
        double dx = dx_q;  // dxr[q], loaded one iteration ahead: off the chain
        if (!FLAT) dx = dx * ((h_q + h_p) * 0.5 + radius);
        const double dh = h_q - h_p;
        double acc = valid ? sqrt_nr(dx * dx + dh * dh) : 0.0;
#pragma unroll
        for (int o = MACRO_ROWS; o < 32; o <<= 1) {
            const double up = __shfl_up_sync(FULL, acc, o);
            if (lane >= o) acc += up;
        }
        const double len_q = path_length + acc;
        // termination: element q is kept unless an element <= q - 2 is past max_distance or below -1000 m
        const bool trig_q = valid && (q >= k_far || h_q < -1000.0);
        const unsigned trig = (__ballot_sync(FULL, trig_q) & rowbits) >> rr;  // bit MACRO_ROWS j' = sub-lane j'
        const bool earlier = done_c || (j >= 1 && trig_e) || (j >= 2 && (trig & ((1u << (MACRO_ROWS * (j - 1))) - 1u)) != 0u);
        const bool emit = valid && writer && !earlier;
        stg_if(o_elev + (size_t)q * PATH_ROWS, h_q, emit);
        stg_if(o_len + (size_t)q * PATH_ROWS, len_q, emit);
        n = emit ? q + 1 : n;
        // carry to element e + m
        const int last = rr + MACRO_ROWS * (m - 1);
        const unsigned before_last = m == 1 ? 0u : (trig & ((1u << (MACRO_ROWS * (m - 1))) - 1u));
        done_c = done_c || trig_e || before_last != 0u;
        trig_e = ((trig >> (MACRO_ROWS * (m - 1))) & 1u) != 0u;
        path_length = __shfl_sync(FULL, len_q, last);
        h_e = __shfl_sync(FULL, h_q, last);
        a = a1, b = b1;
        e += m;
        dx_q = dxr[min(e + j + 1, n_t - 1)];  // this sub-lane's element of the next iteration, whichever kind of step it takes
    }
    for (int o = MACRO_ROWS; o < 32; o <<= 1) n = max(n, __shfl_xor_sync(FULL, n, o));
    if (writer && j == 0) {
        B.p_n[y] = n;
        atomicAdd(B.counters + CNT_PATH_STEPS, (unsigned long long)(n - 1));
    }
}

// Straight rays (-s): the path is a closed form of the step index (no chain through the state), one
// thread per row.
__global__ void __launch_bounds__(128) k_ray_paths_straight(const __grid_constant__ DevScene S, DevBuffers B) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= S.height) return;
    const bool flat = S.flat != 0;
    const double alt = *B.obs_alt;
    Stepper st;
    stepper_init(st, flat, S.radius, alt, to_radians(get_ray_elev(S, y)));
    const size_t row0 = path_index(S.n_t, 0, y);
    B.p_elev[row0] = alt;
    B.p_len[row0] = 0.0;
    RayState prev{0.0, alt};
    double path_length = 0.0;
    int n = 1;
    for (int i = 1; i < S.n_t; ++i) {
        const RayState nw = stepper_next(st, S.atm, flat, 1, S.radius, S.step);
        path_length += calc_dist(flat, S.radius, prev, nw);
        const size_t o = row0 + (size_t)i * PATH_ROWS;
        B.p_elev[o] = nw.h;
        B.p_len[o] = path_length;
        n = i + 1;
        if (prev.x > S.max_distance || prev.h < -1000.0) break;
        prev = nw;
    }
    B.p_n[y] = n;
    atomicAdd(B.counters + CNT_PATH_STEPS, (unsigned long long)(n - 1));
}

// ---------------------------------------------------------------------------------------------
// Min/max pyramids. Level-1 node c covers march steps k in [32c, 32c+31] (k >= 1), i.e. samples
// [32c-1, 32c+31]; level-2 node C covers level-1 nodes [32C, 32C+31]; level 3 is the whole ray.
// A node without any step gets min=+inf, max=-inf and is never a candidate.
// Terrain pyramids are [column][node] (one warp per (column, C), lanes along k); path pyramids are
// node-major [node][row] (one thread per (row, node), lanes along rows) to match the cache layouts.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_terrain_pyramid(const double* __restrict__ vals, const unsigned long long* __restrict__ close,
                                                         int nrows, int n, int n_pad, int n1, int n1_pad, int n2,
                                                         double* __restrict__ min1, double* __restrict__ max1,
                                                         double* __restrict__ min2, double* __restrict__ max2,
                                                         unsigned long long* __restrict__ close1, unsigned long long* __restrict__ close2,
                                                         const unsigned* only_if_set) {
    if (only_if_set && *only_if_set == 0) return;  // fallback launch behind the horizon sweep
    const int lane = threadIdx.x & 31;
    long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= (long long)nrows * n2) return;
    const int row = (int)(warp / n2), C = (int)(warp % n2);
    const double* v = vals + (size_t)row * n_pad;
    const unsigned long long* cl = close ? close + (size_t)row * n_pad : nullptr;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo2 = INF, hi2 = -INF;
    unsigned long long cm2 = 0;
    for (int c1 = 0; c1 < 32; ++c1) {
        int c = C * 32 + c1;
        if (c >= n1) break;
        double lo = INF, hi = -INF;
        unsigned long long cm = 0;
        bool has_steps = n >= 2 && 32 * c <= n - 1;
        if (has_steps) {
            int k = 32 * c + lane;
            if (k < n) {
                double x = v[k];
                lo = fmin(lo, x);
                hi = fmax(hi, x);
                if (cl) cm = cl[k];
            }
            if (lane == 0 && c > 0) {
                double x = v[32 * c - 1];
                lo = fmin(lo, x);
                hi = fmax(hi, x);
                if (cl) cm |= cl[32 * c - 1];
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
            hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
            cm |= __shfl_xor_sync(FULL, cm, o);
        }
        if (lane == 0) {
            min1[(size_t)row * n1_pad + c] = lo;
            max1[(size_t)row * n1_pad + c] = hi;
            if (close1) close1[(size_t)row * n1_pad + c] = cm;
        }
        lo2 = fmin(lo2, lo);
        hi2 = fmax(hi2, hi);
        cm2 |= cm;
    }
    if (lane == 0) {
        min2[(size_t)row * n2 + C] = lo2;
        max2[(size_t)row * n2 + C] = hi2;
        if (close2) close2[(size_t)row * n2 + C] = cm2;
    }
}

// level 3 of the terrain pyramid: one thread per column
__global__ void k_terrain_top(const double* __restrict__ min2, const double* __restrict__ max2, const unsigned long long* __restrict__ close2,
                              int nrows, int n2, double* __restrict__ min3, double* __restrict__ max3, unsigned long long* __restrict__ close3,
                              const unsigned* only_if_set) {
    if (only_if_set && *only_if_set == 0) return;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo = INF, hi = -INF;
    unsigned long long cm = 0;
    for (int C = 0; C < n2; ++C) {
        lo = fmin(lo, min2[(size_t)row * n2 + C]);
        hi = fmax(hi, max2[(size_t)row * n2 + C]);
        if (close2) cm |= close2[(size_t)row * n2 + C];
    }
    min3[row] = lo;
    max3[row] = hi;
    if (close3) close3[row] = cm;
}

// path pyramid level 1: thread per (row, node c); vals is step-major [k][h_pad]
__global__ void __launch_bounds__(256) k_path_pyramid1(const double* __restrict__ vals, const int* __restrict__ lens, int h, int h_pad,
                                                       int n_total, int n1, double* __restrict__ min1, double* __restrict__ max1,
                                                       const unsigned* only_if_set) {
    if (only_if_set && *only_if_set == 0) return;
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (y >= h) return;
    const int n = min(lens[y], n_total);
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo = INF, hi = -INF;
    if (n >= 2 && 32 * c <= n - 1) {
        const int k0 = max(32 * c - 1, 0), k1 = min(32 * c + 31, n - 1);
        for (int k = k0; k <= k1; ++k) {
            double x = vals[path_index(n_total, k, y)];
            lo = fmin(lo, x);
            hi = fmax(hi, x);
        }
    }
    min1[(size_t)c * h_pad + y] = lo;
    max1[(size_t)c * h_pad + y] = hi;
}

// path pyramid levels 2 and 3: thread per row
__global__ void __launch_bounds__(256) k_path_pyramid23(int h, int h_pad, int n1, int n2, const double* __restrict__ min1,
                                                        const double* __restrict__ max1, double* __restrict__ min2,
                                                        double* __restrict__ max2, double* __restrict__ min3, double* __restrict__ max3,
                                                        const unsigned* only_if_set) {
    if (only_if_set && *only_if_set == 0) return;
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= h) return;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo3 = INF, hi3 = -INF;
    for (int C = 0; C < n2; ++C) {
        double lo = INF, hi = -INF;
        for (int c = 32 * C; c < min(32 * C + 32, n1); ++c) {
            lo = fmin(lo, min1[(size_t)c * h_pad + y]);
            hi = fmax(hi, max1[(size_t)c * h_pad + y]);
        }
        min2[(size_t)C * h_pad + y] = lo;
        max2[(size_t)C * h_pad + y] = hi;
        lo3 = fmin(lo3, lo);
        hi3 = fmax(hi3, hi);
    }
    min3[y] = lo3;
    max3[y] = hi3;
}

// ---------------------------------------------------------------------------------------------
// Stage C: the march. One thread per pixel; the 32 lanes of a warp are 32 adjacent rows of ONE
// column, so every terrain-side load (profile, pyramids) is a warp-wide broadcast and every
// path-side load is coalesced ([k][row] layout). Each lane alternates between
//   phase 1: a cheap search for its next event (a step with a terrain crossing, or with a close
//            object) that skips 1024-, 32-step nodes whose min/max intervals cannot cross, and
//   phase 2: the expensive event processing (interpolation, object intersection, colouring,
//            compositing), executed after the warp has reconverged so that all lanes holding an
//            event run it together; finished pixels retire through a warp vote.
// Follows get_single_pixel (utils.rs:201-289) and draw_image (renderer/mod.rs:395-411).
// ---------------------------------------------------------------------------------------------
struct PixelState {
    Rgb8 result;
    double accum_neg_alpha;
    int count;
    int overflows;
    double m_lat, m_lon, m_elev, m_dist;
};

// Crossing march (below): per column, the samples next to a close object keep sin / cos of their latitude and longitude
constexpr unsigned CROSS_NO_SLOT = 0xffffu;
constexpr int CROSS_TRIG_CAP = 1024;  // cached samples per column; the rest evaluates sincos in place

struct MarchOut {
    unsigned char* rgb;          // [h][wl][3]
    atmrt_meta* meta;            // [h][wl]
    int* steps;                  // [h][wl]
    atmrt_trace_point* points;   // [h][wl][max_points] (TRACE)
    int* counts;                 // [h][wl]             (TRACE)
    int max_points;
};

// One emitted trace point: colour it, composite it, remember the first one as the pixel metadata
// (renderer/mod.rs:399-408).
template <bool TRACE>
__device__ __forceinline__ void emit_point(const DevScene& S, const MarchOut& O, size_t pixel, int k, PixelState& st, bool is_terrain,
                                           double lat, double lon, double dist, double elevation, double plen, V3 normal, Color4 color) {
    const double alpha = color.a;
    Rgb8 color1 = color_for_pixel(S.shade, is_terrain, elevation, dist, normal, color);
    Rgb8 color2 = S.shade.fog_enabled ? apply_fog(S.shade.fog_distance, plen, color1) : color1;
    st.result = add_rgb(st.result, color2, st.accum_neg_alpha * alpha);
    st.accum_neg_alpha *= 1.0 - alpha;
    if (st.count == 0) {
        st.m_lat = lat, st.m_lon = lon, st.m_elev = elevation, st.m_dist = dist;
    }
    if (TRACE) {
        if (st.count < O.max_points) {
            atmrt_trace_point& tp = O.points[pixel * O.max_points + st.count];
            tp.lat = lat, tp.lon = lon, tp.distance = dist, tp.elevation = elevation, tp.path_length = plen;
            tp.normal[0] = normal.x, tp.normal[1] = normal.y, tp.normal[2] = normal.z;
            tp.color[0] = color.r, tp.color[1] = color.g, tp.color[2] = color.b, tp.color[3] = alpha;
            tp.is_terrain = is_terrain ? 1 : 0;
            tp.step = k;
        }
    }
    st.count += 1;
}

// The two ends of one step of get_single_pixel's stream: old_tracing_state and new_tracing_state
// (utils.rs:207-219), with dist / path_len of state 0 already forced to 0.
struct StepEnds {
    double lat0, lon0, elev0, ray0, dist0, len0;
    double lat1, lon1, elev1, ray1, dist1, len1;
    unsigned long long mask;  // objects_close of either end
    const double* trig0 = nullptr;  // sin lat, cos lat, sin lon, cos lon of the end (k_thresholds), or null
    const double* trig1 = nullptr;
};

// One step that holds an event (utils.rs:220-285). `normals(&n0, &n1)` yields TerrainData::normal of the two
// ends; it is called only when the terrain is hit. Returns true when the pixel finishes (an alpha == 1 surface).
template <bool OBJECTS, bool TRACE, class NormalFn>
__device__ __forceinline__ bool process_ends(const DevScene& S, const DevBuffers& B, const MarchOut& O, const StepEnds& e, int k, size_t pixel,
                                             PixelState& st, NormalFn normals) {
    const double lat0 = e.lat0, lon0 = e.lon0, elev0 = e.elev0, lat1 = e.lat1, lon1 = e.lon1, elev1 = e.elev1;
    const double ray0 = e.ray0, ray1 = e.ray1, dist0 = e.dist0, dist1 = e.dist1, len0 = e.len0, len1 = e.len1;
    const double diff1 = ray0 - elev0, diff2 = ray1 - elev1;
    const bool terrain_hit = diff1 * diff2 < 0.0;
    bool finish = false;

    if (!OBJECTS) {
        if (!terrain_hit) return false;
        const double prop = diff1 / (diff1 - diff2);
        V3 n0, n1;
        normals(&n0, &n1);
        // TracingState::interpolate, utils.rs:108-125
        emit_point<TRACE>(S, O, pixel, k, st, true, lat0 + (lat1 - lat0) * prop, lon0 + (lon1 - lon0) * prop,
                          dist0 + (dist1 - dist0) * prop, elev0 + (elev1 - elev0) * prop, len0 + (len1 - len0) * prop,
                          n0 + (n1 - n0) * prop, Color4{0.0, 0.0, 0.0, S.shade.terrain_alpha});
        return S.shade.terrain_alpha == 1.0;
    }

    double c_prop[ATMRT_MAX_STEP_POINTS];
    V3 c_normal[ATMRT_MAX_STEP_POINTS];
    Color4 c_color[ATMRT_MAX_STEP_POINTS];
    bool c_terrain[ATMRT_MAX_STEP_POINTS];
    int ncand = 0;
    bool overflow = false;
    if (terrain_hit) {
        double prop = diff1 / (diff1 - diff2);
        V3 n0, n1;
        normals(&n0, &n1);
        c_prop[0] = prop;
        c_normal[0] = n0 + (n1 - n0) * prop;
        c_color[0] = Color4{0.0, 0.0, 0.0, S.shade.terrain_alpha};
        c_terrain[0] = true;
        ncand = 1;
        if (S.shade.terrain_alpha == 1.0) finish = true;
    }
    unsigned long long mask = e.mask;
    if (mask) {
        // (the sines and cosines are the sample's, the same for every row: the crossing march caches them)
        V3 pos1 = e.trig0 ? as_cartesian_sc(S.earth, lat0, ray0, e.trig0[0], e.trig0[1], e.trig0[2], e.trig0[3]) : as_cartesian(S.earth, lat0, lon0, ray0);
        V3 pos2 = e.trig1 ? as_cartesian_sc(S.earth, lat1, ray1, e.trig1[0], e.trig1[1], e.trig1[2], e.trig1[3]) : as_cartesian(S.earth, lat1, lon1, ray1);
        // The reference walks a HashSet (arbitrary order); index order here. Only exact `prop` ties
        // could tell the difference (stable sort below).
        while (mask) {
            int oi = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const DevObject& o = B.objects[oi];
            Collision coll[4];
            int nc = o.kind == ATMRT_OBJECT_FRUSTUM ? frustum_collision(o, pos1, pos2, coll) : billboard_collision(o, pos1, pos2, coll);
            for (int i = 0; i < nc; ++i) {
                if (coll[i].color.a == 0.0) continue;
                if (ncand < ATMRT_MAX_STEP_POINTS) {
                    c_prop[ncand] = coll[i].prop;
                    c_normal[ncand] = coll[i].normal;
                    c_color[ncand] = coll[i].color;
                    c_terrain[ncand] = false;
                    ++ncand;
                } else {
                    overflow = true;
                }
                if (coll[i].color.a == 1.0) {
                    finish = true;
                    break;
                }
            }
        }
    }
    if (overflow) st.overflows += 1;
    // stable sort by prop (step_result.sort_by, utils.rs:281)
    int order[ATMRT_MAX_STEP_POINTS];
    for (int i = 0; i < ncand; ++i) {
        int j = i - 1;
        while (j >= 0 && c_prop[i] < c_prop[order[j]]) {
            order[j + 1] = order[j];
            --j;
        }
        order[j + 1] = i;
    }
    for (int oi = 0; oi < ncand; ++oi) {
        const int i = order[oi];
        const double prop = c_prop[i];
        const double t_elev = elev0 + (elev1 - elev0) * prop;
        const double r_elev = ray0 + (ray1 - ray0) * prop;
        emit_point<TRACE>(S, O, pixel, k, st, c_terrain[i], lat0 + (lat1 - lat0) * prop, lon0 + (lon1 - lon0) * prop,
                          dist0 + (dist1 - dist0) * prop, c_terrain[i] ? t_elev : r_elev, len0 + (len1 - len0) * prop, c_normal[i],
                          c_color[i]);
    }
    return finish;
}

// The Fast generator's step k of pixel (xl, y): the ends come from the two caches (fast.rs:56-64).
template <bool OBJECTS, bool TRACE>
__device__ __forceinline__ bool process_step(const DevScene& S, const DevBuffers& B, const MarchOut& O, int xl, int y, int k,
                                             size_t pixel, PixelState& st, const V3* normals = nullptr, const unsigned short* trig_slot = nullptr,
                                             const double* trig = nullptr) {
    const size_t ti = (size_t)xl * S.n_pad + k;
    const size_t p1 = path_index(S.n_t, k, y), p0 = p1 - PATH_ROWS;
    StepEnds e;
    e.lat0 = B.t_lat[ti - 1], e.lon0 = B.t_lon[ti - 1], e.elev0 = B.t_elev[ti - 1];
    e.lat1 = B.t_lat[ti], e.lon1 = B.t_lon[ti], e.elev1 = B.t_elev[ti];
    e.ray0 = B.p_elev[p0], e.ray1 = B.p_elev[p1];
    e.dist0 = B.path_x[k - 1], e.dist1 = B.path_x[k];  // path_x[0] = 0
    e.len0 = k - 1 == 0 ? 0.0 : B.p_len[p0], e.len1 = B.p_len[p1];
    e.mask = OBJECTS ? (B.t_close[ti - 1] | B.t_close[ti]) : 0ull;
    if (OBJECTS && trig_slot && e.mask) {
        const unsigned s0 = trig_slot[ti - 1], s1 = trig_slot[ti];
        if (s0 != CROSS_NO_SLOT) e.trig0 = trig + ((size_t)xl * CROSS_TRIG_CAP + s0) * 4;
        if (s1 != CROSS_NO_SLOT) e.trig1 = trig + ((size_t)xl * CROSS_TRIG_CAP + s1) * 4;
    }
    return process_ends<OBJECTS, TRACE>(S, B, O, e, k, pixel, st, [&](V3* n0, V3* n1) {
        if (normals) {
            *n0 = normals[0], *n1 = normals[1];
        } else {
            *n0 = sample_normal(S, B.terrain, B, xl, k - 1, e.lat0, e.lon0), *n1 = sample_normal(S, B.terrain, B, xl, k, e.lat1, e.lon1);
        }
    });
}

constexpr int MARCH_THREADS = 64;

__device__ __forceinline__ void init_pixel(PixelState& st) {
    st.result = Rgb8{{0, 0, 0}};
    st.accum_neg_alpha = 1.0;
    st.count = 0;
    st.overflows = 0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    st.m_lat = st.m_lon = st.m_elev = st.m_dist = qnan;
}

// draw_image tail: *px = add(result, def_color, accum_neg_alpha) (renderer/mod.rs:410), the per-pixel
// metadata, and the render counters (one atomic per warp).
// render counters: one atomic per warp
__device__ __forceinline__ void count_pixel(const DevBuffers& B, bool active, const PixelState& st, int consumed) {
    unsigned long long s_steps = active ? (unsigned long long)consumed : 0ull;
    unsigned s_points = active ? (unsigned)st.count : 0u, s_hit = active && st.count > 0 ? 1u : 0u, s_over = active ? (unsigned)st.overflows : 0u;
    for (int o = 16; o > 0; o >>= 1) s_steps += __shfl_xor_sync(FULL, s_steps, o);
    s_points = __reduce_add_sync(FULL, s_points);
    s_hit = __reduce_add_sync(FULL, s_hit);
    s_over = __reduce_add_sync(FULL, s_over);
    if ((threadIdx.x & 31) == 0) {
        if (s_steps) atomicAdd(B.counters + CNT_RAY_STEPS, s_steps);
        if (s_points) atomicAdd(B.counters + CNT_TRACE_POINTS, (unsigned long long)s_points);
        if (s_hit) atomicAdd(B.counters + CNT_PIXELS_HIT, (unsigned long long)s_hit);
        if (s_over) atomicAdd(B.counters + CNT_OVERFLOWS, (unsigned long long)s_over);
    }
}

// draw_image tail: *px = add(result, def_color, accum_neg_alpha) (renderer/mod.rs:410)
__device__ __forceinline__ Rgb8 final_color(const DevScene& S, const PixelState& st) {
    Rgb8 def{{S.shade.def_color[0], S.shade.def_color[1], S.shade.def_color[2]}};
    return add_rgb(st.result, def, st.accum_neg_alpha);
}

template <bool TRACE>
__device__ __forceinline__ void write_pixel(const DevScene& S, const DevBuffers& B, const MarchOut& O, size_t pixel, bool active,
                                            const PixelState& st, int consumed) {
    if (active) {
        Rgb8 px = final_color(S, st);
        if (O.rgb) {
            O.rgb[pixel * 3 + 0] = px.c[0];
            O.rgb[pixel * 3 + 1] = px.c[1];
            O.rgb[pixel * 3 + 2] = px.c[2];
        }
        if (O.meta) {
            atmrt_meta mm{st.m_lat, st.m_lon, st.m_elev, st.m_dist};
            O.meta[pixel] = mm;
        }
        if (O.steps) O.steps[pixel] = consumed;
        if (TRACE && O.counts) O.counts[pixel] = st.count;
    }
    count_pixel(B, active, st, consumed);
}

// When a render is eligible for the horizon sweep (below) the general march is still launched, as the
// fallback: `when` tells it for which outcome of the device-side checks it has to run.
enum MarchWhen { MARCH_ALWAYS = 0, MARCH_IF_NOT_SWEPT = 1, MARCH_FLAGGED_COLUMNS = 2 };

template <bool OBJECTS, bool BRUTE, bool TRACE>
__device__ __forceinline__ void march_column(const DevScene& S, const DevBuffers& B, const MarchOut& O, int xl) {
    const int y = blockIdx.x * MARCH_THREADS + threadIdx.x;
    const bool active = y < S.height;
    const int yy = active ? y : S.height - 1;  // inactive lanes shadow the last row and write nothing
    const int wl = S.x1 - S.x0;
    const size_t pixel = (size_t)yy * wl + xl;
    const int nlim = min(S.n_t, B.p_n[yy]);
    const size_t tbase = (size_t)xl * S.n_pad;
    const size_t hp = (size_t)S.h_pad;
    const size_t prow = path_index(S.n_t, 0, yy);

    PixelState st;
    init_pixel(st);
    int consumed = nlim > 0 ? nlim - 1 : 0;
    bool finished = !active || nlim < 2;
    int k = 1;

    if (!BRUTE && !finished) {  // level 3: can this ray cross this column's terrain at all?
        bool cand = B.rmin3[yy] < B.tmax3[xl] && B.rmax3[yy] > B.tmin3[xl];
        if (OBJECTS) cand = cand || B.close3[xl] != 0;
        if (!cand) finished = true;
    }
    double d_prev = 0.0;
    bool have_prev = false;
    for (;;) {
        // ---- phase 1: search for this lane's next event ----
        int ev = -1;
        if (!finished) {
            while (k < nlim) {
                if (!BRUTE) {
                    if (k == 1 || (k & 1023) == 0) {
                        const int C = k >> 10;
                        const double rmin = B.rmin2[(size_t)C * hp + yy], rmax = B.rmax2[(size_t)C * hp + yy];
                        bool cand = rmin < B.tmax2[(size_t)xl * S.n2 + C] && rmax > B.tmin2[(size_t)xl * S.n2 + C];
                        if (OBJECTS) cand = cand || (B.close2[(size_t)xl * S.n2 + C] != 0 && rmin <= rmax);
                        if (!cand) {
                            k = (C + 1) << 10;
                            have_prev = false;
                            continue;
                        }
                    }
                    if (k == 1 || (k & 31) == 0) {
                        const int c = k >> 5;
                        const double rmin = B.rmin1[(size_t)c * hp + yy], rmax = B.rmax1[(size_t)c * hp + yy];
                        bool cand = rmin < B.tmax1[(size_t)xl * S.n1_pad + c] && rmax > B.tmin1[(size_t)xl * S.n1_pad + c];
                        if (OBJECTS) cand = cand || (B.close1[(size_t)xl * S.n1_pad + c] != 0 && rmin <= rmax);
                        if (!cand) {
                            k = (c + 1) << 5;
                            have_prev = false;
                            continue;
                        }
                    }
                }
                // the reference's test: diff1 * diff2 < 0.0 (utils.rs:220-222); diff1 of step k is diff2 of step k-1
                const double dp = have_prev ? d_prev : B.p_elev[prow + (size_t)(k - 1) * PATH_ROWS] - B.t_elev[tbase + k - 1];
                const double dn = B.p_elev[prow + (size_t)k * PATH_ROWS] - B.t_elev[tbase + k];
                d_prev = dn;
                have_prev = true;
                bool e = dp * dn < 0.0;
                if (OBJECTS) e = e || (B.t_close[tbase + k - 1] | B.t_close[tbase + k]) != 0;
                if (e) {
                    ev = k;
                    break;
                }
                ++k;
            }
            if (ev < 0) finished = true;
        }
        if (__all_sync(FULL, finished)) break;
        // ---- phase 2: lanes holding an event process it together ----
        if (ev >= 0) {
            if (process_step<OBJECTS, TRACE>(S, B, O, xl, yy, ev, pixel, st)) {
                finished = true;
                consumed = ev;
            } else {
                k = ev + 1;
            }
        }
    }

    write_pixel<TRACE>(S, B, O, pixel, active, st, consumed);
}

// grid = (rows / MARCH_THREADS, columns per pass): a block walks the columns blockIdx.y, blockIdx.y +
// gridDim.y, ... so that the fallback launches behind the horizon sweep can use a small grid.
template <bool OBJECTS, bool BRUTE, bool TRACE>
__global__ void __launch_bounds__(MARCH_THREADS, 16) k_march(const __grid_constant__ DevScene S, DevBuffers B, MarchOut O, int when) {
    if (when == MARCH_IF_NOT_SWEPT && B.sweep_flags[0] == 0) return;
    if (when == MARCH_FLAGGED_COLUMNS && (B.sweep_flags[0] != 0 || B.sweep_flags[1] == 0)) return;
    const int wl = S.x1 - S.x0;
    for (int xl = blockIdx.y; xl < wl; xl += gridDim.y) {
        if (when == MARCH_FLAGGED_COLUMNS && B.sweep_col[xl] == 0) continue;
        march_column<OBJECTS, BRUTE, TRACE>(S, B, O, xl);
    }
}

// ---------------------------------------------------------------------------------------------
// Stage C, general case (translucent terrain and / or objects), when the rays of the render do not cross: the
// crossing march.
//
// get_single_pixel (utils.rs:201-289) looks at every step of every ray, and acts at the steps where
// diff1 * diff2 < 0 or an object is close. When the rays of a column are ordered -- r[k][y] >= r[k][y+1], a NaN
// ray only above a finite one, an upper ray at least as long as a lower one: what k_path_check verifies on the
// path cache of THIS render -- "the ray is above the terrain at step k" is a monotone predicate of the row, so
// one number per terrain sample describes a whole column of pixels at that step: Y[x][k], how many rows (from the
// top) have r - t > 0 (or NaN). A pixel (x, y) can only have a sign change at step k if y lies between Y[x][k-1]
// and Y[x][k]; the steps with a close object concern every row.
//
//   k_thresholds    one thread per terrain sample: Y[x][k] by bisection over the rows that are still alive
//                   at step k, and the flag "an object is close at step k - 1 or k"; per window of 32 steps the
//                   range of the thresholds (a band skips the windows that cannot concern it).
//   k_cross_march   one warp per (column, band of CROSS_BAND rows) walks the windows of its column that concern it
//                   and lists the (step, row) pairs of its band in step order. Thirty-two at a time
//                   the pairs are filtered -- the ray is still live, and either its row changed sides or one of
//                   the close objects is within reach of the ray's segment (object_out_of_reach: an exact
//                   rejection, most rows of a step next to an object pass it by) -- and whenever 32 survivors
//                   are queued the lanes take one each and run the reference's step on it (process_step: the
//                   exact product test, interpolation, normals, objects, stable sort, compositing) on the
//                   pixel's state, which lives in shared memory. Pairs of the same row inside one batch run in
//                   queue order. The survivors are a superset of the reference's events (an exact zero or a
//                   NaN on one side changes the predicate without being an event; process_step tests the
//                   product), so the result is the general march's, pixel for pixel.
//
// Both return at once when k_path_check found crossing rays; the general march then runs behind them.
// ---------------------------------------------------------------------------------------------
constexpr unsigned CROSS_CLOSE = 0x8000u;  // Y holds a row count <= 32767 (atmrt_set_params limits the height)
constexpr int CROSS_BAND = 64;             // rows per warp
constexpr int CROSS_WARPS = 4;
constexpr int CROSS_QUEUE = 64;            // >= 31 + 32, a power of two

__global__ void __launch_bounds__(128) k_thresholds(const __grid_constant__ DevScene S, DevBuffers B, unsigned short* __restrict__ thresholds,
                                                    unsigned short* __restrict__ window_min, unsigned short* __restrict__ window_max, int objects,
                                                    unsigned short* __restrict__ trig_slot, double* __restrict__ trig, int* __restrict__ trig_count) {
    if (B.sweep_flags[0] != 0) return;
    const int k = blockIdx.x * 128 + threadIdx.x, xl = blockIdx.y;
    const bool valid = k < S.n_t;
    unsigned v = 0;
    if (valid) {
        // rows alive at step k: p_n[y] > k, and p_n does not increase with y
        int lo = 0, hi = S.height;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (B.p_n[mid] > k) lo = mid + 1;
            else hi = mid;
        }
        const size_t ti = (size_t)xl * S.n_pad + k;
        const double t = B.t_elev[ti];
        hi = lo, lo = 0;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const double d = B.p_elev[path_index(S.n_t, k, mid)] - t;
            if (!(d <= 0.0)) lo = mid + 1;
            else hi = mid;
        }
        v = (unsigned)lo;
        if (objects) {
            const unsigned long long here = B.t_close[ti], before = k >= 1 ? B.t_close[ti - 1] : 0ull, after = k + 1 < S.n_t ? B.t_close[ti + 1] : 0ull;
            if (k >= 1 && (before | here) != 0) v |= CROSS_CLOSE;
            unsigned slot = CROSS_NO_SLOT;
            if ((before | here | after) != 0) {  // an end of a step with a close object: as_cartesian's sines and cosines (mod.rs:59-93)
                const int got = atomicAdd(trig_count + xl, 1);
                if (got < CROSS_TRIG_CAP) {
                    double* t4 = trig + ((size_t)xl * CROSS_TRIG_CAP + got) * 4;
                    double sinlat = 0.0, coslat = 1.0, sinlon, coslon;
                    sincos(to_radians(B.t_lon[ti]), &sinlon, &coslon);
                    if (!S.earth.flat_dirs) sincos(to_radians(B.t_lat[ti]), &sinlat, &coslat);
                    t4[0] = sinlat, t4[1] = coslat, t4[2] = sinlon, t4[3] = coslon;
                    slot = (unsigned)got;
                }
            }
            trig_slot[ti] = (unsigned short)slot;
        }
        thresholds[ti] = (unsigned short)v;
    }
    // the window of 32 steps this warp covers: the range of its thresholds, and whether an object is close in it
    const unsigned y = v & 0x7fffu;
    const unsigned lo_y = __reduce_min_sync(FULL, valid ? y : 0x7fffu), hi_y = __reduce_max_sync(FULL, valid ? y : 0u);
    const bool close = __any_sync(FULL, (v & CROSS_CLOSE) != 0);
    if ((threadIdx.x & 31) == 0 && valid) {
        const size_t wi = (size_t)xl * S.n1_pad + (k >> 5);
        window_min[wi] = (unsigned short)lo_y;
        window_max[wi] = (unsigned short)(hi_y | (close ? CROSS_CLOSE : 0u));
    }
}

struct CrossPixel {
    PixelState st;
    int nlim;    // steps of the ray inside the terrain profile
    int done_k;  // the step the pixel finished at, -1 while it is live
    double pad_; // 72 bytes, not 64: lanes work on adjacent rows of the band in shared memory, and a stride of 16 words puts
                 // every second row on the same banks (2/3 of the kernel's shared-memory wavefronts were conflicts)
};

// Can the segment pos1 -> pos2 touch object o at all? Every point Frustum / Billboard::check_collision returns lies
// between the planes h = 0 and h = height along the object's axis and within max(r1, r2) (half the width) of it. The
// height is linear along the segment and the distance to the axis cannot drop below the nearer end's by more than
// the segment's length, so a segment with both ends beyond one plane, or with both ends farther from the axis than
// that, yields no collision. The 1 mm margins are a million times the rounding of these dot products (coordinates
// of 6.4e6 m, 1e-16 relative). NaN compares false: never rejected.
__device__ __forceinline__ bool object_out_of_reach(const DevObject& o, V3 pos1, V3 pos2, double seg_len) {
    const double margin = 1.0e-3;
    const V3 p1 = pos1 - o.pos, p2 = pos2 - o.pos;
    const double h1 = dot(p1, o.up), h2 = dot(p2, o.up);
    if (fmax(h1, h2) < -margin || fmin(h1, h2) > o.height + margin) return true;
    const double reach = (o.kind == ATMRT_OBJECT_FRUSTUM ? fmax(o.r1, o.r2) : 0.5 * o.width) + seg_len + margin;
    const double q1 = dot(p1, p1) - h1 * h1, q2 = dot(p2, p2) - h2 * h2;  // squared distances to the axis
    return fmin(q1, q2) > reach * reach;
}

template <bool OBJECTS>
__global__ void __launch_bounds__(32 * CROSS_WARPS, 6) k_cross_march(const __grid_constant__ DevScene S, DevBuffers B, MarchOut O,
                                                                     const unsigned short* __restrict__ thresholds,
                                                                     const unsigned short* __restrict__ window_min,
                                                                     const unsigned short* __restrict__ window_max,
                                                                     const unsigned short* __restrict__ trig_slot, const double* __restrict__ trig) {
    if (B.sweep_flags[0] != 0) return;
    __shared__ CrossPixel pix_smem[CROSS_WARPS][CROSS_BAND];
    __shared__ int queue_smem[CROSS_WARPS][CROSS_QUEUE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the warps of a block: the same band of adjacent columns (alike in work, and they read the same rows of the path cache)
    const int wl = S.x1 - S.x0, xl = blockIdx.x * CROSS_WARPS + warp;
    const int b0 = blockIdx.y * CROSS_BAND;
    if (xl >= wl) return;
    const int b1 = min(b0 + CROSS_BAND, S.height);
    CrossPixel* pix = pix_smem[warp];
    int* queue = queue_smem[warp];  // the (step, row) pairs that can have an effect, in step order
    int nmax = 0;
    for (int r = lane; r < CROSS_BAND; r += 32) {
        CrossPixel& c = pix[r];
        init_pixel(c.st);
        c.nlim = b0 + r < b1 ? min(S.n_t, B.p_n[b0 + r]) : 0;
        c.done_k = -1;
        nmax = max(nmax, c.nlim);
    }
    nmax = __reduce_max_sync(FULL, nmax);  // the band's longest ray
    __syncwarp();
    const size_t tbase = (size_t)xl * S.n_pad, wbase = (size_t)xl * S.n1_pad;
    const unsigned short* Y = thresholds + tbase;
    const int nwin = (nmax + 31) >> 5;
    int head = 0, tail = 0;  // queue positions (monotone; the slot is position & (CROSS_QUEUE - 1))
    static_assert(CROSS_BAND == 64, "a queued pair is (step << 6) | row of the band");
    // the enumeration, outermost first: 32 windows at a time (w0, wmask) -> one window (kb; per lane the rows
    // [first, last) of step kb + lane, [t_first, t_last) of them because the row changed sides; `any`: the steps that
    // have rows) -> one step (cur_k and its rows cur_f + [cur_i, cur_n))
    int w0 = -32, kb = 0, first = 0, last = 0, t_first = 0, t_last = 0;
    unsigned wmask = 0, any = 0;
    int cur_k = 0, cur_f = 0, cur_n = 0, cur_i = 0, cur_tf = 0, cur_tl = 0;
    bool enumerated = false;

    for (;;) {
        if (tail - head >= 32 || (tail > head && enumerated)) {
            // ---- the lanes take one queued pair each and run the reference's step; pairs of one row run in queue order ----
            const int n = min(32, tail - head);
            const bool have = lane < n;
            const int ev = have ? queue[(head + lane) & (CROSS_QUEUE - 1)] : -1 - lane;
            const int r = ev & (CROSS_BAND - 1), k = ev >> 6;
            const unsigned same = __match_any_sync(FULL, have ? r : -1 - lane);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(FULL, have ? rank : 0);
            for (int round = 0; round <= rounds; ++round) {
                if (have && rank == round) {
                    CrossPixel& c = pix[r];
                    if (c.done_k < 0 && k < c.nlim) {
                        const int y = b0 + r;
                        PixelState st = c.st;
                        if (process_step<OBJECTS, false>(S, B, O, xl, y, k, (size_t)y * wl + xl, st, nullptr, trig_slot, trig)) c.done_k = k;
                        c.st = st;
                    }
                }
                __syncwarp();
            }
            head += n;
            continue;
        }
        if (enumerated) break;
        if (cur_i >= cur_n) {  // the next step that has rows
            if (!any) {        // ... in the next window that can have any
                if (!wmask) {
                    w0 += 32;
                    if (w0 >= nwin) {
                        enumerated = true;
                        continue;
                    }
                    // a window matters to the band if its thresholds (and the one before its first step: the window before
                    // it is included) reach into the band, or an object is close somewhere in it
                    const int w = w0 + lane;
                    bool matters = false;
                    if (w < nwin) {
                        const unsigned hi_w = window_max[wbase + w], lo_w = window_min[wbase + w];
                        const unsigned hi_p = w > 0 ? window_max[wbase + w - 1] : hi_w, lo_p = w > 0 ? window_min[wbase + w - 1] : lo_w;
                        matters = (OBJECTS && (hi_w & CROSS_CLOSE)) ||
                                  ((int)min(lo_w, lo_p) < b1 && (int)(max(hi_w, hi_p) & 0x7fffu) > b0);
                    }
                    wmask = __ballot_sync(FULL, matters);
                    continue;
                }
                kb = (w0 + __ffs(wmask) - 1) << 5;
                wmask &= wmask - 1;
                const int k = kb + lane;
                const bool in = k >= 1 && k < nmax;
                const unsigned v = in ? Y[k] : 0u;
                unsigned before = __shfl_up_sync(FULL, v, 1);
                if (lane == 0 && in) before = Y[k - 1];
                const int yk = (int)(v & 0x7fffu), yp = (int)(before & 0x7fffu);
                t_first = max(min(yk, yp), b0), t_last = min(max(yk, yp), b1);  // rows [t_first, t_last) changed sides
                first = t_first, last = t_last;
                if (OBJECTS && (v & CROSS_CLOSE)) first = b0, last = b1;
                if (!in) last = first;
                any = __ballot_sync(FULL, last > first);
                continue;
            }
            const int j = __ffs(any) - 1;
            any &= any - 1;
            cur_k = kb + j, cur_i = 0;
            cur_f = __shfl_sync(FULL, first, j), cur_n = __shfl_sync(FULL, last, j) - cur_f;
            cur_tf = __shfl_sync(FULL, t_first, j), cur_tl = __shfl_sync(FULL, t_last, j);
            continue;
        }
        // ---- the next 32 rows of the step: which can have an effect? a live ray, and a possible sign change or an object within reach ----
        {
            const int row = cur_f + cur_i + lane, k = cur_k;
            bool keep = false;
            if (cur_i + lane < cur_n) {
                const CrossPixel& c = pix[row - b0];
                if (c.done_k < 0 && k < c.nlim) {
                    keep = true;
                    if (OBJECTS && !(row >= cur_tf && row < cur_tl)) {
                        const size_t ti = tbase + k;
                        const unsigned s0 = trig_slot[ti - 1], s1 = trig_slot[ti];
                        if (s0 != CROSS_NO_SLOT && s1 != CROSS_NO_SLOT) {
                            const double* q0 = trig + ((size_t)xl * CROSS_TRIG_CAP + s0) * 4;
                            const double* q1 = trig + ((size_t)xl * CROSS_TRIG_CAP + s1) * 4;
                            const size_t p1 = path_index(S.n_t, k, row);
                            const V3 pos1 = as_cartesian_sc(S.earth, B.t_lat[ti - 1], B.p_elev[p1 - PATH_ROWS], q0[0], q0[1], q0[2], q0[3]);
                            const V3 pos2 = as_cartesian_sc(S.earth, B.t_lat[ti], B.p_elev[p1], q1[0], q1[1], q1[2], q1[3]);
                            const V3 w = pos2 - pos1;
                            const double seg_len = sqrt(dot(w, w));
                            unsigned long long mask = B.t_close[ti - 1] | B.t_close[ti];
                            keep = false;
                            while (mask && !keep) {
                                const int oi = __ffsll((long long)mask) - 1;
                                mask &= mask - 1;
                                keep = !object_out_of_reach(B.objects[oi], pos1, pos2, seg_len);
                            }
                        }
                    }
                }
            }
            const unsigned kept = __ballot_sync(FULL, keep);
            if (keep) queue[(tail + __popc(kept & ((1u << lane) - 1u))) & (CROSS_QUEUE - 1)] = (k << 6) | (row - b0);
            tail += __popc(kept);
            cur_i += 32;
            __syncwarp();
        }
    }

    for (int r = lane; r < CROSS_BAND; r += 32) {  // all lanes: write_pixel votes
        const CrossPixel& c = pix[r];
        const bool active = b0 + r < b1;
        const int y = active ? b0 + r : b1 - 1;
        const int consumed = c.done_k >= 0 ? c.done_k : (c.nlim > 0 ? c.nlim - 1 : 0);
        write_pixel<false>(S, B, O, (size_t)y * wl + xl, active, c.st, consumed);
    }
}

// ---------------------------------------------------------------------------------------------
// The Rectilinear generator (generators/rectilinear.rs): a rectilinear projection, so every pixel has its
// own elevation AND azimuth -- one ray integration and one azimuth walk per pixel, nothing to cache. One
// thread per pixel runs the reference's PathIterator (rectilinear.rs:112-186) fused with get_single_pixel:
// per step one RK4 step (g(h) from the table in shared memory, device_paths.cuh), one DirectionalCalc walk
// and one bilinear terrain tap (plus objects_close when the scene has objects); the normals are evaluated at
// the hits only. The 32 lanes of a warp are 32 adjacent pixels of one row: neighbouring azimuths and the
// same elevation to within a pixel, so their terrain taps share the 128-byte micro-tiles.
// ---------------------------------------------------------------------------------------------
// RectilinearGenerator::get_ray_params, rectilinear.rs:80-105 (nalgebra: Rz(yaw) Ry(pitch) Rx(roll), the
// product accumulated column by column).
__device__ __forceinline__ void get_ray_params(const DevScene& S, int x, int y, double* elevation, double* direction) {
    const double width = (double)S.width;
    const double xf = (double)(short)((short)x - (short)S.width / 2);
    const double yf = (double)(short)((short)y - (short)S.height / 2);
    const double z = width / 2.0 / tan(to_radians(S.fov) / 2.0);
    double sr, cr, sp, cp, sy, cy;
    sincos(0.0, &sr, &cr);
    sincos(-to_radians(S.tilt), &sp, &cp);
    sincos(to_radians(S.direction), &sy, &cy);
    const double m[3][3] = {{cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr},
                            {sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr},
                            {-sp, cp * sr, cp * cr}};
    const double v[3] = {z, xf, -yf};  // [forward, right, up]
    double d[3];
    for (int i = 0; i < 3; ++i) d[i] = (m[i][0] * v[0] + m[i][1] * v[1]) + m[i][2] * v[2];
    const double n = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int i = 0; i < 3; ++i) d[i] = d[i] / n;
    *elevation = asin(d[2]);
    *direction = atan2(d[1], d[0]);
}

// ResultPixel.elevation_angle / azimuth of every pixel of the column block ([H][x1 - x0] each; either may be
// null). Fast generator, fast.rs:67-76: get_ray_elev(y) and get_ray_dir(x) wrapped ONCE into [0, 360);
// Rectilinear generator, rectilinear.rs:78-116: the pixel's own (elevation, direction).to_degrees(), not wrapped.
__global__ void __launch_bounds__(128) k_pixel_angles(const __grid_constant__ DevScene S, double* __restrict__ elevation_angle,
                                                      double* __restrict__ azimuth) {
    const int wl = S.x1 - S.x0;
    const int xl = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xl >= wl) return;
    double el, az;
    if (S.generator == ATMRT_GENERATOR_RECTILINEAR) {
        get_ray_params(S, S.x0 + xl, y, &el, &az);
        el = to_degrees(el), az = to_degrees(az);
    } else {
        el = get_ray_elev(S, y);
        az = get_ray_dir(S, S.x0 + xl);
        if (az < 0.0) az += 360.0;
        else if (az >= 360.0) az -= 360.0;
    }
    const size_t pixel = (size_t)y * wl + xl;
    if (elevation_angle) elevation_angle[pixel] = el;
    if (azimuth) azimuth[pixel] = az;
}

// ---------------------------------------------------------------------------------------------
// The InterpolatingRectilinear generator (generators/interpolating_rectilinear.rs): the rectilinear image's rays,
// served from a regular grid in (elevation, direction) whose points are Fast-generator pixels. The grid itself is an
// ordinary Fast render with explicit angle tables (DevScene::row_elev_deg / col_dir_deg) that keeps its trace lists;
// these kernels find the grid's steps and extent and blend the lists.
// ---------------------------------------------------------------------------------------------
struct InterpGrid {
    double elev_step, dir_step;  // FovData::min_elev_step / min_dir_step (radians)
    int elev_top;                // elevation index of grid row 0 (rows run downwards, like an image)
    int dir_left;                // direction index of grid column 0
    int rows, cols;
    int max_points;              // trace points kept per grid pixel
    const atmrt_trace_point* points;  // [rows][cols][max_points]
    const int* counts;                // [rows][cols], the true counts
};

// gen_fov_data, interpolating_rectilinear.rs:432-521: the smallest |difference| of elevation between vertical neighbours
// and of direction between horizontal neighbours over the WHOLE image, each difference raised to fov / width / 3. The
// differences are positive doubles: their bit patterns order like the values. mins[0], mins[1] start at 2 pi.
__global__ void __launch_bounds__(128) k_interp_steps(const __grid_constant__ DevScene S, unsigned long long* __restrict__ mins) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    double d_elev = 1.0e300, d_dir = 1.0e300;
    if (x < S.width) {
        const double two_pi = 360.0 * (PI / 180.0), min_diff = to_radians(S.fov) / (double)S.width / 3.0;
        double el, az, el2, az2;
        get_ray_params(S, x, y, &el, &az);
        if (y >= 1) {
            get_ray_params(S, x, y - 1, &el2, &az2);
            d_elev = fmax(fabs(el - el2), min_diff);  // (`if diff < min_diff { diff = min_diff }`; a NaN difference is never the minimum)
            if (!(d_elev == d_elev)) d_elev = 1.0e300;
        }
        if (x >= 1) {
            get_ray_params(S, x - 1, y, &el2, &az2);
            double diff = fabs(az - az2);
            if (diff > two_pi) diff -= two_pi;
            d_dir = fmax(diff, min_diff);
            if (!(d_dir == d_dir)) d_dir = 1.0e300;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        d_elev = fmin(d_elev, __shfl_xor_sync(FULL, d_elev, o));
        d_dir = fmin(d_dir, __shfl_xor_sync(FULL, d_dir, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mins + 0, (unsigned long long)__double_as_longlong(d_elev));
        atomicMin(mins + 1, (unsigned long long)__double_as_longlong(d_dir));
    }
}

// FovData::cache_coords, :185-204, of one pixel
struct InterpCorner {
    int elev_index, dir_index;
    double rem_elev, rem_dir;
};
__device__ __forceinline__ InterpCorner interp_corner(const DevScene& S, double elev_step, double dir_step, int x, int y) {
    double el, az;
    get_ray_params(S, x, y, &el, &az);
    const double ef = el / elev_step, df = az / dir_step;
    InterpCorner c;
    c.elev_index = (int)floor(ef), c.dir_index = (int)floor(df);
    c.rem_elev = ef - (double)c.elev_index, c.rem_dir = df - (double)c.dir_index;
    return c;
}

// the index ranges the column block needs: range = {min elev, max elev, min dir, max dir} (starting at +-INT_MAX)
__global__ void __launch_bounds__(128) k_interp_range(const __grid_constant__ DevScene S, double elev_step, double dir_step, int* __restrict__ range) {
    const int xl = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    int e_lo = INT_MAX, e_hi = -INT_MAX, d_lo = INT_MAX, d_hi = -INT_MAX;
    if (xl < S.x1 - S.x0) {
        const InterpCorner c = interp_corner(S, elev_step, dir_step, S.x0 + xl, y);
        e_lo = e_hi = c.elev_index, d_lo = d_hi = c.dir_index;
    }
    e_lo = __reduce_min_sync(FULL, e_lo), e_hi = __reduce_max_sync(FULL, e_hi);
    d_lo = __reduce_min_sync(FULL, d_lo), d_hi = __reduce_max_sync(FULL, d_hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(range + 0, e_lo), atomicMax(range + 1, e_hi);
        atomicMin(range + 2, d_lo), atomicMax(range + 3, d_hi);
    }
}

// TracePoint::interpolate / PixelColor::interpolate (generators/mod.rs:32-80): self * (1 - coeff) + other * coeff
struct BlendPoint {
    double lat, lon, distance, elevation, path_length;
    V3 normal;
    Color4 color;
    bool is_terrain;
    int step;
};
__device__ __forceinline__ BlendPoint load_point(const atmrt_trace_point& t) {
    return {t.lat, t.lon, t.distance, t.elevation, t.path_length, V3{t.normal[0], t.normal[1], t.normal[2]},
            Color4{t.color[0], t.color[1], t.color[2], t.color[3]}, t.is_terrain != 0, t.step};
}
__device__ __forceinline__ BlendPoint blend(const BlendPoint& a, const BlendPoint& b, double coeff) {
    const double ca = 1.0 - coeff;
    BlendPoint r;
    r.lat = a.lat * ca + b.lat * coeff, r.lon = a.lon * ca + b.lon * coeff, r.distance = a.distance * ca + b.distance * coeff;
    r.elevation = a.elevation * ca + b.elevation * coeff, r.path_length = a.path_length * ca + b.path_length * coeff;
    r.normal = a.normal * ca + b.normal * coeff;
    r.is_terrain = a.is_terrain || b.is_terrain;
    if (a.is_terrain && b.is_terrain) r.color = Color4{0.0, 0.0, 0.0, a.color.a * ca + b.color.a * coeff};
    else if (!a.is_terrain && !b.is_terrain)
        r.color = Color4{a.color.r * ca + b.color.r * coeff, a.color.g * ca + b.color.g * coeff, a.color.b * ca + b.color.b * coeff, a.color.a * ca + b.color.a * coeff};
    else r.color = Color4{0.0, 0.0, 0.0, a.is_terrain ? a.color.a : b.color.a};
    r.step = a.step;
    return r;
}

// interpolate_trace_points, :274-345: `e[i]` the group's point of grid pixel SEQUENCE[i] = (elev + i / 2, dir + i % 2)
__device__ inline bool blend_group(const atmrt_trace_point* const e[4], double re, double rd, BlendPoint* out) {
    const int have = (e[0] ? 1 : 0) | (e[1] ? 2 : 0) | (e[2] ? 4 : 0) | (e[3] ? 8 : 0);
    int a = -1, b = -1, c = -1;  // the points in the order the rule takes them
    double r_elev = re, r_dir = rd;
    int rule;  // 1: single, 2: two adjacent, 3: two diagonal, 4: three, 5: four
    switch (have) {
        case 1: rule = 1, a = 0; if (!(re < 0.5 && rd < 0.5)) return false; break;
        case 2: rule = 1, a = 1; if (!(re < 0.5 && rd >= 0.5)) return false; break;
        case 4: rule = 1, a = 2; if (!(re >= 0.5 && rd < 0.5)) return false; break;
        case 8: rule = 1, a = 3; if (!(re >= 0.5 && rd >= 0.5)) return false; break;
        case 1 | 2: rule = 2, a = 0, b = 1; break;
        case 1 | 4: rule = 2, a = 0, b = 2, r_elev = rd, r_dir = re; break;
        case 1 | 8: rule = 3, a = 0, b = 3; break;
        case 2 | 4: rule = 3, a = 1, b = 2, r_dir = 1.0 - rd; break;
        case 2 | 8: rule = 2, a = 1, b = 3, r_elev = 1.0 - rd, r_dir = re; break;
        case 4 | 8: rule = 2, a = 2, b = 3, r_elev = 1.0 - re; break;
        case 1 | 2 | 4: rule = 4, a = 0, b = 1, c = 2; break;
        case 1 | 2 | 8: rule = 4, a = 1, b = 0, c = 3, r_dir = 1.0 - rd; break;
        case 1 | 4 | 8: rule = 4, a = 0, b = 3, c = 2, r_elev = 1.0 - re; break;
        case 2 | 4 | 8: rule = 4, a = 3, b = 2, c = 1, r_elev = 1.0 - re, r_dir = 1.0 - rd; break;
        case 15: rule = 5; break;
        default: return false;
    }
    if (rule == 1) {
        *out = load_point(*e[a]);
    } else if (rule == 2) {  // interpolate_two_adjacent
        if (r_elev >= 0.5) return false;
        *out = blend(load_point(*e[a]), load_point(*e[b]), r_dir);
    } else if (rule == 3) {  // interpolate_two_diagonal
        if ((r_elev >= 0.5 && r_dir < 0.5) || (r_elev < 0.5 && r_dir >= 0.5)) return false;
        const double coeff = r_elev * r_dir / (r_elev * r_dir + (1.0 - r_elev) * (1.0 - r_dir));
        *out = blend(load_point(*e[a]), load_point(*e[b]), coeff);
    } else if (rule == 4) {  // interpolate_three
        if (r_elev >= 0.5 && r_dir >= 0.5) return false;
        const double sum = 1.0 - r_elev + r_elev * (1.0 - r_dir);
        *out = blend(blend(load_point(*e[a]), load_point(*e[b]), r_dir), load_point(*e[c]), r_elev * (1.0 - r_dir) / sum);
    } else {  // interpolate_four
        *out = blend(blend(load_point(*e[0]), load_point(*e[1]), rd), blend(load_point(*e[2]), load_point(*e[3]), rd), re);
    }
    return true;
}

constexpr int INTERP_MAX_POINTS = 16;  // trace points kept per grid pixel (a grid pixel with more is counted in step_overflows)

// interpolate (:394-419) + draw_image for one image pixel: group the trace points of the four grid pixels
// (collect_trace_points, :213-243: a point joins the FIRST group holding a point of its class closer than one simulation
// step in distance, else opens one; a later point of the same grid pixel replaces the earlier one in the group's slot),
// blend every group and composite the results in group order.
template <bool TRACE>
__global__ void __launch_bounds__(128) k_interp_blend(const __grid_constant__ DevScene S, DevBuffers B, MarchOut O, InterpGrid G) {
    const int wl = S.x1 - S.x0;
    const int xl = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const bool active = xl < wl;
    const int xx = min(xl, wl - 1);
    const size_t pixel = (size_t)y * wl + xx;
    const InterpCorner c = interp_corner(S, G.elev_step, G.dir_step, S.x0 + xx, y);
    PixelState st;
    init_pixel(st);
    signed char group_of[4][INTERP_MAX_POINTS];  // the group of every point of the four lists
    const atmrt_trace_point* list[4];
    int n[4], ngroups = 0;
    bool overflow = false;
    for (int q = 0; q < 4; ++q) {
        const int row = G.elev_top - (c.elev_index + q / 2), col = c.dir_index + q % 2 - G.dir_left;
        const bool inside = row >= 0 && row < G.rows && col >= 0 && col < G.cols;  // (always, for the block the grid was made for)
        const size_t g = inside ? (size_t)row * G.cols + col : 0;
        list[q] = G.points + g * G.max_points;
        const int cnt = inside ? G.counts[g] : 0;
        overflow |= cnt > G.max_points;
        n[q] = min(min(cnt, G.max_points), INTERP_MAX_POINTS);
    }
    for (int q = 0; q < 4; ++q)
        for (int i = 0; i < n[q]; ++i) {
            const double dist = list[q][i].distance;
            const int cls = list[q][i].is_terrain;
            int g = ngroups;  // the first group (in order of creation) that holds a close point of the same class
            for (int q2 = 0; q2 <= q; ++q2)
                for (int i2 = 0; i2 < (q2 == q ? i : n[q2]); ++i2)
                    if (group_of[q2][i2] < g && fabs(dist - list[q2][i2].distance) < S.step && cls == list[q2][i2].is_terrain) g = group_of[q2][i2];
            if (g == ngroups) ++ngroups;
            group_of[q][i] = (signed char)g;
        }
    for (int g = 0; g < ngroups; ++g) {
        const atmrt_trace_point* e[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int q = 0; q < 4; ++q)
            for (int i = 0; i < n[q]; ++i)
                if (group_of[q][i] == g) e[q] = list[q] + i;  // match_sequence, :245-266: the last one wins the slot
        BlendPoint p;
        if (blend_group(e, c.rem_elev, c.rem_dir, &p))
            emit_point<TRACE>(S, O, pixel, p.step, st, p.is_terrain, p.lat, p.lon, p.distance, p.elevation, p.path_length, p.normal, p.color);
    }
    if (overflow) st.overflows += 1;
    write_pixel<TRACE>(S, B, O, pixel, active, st, 0);
}

// ResultPixel.elevation_angle / azimuth of the blended pixel (interpolate, :405-417): the four grid pixels' angles (azimuth
// wrapped once into [0, 360), Cache::get_pixel :97-103) weighted with the remainders
__global__ void __launch_bounds__(128) k_interp_angles(const __grid_constant__ DevScene S, double elev_step, double dir_step, double* __restrict__ elevation_angle,
                                                       double* __restrict__ azimuth) {
    const int wl = S.x1 - S.x0;
    const int xl = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (xl >= wl) return;
    const InterpCorner c = interp_corner(S, elev_step, dir_step, S.x0 + xl, y);
    double e4[4], a4[4];
    for (int q = 0; q < 4; ++q) {
        e4[q] = to_degrees((double)(c.elev_index + q / 2) * elev_step);
        double a = to_degrees((double)(c.dir_index + q % 2) * dir_step);
        if (a < 0.0) a += 360.0;
        else if (a >= 360.0) a -= 360.0;
        a4[q] = a;
    }
    const double re = c.rem_elev, rd = c.rem_dir;
    const size_t pixel = (size_t)y * wl + xl;
    if (elevation_angle) elevation_angle[pixel] = e4[0] * (1.0 - re) * (1.0 - rd) + e4[1] * (1.0 - re) * rd + e4[2] * re * (1.0 - rd) + e4[3] * re * rd;
    if (azimuth) azimuth[pixel] = a4[0] * (1.0 - re) * (1.0 - rd) + a4[1] * (1.0 - re) * rd + a4[2] * re * (1.0 - rd) + a4[3] * re * rd;
}

constexpr int RECT_THREADS = 128;

template <bool FLAT, bool OBJECTS>
__global__ void __launch_bounds__(RECT_THREADS, 4) k_rectilinear(const __grid_constant__ DevScene S, DevBuffers B, MarchOut O, int libm_only) {
    __shared__ double tab_smem[ATM_FIELDS * ATM_CELLS];
#pragma unroll 4
    for (int i = threadIdx.x; i < ATM_FIELDS * ATM_CELLS; i += RECT_THREADS) tab_smem[i] = B.atm_cells[i];
    __syncthreads();
    const GSource gs{(unsigned)__cvta_generic_to_shared(tab_smem), B.atm_pieces, B.n_atm_pieces};
    const int wl = S.x1 - S.x0;
    const int xl = blockIdx.x * RECT_THREADS + threadIdx.x, y = blockIdx.y;
    const bool active = xl < wl;
    const int xx = min(xl, wl - 1);
    const size_t pixel = (size_t)y * wl + xx;
    double elevation, direction;
    get_ray_params(S, S.x0 + xx, y, &elevation, &direction);
    double cc[8];
    direction_calc(S, to_degrees(direction), cc);  // params.model.coords_at_dist_calc(position, direction.to_degrees())
    const double alt = *B.obs_alt;
    const double radius = S.radius;
    const double d = FLAT ? S.step : S.step / radius;
    const double hd = 0.5 * d, d6 = d / 6.0;
    const double shift = FLAT ? ATM_BASE : radius + ATM_BASE;
    // env.cast_ray_stepper(alt, elevation, straight_rays)
    Stepper line;
    stepper_init(line, FLAT, radius, alt, elevation);
    double a = line.a, b = line.b;
    PathBase bE = path_base<FLAT>(gs.tab, shift, a), bM = path_base<FLAT>(gs.tab, shift, fma(hd, b, a)), bN = path_base<FLAT>(gs.tab, shift, fma(d, b, a));
    const double d15 = 1.5 * d, d2 = 2.0 * d;

    PixelState st;
    init_pixel(st);
    int consumed = 0;
    // point 0: the observer's state over the terrain under it
    double lat0, lon0;
    V3 fpos0{0.0, 0.0, 0.0};
    walk_coords(S, cc, 0.0, &lat0, &lon0, &fpos0);
    double elev0 = elev_or_zero(B.terrain, lat0, lon0);
    unsigned long long mask0 = OBJECTS ? objects_close(S, B, fpos0, lat0, lon0) : 0ull;
    double ray0 = alt, len0 = 0.0;
    RayState prev{0.0, alt};
    if (!(0.0 > S.max_distance || alt < -1000.0)) {
#pragma unroll 1
        for (int k = 1; k < S.n_x; ++k) {
            // self.ray.next(): state k
            double x1, h1;
            if (S.straight) {
                const RayState nw = stepper_next(line, S.atm, FLAT, 1, radius, S.step);
                x1 = nw.x, h1 = nw.h;
            } else {
                double a_new, b_new;
                bool ok = false;
                if (!libm_only) {
                    ok = rk4_step_shared<FLAT>(d, hd, d6, a, b, bE, bM, bN, &a_new, &b_new) || a != a || b != b;
                    bE = bN;
                    bM = path_base<FLAT>(gs.tab, shift, fma(d15, b, a));
                    bN = path_base<FLAT>(gs.tab, shift, fma(d2, b, a));
                }
                if (!ok) {
                    if (libm_only) rk4_step<FLAT, 2>(S.atm, gs, radius, d, hd, d6, a, b, &a_new, &b_new);
                    else rk4_step<FLAT, 1>(S.atm, gs, radius, d, hd, d6, a, b, &a_new, &b_new);
                }
                a = a_new, b = b_new;
                x1 = B.path_x[k];
                h1 = FLAT ? a : a - radius;
            }
            const double len1 = len0 + calc_dist(FLAT, radius, prev, RayState{x1, h1});
            // the point exists unless it is past max_distance or below -1000 m (rectilinear.rs:175-177)
            if (x1 > S.max_distance || h1 < -1000.0) break;
            consumed = k;
            double lat1, lon1;
            V3 fpos1{0.0, 0.0, 0.0};
            walk_coords(S, cc, x1, &lat1, &lon1, &fpos1);
            const double elev1 = elev_or_zero(B.terrain, lat1, lon1);
            const unsigned long long mask1 = OBJECTS ? objects_close(S, B, fpos1, lat1, lon1) : 0ull;
            const double diff1 = ray0 - elev0, diff2 = h1 - elev1;
            if (diff1 * diff2 < 0.0 || (OBJECTS && (mask0 | mask1))) {
                const StepEnds e{lat0, lon0, elev0, ray0, prev.x, len0, lat1, lon1, elev1, h1, x1, len1, mask0 | mask1};
                const bool finish = process_ends<OBJECTS, false>(S, B, O, e, k, pixel, st, [&](V3* n0, V3* n1) {
                    const SampleTrig t0 = sample_trig(S, fpos0, lat0, lon0), t1 = sample_trig(S, fpos1, lat1, lon1);
                    *n0 = find_normal(S, B.terrain, lat0, lon0, t0.sinlat, t0.coslat, t0.sinlon, t0.coslon);
                    *n1 = find_normal(S, B.terrain, lat1, lon1, t1.sinlat, t1.coslat, t1.sinlon, t1.coslon);
                });
                if (finish) break;
            }
            lat0 = lat1, lon0 = lon1, elev0 = elev1, mask0 = mask1, fpos0 = fpos1;
            ray0 = h1, len0 = len1;
            prev = RayState{x1, h1};
        }
    }
    write_pixel<false>(S, B, O, pixel, active, st, consumed);
    // points the stream produced (states integrated + 1), the Rectilinear generator's count of path steps
    unsigned long long pts = active && !(0.0 > S.max_distance || alt < -1000.0) ? (unsigned long long)consumed + 1ull : 0ull;
    for (int o = 16; o > 0; o >>= 1) pts += __shfl_xor_sync(FULL, pts, o);
    if ((threadIdx.x & 31) == 0 && pts) atomicAdd(B.counters + CNT_PATH_STEPS, pts);
}

// ---------------------------------------------------------------------------------------------
// Stage C for opaque terrain without objects: the horizon sweep.
//
// With terrain_alpha == 1 and no objects a pixel ends at its FIRST sign change of ray - terrain
// (utils.rs:220-239). If the rays of one column never cross each other -- r[k][y] >= r[k][y+1] for
// every step k, row y above row y+1 -- and every ray starts above the terrain, then the first-hit step
// is monotone in the row: while row y+1 has not hit (its differences are all > 0), the differences of
// row y are larger still, so row y's first crossing cannot come before row y+1's. One thread per column
// therefore walks k forward and y upward, testing exactly the reference's product on the cells it
// visits: O(N_t + H) work per column instead of a search per pixel. k_path_check verifies the
// monotonicity on the path cache of THIS render (any atmosphere: ducting makes rays cross); a column
// where a visited difference is exactly 0, the start is not above the terrain, or a hit does not go
// from above to below is flagged. The general march (k_march) runs for the whole image when the check
// fails and for the flagged columns otherwise, so the result is always the reference's.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_path_check(const double* __restrict__ elev, const int* __restrict__ lens, int h, int h_pad, int n_t,
                                                    unsigned* __restrict__ flags) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;  // compares row y (above) with row y + 1
    const int k0 = blockIdx.y * 64;
    bool bad = false;
    if (y + 1 < h) {
        const int n_up = min(n_t, lens[y]), n_dn = min(n_t, lens[y + 1]);
        if (n_up < n_dn) bad = true;  // the upper ray must live at least as long
        const int k1 = min(k0 + 64, n_dn);
        for (int k = k0; k < k1; ++k) {
            const double up = elev[path_index(n_t, k, y)], dn = elev[path_index(n_t, k, y + 1)];
            // A NaN ray (it left the atmosphere model) never hits; it must be the upper one of the pair.
            if (up < dn || (dn != dn && up == up)) bad = true;
        }
    }
    if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, 1u);
}

constexpr int SWEEP_ROWS = 4;  // rows resolved per window from one 32-byte load per lane

// ---------------------------------------------------------------------------------------------
// Stage C for opaque terrain without objects, three kernels:
//
//   k_sweep_bits   the horizon sweep with the sign tests of a whole window taken as warp votes: the lanes hold 32
//                  consecutive steps of a group of four rows, three ballots per row (ray above / below / exactly on the
//                  terrain) turn the window into bit masks, and every row of the group is then resolved by a handful of
//                  warp-uniform integer operations -- the first set bit of (above << 1) & below is the reference's first
//                  d1 * d2 < 0 seen from above (utils.rs:220-222). Per column it also lists the DISTINCT terrain samples
//                  its hits touch (sample k - 1 and k of every first-hit step k, in ascending order): the hit plane
//                  stores, per pixel, the position of its k in that list.
//   k_hit_normals  TerrainData::normal (find_normal, utils.rs:15-40) of every listed sample, one thread per sample: the
//                  deferred normals at full lanes, each distinct sample once (rows in the foreground share steps).
//   k_shade_tiles  interpolation, colouring and the metadata of the pixels; warps along the rows of one column (shared
//                  sectors in every cache), written row-major through shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int BITS_WARPS = 4;  // adjacent columns share a block (and the path windows in L1)

struct SweepLists {
    int* list;        // [wl][bands][cap]: the distinct samples of a band's hits, ascending; list[s - 1] == list[s] - 1 for every slot s a pixel refers to
    int* count;       // [wl][bands]
    double* normals;  // [wl][bands][cap][3]
    int cap;          // entries per (column, band)
    int bands;        // a column is swept in row bands, numbered from the top of the image ...
    int band_rows;    // ... of this many rows (a multiple of 32)
    int split;        // or, when > 0: two bands, rows [0, split) and [split, height) (a multiple of 32)
};
__host__ __device__ __forceinline__ int sweep_band_of_row(const SweepLists& L, int y) { return L.split > 0 ? (y >= L.split ? 1 : 0) : y / L.band_rows; }
// The hit plane holds, per pixel, 1 + the slot of its first-hit step in the band's list (0: no hit) -- and, when both fit 16 bits,
// the step itself above it: (step << 16) | (slot + 1). The shading then has the step after ONE dependent load instead of two
// (hit plane -> list -> caches was the chain that bound it). Warp-uniform, the same in the sweep and in its readers.
__device__ __forceinline__ bool sweep_hit_packed(const DevScene& S, const SweepLists& L) { return S.n_t <= 32767 && L.cap <= 65535; }

// predicated global stores (no branch): the sweep's control flow is warp-uniform, only lane 0 writes
__device__ __forceinline__ void st_if_s32(int* p, int v, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.s32 [%1], %2;\n\t}" ::"r"((unsigned)ok), "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_if_v4s32(int* p, int a, int b, int c, int d, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.v4.s32 [%1], {%2, %3, %4, %5};\n\t}" ::"r"((unsigned)ok), "l"(p), "r"(a), "r"(b),
                 "r"(c), "r"(d)
                 : "memory");
}

template <bool SPLIT>  // SPLIT: one band of a frame split in two (L.split), `only_band` of them; else every band of L.bands uniform ones
__global__ void __launch_bounds__(32 * BITS_WARPS, 8) k_sweep_bits(const __grid_constant__ DevScene S, DevBuffers B, SweepLists L, int col0, int col1,
                                                                   int only_band) {
    if (B.sweep_flags[0] != 0) return;
    const int lane = threadIdx.x & 31;
    const bool lane0 = lane == 0;
    const int xl = col0 + blockIdx.x * BITS_WARPS + (threadIdx.x >> 5);
    if (xl >= col1) return;
    // The walk of a column is one dependent chain; cut into row bands it is several shorter ones, which is what keeps a
    // narrow column block (one of eight GPUs) busy. By the monotonicity the sweep rests on, the walk enters a band in the
    // state the row just below the band leaves behind -- its first-hit step, or the end of its path -- which the band
    // finds by scanning that ONE row from step 1: the same cells, the same tests. blockIdx.y = 0 is the bottom band.
    // (`only_band` >= 0: a launch for that one band -- the lower band of a split frame is swept while the rays of the upper
    // one are still being integrated)
    const int band = SPLIT ? only_band : L.bands - 1 - (int)blockIdx.y;
    const int y_lo = SPLIT ? (band ? L.split : 0) : band * L.band_rows;
    const int y_hi = (SPLIT ? (band ? S.height : L.split) : min(S.height, y_lo + L.band_rows)) - 1;
    const double* __restrict__ te = B.t_elev + (size_t)xl * S.n_pad;
    int* __restrict__ hit = B.sweep_hit + (size_t)xl * S.h_pad;  // per pixel: 1 + the slot of its first-hit step in the band's list (0: no hit), the step above it (sweep_hit_packed)
    const int hit_mul = sweep_hit_packed(S, L) ? 65536 : 0;
    int* const list = L.list + ((size_t)xl * L.bands + band) * L.cap;
    const int n_t = S.n_t, k_last = n_t - 1;
    static_assert(SWEEP_ROWS == PATH_ROWS && SWEEP_ROWS == 4, "the sweep reads one row group of the path cache per 32-byte load");
    bool flagged = !(B.p_elev[0] - te[0] > 0.0);  // every ray starts at the observer altitude (element 0 of every row)
    // The window: lane j holds step kb + j; lane 0 is the "before" side of the first step a row may cross at, so a
    // window serves the steps kb + 1 .. kb + 31 and advances by 31. `lo`: the current row may cross at lanes >= lo only
    // (kb + lo is the first-hit step of the row below it, or where that row's path ended).
    int kb = 0, lo = 1;
    int* lp = list;     // end of the list ...
    int cnt = 0;        // ... and its length
    int last = -1;      // its last entry: the step of the latest hit
    unsigned bad = 0u;  // odd cells met so far
    unsigned m4;        // this lane's step: the rows of the group whose ray is below the terrain there (bit R = row R)
    bool finished = false;  // the walk reached the end of the caches: every row above sees only sky
    const size_t gstride = (size_t)n_t * PATH_ROWS;  // one row group of the path cache
    if (y_hi + 1 < S.height && !flagged) {
        // ---- entering the band: the row below it, four windows at a time ----
        const int yb = y_hi + 1, nlim = min(n_t, B.p_n[yb]);
        const double* pr = B.p_elev + path_index(n_t, 0, yb);
        unsigned entry = 0u;
        for (;;) {
            unsigned stop = 0u, oddf = 0u, ent[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int kw = kb + 31 * w, kk = min(kw + lane, k_last);
                const double dw = pr[(size_t)kk * PATH_ROWS] - te[kk];
                const unsigned ab = __ballot_sync(FULL, dw > 0.0), be = __ballot_sync(FULL, dw < 0.0), ze = __ballot_sync(FULL, dw == 0.0);
                const int sp = nlim - kw;
                const unsigned vis = (sp >= 32 ? FULL : sp <= 0 ? 0u : (1u << sp) - 1u) & ~1u;
                ent[w] = (ab << 1) & be & vis;
                const unsigned od = (((be << 1) & ab) | ze) & vis;
                stop |= ent[w] != 0u || kw + 32 >= nlim ? 1u << w : 0u;
                // odd cells the row visits: all of a window without a crossing, those before the crossing otherwise
                oddf |= (ent[w] ? od & ((1u << (__ffs(ent[w]) - 1)) - 1u) : od) != 0u ? 1u << w : 0u;
            }
            const int nskip = stop ? __ffs(stop) - 1 : 4;     // windows passed without an event
            bad |= oddf & (stop ? (2u << nskip) - 1u : 0xfu);  // ... and the one that holds the event, up to it
            kb += 31 * nskip;
            if (stop) {
                entry = nskip == 0 ? ent[0] : nskip == 1 ? ent[1] : nskip == 2 ? ent[2] : ent[3];
                break;
            }
        }
        if (entry) {  // the row below hits at step kb + hl: the band's rows go on from there
            const int hl = __ffs(entry) - 1, kh = kb + hl;
            st_if_s32(lp, kh - 1, lane0);
            st_if_s32(lp + 1, kh, lane0);
            lp += 2, cnt = 2, last = kh, lo = hl;
        } else {  // its path ends without a sign change (or it sees only sky: kb + lo == n_t)
            kb = ((nlim - 1) / 31) * 31, lo = nlim - kb;
        }
    }
    const int g_top = y_hi / SWEEP_ROWS;
    const double* pg = B.p_elev + (size_t)g_top * gstride;          // this group's rows, [k][row % 4]
    const int* pn = B.p_n + g_top * SWEEP_ROWS;
    int kk4 = min(kb + lane, k_last) * PATH_ROWS;                   // this lane's step of the window, as an offset into pg
    double t_cur = te[min(kb + lane, k_last)];
    double4 cur = *reinterpret_cast<const double4*>(pg + kk4);
    double4 nxt = cur;  // the window of the NEXT group at the same steps, loaded one group ahead (the walk's latency chain)
    int nxt_kb = -1;
#define ATMRT_BITS_M4()                                                                                              \
    m4 = (cur.x - t_cur < 0.0 ? 1u : 0u) | (cur.y - t_cur < 0.0 ? 2u : 0u) | (cur.z - t_cur < 0.0 ? 4u : 0u) |       \
         (cur.w - t_cur < 0.0 ? 8u : 0u);
    for (int g = g_top; g * SWEEP_ROWS >= y_lo && bad == 0u && !flagged && !finished; --g, pg -= gstride, pn -= SWEEP_ROWS) {
        const int ybase = g * SWEEP_ROWS;
        const int rmax = min(SWEEP_ROWS - 1, y_hi - ybase);
        const int4 len = *reinterpret_cast<const int4*>(pn);  // p_n is padded to h_pad entries
        if (g != g_top) cur = nxt_kb == kb ? nxt : *reinterpret_cast<const double4*>(pg + kk4);
        if (ybase > y_lo) nxt = *reinterpret_cast<const double4*>(pg - gstride + kk4), nxt_kb = kb;
        ATMRT_BITS_M4()
        // Rows that cross at the very step the row below them hit at need no search: where that row crossed from above,
        // every ray above it was above the terrain one step earlier too (the rays do not cross: k_path_check), so such
        // a row hits there iff it is below the terrain at that step -- one bit of m4 at the hit's lane.
        unsigned same = lo < 32 && kb + lo == last ? __shfl_sync(FULL, m4, lo & 31) : 0u;
        int h0 = 0, h1 = 0, h2 = 0, h3 = 0;
        // One row: walk the windows until the row crosses from above (its hit), its path ends (no hit; the row above
        // goes on from there: a path never outlives the one above it, k_path_check) or the caches end (sky).
#define ATMRT_BITS_ROW(C, R, LEN, HOUT)                                                                          \
    if (R <= rmax) {                                                                                             \
        if ((same >> R) & 1u) {                                                                                  \
            HOUT = last * hit_mul + cnt; /* the same step, the same slot */                                      \
        } else {                                                                                                 \
            const int nlim = min(n_t, LEN);                                                                      \
            int adv = 0; /* windows this row has moved on without a hit */                                       \
            same = 0u;                                                                                           \
            for (;;) {                                                                                           \
                if (kb + lo >= nlim) {                                                                           \
                    finished = finished || kb + lo >= n_t;                                                       \
                    break;                                                                                       \
                }                                                                                                \
                if (lo < 32) {                                                                                   \
                    const double d = cur.C - t_cur; /* ray - terrain at step kb + lane */                        \
                    const unsigned above = __ballot_sync(FULL, d > 0.0), below = __ballot_sync(FULL, d < 0.0);   \
                    const unsigned zero = __ballot_sync(FULL, d == 0.0);                                         \
                    const int span = nlim - kb; /* lanes below `span` are steps of this row's path */           \
                    const unsigned visit = (span >= 32 ? FULL : (1u << span) - 1u) & (FULL << lo);               \
                    const unsigned entry = (above << 1) & below & visit; /* d1 > 0 > d2: utils.rs:220-222, from above */ \
                    /* an exit from below (the ray started under the surface) or an exact zero among the cells the  \
                       row visits before its hit sends the column to the general march */                        \
                    const unsigned odd = (((below << 1) & above) | zero) & visit;                                \
                    if (entry) {                                                                                 \
                        const int hl = __ffs(entry) - 1, kh = kb + hl;                                           \
                        bad |= odd & ((1u << hl) - 1u);                                                          \
                        /* the distinct samples of the hits: kh - 1 and kh, appended unless they are the last entries */ \
                        const bool new1 = last != kh, new0 = new1 && last != kh - 1;                             \
                        st_if_s32(lp, kh - 1, lane0 && new0);                                                    \
                        lp += new0 ? 1 : 0, cnt += new0 ? 1 : 0;                                                 \
                        st_if_s32(lp, kh, lane0 && new1);                                                        \
                        lp += new1 ? 1 : 0, cnt += new1 ? 1 : 0;                                                 \
                        last = kh;                                                                               \
                        HOUT = kh * hit_mul + cnt; /* 1 + slot of kh (and kh) */                                 \
                        lo = hl;    /* the row above continues from the same step */                             \
                        same = __shfl_sync(FULL, m4, hl);                                                        \
                        break;                                                                                   \
                    }                                                                                            \
                    bad |= odd;                                                                                  \
                    if (kb + 32 >= nlim) { /* the row ends inside the window */                                  \
                        lo = nlim - kb;                                                                          \
                        break;                                                                                   \
                    }                                                                                            \
                    kb += 31, lo = 1;                                                                            \
                    /* Far field: a row crosses many windows without a hit. From its second empty window on, four    \
                       windows of THIS row at once -- four loads in flight instead of a chain of four -- and on past  \
                       them while none holds a crossing or the end of the row's path; the window that does is then  \
                       taken up as usual. */                                                                     \
                    if (adv++ > 0) for (;;) {                                                                    \
                        unsigned stop = 0u, oddf = 0u;                                                           \
                        _Pragma("unroll") for (int w = 0; w < 4; ++w) {                                          \
                            const int kw = kb + 31 * w, kk = min(kw + lane, k_last);                             \
                            const double dw = pg[(size_t)kk * PATH_ROWS + R] - te[kk];                           \
                            const unsigned ab = __ballot_sync(FULL, dw > 0.0), be = __ballot_sync(FULL, dw < 0.0); \
                            const unsigned ze = __ballot_sync(FULL, dw == 0.0);                                  \
                            const int sp = nlim - kw;                                                            \
                            const unsigned vis = (sp >= 32 ? FULL : sp <= 0 ? 0u : (1u << sp) - 1u) & ~1u;       \
                            const bool here = ((ab << 1) & be & vis) != 0u || kw + 32 >= nlim;                   \
                            stop |= here ? 1u << w : 0u;                                                         \
                            oddf |= ((((be << 1) & ab) | ze) & vis) != 0u ? 1u << w : 0u;                        \
                        }                                                                                        \
                        const int nskip = stop ? __ffs(stop) - 1 : 4; /* whole windows without an event */       \
                        bad |= oddf & ((1u << nskip) - 1u);                                                      \
                        kb += 31 * nskip;                                                                        \
                        if (stop) break;                                                                         \
                    }                                                                                            \
                } else { /* the row below ended on the last step of this window: move on (kb + lo is unchanged) */ \
                    kb += 31, lo -= 31;                                                                          \
                }                                                                                                \
                kk4 = min(kb + lane, k_last) * PATH_ROWS;                                                        \
                t_cur = te[min(kb + lane, k_last)];                                                              \
                cur = *reinterpret_cast<const double4*>(pg + kk4);                                               \
                ATMRT_BITS_M4()                                                                                  \
            }                                                                                                    \
        }                                                                                                        \
    }
        ATMRT_BITS_ROW(w, 3, len.w, h3)
        ATMRT_BITS_ROW(z, 2, len.z, h2)
        ATMRT_BITS_ROW(y, 1, len.y, h1)
        ATMRT_BITS_ROW(x, 0, len.x, h0)
#undef ATMRT_BITS_ROW
        st_if_v4s32(hit + ybase, h0, h1, h2, h3, lane0);  // (the plane is padded to h_pad rows)
        if (finished) {
            // The first row that sees only sky scanned to the end, and no row above it can cross any more (no path is
            // longer than n_t). All of them at once.
            for (int yy = ybase - 1 - lane; yy >= y_lo; yy -= 32) hit[yy] = 0;
        }
    }
#undef ATMRT_BITS_M4
    flagged = flagged || bad != 0u;
    if (lane0) {
        L.count[(size_t)xl * L.bands + band] = cnt;
        if (flagged) {
            B.sweep_col[xl] = 1;              // (zeroed before the launch; every band that flags writes the same 1)
            atomicOr(B.sweep_flags + 1, 1u);  // some column needs the brute-force march
        }
    }
}

// TerrainData::normal of the listed samples (sample_normal: find_normal at the sample's cached coordinates).
template <int W>
__global__ void __launch_bounds__(128, 8) k_hit_normals(const __grid_constant__ DevScene S, DevBuffers B, SweepLists L, int parts, int only_band) {
    if (B.sweep_flags[0] != 0) return;
    const int part = blockIdx.x % parts;
    const int seg = only_band >= 0 ? (blockIdx.x / parts) * L.bands + only_band : blockIdx.x / parts;  // seg = column * bands + band
    const int xl = seg / L.bands;
    if (B.sweep_col[xl] != 0) return;  // flagged columns belong to the brute-force march
    const int cnt = L.count[seg];
    const int* __restrict__ list = L.list + (size_t)seg * L.cap;
    double* __restrict__ out = L.normals + (size_t)seg * L.cap * 3;
    for (int s = part * 128 + threadIdx.x; s < cnt; s += parts * 128) {
        const int smp = list[s];
        const size_t ti = (size_t)xl * S.n_pad + smp;
        const V3 n = sample_normal<W>(S, B.terrain, B, xl, smp, B.t_lat[ti], B.t_lon[ti]);
        out[3 * s] = n.x, out[3 * s + 1] = n.y, out[3 * s + 2] = n.z;
    }
}

// get_single_pixel's hit (utils.rs:220-236) and draw_image (renderer/mod.rs:395-411) of the swept pixels. A block is a
// tile of 32 rows x TILE_COLS columns: each warp shades 32 adjacent rows of ONE column -- adjacent rows hit adjacent
// steps and adjacent list slots, so the reads of the hit plane, the list, the normals and both caches share sectors --
// and the results are staged in shared memory and written with the lanes running along x, so that the row-major image
// and metadata are stored as contiguous row segments.
// With terrain_alpha == 1 the compositing of ONE opaque trace point is  result = add([0,0,0], c, 1.0 * 1.0)  and then
// add(result, default, 0.0), and ((0/255 + c/255 * 1) * 255) as u8 == c, ((c/255 + d/255 * 0) * 255) as u8 == c for all
// 256 values of c (tests/test_oracle_known_answers.py: test_identity_requantisation_all_values): the pixel IS the
// (fogged) colour of its trace point, and a pixel without one is add([0,0,0], default, 1.0) == default.
constexpr int TILE_COLS = 8;

__global__ void __launch_bounds__(32 * TILE_COLS, 6) k_shade_tiles(const __grid_constant__ DevScene S, DevBuffers B, MarchOut O, SweepLists L, int row0,
                                                                   int count) {
    if (B.sweep_flags[0] != 0) return;
    // staging rows padded: a warp stages 32 ROWS of one column, so the row stride decides the banks -- 32 doubles put every lane
    // of a store on one bank (32 wavefronts per store); 34 doubles (16-byte aligned rows for the double2 read-out) spread them
    // over 8 bank pairs, 7 words of colour and 9 of steps over all 32 banks
    __shared__ __align__(16) double s_meta[32][TILE_COLS * 4 + 2];
    __shared__ __align__(16) unsigned char s_rgb[32][TILE_COLS * 3 + 4];
    __shared__ int s_steps[32][TILE_COLS + 1];
    __shared__ unsigned char s_skip[TILE_COLS];
    __shared__ unsigned long long s_cnt[TILE_COLS];
    __shared__ unsigned s_hits[TILE_COLS];
    const int wl = S.x1 - S.x0, n_t = S.n_t;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = blockIdx.y * TILE_COLS, y0 = row0 + blockIdx.x * 32;  // row0: a multiple of 32 (row bands)
    const int xl = c0 + w, y = y0 + lane;
    const bool col_ok = xl < wl && B.sweep_col[min(xl, wl - 1)] == 0;  // flagged columns belong to the brute-force march
    const bool active = col_ok && y < S.height;
    if (lane == 0) s_skip[w] = col_ok ? 0 : 1;
    const int xx = min(xl, wl - 1), yy = min(y, S.height - 1);
    const int nlim = min(n_t, B.p_n[yy]);
    const bool packed = sweep_hit_packed(S, L);
    const int hv = active ? B.sweep_hit[(size_t)xx * S.h_pad + yy] : 0;
    const int slot1 = packed ? hv & 0xffff : hv;
    const bool hit = slot1 > 0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    Rgb8 px{{S.shade.def_color[0], S.shade.def_color[1], S.shade.def_color[2]}};
    double m_lat = qnan, m_lon = qnan, m_elev = qnan, m_dist = qnan;
    int consumed = nlim > 0 ? nlim - 1 : 0;
    if (hit) {
        const int s = slot1 - 1;
        const size_t seg = ((size_t)xx * L.bands + sweep_band_of_row(L, yy)) * L.cap;  // the lists of this pixel's band
        const int kh = packed ? hv >> 16 : L.list[seg + s];
        const double* __restrict__ nr = L.normals + (seg + s - 1) * 3;  // slots s - 1, s: samples kh - 1, kh
        const V3 n0{nr[0], nr[1], nr[2]}, n1{nr[3], nr[4], nr[5]};
        const size_t ti = (size_t)xx * S.n_pad + kh;
        const size_t p1 = path_index(n_t, kh, yy), p0 = p1 - PATH_ROWS;
        const double lat0 = B.t_lat[ti - 1], lon0 = B.t_lon[ti - 1], elev0 = B.t_elev[ti - 1];
        const double lat1 = B.t_lat[ti], lon1 = B.t_lon[ti], elev1 = B.t_elev[ti];
        const double ray0 = B.p_elev[p0], ray1 = B.p_elev[p1];
        const double dist0 = B.path_x[kh - 1], dist1 = B.path_x[kh];  // path_x[0] = 0
        const double len0 = kh - 1 == 0 ? 0.0 : B.p_len[p0], len1 = B.p_len[p1];
        const double diff1 = ray0 - elev0, diff2 = ray1 - elev1;
        const double prop = diff1 / (diff1 - diff2);
        // TracingState::interpolate, utils.rs:108-125
        m_lat = lat0 + (lat1 - lat0) * prop, m_lon = lon0 + (lon1 - lon0) * prop;
        m_dist = dist0 + (dist1 - dist0) * prop, m_elev = elev0 + (elev1 - elev0) * prop;
        const double plen = len0 + (len1 - len0) * prop;
        const V3 normal = n0 + (n1 - n0) * prop;
        px = color_for_pixel(S.shade, true, m_elev, m_dist, normal, Color4{0.0, 0.0, 0.0, 1.0});
        if (S.shade.fog_enabled) px = apply_fog(S.shade.fog_distance, plen, px);
        consumed = kh;
    }
    s_rgb[lane][w * 3 + 0] = px.c[0], s_rgb[lane][w * 3 + 1] = px.c[1], s_rgb[lane][w * 3 + 2] = px.c[2];
    s_meta[lane][w * 4 + 0] = m_lat, s_meta[lane][w * 4 + 1] = m_lon, s_meta[lane][w * 4 + 2] = m_elev, s_meta[lane][w * 4 + 3] = m_dist;
    s_steps[lane][w] = consumed;
    {  // render counters: one atomic set per block
        unsigned long long steps = active ? (unsigned long long)consumed : 0ull;
        for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(FULL, steps, o);
        const unsigned hits = __popc(__ballot_sync(FULL, active && hit));
        if (lane == 0) s_cnt[w] = steps, s_hits[w] = hits;
    }
    const bool all_cols = __syncthreads_and(col_ok ? 1 : 0) != 0;  // also the barrier between staging and write-out
    const int rows = min(32, S.height - y0);
    // Write-out with the lanes along x: thread t handles pixel (row t / TILE_COLS, column t % TILE_COLS) of the tile -- its
    // 32 bytes of metadata as two 16-byte stores -- and, for the colour, word t of the tile's 32 x 6 four-byte words
    // when the row segments are word-aligned.
    {
        const int r = threadIdx.x / TILE_COLS, c = threadIdx.x % TILE_COLS;
        const bool live = r < rows && c0 + c < wl && !s_skip[c];
        if (O.meta && live) {
            double2* out = reinterpret_cast<double2*>(O.meta + (size_t)(y0 + r) * wl + c0 + c);
            out[0] = make_double2(s_meta[r][c * 4 + 0], s_meta[r][c * 4 + 1]);
            out[1] = make_double2(s_meta[r][c * 4 + 2], s_meta[r][c * 4 + 3]);
        }
        if (O.steps && live) O.steps[(size_t)(y0 + r) * wl + c0 + c] = s_steps[r][c];
    }
    if (O.rgb) {
        if (all_cols && (wl & 3) == 0) {  // every row segment of the tile is 24 aligned bytes
            constexpr int WORDS = TILE_COLS * 3 / 4;
            if ((int)threadIdx.x < rows * WORDS) {
                const int r = threadIdx.x / WORDS, q = threadIdx.x % WORDS;
                const unsigned v = *reinterpret_cast<const unsigned*>(&s_rgb[r][q * 4]);
                *reinterpret_cast<unsigned*>(O.rgb + ((size_t)(y0 + r) * wl + c0) * 3 + q * 4) = v;
            }
        } else {
            for (int e = threadIdx.x; e < rows * TILE_COLS * 3; e += 32 * TILE_COLS) {
                const int r = e / (TILE_COLS * 3), q = e % (TILE_COLS * 3), c = q / 3;
                if (c0 + c < wl && !s_skip[c]) O.rgb[((size_t)(y0 + r) * wl + c0) * 3 + q] = s_rgb[r][q];
            }
        }
    }
    // (`count` == 0: the lower band of a split frame is shaded before every check of the frame is in; its pixels are counted
    // afterwards, by k_count_swept)
    if (threadIdx.x == 0 && count) {
        unsigned long long st = 0ull, ht = 0ull;
        for (int i = 0; i < TILE_COLS; ++i) st += s_cnt[i], ht += s_hits[i];
        if (st) atomicAdd(B.counters + CNT_RAY_STEPS, st);
        if (ht) atomicAdd(B.counters + CNT_TRACE_POINTS, ht), atomicAdd(B.counters + CNT_PIXELS_HIT, ht);
    }
}

// The render counters of the swept pixels of rows [row0, row1): what k_shade_tiles adds when it counts itself -- the steps
// a pixel consumed (its first-hit step, or its ray's last) and the hits -- for the columns the sweep did not hand to the
// brute-force march, unless the whole image went to the general march. One thread per pixel, lanes along the rows.
__global__ void __launch_bounds__(256) k_count_swept(const __grid_constant__ DevScene S, DevBuffers B, SweepLists L, int row0, int row1) {
    if (B.sweep_flags[0] != 0) return;
    __shared__ unsigned long long s_steps[8];
    __shared__ unsigned s_hits[8];
    const int wl = S.x1 - S.x0, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool packed = sweep_hit_packed(S, L);
    const int xl = blockIdx.x * 8 + w;  // a warp per column, a block per eight columns: one set of atomics per block
    unsigned long long steps = 0ull;
    unsigned hits = 0u;
    if (xl < wl && B.sweep_col[xl] == 0) {
        const int* __restrict__ hit = B.sweep_hit + (size_t)xl * S.h_pad;
        for (int y = row0 + lane; y < row1; y += 32) {
            const int nlim = min(S.n_t, B.p_n[y]);
            const int hv = hit[y];
            const int slot1 = packed ? hv & 0xffff : hv;
            if (slot1 > 0) {
                steps += (unsigned long long)(packed ? hv >> 16 : L.list[((size_t)xl * L.bands + sweep_band_of_row(L, y)) * L.cap + slot1 - 1]);
                hits += 1u;
            } else {
                steps += (unsigned long long)(nlim > 0 ? nlim - 1 : 0);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(FULL, steps, o);
    hits = __reduce_add_sync(FULL, hits);
    if (lane == 0) s_steps[w] = steps, s_hits[w] = hits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long st = 0ull, ht = 0ull;
        for (int i = 0; i < 8; ++i) st += s_steps[i], ht += s_hits[i];
        if (st) atomicAdd(B.counters + CNT_RAY_STEPS, st);
        if (ht) atomicAdd(B.counters + CNT_TRACE_POINTS, ht), atomicAdd(B.counters + CNT_PIXELS_HIT, ht);
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 pipe micro-benchmark (the roofline denominator for this path: MEASURED_PEAKS.json has no
// FP64 figure). 8 independent DFMA (or DADD) chains per thread.
// ---------------------------------------------------------------------------------------------
template <bool FMA>
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double b, double c) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    for (int i = 0; i < iters; ++i) {
        if (FMA) {
            a0 = fma(a0, b, c), a1 = fma(a1, b, c), a2 = fma(a2, b, c), a3 = fma(a3, b, c);
            a4 = fma(a4, b, c), a5 = fma(a5, b, c), a6 = fma(a6, b, c), a7 = fma(a7, b, c);
        } else {
            a0 = a0 + c, a1 = a1 + c, a2 = a2 + c, a3 = a3 + c;
            a4 = a4 + c, a5 = a5 + c, a6 = a6 + c, a7 = a7 + c;
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;  // keep the chains alive
}

}  // namespace atmrt
