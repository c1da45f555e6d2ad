// atmrt_lib.cu -- context management and the C ABI of include/atmrt.h.
//
// One context per GPU. A render is three stages on two streams (terrain profile || ray paths, then
// the march); all buffers live in HBM and are reused between renders. There is no CPU fallback:
// without a CUDA device every entry point fails with ATMRT_ERR_NO_DEVICE / ATMRT_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <set>
#include <string>
#include <functional>
#include <thread>
#include <vector>

#include "kernels.cuh"

using namespace atmrt;

static_assert(sizeof(DevScene) < 8000, "DevScene must fit the kernel parameter space (32764 bytes on sm_70+ with CUDA >= 12.1)");

namespace {

std::string g_create_error;
std::mutex g_pageable_mutex;
std::set<void*> g_pageable;  // atmrt_host_alloc blocks that are ordinary (not page-locked) memory

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct atmrt_ctx {
    int device = 0;
    std::string err;
    cudaStream_t s_a = nullptr, s_b = nullptr, s_main = nullptr;
    cudaEvent_t ev_band[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_prep = nullptr, ev_a = nullptr, ev_b = nullptr;
    cudaEvent_t ev_last = nullptr;  // end of the most recent atmrt_render_device: the next one, on whatever stream, starts behind it
    bool last_recorded = false;
    // Stage timing: one set of events per render since the last harvest (atmrt_stage_times), so a
    // whole timed region of asynchronous renders can be averaged without synchronising inside it.
    struct StageEvents {
        cudaEvent_t a0, a1, b0, b1, c0, c1, t0, t1;
        cudaEvent_t k0[ATMRT_KERNEL_COUNT], k1[ATMRT_KERNEL_COUNT];  // around the hot kernels (atmrt_kernel_times)
        unsigned kmask;                                             // which of them this render launched
    };
    static constexpr size_t RING = 64;  // renders kept for the averages: the oldest is overwritten
    std::vector<StageEvents> ring;
    size_t ring_next = 0, ring_used = 0;
    cudaEvent_t t_0 = nullptr, t_1 = nullptr;  // fp64 micro-benchmark
    int num_sms = 148;

    // terrain
    std::vector<atmrt_tile_desc> tile_descs;
    std::vector<DevTile> tiles_host;
    void* terrain_owned = nullptr;
    DevTerrain terrain{};
    bool has_terrain = false;

    // scene
    atmrt_params params{};
    bool has_params = false;
    std::vector<atmrt_object> objects;
    std::vector<DevBuf> textures;
    std::vector<int> tex_w, tex_h;
    DevBuf d_objects_in, d_objects;

    // render state
    DevScene scene{};
    DevBuffers buf{};
    bool rendered = false;
    int march_mode = 0;
    std::vector<double> dist_k;
    std::vector<double> walk_sc;  // (sin, cos)(dist_k / radius), cached
    int walk_sc_n = -1;
    double walk_sc_step = 0.0, walk_sc_radius = 0.0;
    std::vector<double> path_x;  // [2][n_t]: x of path element k (the same for every row), and (x_k - x_{k-1}) / R
    int path_k_far = 0;
    std::vector<double> atm_cells;  // g(h) table of the ray-path stage, [ATM_FIELDS][ATM_CELLS]
    std::vector<DevGPiece> atm_pieces;  // its cells that hold the start of a temperature function
    std::vector<double> atm_bnd;           // sorted altitudes where g is not smooth: the starts of the temperature functions, the table's ends
    std::vector<unsigned char> atm_first;  // per table cell: index of the first of them at or above the cell's lower edge
    DevBuf d_atm_aux;
    int atm_cells_served = 0;
    bool atm_table_valid = false;
    atmrt_atmosphere_def atm_table_def{};
    double atm_table_wavelength = 0.0;
    int path_mode = 0;  // 0: g(h) from the table, 1: every evaluation through libm (validation)
    DevBuf d_atm_cells;
    DevBuf d_sweep_flags, d_sweep_col, d_sweep_hit, d_cross, d_cross_trig;
    DevBuf d_anchor;  // walk anchors of stage A, [wl][n_anchor]
    bool walk_anchors = true;
    DevBuf d_list, d_count, d_normals;  // stage C: the distinct hit samples of every column and their normals
    int sweep_bands = 0;                // 0: chosen per render (launch_render)
    DevBuf d_stage;  // raw posts of pack_terrain on their way to the tiled layout
    bool sweep_enabled = true;
    struct Uploaded {  // what upload_scene_inputs sent last, and where
        void* dst = nullptr;
        std::vector<char> bytes;
    };
    Uploaded up[6];
    // the terrain on its way (atmrt_group_render_tiles): uploaded, retiled and gathered on s_t / s_r; whoever reads the terrain
    // waits for ev_terrain
    cudaStream_t s_t = nullptr, s_r = nullptr;
    // a frame split below the horizon (launch_render): the lower rows' rays on s_b2, the lower band's sweep on s_c
    cudaStream_t s_b2 = nullptr, s_c = nullptr;
    cudaEvent_t ev_b1 = nullptr, ev_s1 = nullptr;
    cudaEvent_t ev_terrain = nullptr, ev_tile = nullptr;
    bool terrain_in_flight = false;
    // InterpolatingRectilinear generator: the grid's angle tables (device) while its Fast render is prepared, its trace lists,
    // the statistics of that render
    const double* grid_row_elev = nullptr;
    const double* grid_col_dir = nullptr;
    DevBuf d_interp, d_grid_angles, d_grid_points, d_grid_counts;
    atmrt_stats grid_stats{};
    int grid_overflows = 0;
    DevBuf d_dist, d_colcalc, d_tlat, d_tlon, d_telev, d_tclose;
    DevBuf d_pdist, d_pelev, d_plen, d_pn;
    DevBuf d_tmin1, d_tmax1, d_tmin2, d_tmax2, d_tmin3, d_tmax3, d_close1, d_close2, d_close3;
    DevBuf d_rmin1, d_rmax1, d_rmin2, d_rmax2, d_rmin3, d_rmax3;
    DevBuf d_obs, d_counters;
    DevBuf d_rgb, d_meta, d_steps, d_points, d_counts;
    DevBuf d_probe_a, d_probe_b, d_probe_c, d_probe_d;
    int launches = 0;
};

namespace {

constexpr int SHADE_BANDS = 8;  // row bands of the shading when the image goes straight to host memory

int fail(atmrt_ctx* ctx, int code, const std::string& msg) {
    if (ctx)
        ctx->err = msg;
    else
        g_create_error = msg;
    return code;
}

#define CUDA_TRY(ctx, call)                                                                                  \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, ATMRT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

int ensure(atmrt_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return 0;
    if (b.p) CUDA_TRY(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CUDA_TRY(ctx, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

// ---- terrain layout -------------------------------------------------------------------------
struct TerrainLayout {
    int lat_min = 0, lon_min = 0, nlat_tiles = 0, nlon_tiles = 0;
    size_t off_tiles = 0, off_lookup = 0, off_posts = 0, total = 0;
    std::vector<DevTile> tiles;
    std::vector<int> lookup;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int make_layout(atmrt_ctx* ctx, const atmrt_tile_desc* descs, int n, TerrainLayout* L) {
    if (n < 0 || (n > 0 && !descs)) return fail(ctx, ATMRT_ERR_INVALID, "bad tile list");
    int lat_lo = 0, lat_hi = -1, lon_lo = 0, lon_hi = -1;
    for (int i = 0; i < n; ++i) {
        const atmrt_tile_desc& d = descs[i];
        if (d.nlat < 2 || d.nlon < 2 || !(d.lat_interval > 0.0) || !(d.lon_interval > 0.0))
            return fail(ctx, ATMRT_ERR_INVALID, "tile with fewer than 2x2 posts or a non-positive interval");
        if (i == 0) {
            lat_lo = lat_hi = d.lat0;
            lon_lo = lon_hi = d.lon0;
        } else {
            lat_lo = std::min(lat_lo, d.lat0), lat_hi = std::max(lat_hi, d.lat0);
            lon_lo = std::min(lon_lo, d.lon0), lon_hi = std::max(lon_hi, d.lon0);
        }
    }
    L->lat_min = lat_lo, L->lon_min = lon_lo;
    L->nlat_tiles = n ? lat_hi - lat_lo + 1 : 0;
    L->nlon_tiles = n ? lon_hi - lon_lo + 1 : 0;
    if ((long long)L->nlat_tiles * L->nlon_tiles > (1 << 22)) return fail(ctx, ATMRT_ERR_INVALID, "tile index span too large");
    L->lookup.assign((size_t)L->nlat_tiles * L->nlon_tiles, -1);
    L->tiles.resize(n);
    long long off = 0;
    for (int i = 0; i < n; ++i) {
        const atmrt_tile_desc& d = descs[i];
        DevTile& t = L->tiles[i];
        t.min_lat = d.min_lat, t.min_lon = d.min_lon;
        t.lat_interval = d.lat_interval, t.lon_interval = d.lon_interval;
        t.inv_lat_interval = 1.0 / d.lat_interval, t.inv_lon_interval = 1.0 / d.lon_interval;
        t.max_lat = d.min_lat + (double)(d.nlat - 1) * d.lat_interval / 3600.0;
        t.max_lon = d.min_lon + (double)(d.nlon - 1) * d.lon_interval / 3600.0;
        t.nlat = d.nlat, t.nlon = d.nlon;
        t.mt_lat = (d.nlat + MT - 1) / MT;
        t._pad = 0;
        t.post_offset = off;
        off += (long long)((d.nlon + MT - 1) / MT) * t.mt_lat * (MT * MT);
        // HashMap::insert: a later tile with the same key replaces the earlier one (terrain/mod.rs:93-97)
        L->lookup[(size_t)(d.lat0 - lat_lo) * L->nlon_tiles + (d.lon0 - lon_lo)] = i;
    }
    L->off_tiles = 0;
    L->off_lookup = align_up(sizeof(DevTile) * (size_t)std::max(n, 1), 256);
    L->off_posts = align_up(L->off_lookup + sizeof(int) * std::max<size_t>(L->lookup.size(), 1), 256);
    L->total = align_up(L->off_posts + sizeof(int16_t) * (size_t)std::max<long long>(off, 1), 256);
    return 0;
}

// ---- atmosphere lowering (Atmosphere::from_def; host libm, same as the reference's CPU lowering) --
constexpr double ATM_G = 9.80665, ATM_M = 0.0289644, ATM_R = 8.31432;  // R* of US Standard Atmosphere 1976

double host_cubic_temperature(const DevAtmLayer& l, double h) {
    const double x = h - l.x0;
    return std::fma(std::fma(std::fma(l.c3, x, l.c2), x, l.c1), x, l.c0);
}
double host_layer_temperature(const DevAtmLayer& l, double h) {
    if (l.cubic) return host_cubic_temperature(l, h);
    return l.t_ref + l.gradient * (h - l.h_ref);
}
// device_atm.cuh: cubic_inverse_integral
double host_cubic_inverse_integral(const DevAtmLayer& l, double h) {
    const double gx[4] = ATMRT_GL8_X, gw[4] = ATMRT_GL8_W;
    const double half = 0.5 * (h - l.h_ref), mid = 0.5 * (h + l.h_ref);
    double acc = 0.0;
    for (int i = 3; i >= 0; --i) acc += gw[i] / host_cubic_temperature(l, mid - half * gx[i]);
    for (int i = 0; i < 4; ++i) acc += gw[i] / host_cubic_temperature(l, mid + half * gx[i]);
    return acc * half;
}
double host_layer_pressure(const DevAtmLayer& l, double h) {
    if (l.cubic) return l.p_ref * std::exp(-ATM_G * ATM_M / ATM_R * host_cubic_inverse_integral(l, h));
    if (l.gradient != 0.0) {
        double t = host_layer_temperature(l, h);
        return l.p_ref * std::pow(t / l.t_ref, -ATM_G * ATM_M / (ATM_R * l.gradient));
    }
    return l.p_ref * std::exp(-ATM_G * ATM_M * (h - l.h_ref) / (ATM_R * l.t_ref));
}

// Second derivatives M[i] of the cubic spline through (x[i], y[i]) under one of the reference's three boundary
// conditions (README.md:303-309), by the Thomas algorithm on the usual tridiagonal system
//   h[i-1] M[i-1] + 2 (h[i-1] + h[i]) M[i] + h[i] M[i+1] = 6 ((y[i+1] - y[i]) / h[i] - (y[i] - y[i-1]) / h[i-1]).
void spline_second_derivatives(const std::vector<double>& x, const std::vector<double>& y, int boundary, double a, double b, std::vector<double>& M) {
    const int m = (int)x.size();
    std::vector<double> sub(m, 0.0), diag(m, 1.0), sup(m, 0.0), rhs(m, 0.0);
    for (int i = 1; i + 1 < m; ++i) {
        const double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
        sub[i] = h0, diag[i] = 2.0 * (h0 + h1), sup[i] = h1;
        rhs[i] = 6.0 * ((y[i + 1] - y[i]) / h1 - (y[i] - y[i - 1]) / h0);
    }
    if (boundary == ATMRT_SPLINE_DERIVATIVES) {
        const double h0 = x[1] - x[0], hl = x[m - 1] - x[m - 2];
        diag[0] = 2.0 * h0, sup[0] = h0, rhs[0] = 6.0 * ((y[1] - y[0]) / h0 - a);
        sub[m - 1] = hl, diag[m - 1] = 2.0 * hl, rhs[m - 1] = 6.0 * (b - (y[m - 1] - y[m - 2]) / hl);
    } else if (boundary == ATMRT_SPLINE_SECOND_DERIVATIVES) {
        rhs[0] = a, rhs[m - 1] = b;
    }
    for (int i = 1; i < m; ++i) {
        const double w = sub[i] / diag[i - 1];
        diag[i] -= w * sup[i - 1];
        rhs[i] -= w * rhs[i - 1];
    }
    M.assign(m, 0.0);
    M[m - 1] = rhs[m - 1] / diag[m - 1];
    for (int i = m - 2; i >= 0; --i) M[i] = (rhs[i] - sup[i] * M[i + 1]) / diag[i];
}

// ---- the g(h) table of the ray-path stage (device_paths.cuh) -----------------------------------
// The reference's atmosphere function as a function of real numbers, evaluated in x87 extended
// precision (64-bit significand): the same layer lookup, hydrostatic law (layer_pressure), Ciddor terms
// (air_index) and central difference (env_dn, eps = 0.01 m) as oracle/atmrt_oracle.cpp / device_atm.cuh,
// with n - 1 kept apart from the 1 so that the difference does not cancel against it.
struct LdAtmosphere {
    const DevAtmosphere* a;
    long double r_axs, r_vs, m_a, rho_axs;  // wavelength terms, re-lowered in extended precision
    int forced_layer;
};

long double ld_n_minus_1(const LdAtmosphere& A, long double h) {
    const DevAtmosphere& a = *A.a;
    int idx = 0;
    for (int i = 1; i < a.n; ++i)
        if (h >= (long double)a.layer[i].start) idx = i;
    if (A.forced_layer >= 0) idx = A.forced_layer;  // that temperature function's law, continued past its boundaries
    const DevAtmLayer& l = a.layer[idx];
    auto cubic_t = [&](long double hh) {
        const long double x = hh - (long double)l.x0;
        return (long double)l.c0 + x * ((long double)l.c1 + x * ((long double)l.c2 + x * (long double)l.c3));
    };
    const long double t = l.cubic ? cubic_t(h) : (long double)l.t_ref + (long double)l.gradient * (h - (long double)l.h_ref);
    long double p;
    if (l.cubic) {  // the same 8-point rule as cubic_inverse_integral, as a function of real numbers
        static const long double gx[4] = {0.1834346424956498049394761L, 0.5255324099163289858177390L, 0.7966664774136267395915539L, 0.9602898564975362316835609L};
        static const long double gw[4] = {0.3626837833783619829651504L, 0.3137066458778872873379622L, 0.2223810344533744705443560L, 0.1012285362903762591525314L};
        const long double half = 0.5L * (h - (long double)l.h_ref), mid = 0.5L * (h + (long double)l.h_ref);
        long double acc = 0.0L;
        for (int i = 0; i < 4; ++i) acc += gw[i] / cubic_t(mid - half * gx[i]) + gw[i] / cubic_t(mid + half * gx[i]);
        p = (long double)l.p_ref * expl(-(long double)ATM_G * ATM_M / (long double)ATM_R * acc * half);
    } else if (l.gradient != 0.0)
        p = (long double)l.p_ref * powl(t / (long double)l.t_ref, -(long double)ATM_G * ATM_M / ((long double)ATM_R * l.gradient));
    else
        p = (long double)l.p_ref * expl(-(long double)ATM_G * ATM_M * (h - (long double)l.h_ref) / ((long double)ATM_R * l.t_ref));
    const long double a0 = 1.58123e-6L, a1 = -2.9331e-8L, a2 = 1.1043e-10L, b0 = 5.707e-6L, b1 = -2.051e-8L, c0 = 1.9898e-4L, c1 = -2.376e-6L;
    const long double d = 1.83e-11L, e = -0.765e-8L, rho_vs = 0.00985938L, gas_r = 8.314510L, m_v = 0.018015L;
    const long double alpha = 1.00062L, beta = 3.14e-8L, gamma = 5.6e-7L;
    const long double sa = 1.2378847e-5L, sb = -1.9121316e-2L, sc = 33.93711047L, sd = -6.3431645e3L;
    const long double t_c = t - 273.15L;
    long double x_v = 0.0L;
    if (a.humidity != 0.0) {
        const long double svp = expl(sa * t * t + sb * t + sc + sd / t);
        const long double f = alpha + beta * p + gamma * t_c * t_c;
        x_v = (long double)a.humidity * f * svp / p;
    }
    const long double pt = p / t;
    const long double z_m = 1.0L - pt * (a0 + a1 * t_c + a2 * t_c * t_c + (b0 + b1 * t_c) * x_v + (c0 + c1 * t_c) * x_v * x_v) + pt * pt * (d + e * x_v * x_v);
    const long double rho_v = x_v * p * m_v / (z_m * gas_r * t);
    const long double rho_a = (1.0L - x_v) * p * A.m_a / (z_m * gas_r * t);
    return (rho_a / A.rho_axs) * A.r_axs + (rho_v / rho_vs) * A.r_vs;
}

// g = n'/n. The derivative is taken with the five-point stencil at 2 m spacing instead of the reference's
// two-point one at 0.01 m: both approximate the same n'(h) (truncation (2 m / H)^4 / 30 < 1e-13 and
// (0.01 m / H)^2 / 6 < 1e-12 relative, H > 1.7 km the density scale height), but the wide stencil does not
// amplify the 1e-18 rounding of powl by H / 0.01 m.
long double ld_g(const LdAtmosphere& A, long double h) {
    const long double s = 2.0L;
    const long double d1 = ld_n_minus_1(A, h + s) - ld_n_minus_1(A, h - s), d2 = ld_n_minus_1(A, h + 2.0L * s) - ld_n_minus_1(A, h - 2.0L * s);
    return (8.0L * d1 - d2) / (12.0L * s) / (1.0L + ld_n_minus_1(A, h));
}

// Degree-6 interpolant of g at the 7 Chebyshev nodes of [centre - half, centre + half], as monomial
// coefficients in u = (h - centre) / half; kept only if it reproduces g at 33 check points to 1e-10
// relative or 1e-19 /m absolute (g is 3.5e-8 /m at sea level; either error bends a 400 km ray by less
// than 1e-6 m -- the reference's own evaluation of g carries 3e-7 relative noise).
bool fit_g_cell(const LdAtmosphere& A, double centre, double half_width, double* cd) {
    constexpr int M = ATM_FIELDS;  // nodes = coefficients
    static_assert(M == 7, "the Chebyshev-to-monomial table below is for degree 6");
    static const long double cheb[7][7] = {{1, 0, 0, 0, 0, 0, 0},  {0, 1, 0, 0, 0, 0, 0},    {-1, 0, 2, 0, 0, 0, 0},    {0, -3, 0, 4, 0, 0, 0},
                                           {1, 0, -8, 0, 8, 0, 0}, {0, 5, 0, -20, 0, 16, 0}, {-1, 0, 18, 0, -48, 0, 32}};  // T_k(u) as monomials
    const long double pi = 3.14159265358979323846264338327950288L, half = half_width;
    long double f[M], c[M];
    for (int k = 0; k < M; ++k) {
        f[k] = ld_g(A, (long double)centre + half * cosl(pi * (k + 0.5L) / M));
        if (!std::isfinite((double)f[k])) return false;
    }
    for (int m = 0; m < M; ++m) {
        long double acc = 0;
        for (int k = 0; k < M; ++k) acc += f[k] * cosl(pi * m * (k + 0.5L) / M);
        c[m] = acc * (m == 0 ? 1.0L : 2.0L) / M;
    }
    for (int q = 0; q < M; ++q) {
        long double mono = 0;
        for (int m = 0; m < M; ++m) mono += c[m] * cheb[m][q];
        cd[q] = (double)mono;
    }
    for (int k = 0; k <= 32; ++k) {  // the device's evaluation order, in f64
        const double u = -1.0 + k / 16.0;
        const double u2 = u * u, u4 = u2 * u2;
        const double v = std::fma(std::fma(cd[6], u2, std::fma(cd[5], u, cd[4])), u4, std::fma(std::fma(cd[3], u, cd[2]), u2, std::fma(cd[1], u, cd[0])));
        const long double want = ld_g(A, (long double)centre + half * u);
        const long double err = fabsl((long double)v - want);
        if (!(std::isfinite(v) && (err <= 1e-10L * fabsl(want) || err <= 1e-19L))) {
            if (getenv("ATMRT_DEBUG_TABLE"))
                fprintf(stderr, "fit_g_cell: centre %.1f +- %.1f function %d u %.4f got %.17g want %.17Lg rel %.3Lg\n", centre, half_width, A.forced_layer, u, v, want,
                        err / fabsl(want));
            return false;
        }
    }
    return true;
}

// The table, [ATM_FIELDS][ATM_CELLS] (cells that cannot be served hold NaN), and the pieces that serve the
// cells holding the start of a temperature function.
void build_g_table(const DevAtmosphere& a, double wavelength, std::vector<double>& cells, std::vector<DevGPiece>& pieces, int* served_out) {
    const double nan = std::numeric_limits<double>::quiet_NaN(), inf = std::numeric_limits<double>::infinity();
    cells.assign((size_t)ATM_FIELDS * ATM_CELLS, nan);
    pieces.clear();
    LdAtmosphere A{&a, 0, 0, 0, 0, -1};
    {  // air_index's wavelength terms (oracle: air_index; lower_atmosphere does the same in f64)
        const long double lambda_um = (long double)wavelength * 1.0e6L, s = 1.0L / (lambda_um * lambda_um);
        const long double r_as = 1.0e-8L * (5792105.0L / (238.0185L - s) + 167917.0L / (57.362L - s));
        A.r_vs = 1.022e-8L * (295.235L + 2.6422L * s + -0.032380L * s * s + 0.004028L * s * s * s);
        A.m_a = 0.0289635L + 1.2011e-8L * (450.0L - 400.0L);
        A.r_axs = r_as * (1.0L + 5.34e-7L * (450.0L - 450.0L));
        A.rho_axs = 101325.0L * A.m_a / (0.9995922115L * 8.314510L * 288.15L);
    }
    int served = 0;
    // [lo, hi) inside one temperature function: one piece if the fit holds, else halves down to 1 m (what
    // stays unserved -- e.g. the last metres before a temperature function reaches 0 K -- is left to libm)
    std::function<void(double, double, int)> add_pieces = [&](double lo, double hi, int depth) {
        if (pieces.size() >= (size_t)ATM_MAX_PIECES) return;
        if (!std::isfinite((double)ld_g(A, lo)) && !std::isfinite((double)ld_g(A, 0.5 * (lo + hi))) && !std::isfinite((double)ld_g(A, hi))) return;  // no law here (T <= 0)
        DevGPiece pc{};
        pc.h_lo = lo, pc.h_hi = hi, pc.centre = 0.5 * (lo + hi);
        const double half_width = 0.5 * (hi - lo);
        pc.inv_half = 1.0 / half_width;
        if (fit_g_cell(A, pc.centre, half_width, pc.c)) {
            pieces.push_back(pc);
        } else if (depth < 8 && hi - lo > 1.0) {
            add_pieces(lo, pc.centre, depth + 1);
            add_pieces(pc.centre, hi, depth + 1);
        }
    };
    for (int j = 1; j + 1 < ATM_CELLS; ++j) {  // the first and the last cell catch out-of-range and NaN altitudes
        const double hj = ATM_BASE + (double)j * ATM_CELL, half = 0.5 * ATM_CELL, lo = hj - half, hi = hj + half;
        // temperature functions that own a part of the cell's interior (a function that starts on an edge,
        // give or take the rounding of the cell index, does not split the cell)
        int first = 0, last = 0;
        for (int i = 1; i < a.n; ++i) {
            if (lo + 1e-6 >= a.layer[i].start) first = i;
            if (hi - 1e-6 > a.layer[i].start) last = i;
        }
        double cd[ATM_FIELDS];
        A.forced_layer = first;  // its law, continued past its ends for the stencil of ld_g
        if (first == last && fit_g_cell(A, hj, half, cd)) {
            for (int q = 0; q < ATM_FIELDS; ++q) cells[(size_t)q * ATM_CELLS + j] = cd[q];
            ++served;
            continue;
        }
        // Not served as a whole: pieces, one or more per temperature function over the part of the cell that
        // function owns (the cell index rounds: overlap the neighbours by 1 m). The reference's difference
        // quotient blends two laws within 0.01 m of a start; here g switches at the start itself.
        for (int i = first; i <= last; ++i) {
            const double p_lo = std::max(i == first ? -inf : a.layer[i].start, lo - 1.0);
            const double p_hi = std::min(i == last ? inf : a.layer[i + 1].start, hi + 1.0);
            if (!(p_hi > p_lo)) continue;
            A.forced_layer = i;  // that function's law, continued past its ends for the stencil
            add_pieces(p_lo, p_hi, 0);
        }
        A.forced_layer = -1;
    }
    if (served_out) *served_out = served;
}

int lower_atmosphere(atmrt_ctx* ctx, const atmrt_atmosphere_def& def, double wavelength, DevAtmosphere* out) {
    const int nf = def.n_functions;
    const double inf = std::numeric_limits<double>::infinity();
    if (nf < 1 || nf > ATMRT_MAX_ATM_FUNCTIONS) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: n_functions out of range");
    DevAtmosphere a{};
    a.humidity = def.humidity;
    // One layer per Linear function, one per segment of a Spline function that reaches into the function's range
    // [its start, the next function's start); the end segments continue the spline beyond its points.
    bool any_spline = false;
    int n = 0;
    for (int f = 0; f < nf; ++f) {
        if (f >= 2 && !(def.fn_start_altitude[f] > def.fn_start_altitude[f - 1]))
            return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: function altitudes must increase");
        const double fs = f == 0 ? -inf : def.fn_start_altitude[f], fe = f + 1 < nf ? def.fn_start_altitude[f + 1] : inf;
        if (def.fn_kind[f] == ATMRT_FUNCTION_LINEAR) {
            if (n >= ATM_MAX_LAYERS) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: more than 32 functions and spline segments");
            a.layer[n].start = fs;
            a.layer[n].gradient = def.fn_gradient[f];
            ++n;
            continue;
        }
        if (def.fn_kind[f] != ATMRT_FUNCTION_SPLINE) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: unknown temperature function kind");
        any_spline = true;
        const int first = def.fn_first_point[f], m = def.fn_n_points[f];
        if (m < 2 || first < 0 || first + m > def.n_spline_points || def.n_spline_points > ATMRT_MAX_SPLINE_POINTS)
            return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: a Spline needs at least two points inside spline_points");
        if (def.fn_boundary[f] < ATMRT_SPLINE_NATURAL || def.fn_boundary[f] > ATMRT_SPLINE_SECOND_DERIVATIVES)
            return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: unknown Spline boundary condition");
        std::vector<double> x(m), y(m), M;
        for (int i = 0; i < m; ++i) {
            x[i] = def.spline_points[first + i][0], y[i] = def.spline_points[first + i][1];
            if (i > 0 && !(x[i] > x[i - 1])) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: Spline point altitudes must increase");
        }
        spline_second_derivatives(x, y, def.fn_boundary[f], def.fn_boundary_values[f][0], def.fn_boundary_values[f][1], M);
        for (int i = 0; i + 1 < m; ++i) {
            const double lo = i == 0 ? -inf : x[i], hi = i + 2 == m ? inf : x[i + 1];
            if (!(hi > fs) || !(lo < fe)) continue;  // the segment lies outside the function's range
            if (n >= ATM_MAX_LAYERS) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere: more than 32 functions and spline segments");
            const double h = x[i + 1] - x[i];
            DevAtmLayer& l = a.layer[n++];
            l.start = std::max(fs, lo);
            l.cubic = 1;
            l.x0 = x[i];
            l.c0 = y[i];
            l.c1 = (y[i + 1] - y[i]) / h - h * (2.0 * M[i] + M[i + 1]) / 6.0;
            l.c2 = 0.5 * M[i];
            l.c3 = (M[i + 1] - M[i]) / (6.0 * h);
        }
    }
    a.n = n;
    auto find = [&](double h) {
        int idx = 0;
        for (int i = 1; i < n; ++i)
            if (h >= a.layer[i].start) idx = i;
        return idx;
    };
    // Temperatures of the Linear layers: anchored at (ha[i], ta[i]) by continuity with a neighbour whose temperature
    // is known -- the layer holding the temperature fixed point when every function is Linear, the Spline segments
    // otherwise. Above a known layer the anchor is the lower boundary, below it the upper one.
    std::vector<double> ha(n, 0.0), ta(n, 0.0);
    std::vector<char> known(n, 0);
    auto temp_in = [&](int i, double h) { return a.layer[i].cubic ? host_cubic_temperature(a.layer[i], h) : ta[i] + a.layer[i].gradient * (h - ha[i]); };
    if (any_spline) {
        for (int i = 0; i < n; ++i) known[i] = (char)a.layer[i].cubic;
    } else {
        const int jt = find(def.temperature_altitude);
        ha[jt] = def.temperature_altitude, ta[jt] = def.temperature, known[jt] = 1;
    }
    for (int i = 0; i + 1 < n; ++i)  // upwards
        if (known[i] && !known[i + 1]) ha[i + 1] = a.layer[i + 1].start, ta[i + 1] = temp_in(i, a.layer[i + 1].start), known[i + 1] = 1;
    for (int i = n - 1; i >= 1; --i)  // downwards
        if (known[i] && !known[i - 1]) ha[i - 1] = a.layer[i].start, ta[i - 1] = temp_in(i, a.layer[i].start), known[i - 1] = 1;
    // reference point per layer: the pressure fixed point in its layer, the lower boundary above it,
    // the upper boundary below it; pressures propagate hydrostatically from the fixed point.
    const int jp = find(def.pressure_altitude);
    a.layer[jp].h_ref = def.pressure_altitude;
    a.layer[jp].t_ref = temp_in(jp, def.pressure_altitude);
    a.layer[jp].p_ref = def.pressure;
    for (int i = jp + 1; i < n; ++i) {
        a.layer[i].h_ref = a.layer[i].start;
        a.layer[i].t_ref = temp_in(i, a.layer[i].start);
        a.layer[i].p_ref = host_layer_pressure(a.layer[i - 1], a.layer[i].start);
    }
    for (int i = jp - 1; i >= 0; --i) {
        a.layer[i].h_ref = a.layer[i + 1].start;
        a.layer[i].t_ref = temp_in(i, a.layer[i + 1].start);
        a.layer[i].p_ref = host_layer_pressure(a.layer[i + 1], a.layer[i + 1].start);
    }
    for (int i = 0; i < n; ++i) {
        DevAtmLayer& l = a.layer[i];
        l.gm = -ATM_G * ATM_M;
        l.rt = ATM_R * l.t_ref;
        l.expo = l.cubic ? -ATM_G * ATM_M / ATM_R : l.gradient != 0.0 ? -ATM_G * ATM_M / (ATM_R * l.gradient) : 0.0;
    }
    // Ciddor (1996) wavelength-only terms; 450 ppm CO2.
    const double w0 = 295.235, w1 = 2.6422, w2 = -0.032380, w3 = 0.004028;
    const double k0 = 238.0185, k1 = 5792105.0, k2 = 57.362, k3 = 167917.0;
    const double p_r1 = 101325.0, t_r1 = 288.15, z_a = 0.9995922115, gas_r = 8.314510, x_c = 450.0;
    double lambda_um = wavelength * 1.0e6;
    double s = 1.0 / (lambda_um * lambda_um);
    double r_as = 1.0e-8 * (k1 / (k0 - s) + k3 / (k2 - s));
    a.r_vs = 1.022e-8 * (w0 + w1 * s + w2 * s * s + w3 * s * s * s);
    a.m_a = 0.0289635 + 1.2011e-8 * (x_c - 400.0);
    a.r_axs = r_as * (1.0 + 5.34e-7 * (x_c - 450.0));
    a.rho_axs = p_r1 * a.m_a / (z_a * gas_r * t_r1);
    a.k_dry = a.m_a * a.r_axs / (gas_r * a.rho_axs);  // n - 1 = (p/T) k_dry / Z for dry air
    *out = a;
    return 0;
}

int validate_params(atmrt_ctx* ctx, const atmrt_params& p) {
    if (p.width <= 0 || p.height <= 0 || p.width > 32767 || p.height > 32767)
        return fail(ctx, ATMRT_ERR_INVALID, "width/height must be in 1..32767 (i16 pixel centring, fast.rs:116,122)");
    if (p.x0 < 0 || p.x1 > p.width || p.x0 >= p.x1) return fail(ctx, ATMRT_ERR_INVALID, "column block [x0,x1) out of range");
    if (!(p.simulation_step > 0.0)) return fail(ctx, ATMRT_ERR_INVALID, "simulation_step must be positive");
    if (!(p.max_distance > 0.0)) return fail(ctx, ATMRT_ERR_INVALID, "max_distance must be positive");
    if (p.max_distance / p.simulation_step > 4.0e6) return fail(ctx, ATMRT_ERR_INVALID, "more than 4e6 samples per ray");
    if (p.earth_model < ATMRT_EARTH_SPHERICAL || p.earth_model > ATMRT_EARTH_OBSERVER_AE) return fail(ctx, ATMRT_ERR_INVALID, "unknown earth model");
    if ((p.earth_model == ATMRT_EARTH_SPHERICAL || p.earth_model == ATMRT_EARTH_ELLIPSOID || p.earth_model == ATMRT_EARTH_OBSERVER_AE) && !(p.radius > 0.0))
        return fail(ctx, ATMRT_ERR_INVALID, "radius (Spherical radius, Ellipsoid a, ObserverAe proj_radius) must be positive");
    if (p.earth_model == ATMRT_EARTH_ELLIPSOID && !(p.ellipsoid_b > 0.0)) return fail(ctx, ATMRT_ERR_INVALID, "ellipsoid_b must be positive");
    if (p.coloring != ATMRT_COLORING_SIMPLE && p.coloring != ATMRT_COLORING_SHADING) return fail(ctx, ATMRT_ERR_INVALID, "unknown coloring");
    if (p.generator != ATMRT_GENERATOR_FAST && p.generator != ATMRT_GENERATOR_RECTILINEAR && p.generator != ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR)
        return fail(ctx, ATMRT_ERR_INVALID, "unknown generator");
    return 0;
}

// Build DevScene + size all render buffers.
int prepare_render(atmrt_ctx* ctx) {
    if (!ctx->has_terrain) return fail(ctx, ATMRT_ERR_STATE, "render before set_terrain/bind_terrain");
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "render before set_params");
    const atmrt_params& p = ctx->params;
    DevScene& S = ctx->scene;
    S = DevScene{};
    S.lat0 = p.latitude, S.lon0 = p.longitude;
    S.direction = p.direction, S.tilt = p.tilt, S.fov = p.fov, S.max_distance = p.max_distance;
    S.step = p.simulation_step;
    S.altitude = p.altitude;
    S.earth_model = p.earth_model;
    {
        DevEarth& E = S.earth;
        E.model = p.earth_model;
        E.flat_dirs = p.earth_model == ATMRT_EARTH_FLAT_DISTORTED || p.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT || p.earth_model == ATMRT_EARTH_OBSERVER_AE;
        E.walker = p.earth_model == ATMRT_EARTH_FLAT_DISTORTED ? WALK_FLDS
                   : p.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT ? WALK_AZEQ
                   : p.earth_model == ATMRT_EARTH_ELLIPSOID ? WALK_ELLIPSOID : WALK_SPHERICAL;
        E.radius = p.radius;
        E.b = p.ellipsoid_b;
        if (p.earth_model == ATMRT_EARTH_ELLIPSOID) {
            E.f = (p.radius - p.ellipsoid_b) / p.radius;
            E.e2 = 1.0 - (p.ellipsoid_b * p.ellipsoid_b) / (p.radius * p.radius);
        }
        // EarthModel::to_shape, earth_model/mod.rs:95-112: the ray physics sees a plane or a sphere
        S.flat = E.flat_dirs;
        S.radius = p.earth_model == ATMRT_EARTH_ELLIPSOID ? (2.0 * p.radius + p.ellipsoid_b) / 3.0 : p.radius;
    }
    S.straight = p.straight_rays != 0;
    S.width = p.width, S.height = p.height, S.x0 = p.x0, S.x1 = p.x1;
    if (S.earth.walker == WALK_SPHERICAL) {
        S.sin_diff = std::sin(NORMAL_DIFF / p.radius);
        S.cos_diff = std::cos(NORMAL_DIFF / p.radius);
        const double delta = NORMAL_DIFF / p.radius;
        S.tan_diff = std::tan(delta);
        S.versin_diff = 2.0 * std::sin(0.5 * delta) * std::sin(0.5 * delta);  // 1 - cos(delta) without cancellation
        S.diff_deg = to_degrees(delta);
    }
    int rc = lower_atmosphere(ctx, p.atmosphere, p.wavelength, &S.atm);
    if (rc) return rc;
    DevShade& sh = S.shade;
    sh.coloring = p.coloring, sh.palette = p.palette, sh.fog_enabled = p.fog_enabled;
    sh.water_level = p.water_level, sh.ambient_light = p.ambient_light;
    sh.light[0] = p.light_dir[0], sh.light[1] = p.light_dir[1], sh.light[2] = p.light_dir[2];
    sh.simple_max_distance = p.simple_max_distance;
    sh.fog_distance = p.fog_distance;
    sh.terrain_alpha = p.terrain_alpha;
    auto q = [](double v) -> unsigned char { return v != v || v <= 0.0 ? 0 : (v >= 255.0 ? 255 : (unsigned char)v); };
    if (p.fog_enabled) {  // ColoringMethod::fog_color
        sh.def_color[0] = sh.def_color[1] = sh.def_color[2] = 160;
    } else if (p.coloring == ATMRT_COLORING_SIMPLE) {  // simple.rs:46-48
        sh.def_color[0] = sh.def_color[1] = sh.def_color[2] = 28;
    } else {  // shading.rs:134-142
        const double legacy[3] = {0.11, 0.11, 0.11}, improved[3] = {0.23, 0.41, 0.55};
        const double* c = p.palette == ATMRT_PALETTE_LEGACY ? legacy : improved;
        for (int i = 0; i < 3; ++i) sh.def_color[i] = q(c[i] * 255.0);
    }
    // gen_terrain_cache's running sum: distance = 0; while distance < max { ...; distance += step }
    ctx->dist_k.clear();
    for (double d = 0.0; d < p.max_distance; d += p.simulation_step) ctx->dist_k.push_back(d);
    const int n_t = (int)ctx->dist_k.size();
    S.n_t = n_t;
    // SphericalCalc::coords_at_dist's sin / cos of dist / radius (directional_calc.rs:72-76) do not depend on the
    // column: tabulated behind dist_k (at an even offset: they are read as double2), once per (n_t, step, radius)
    const size_t sc_off = ((size_t)n_t + 1) & ~(size_t)1;
    S.n_sc_off = (int)sc_off;
    ctx->dist_k.resize(sc_off + (size_t)2 * n_t, 0.0);
    if (p.earth_model == ATMRT_EARTH_SPHERICAL || p.earth_model == ATMRT_EARTH_OBSERVER_AE) {
        if (ctx->walk_sc_n != n_t || ctx->walk_sc_step != p.simulation_step || ctx->walk_sc_radius != p.radius) {
            ctx->walk_sc.resize((size_t)2 * n_t);
            for (int k = 0; k < n_t; ++k) {
                const double ang = ctx->dist_k[k] / p.radius;
                ctx->walk_sc[2 * k] = std::sin(ang), ctx->walk_sc[2 * k + 1] = std::cos(ang);
            }
            ctx->walk_sc_n = n_t, ctx->walk_sc_step = p.simulation_step, ctx->walk_sc_radius = p.radius;
        }
        std::copy(ctx->walk_sc.begin(), ctx->walk_sc.end(), ctx->dist_k.begin() + sc_off);
    }
    {
        // PathElem::dist does not depend on the row: the stepper's independent variable advances by the same
        // step for every ray (phi += step / R, x = phi * R; flat and straight rays: x += step). The same
        // running sums here, element by element.
        const bool sph_rk4 = !S.flat && !S.straight;
        const double d = sph_rk4 ? p.simulation_step / S.radius : p.simulation_step;
        const double inv_radius = S.flat ? 0.0 : 1.0 / S.radius;
        // two entries more than the Fast generator's caches hold: the Rectilinear generator's stream ends at the
        // first state PAST max_distance (rectilinear.rs:175-177)
        const int n_x = n_t + 2;
        S.n_x = n_x;
        ctx->path_x.assign((size_t)2 * n_x, 0.0);
        double t = 0.0;
        ctx->path_k_far = n_t;
        for (int k = 1; k < n_x; ++k) {
            t += d;
            const double x = sph_rk4 ? t * S.radius : t;
            ctx->path_x[k] = x;
            const double dx = x - ctx->path_x[k - 1];
            ctx->path_x[(size_t)n_x + k] = S.flat ? dx : dx * inv_radius;  // calc_dist's dx / R (utils.rs:49)
        }
        for (int k = n_t - 1; k >= 0; --k)
            if (ctx->path_x[k] > p.max_distance) ctx->path_k_far = k;  // first element past max_distance (utils.rs:167)
        S.path_k_far = ctx->path_k_far;
    }
    S.n_pad = (n_t + 31) / 32 * 32;
    S.n1 = (n_t + CHUNK - 1) / CHUNK;
    S.n1_pad = (S.n1 + 31) / 32 * 32;
    S.n2 = (S.n1 + 31) / 32;
    S.h_pad = (p.height + 31) / 32 * 32;
    S.nobjects = (int)ctx->objects.size();
    // Walk anchors of stage A (kernels.cuh: anchored_lat_lon): a power of two of samples per anchor whose half span
    // stays below 2.5e-3 rad of arc (1024 samples at 25 m, 512 at 50 m); none for coarse steps or other walkers.
    S.anchor_shift = 0, S.n_anchor = 0;
    if (S.earth.walker == WALK_SPHERICAL && ctx->walk_anchors && p.generator == ATMRT_GENERATOR_FAST) {
        const double max_samples = 5.0e-3 * p.radius / p.simulation_step;
        int shift = 0;
        while (shift < 10 && (double)(2 << shift) <= max_samples) ++shift;
        if (shift >= 6) S.anchor_shift = shift, S.n_anchor = (n_t + (1 << shift) - 1) >> shift;
    }

    const size_t wl = (size_t)(p.x1 - p.x0), h = (size_t)p.height, np = (size_t)S.n_pad;
    const size_t f8 = sizeof(double);
    int e = 0;
    // Rectilinear: one ray and one walk per pixel, no caches; InterpolatingRectilinear: the caches belong to its grid,
    // which is prepared as a Fast render of its own (launch_interpolating)
    const bool rect = p.generator != ATMRT_GENERATOR_FAST;
    S.generator = p.generator;
    S.row_elev_deg = ctx->grid_row_elev, S.col_dir_deg = ctx->grid_col_dir;
    e |= ensure(ctx, ctx->d_dist, f8 * (3 * n_t + 2));  // dist_k, then (sin, cos)(dist_k / R)
    e |= ensure(ctx, ctx->d_pdist, f8 * 2 * S.n_x);
    if (!rect) {
    e |= ensure(ctx, ctx->d_colcalc, f8 * 8 * wl);
    if (S.n_anchor > 0) e |= ensure(ctx, ctx->d_anchor, sizeof(WalkAnchor) * wl * (size_t)S.n_anchor);
    e |= ensure(ctx, ctx->d_tlat, f8 * wl * np);
    e |= ensure(ctx, ctx->d_tlon, f8 * wl * np);
    e |= ensure(ctx, ctx->d_telev, f8 * wl * np);
    if (S.nobjects > 0) e |= ensure(ctx, ctx->d_tclose, 8 * wl * np);
    const size_t hp = (size_t)S.h_pad;
    e |= ensure(ctx, ctx->d_pdist, f8 * 2 * S.n_x);
    e |= ensure(ctx, ctx->d_pelev, f8 * hp * n_t);
    e |= ensure(ctx, ctx->d_plen, f8 * hp * n_t);
    e |= ensure(ctx, ctx->d_pn, sizeof(int) * hp);
    e |= ensure(ctx, ctx->d_tmin1, f8 * wl * S.n1_pad);
    e |= ensure(ctx, ctx->d_tmax1, f8 * wl * S.n1_pad);
    e |= ensure(ctx, ctx->d_tmin2, f8 * wl * S.n2);
    e |= ensure(ctx, ctx->d_tmax2, f8 * wl * S.n2);
    e |= ensure(ctx, ctx->d_tmin3, f8 * wl);
    e |= ensure(ctx, ctx->d_tmax3, f8 * wl);
    if (S.nobjects > 0) {
        e |= ensure(ctx, ctx->d_close1, 8 * wl * S.n1_pad);
        e |= ensure(ctx, ctx->d_close2, 8 * wl * S.n2);
        e |= ensure(ctx, ctx->d_close3, 8 * wl);
    }
    e |= ensure(ctx, ctx->d_rmin1, f8 * hp * S.n1);
    e |= ensure(ctx, ctx->d_rmax1, f8 * hp * S.n1);
    e |= ensure(ctx, ctx->d_rmin2, f8 * hp * S.n2);
    e |= ensure(ctx, ctx->d_rmax2, f8 * hp * S.n2);
    e |= ensure(ctx, ctx->d_rmin3, f8 * hp);
    e |= ensure(ctx, ctx->d_rmax3, f8 * hp);
    }
    (void)h;
    e |= ensure(ctx, ctx->d_obs, f8);
    e |= ensure(ctx, ctx->d_atm_cells, f8 * ATM_FIELDS * ATM_CELLS + sizeof(DevGPiece) * ATM_MAX_PIECES);
    e |= ensure(ctx, ctx->d_atm_aux, f8 * ATM_MAX_BND + ATM_CELLS);
    e |= ensure(ctx, ctx->d_sweep_flags, 16);
    if (!rect) {
        e |= ensure(ctx, ctx->d_sweep_col, wl);
        e |= ensure(ctx, ctx->d_sweep_hit, sizeof(int) * wl * (size_t)S.h_pad);
    }
    // the g(h) table depends on the atmosphere and the wavelength only: rebuilt when they change
    if (!ctx->atm_table_valid || memcmp(&ctx->atm_table_def, &p.atmosphere, sizeof(p.atmosphere)) != 0 || ctx->atm_table_wavelength != p.wavelength) {
        build_g_table(S.atm, p.wavelength, ctx->atm_cells, ctx->atm_pieces, &ctx->atm_cells_served);
        {  // where a macro step of the ray-path stage must not reach across (kernels.cuh: k_ray_paths_macro)
            std::vector<double>& bnd = ctx->atm_bnd;
            bnd.clear();
            bnd.push_back(ATM_BASE - 0.5 * ATM_CELL);
            bnd.push_back(ATM_BASE + (ATM_CELLS - 0.5) * ATM_CELL);
            for (int i = 1; i < S.atm.n; ++i) bnd.push_back(S.atm.layer[i].start);
            std::sort(bnd.begin(), bnd.end());
            bnd.resize(ATM_MAX_BND, std::numeric_limits<double>::infinity());
            ctx->atm_first.assign(ATM_CELLS, 0);
            for (int j = 0; j < ATM_CELLS; ++j) {
                const double lower = ATM_BASE + ((double)j - 0.5) * ATM_CELL;
                int t = 0;
                while (bnd[t] < lower) ++t;
                ctx->atm_first[j] = (unsigned char)t;
            }
        }
        ctx->atm_table_def = p.atmosphere;
        ctx->atm_table_wavelength = p.wavelength;
        ctx->atm_table_valid = true;
    }
    e |= ensure(ctx, ctx->d_counters, 8 * CNT_COUNT);
    e |= ensure(ctx, ctx->d_objects, sizeof(DevObject) * std::max(1, S.nobjects));
    e |= ensure(ctx, ctx->d_objects_in, sizeof(atmrt_object) * std::max(1, S.nobjects));
    if (e) return ATMRT_ERR_CUDA;

    DevBuffers& B = ctx->buf;
    B = DevBuffers{};
    B.dist_k = (const double*)ctx->d_dist.p;
    B.walk_sc = S.earth.walker == WALK_SPHERICAL ? (const double2*)((const double*)ctx->d_dist.p + S.n_sc_off) : nullptr;
    B.colcalc = (double*)ctx->d_colcalc.p;
    B.walk_anchor = S.n_anchor > 0 ? (WalkAnchor*)ctx->d_anchor.p : nullptr;
    B.t_lat = (double*)ctx->d_tlat.p, B.t_lon = (double*)ctx->d_tlon.p, B.t_elev = (double*)ctx->d_telev.p;
    B.terrain = ctx->terrain;
    B.t_close = S.nobjects > 0 ? (unsigned long long*)ctx->d_tclose.p : nullptr;
    B.path_x = (const double*)ctx->d_pdist.p, B.path_dxr = B.path_x + S.n_x;
    B.p_elev = (double*)ctx->d_pelev.p, B.p_len = (double*)ctx->d_plen.p;
    B.p_n = (int*)ctx->d_pn.p;
    B.tmin1 = (double*)ctx->d_tmin1.p, B.tmax1 = (double*)ctx->d_tmax1.p;
    B.tmin2 = (double*)ctx->d_tmin2.p, B.tmax2 = (double*)ctx->d_tmax2.p;
    B.tmin3 = (double*)ctx->d_tmin3.p, B.tmax3 = (double*)ctx->d_tmax3.p;
    B.close3 = S.nobjects > 0 ? (unsigned long long*)ctx->d_close3.p : nullptr;
    B.close1 = S.nobjects > 0 ? (unsigned long long*)ctx->d_close1.p : nullptr;
    B.close2 = S.nobjects > 0 ? (unsigned long long*)ctx->d_close2.p : nullptr;
    B.rmin1 = (double*)ctx->d_rmin1.p, B.rmax1 = (double*)ctx->d_rmax1.p;
    B.rmin2 = (double*)ctx->d_rmin2.p, B.rmax2 = (double*)ctx->d_rmax2.p;
    B.rmin3 = (double*)ctx->d_rmin3.p, B.rmax3 = (double*)ctx->d_rmax3.p;
    B.obs_alt = (double*)ctx->d_obs.p;
    B.objects = (DevObject*)ctx->d_objects.p;
    B.counters = (unsigned long long*)ctx->d_counters.p;
    B.atm_cells = (const double*)ctx->d_atm_cells.p;
    B.atm_pieces = (const DevGPiece*)((const char*)ctx->d_atm_cells.p + f8 * ATM_FIELDS * ATM_CELLS);
    B.n_atm_pieces = (int)ctx->atm_pieces.size();
    B.atm_bnd = (const double*)ctx->d_atm_aux.p;
    B.atm_first = (const unsigned char*)ctx->d_atm_aux.p + f8 * ATM_MAX_BND;
    B.sweep_flags = (unsigned*)ctx->d_sweep_flags.p;
    B.sweep_col = (unsigned char*)ctx->d_sweep_col.p;
    B.sweep_hit = (int*)ctx->d_sweep_hit.p;
    return 0;
}

// A copy from pageable memory is staged by the driver and holds up the head of the frame; the tables below change only
// with the parameters, so each is sent again only when its bytes (or its place on the device) differ from what was sent last.
static int upload_if_changed(atmrt_ctx* ctx, atmrt_ctx::Uploaded& u, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (u.dst == dst && u.bytes.size() == bytes && (bytes == 0 || memcmp(u.bytes.data(), src, bytes) == 0)) return 0;
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    u.dst = dst;
    u.bytes.assign((const char*)src, (const char*)src + bytes);
    return 0;
}

int upload_scene_inputs(atmrt_ctx* ctx, cudaStream_t s) {
    const DevScene& S = ctx->scene;
    int rc;
    if ((rc = upload_if_changed(ctx, ctx->up[0], ctx->d_dist.p, ctx->dist_k.data(), sizeof(double) * ctx->dist_k.size(), s))) return rc;
    if ((rc = upload_if_changed(ctx, ctx->up[1], ctx->d_pdist.p, ctx->path_x.data(), sizeof(double) * 2 * S.n_x, s))) return rc;
    if ((rc = upload_if_changed(ctx, ctx->up[2], ctx->d_atm_cells.p, ctx->atm_cells.data(), sizeof(double) * ATM_FIELDS * ATM_CELLS, s))) return rc;
    if ((rc = upload_if_changed(ctx, ctx->up[3], (char*)ctx->d_atm_cells.p + sizeof(double) * ATM_FIELDS * ATM_CELLS, ctx->atm_pieces.data(),
                                sizeof(DevGPiece) * ctx->atm_pieces.size(), s)))
        return rc;
    if ((rc = upload_if_changed(ctx, ctx->up[4], ctx->d_atm_aux.p, ctx->atm_bnd.data(), sizeof(double) * ATM_MAX_BND, s))) return rc;
    if ((rc = upload_if_changed(ctx, ctx->up[5], (char*)ctx->d_atm_aux.p + sizeof(double) * ATM_MAX_BND, ctx->atm_first.data(), ATM_CELLS, s))) return rc;
    if (S.nobjects > 0) {
        std::vector<DevObject> host(S.nobjects);
        for (int i = 0; i < S.nobjects; ++i) {
            const atmrt_object& o = ctx->objects[i];
            DevObject& d = host[i];
            memset(&d, 0, sizeof(d));
            d.kind = o.kind;
            d.tex_w = o.texture_width, d.tex_h = o.texture_height;
            d.r1 = o.r1, d.r2 = o.r2, d.width = o.width, d.height = o.height;
            d.close_r = o.kind == ATMRT_OBJECT_FRUSTUM ? std::fmax(o.r1, o.r2) : o.width;
            d.color = Color4{o.color[0], o.color[1], o.color[2], o.color[3]};
            d.tex = (const uint8_t*)ctx->textures[i].p;
        }
        // pageable -> synchronous staging inside cudaMemcpyAsync, so `host` may go out of scope
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_objects.p, host.data(), sizeof(DevObject) * S.nobjects, cudaMemcpyHostToDevice, s));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_objects_in.p, ctx->objects.data(), sizeof(atmrt_object) * S.nobjects, cudaMemcpyHostToDevice, s));
    }
    return 0;
}

struct RenderTargets {
    unsigned char* rgb = nullptr;
    atmrt_meta* meta = nullptr;
    int* steps = nullptr;
    atmrt_trace_point* points = nullptr;
    int* counts = nullptr;
    int max_points = 0;
    // host destinations (atmrt_render): when the horizon sweep serves the image, the shading is launched in
    // row bands and every finished band travels to the host while the next one is shaded
    unsigned char* host_rgb = nullptr;
    atmrt_meta* host_meta = nullptr;
    int* host_steps = nullptr;
    // the host image may be wider than this context's column block (a shard of a multi-GPU frame written in place):
    // `host_width` pixels per host row (0: the block's own width), the block starts at column `host_x0` of it
    int host_width = 0, host_x0 = 0;
    bool host_copied = false;  // out: the bands were copied (valid unless a device-side fallback ran afterwards)
};

// Rows [r0, r1) of the column block from the device planes to the host image (row-major; the block may be a
// sub-rectangle of a wider host image): strided 2-D copies, asynchronous on `st` when the host memory is pinned.
int copy_rows_to_host(atmrt_ctx* ctx, const RenderTargets& rt, int wl, int r0, int r1, cudaStream_t st) {
    const size_t hw = rt.host_width > 0 ? (size_t)rt.host_width : (size_t)wl;
    const size_t src = (size_t)r0 * wl, dst = (size_t)r0 * hw + (size_t)rt.host_x0;
    const size_t rows = (size_t)(r1 - r0);
    if (rt.host_rgb)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(rt.host_rgb + dst * 3, hw * 3, rt.rgb + src * 3, (size_t)wl * 3, (size_t)wl * 3, rows, cudaMemcpyDeviceToHost, st));
    if (rt.host_meta)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(rt.host_meta + dst, hw * sizeof(atmrt_meta), rt.meta + src, (size_t)wl * sizeof(atmrt_meta), (size_t)wl * sizeof(atmrt_meta),
                                        rows, cudaMemcpyDeviceToHost, st));
    if (rt.host_steps)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(rt.host_steps + dst, hw * sizeof(int), rt.steps + src, (size_t)wl * sizeof(int), (size_t)wl * sizeof(int), rows,
                                        cudaMemcpyDeviceToHost, st));
    return 0;
}

// The ray-path stage with macro steps: as many simulation steps per macro step as keep it within
// MACRO_MAX_METRES (16 at 25 or 50 m; fewer for coarser simulation steps, none above 400 m).
template <bool FLAT>
int launch_macro_paths(atmrt_ctx* ctx, const DevScene& S, const DevBuffers& B, int row0, int row1, cudaStream_t st, atmrt_ctx::StageEvents* E) {
    const int warps = MACRO_THREADS / 32, rows = row1 - row0;
    auto blocks = [&](int m) { return (rows + warps * (32 / m) - 1) / (warps * (32 / m)); };
    if (E) CUDA_TRY(ctx, cudaEventRecord(E->k0[ATMRT_KERNEL_RAY_PATHS], st));
    if (16.0 * S.step <= MACRO_MAX_METRES) k_ray_paths_macro<FLAT, 16><<<blocks(16), MACRO_THREADS, 0, st>>>(S, B, row0, row1);
    else if (8.0 * S.step <= MACRO_MAX_METRES) k_ray_paths_macro<FLAT, 8><<<blocks(8), MACRO_THREADS, 0, st>>>(S, B, row0, row1);
    else if (4.0 * S.step <= MACRO_MAX_METRES) k_ray_paths_macro<FLAT, 4><<<blocks(4), MACRO_THREADS, 0, st>>>(S, B, row0, row1);
    else if (2.0 * S.step <= MACRO_MAX_METRES) k_ray_paths_macro<FLAT, 2><<<blocks(2), MACRO_THREADS, 0, st>>>(S, B, row0, row1);
    else return 1;  // steps too long for macro steps: the caller integrates all rows with k_ray_paths
    if (E) {
        CUDA_TRY(ctx, cudaEventRecord(E->k1[ATMRT_KERNEL_RAY_PATHS], st));
        E->kmask |= 1u << ATMRT_KERNEL_RAY_PATHS;
    }
    return 0;
}

// Launch the whole render on (s_a || s_b) -> main. Asynchronous.
int next_stage_events(atmrt_ctx* ctx, atmrt_ctx::StageEvents** out) {
    if (ctx->ring.size() < atmrt_ctx::RING) {
        atmrt_ctx::StageEvents e{};
        cudaEvent_t* evs[] = {&e.a0, &e.a1, &e.b0, &e.b1, &e.c0, &e.c1, &e.t0, &e.t1};
        for (cudaEvent_t* ev : evs) CUDA_TRY(ctx, cudaEventCreate(ev));
        for (int i = 0; i < ATMRT_KERNEL_COUNT; ++i) {
            CUDA_TRY(ctx, cudaEventCreate(&e.k0[i]));
            CUDA_TRY(ctx, cudaEventCreate(&e.k1[i]));
        }
        ctx->ring.push_back(e);
    }
    atmrt_ctx::StageEvents* E = &ctx->ring[ctx->ring_next % atmrt_ctx::RING];
    E->kmask = 0u;
    ctx->ring_next++;
    ctx->ring_used = std::min(ctx->ring_used + 1, atmrt_ctx::RING);
    *out = E;
    return 0;
}

// events around one hot kernel (or one group of launches of it) of this render
#define KT_BEGIN(id, st)                                     \
    {                                                        \
        CUDA_TRY(ctx, cudaEventRecord(E->k0[id], st));       \
        E->kmask |= 1u << (id);                              \
    }
#define KT_END(id, st) CUDA_TRY(ctx, cudaEventRecord(E->k1[id], st));

int launch_interpolating(atmrt_ctx* ctx, RenderTargets& rt, cudaStream_t main);
int collect_stats(atmrt_ctx* ctx, atmrt_stats* stats);

int launch_render(atmrt_ctx* ctx, RenderTargets& rt, cudaStream_t main) {
    if (ctx->scene.generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR) return launch_interpolating(ctx, rt, main);
    const DevScene& S = ctx->scene;
    const DevBuffers& B = ctx->buf;
    const int wl = S.x1 - S.x0, h = S.height;
    ctx->launches = 0;
    atmrt_ctx::StageEvents* E = nullptr;
    int erc = next_stage_events(ctx, &E);
    if (erc) return erc;
    const bool timed = true;
    CUDA_TRY(ctx, cudaEventRecord(E->t0, main));
    int rc = upload_scene_inputs(ctx, main);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, 8 * CNT_COUNT, main));
    // A terrain that is still on its way (atmrt_group_render_tiles): the scene preparation reads it only for Relative altitudes
    // (Altitude::abs, params.rs:23-30) and the ray paths not at all -- they are integrated while the tiles arrive. The
    // Rectilinear generator's one kernel reads it from the start.
    bool early_terrain = false;
    if (ctx->terrain_in_flight) {
        early_terrain = S.altitude.kind == ATMRT_ALT_RELATIVE || S.generator == ATMRT_GENERATOR_RECTILINEAR;
        for (const atmrt_object& o : ctx->objects) early_terrain = early_terrain || o.altitude.kind == ATMRT_ALT_RELATIVE;
        if (early_terrain) CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_terrain, 0));
    }
    k_prepare_scene<<<(S.nobjects + 1 + 63) / 64, 64, 0, main>>>(S, ctx->terrain, B, (const atmrt_object*)ctx->d_objects_in.p);
    ctx->launches++;

    if (S.generator == ATMRT_GENERATOR_RECTILINEAR) {
        // generators/rectilinear.rs: one ray and one azimuth walk per pixel, a single kernel (the stage events
        // bracket it as the march; trace lists are a Fast-generator probe)
        if (rt.points || rt.counts) return fail(ctx, ATMRT_ERR_INVALID, "render_trace is not available with the Rectilinear generator");
        MarchOut O{rt.rgb, rt.meta, rt.steps, nullptr, nullptr, 0};
        cudaEvent_t marks[] = {E->a0, E->a1, E->b0, E->b1, E->c0};
        for (cudaEvent_t ev : marks) CUDA_TRY(ctx, cudaEventRecord(ev, main));
        const dim3 grid((wl + RECT_THREADS - 1) / RECT_THREADS, h);
        const int libm = ctx->path_mode == 1 ? 1 : 0;
        KT_BEGIN(ATMRT_KERNEL_RECTILINEAR, main)
        if (S.flat) {
            if (S.nobjects > 0) k_rectilinear<true, true><<<grid, RECT_THREADS, 0, main>>>(S, B, O, libm);
            else k_rectilinear<true, false><<<grid, RECT_THREADS, 0, main>>>(S, B, O, libm);
        } else {
            if (S.nobjects > 0) k_rectilinear<false, true><<<grid, RECT_THREADS, 0, main>>>(S, B, O, libm);
            else k_rectilinear<false, false><<<grid, RECT_THREADS, 0, main>>>(S, B, O, libm);
        }
        KT_END(ATMRT_KERNEL_RECTILINEAR, main)
        ctx->launches++;
        CUDA_TRY(ctx, cudaEventRecord(E->c1, main));
        CUDA_TRY(ctx, cudaEventRecord(E->t1, main));
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->rendered = true;
        return 0;
    }

    // Opaque terrain without objects ends every pixel at its first crossing: the horizon sweep
    // (kernels.cuh) finds those in O(N_t + H) per column when the rays of this render do not cross each
    // other, which k_path_check verifies on the device; the general march is launched behind it as the
    // fallback (it returns at once when the sweep served the image).
    const bool trace = rt.points != nullptr || rt.counts != nullptr;
    const bool brute = ctx->march_mode == 1, objs = S.nobjects > 0;
    const bool sweep = ctx->sweep_enabled && !objs && !trace && !brute && ctx->params.terrain_alpha == 1.0;
    // Translucent terrain and / or objects: the crossing march (kernels.cuh) under the same device-side check, with the
    // hierarchical march behind it.
    const bool cross = ctx->sweep_enabled && !sweep && !trace && !brute;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_sweep_flags.p, 0, 16, main));
    if (sweep) CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_sweep_col.p, 0, (size_t)wl, main));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_prep, main));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_a, ctx->ev_prep, 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_b, ctx->ev_prep, 0));

    // A narrow column block (one of several GPUs) is bound by the chain of stage B: the frame is split at a row below the
    // horizon from which every ray ends within a quarter of the distance (it dives to -1000 m: utils.rs:167-170). The rays
    // below the split are short chains, done long before the others; the sweep's lower band -- most of the hits -- needs
    // only them and the terrain, and runs while the long rays are still being integrated. The upper band then enters where
    // the lower one's top row left off (the row bands of k_sweep_bits), so the image is the unsplit one's, bit for bit.
    int split = 0;
    if (sweep && !S.straight && ctx->path_mode == 0 && 2.0 * S.step <= MACRO_MAX_METRES && S.altitude.kind == ATMRT_ALT_ABSOLUTE &&
        S.row_elev_deg == nullptr && (ctx->sweep_bands == -1 || (ctx->sweep_bands == 0 && wl <= ctx->num_sms * 32))) {
        const double drop = S.altitude.value + 1000.0, reach = 0.25 * S.max_distance;
        if (drop > 0.0 && reach > 0.0) {
            const double e0 = to_degrees(std::atan(drop / reach));  // rays steeper than this end within `reach`
            // get_ray_elev(y) = tilt - (y - h / 2) / h * fov / (w / h) (fast.rs:118-125)
            const double per_row = S.fov / (double)S.width;
            const double y0 = (double)(h / 2) + (S.tilt + e0) / per_row;
            if (y0 > 0.0 && y0 < (double)h) {
                const int cand = ((int)std::ceil(y0) + 31) / 32 * 32;
                if (cand >= 256 && cand <= h - 256) split = cand;
            }
        }
    }
    SweepLists L{};
    size_t segs = 0;
    if (sweep) {
        // Row bands: one walk per column unless the column block is so narrow that its walks could not fill a fraction of
        // the machine. (Measured at config 5: a band costs the scan of one whole row, +0.9 ms per extra band on the full
        // panorama and +0.1 ms on a 2048-column block of an 8-GPU frame; uniform bands pay below ~1000 columns.)
        int bands = ctx->sweep_bands > 0 ? ctx->sweep_bands : ctx->sweep_bands < 0 ? 1 : (int)(((long long)ctx->num_sms * 8) / std::max(wl, 1));
        bands = std::max(1, std::min(bands, std::max(1, h / 128)));
        if (bands > 1) split = 0;
        L.band_rows = ((h + bands - 1) / bands + 31) / 32 * 32;
        L.bands = (h + L.band_rows - 1) / L.band_rows;
        L.split = split;
        if (split) L.bands = 2, L.band_rows = std::max(split, h - split);
        L.cap = std::min(2 * L.band_rows + 2, S.n_pad);
        segs = (size_t)wl * L.bands;
        if ((rc = ensure(ctx, ctx->d_list, sizeof(int) * segs * L.cap))) return rc;
        if ((rc = ensure(ctx, ctx->d_count, sizeof(int) * segs))) return rc;
        if ((rc = ensure(ctx, ctx->d_normals, sizeof(double) * 3 * segs * L.cap))) return rc;
        L.list = (int*)ctx->d_list.p, L.count = (int*)ctx->d_count.p, L.normals = (double*)ctx->d_normals.p;
    }
    MarchOut O{rt.rgb, rt.meta, rt.steps, rt.points, rt.counts, rt.max_points};
    // the normals of the listed hit samples (one band or all) and the shading of rows [r0, r1) on a stream; every shaded band
    // of rows travels to the host on the (by then idle) stage-A stream while the next one is shaded
    const bool to_host = rt.host_rgb || rt.host_meta || rt.host_steps;
    int shade_band = 0;
    auto launch_normals = [&](cudaStream_t st, int only_band) {
        const size_t nseg = only_band >= 0 ? (size_t)wl : segs;
        // enough blocks per (column, band) that a narrow column block still fills the machine
        const int parts = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)ctx->num_sms * 16 + nseg - 1) / nseg));
        if (S.earth.walker == WALK_SPHERICAL) k_hit_normals<WALK_SPHERICAL><<<(unsigned)(nseg * parts), 128, 0, st>>>(S, B, L, parts, only_band);
        else k_hit_normals<-1><<<(unsigned)(nseg * parts), 128, 0, st>>>(S, B, L, parts, only_band);
        ctx->launches++;
    };
    auto launch_shade = [&](cudaStream_t st, int row0, int row1, int nbands, int count) -> int {
        const int band_rows = ((row1 - row0 + nbands - 1) / nbands + 31) / 32 * 32;
        for (int r0 = row0; r0 < row1; r0 += band_rows) {
            const int r1 = std::min(row1, r0 + band_rows);
            k_shade_tiles<<<dim3((r1 - r0 + 31) / 32, (wl + TILE_COLS - 1) / TILE_COLS), 32 * TILE_COLS, 0, st>>>(S, B, O, L, r0, count);
            ctx->launches++;
            if (to_host) {
                cudaEvent_t ev = ctx->ev_band[shade_band++ % SHADE_BANDS];
                CUDA_TRY(ctx, cudaEventRecord(ev, st));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_a, ev, 0));
                const int crc = copy_rows_to_host(ctx, rt, wl, r0, r1, ctx->s_a);
                if (crc) return crc;
            }
        }
        return 0;
    };
    // Stage B on s_b: all rows (every column needs every row)
    if (timed) CUDA_TRY(ctx, cudaEventRecord(E->b0, ctx->s_b));
    {
        const int rb = (h + 31) / 32;
        if (S.straight) {
            k_ray_paths_straight<<<(h + 127) / 128, 128, 0, ctx->s_b>>>(S, B);
        } else if (ctx->path_mode == 1 || ctx->path_mode == 2) {
            if (S.flat) {
                if (ctx->path_mode == 1) k_ray_paths<true, true><<<rb, 32, 0, ctx->s_b>>>(S, B);
                else k_ray_paths<true, false><<<rb, 32, 0, ctx->s_b>>>(S, B);
            } else {
                if (ctx->path_mode == 1) k_ray_paths<false, true><<<rb, 32, 0, ctx->s_b>>>(S, B);
                else k_ray_paths<false, false><<<rb, 32, 0, ctx->s_b>>>(S, B);
            }
        } else {
            // the long rays first, on the stream whose blocks are placed first
            int mrc = S.flat ? launch_macro_paths<true>(ctx, S, B, 0, split ? split : h, ctx->s_b, E) : launch_macro_paths<false>(ctx, S, B, 0, split ? split : h, ctx->s_b, E);
            if (mrc < 0) return mrc;
            if (mrc > 0) {  // steps too long for macro steps (never with a split frame: the condition above)
                if (S.flat) k_ray_paths<true, false><<<rb, 32, 0, ctx->s_b>>>(S, B);
                else k_ray_paths<false, false><<<rb, 32, 0, ctx->s_b>>>(S, B);
            } else if (split) {
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_b2, ctx->ev_prep, 0));
                mrc = S.flat ? launch_macro_paths<true>(ctx, S, B, split, h, ctx->s_b2, nullptr) : launch_macro_paths<false>(ctx, S, B, split, h, ctx->s_b2, nullptr);
                if (mrc) return mrc < 0 ? mrc : fail(ctx, ATMRT_ERR_CUDA, "stage B: inconsistent macro step");
                CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b1, ctx->s_b2));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_b, ctx->ev_b1, 0));  // the path check below looks at every row
                ctx->launches++;
            }
        }
        ctx->launches++;
    }
    auto path_pyramids = [&](cudaStream_t st, int only_if_not_swept) {
        k_path_pyramid1<<<dim3((h + 255) / 256, S.n1), 256, 0, st>>>(B.p_elev, B.p_n, h, S.h_pad, S.n_t, S.n1, B.rmin1, B.rmax1,
                                                                     only_if_not_swept ? B.sweep_flags : nullptr);
        k_path_pyramid23<<<(h + 255) / 256, 256, 0, st>>>(h, S.h_pad, S.n1, S.n2, B.rmin1, B.rmax1, B.rmin2, B.rmax2, B.rmin3, B.rmax3,
                                                          only_if_not_swept ? B.sweep_flags : nullptr);
        ctx->launches += 2;
    };
    if (sweep || cross) {
        k_path_check<<<dim3((h + 255) / 256, (S.n_t + 63) / 64), 256, 0, ctx->s_b>>>(B.p_elev, B.p_n, h, S.h_pad, S.n_t, B.sweep_flags);
        ctx->launches++;
    } else {
        path_pyramids(ctx->s_b, 0);
    }
    if (timed) CUDA_TRY(ctx, cudaEventRecord(E->b1, ctx->s_b));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->s_b));

    // Stage A on s_a
    if (ctx->terrain_in_flight && !early_terrain) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_a, ctx->ev_terrain, 0));
    if (timed) CUDA_TRY(ctx, cudaEventRecord(E->a0, ctx->s_a));
    k_column_setup<<<(wl + 127) / 128, 128, 0, ctx->s_a>>>(S, B);
    if (S.n_anchor > 0) {
        k_walk_anchors<<<dim3((S.n_anchor + 127) / 128, wl), 128, 0, ctx->s_a>>>(S, B);
        ctx->launches++;
    }
    KT_BEGIN(ATMRT_KERNEL_TERRAIN_PROFILE, ctx->s_a)
    if (S.earth.walker == WALK_SPHERICAL) k_terrain_profile<WALK_SPHERICAL><<<dim3((S.n_t + 127) / 128, wl), 128, 0, ctx->s_a>>>(S, ctx->terrain, B, 0);
    else k_terrain_profile<-1><<<dim3((S.n_t + 127) / 128, wl), 128, 0, ctx->s_a>>>(S, ctx->terrain, B, 0);
    KT_END(ATMRT_KERNEL_TERRAIN_PROFILE, ctx->s_a)
    ctx->launches += 2;
    auto terrain_pyramids = [&](cudaStream_t st, int only_if_not_swept) {
        const long long warps = (long long)wl * S.n2;
        k_terrain_pyramid<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(B.t_elev, B.t_close, wl, S.n_t, S.n_pad, S.n1, S.n1_pad, S.n2, B.tmin1,
                                                                                B.tmax1, B.tmin2, B.tmax2, B.close1, B.close2,
                                                                                only_if_not_swept ? B.sweep_flags : nullptr);
        k_terrain_top<<<(wl + 127) / 128, 128, 0, st>>>(B.tmin2, B.tmax2, B.close2, wl, S.n2, B.tmin3, B.tmax3, B.close3,
                                                        only_if_not_swept ? B.sweep_flags : nullptr);
        ctx->launches += 2;
    };
    if (!sweep && !cross) terrain_pyramids(ctx->s_a, 0);
    if (timed) CUDA_TRY(ctx, cudaEventRecord(E->a1, ctx->s_a));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->s_a));
    if (split) {  // the lower band of the sweep: the terrain and the short rays are all it reads
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_c, ctx->ev_a, 0));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_c, ctx->ev_b1, 0));
        k_sweep_bits<true><<<dim3((wl + BITS_WARPS - 1) / BITS_WARPS, 1), 32 * BITS_WARPS, 0, ctx->s_c>>>(S, B, L, 0, wl, 1);
        ctx->launches++;
        launch_normals(ctx->s_c, 1);
        if ((rc = launch_shade(ctx->s_c, split, h, to_host ? SHADE_BANDS / 2 : 1, 0))) return rc;  // (counted below, once every check is in)
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_s1, ctx->s_c));
    }

    // Stage C on main
    CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_a, 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_b, 0));
    if (timed) CUDA_TRY(ctx, cudaEventRecord(E->c0, main));
    const dim3 grid((h + MARCH_THREADS - 1) / MARCH_THREADS, wl);
    if (sweep) {
        // Bit-mask sweep -> normals of the distinct hit samples -> row-major shading (kernels.cuh).
        // Row bands: one walk per column unless the column block is so narrow that its walks could not fill a fraction of
        // the machine. (Measured at config 5: a band costs the scan of one whole row, +0.9 ms per extra band on the full
        // panorama and +0.1 ms on a 2048-column block of an 8-GPU frame; bands pay below ~1000 columns.)
        KT_BEGIN(ATMRT_KERNEL_SWEEP, main)
        // (a split frame: the upper band here, after every ray and the terrain; the lower one went out on s_c behind stage A
        // and the short rays, with its normals and its shading)
        if (split) k_sweep_bits<true><<<dim3((wl + BITS_WARPS - 1) / BITS_WARPS, 1), 32 * BITS_WARPS, 0, main>>>(S, B, L, 0, wl, 0);
        else k_sweep_bits<false><<<dim3((wl + BITS_WARPS - 1) / BITS_WARPS, L.bands), 32 * BITS_WARPS, 0, main>>>(S, B, L, 0, wl, -1);
        KT_END(ATMRT_KERNEL_SWEEP, main)
        ctx->launches++;
        KT_BEGIN(ATMRT_KERNEL_HIT_NORMALS, main)
        launch_normals(main, split ? 0 : -1);
        KT_END(ATMRT_KERNEL_HIT_NORMALS, main)
        KT_BEGIN(ATMRT_KERNEL_SHADE, main)
        if ((rc = launch_shade(main, 0, split ? split : h, to_host && h >= 256 ? (split ? SHADE_BANDS / 2 : SHADE_BANDS) : 1, 1))) return rc;
        KT_END(ATMRT_KERNEL_SHADE, main)
        if (split) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_s1, 0));  // the lower band's pixels
            k_count_swept<<<(wl + 7) / 8, 256, 0, main>>>(S, B, L, split, h);
            ctx->launches++;
        }
        if (to_host) {
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->s_a));
            CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_a, 0));
            rt.host_copied = true;
        }
    }
    if (sweep) {
        // fallbacks, no-ops unless the device-side checks ask for them: pyramids + hierarchical march of the
        // whole image (rays cross), brute-force march of flagged columns
        terrain_pyramids(main, 1);
        path_pyramids(main, 1);
        const dim3 fgrid(grid.x, std::min(wl, 64));
        k_march<false, false, false><<<fgrid, MARCH_THREADS, 0, main>>>(S, B, O, MARCH_IF_NOT_SWEPT);
        k_march<false, true, false><<<fgrid, MARCH_THREADS, 0, main>>>(S, B, O, MARCH_FLAGGED_COLUMNS);
        ctx->launches += 2;
    } else if (cross) {
        const size_t samples = (size_t)wl * S.n_pad;
        // thresholds [column][n_pad] | per window of 32 steps: min, max of the thresholds [column][n1_pad] each
        const size_t off_win = align_up(sizeof(unsigned short) * samples, 256), win_bytes = align_up(sizeof(unsigned short) * (size_t)wl * S.n1_pad, 256);
        if ((rc = ensure(ctx, ctx->d_cross, off_win + 2 * win_bytes))) return rc;
        unsigned short* thresholds = (unsigned short*)ctx->d_cross.p;
        unsigned short* window_min = (unsigned short*)((char*)ctx->d_cross.p + off_win);
        unsigned short* window_max = (unsigned short*)((char*)ctx->d_cross.p + off_win + win_bytes);
        unsigned short* trig_slot = nullptr;
        double* trig = nullptr;
        int* trig_count = nullptr;
        if (objs) {  // per column: slot of every sample next to a close object | its four sines and cosines | the slot counter
            const size_t off_trig = align_up(sizeof(unsigned short) * samples, 256), off_count = off_trig + sizeof(double) * 4 * CROSS_TRIG_CAP * (size_t)wl;
            if ((rc = ensure(ctx, ctx->d_cross_trig, off_count + sizeof(int) * (size_t)wl))) return rc;
            trig_slot = (unsigned short*)ctx->d_cross_trig.p;
            trig = (double*)((char*)ctx->d_cross_trig.p + off_trig);
            trig_count = (int*)((char*)ctx->d_cross_trig.p + off_count);
            CUDA_TRY(ctx, cudaMemsetAsync(trig_count, 0, sizeof(int) * (size_t)wl, main));
        }
        const int bands = (h + CROSS_BAND - 1) / CROSS_BAND;
        KT_BEGIN(ATMRT_KERNEL_MARCH, main)
        k_thresholds<<<dim3((S.n_t + 127) / 128, wl), 128, 0, main>>>(S, B, thresholds, window_min, window_max, objs ? 1 : 0, trig_slot, trig, trig_count);
        if (objs) k_cross_march<true><<<dim3((wl + CROSS_WARPS - 1) / CROSS_WARPS, bands), 32 * CROSS_WARPS, 0, main>>>(S, B, O, thresholds, window_min, window_max, trig_slot, trig);
        else k_cross_march<false><<<dim3((wl + CROSS_WARPS - 1) / CROSS_WARPS, bands), 32 * CROSS_WARPS, 0, main>>>(S, B, O, thresholds, window_min, window_max, nullptr, nullptr);
        // the fallback, a no-op unless the rays of this render cross
        terrain_pyramids(main, 1);
        path_pyramids(main, 1);
        if (objs) k_march<true, false, false><<<dim3(grid.x, std::min(wl, 64)), MARCH_THREADS, 0, main>>>(S, B, O, MARCH_IF_NOT_SWEPT);
        else k_march<false, false, false><<<dim3(grid.x, std::min(wl, 64)), MARCH_THREADS, 0, main>>>(S, B, O, MARCH_IF_NOT_SWEPT);
        KT_END(ATMRT_KERNEL_MARCH, main)
        ctx->launches += 3;
    } else {
        KT_BEGIN(ATMRT_KERNEL_MARCH, main)
#define ATMRT_LAUNCH_MARCH(OB, BR, TR) k_march<OB, BR, TR><<<grid, MARCH_THREADS, 0, main>>>(S, B, O, MARCH_ALWAYS)
        if (objs) {
            if (brute) { if (trace) ATMRT_LAUNCH_MARCH(true, true, true); else ATMRT_LAUNCH_MARCH(true, true, false); }
            else       { if (trace) ATMRT_LAUNCH_MARCH(true, false, true); else ATMRT_LAUNCH_MARCH(true, false, false); }
        } else {
            if (brute) { if (trace) ATMRT_LAUNCH_MARCH(false, true, true); else ATMRT_LAUNCH_MARCH(false, true, false); }
            else       { if (trace) ATMRT_LAUNCH_MARCH(false, false, true); else ATMRT_LAUNCH_MARCH(false, false, false); }
        }
#undef ATMRT_LAUNCH_MARCH
        KT_END(ATMRT_KERNEL_MARCH, main)
        ctx->launches++;
    }
    if (timed) {
        CUDA_TRY(ctx, cudaEventRecord(E->c1, main));
        CUDA_TRY(ctx, cudaEventRecord(E->t1, main));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->rendered = true;
    return 0;
}

// FovData::min_elev_step / min_dir_step (interpolating_rectilinear.rs:432-521) of the image `S` describes, in radians
int interpolating_steps(atmrt_ctx* ctx, const DevScene& S, cudaStream_t main, double* elev_step, double* dir_step) {
    int rc = ensure(ctx, ctx->d_interp, 64);
    if (rc) return rc;
    const double two_pi = 360.0 * (PI / 180.0);
    unsigned long long init[2];
    memcpy(&init[0], &two_pi, 8), init[1] = init[0];
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_interp.p, init, sizeof(init), cudaMemcpyHostToDevice, main));
    k_interp_steps<<<dim3((S.width + 127) / 128, S.height), 128, 0, main>>>(S, (unsigned long long*)ctx->d_interp.p);
    CUDA_TRY(ctx, cudaGetLastError());
    double mins[2];
    CUDA_TRY(ctx, cudaMemcpyAsync(mins, ctx->d_interp.p, sizeof(mins), cudaMemcpyDeviceToHost, main));
    CUDA_TRY(ctx, cudaStreamSynchronize(main));
    const double scale = 1.5;
    *elev_step = mins[0] * scale, *dir_step = mins[1] * scale;
    return 0;
}

// InterpolatingRectilinearGenerator::generate (interpolating_rectilinear.rs:121-170). The reference fills its caches
// lazily, grid point by grid point; here the grid that covers the column block is one Fast render with explicit angle
// tables that keeps its trace lists on the device, and one kernel blends them into the image.
int launch_interpolating(atmrt_ctx* ctx, RenderTargets& rt, cudaStream_t main) {
    const atmrt_params image_params = ctx->params;
    const DevScene S = ctx->scene;  // the image: frame, size, column block, shading
    const int wl = S.x1 - S.x0, h = S.height;
    double elev_step, dir_step;
    int rc = interpolating_steps(ctx, S, main, &elev_step, &dir_step);
    if (rc) return rc;
    if (!(elev_step > 0.0) || !(dir_step > 0.0)) return fail(ctx, ATMRT_ERR_INVALID, "InterpolatingRectilinear: the field of view yields no grid step");
    // the grid points the column block reads: floor(angle / step) and the next one, per pixel (FovData::cache_coords, :185-204)
    int range[4] = {INT_MAX, -INT_MAX, INT_MAX, -INT_MAX};
    int* d_range = (int*)((char*)ctx->d_interp.p + 16);
    CUDA_TRY(ctx, cudaMemcpyAsync(d_range, range, sizeof(range), cudaMemcpyHostToDevice, main));
    k_interp_range<<<dim3((wl + 127) / 128, h), 128, 0, main>>>(S, elev_step, dir_step, d_range);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(range, d_range, sizeof(range), cudaMemcpyDeviceToHost, main));
    CUDA_TRY(ctx, cudaStreamSynchronize(main));
    const long long rows = (long long)range[1] + 1 - range[0] + 1, cols = (long long)range[3] + 1 - range[2] + 1;
    if (rows < 2 || cols < 2 || rows > 32767 || cols > 32767) return fail(ctx, ATMRT_ERR_INVALID, "InterpolatingRectilinear: the grid does not fit 32767 x 32767 points");
    InterpGrid G{};
    G.elev_step = elev_step, G.dir_step = dir_step;
    G.elev_top = range[1] + 1, G.dir_left = range[2];
    G.rows = (int)rows, G.cols = (int)cols, G.max_points = INTERP_MAX_POINTS;
    // Cache::get_path_cache / get_terrain_cache, :46-84: index * step, in degrees; rows from the highest elevation down
    std::vector<double> angles((size_t)(rows + cols));
    for (int j = 0; j < G.rows; ++j) angles[(size_t)j] = to_degrees((double)(G.elev_top - j) * elev_step);
    for (int i = 0; i < G.cols; ++i) angles[(size_t)G.rows + i] = to_degrees((double)(G.dir_left + i) * dir_step);
    if ((rc = ensure(ctx, ctx->d_grid_angles, sizeof(double) * angles.size()))) return rc;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_grid_angles.p, angles.data(), sizeof(double) * angles.size(), cudaMemcpyHostToDevice, main));
    CUDA_TRY(ctx, cudaStreamSynchronize(main));  // `angles` is pageable
    const size_t npoints = (size_t)rows * (size_t)cols;
    if ((rc = ensure(ctx, ctx->d_grid_points, npoints * sizeof(atmrt_trace_point) * INTERP_MAX_POINTS))) return rc;
    if ((rc = ensure(ctx, ctx->d_grid_counts, npoints * sizeof(int)))) return rc;

    // ---- the grid: a Fast render of rows x cols pixels with these angles, trace lists kept ----
    atmrt_params grid_params = image_params;
    grid_params.generator = ATMRT_GENERATOR_FAST;
    grid_params.width = G.cols, grid_params.height = G.rows, grid_params.x0 = 0, grid_params.x1 = G.cols;
    ctx->params = grid_params;
    ctx->grid_row_elev = (const double*)ctx->d_grid_angles.p;
    ctx->grid_col_dir = (const double*)ctx->d_grid_angles.p + G.rows;
    rc = prepare_render(ctx);
    if (!rc) {
        RenderTargets grid;
        grid.points = (atmrt_trace_point*)ctx->d_grid_points.p;
        grid.counts = (int*)ctx->d_grid_counts.p;
        grid.max_points = INTERP_MAX_POINTS;
        rc = launch_render(ctx, grid, main);
    }
    if (!rc && cudaStreamSynchronize(main) != cudaSuccess) rc = fail(ctx, ATMRT_ERR_CUDA, "InterpolatingRectilinear: the grid render failed");
    if (!rc) rc = collect_stats(ctx, &ctx->grid_stats);  // ray steps, path steps, terrain samples are the grid's
    const int grid_launches = ctx->launches;
    ctx->params = image_params;
    ctx->grid_row_elev = ctx->grid_col_dir = nullptr;
    ctx->scene = S;
    if (rc) return rc;

    // ---- the image: blend the four grid pixels around every pixel's own ray, colour and composite ----
    G.points = (const atmrt_trace_point*)ctx->d_grid_points.p;
    G.counts = (const int*)ctx->d_grid_counts.p;
    const DevBuffers& B = ctx->buf;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, 8 * CNT_COUNT, main));
    MarchOut O{rt.rgb, rt.meta, rt.steps, rt.points, rt.counts, rt.max_points};
    const dim3 grid_dim((wl + 127) / 128, h);
    if (rt.points || rt.counts) k_interp_blend<true><<<grid_dim, 128, 0, main>>>(S, B, O, G);
    else k_interp_blend<false><<<grid_dim, 128, 0, main>>>(S, B, O, G);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches = grid_launches + 3;
    ctx->rendered = true;
    return 0;
}

int collect_stats(atmrt_ctx* ctx, atmrt_stats* stats) {
    if (!stats) return 0;
    const DevScene& S = ctx->scene;
    unsigned long long c[CNT_COUNT];
    CUDA_TRY(ctx, cudaMemcpy(c, ctx->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost));
    std::vector<int> pn(S.height, 0);
    if (S.generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR) {
        // the marches are the grid's; trace points and hit pixels are the blended image's; a grid pixel with more trace
        // points than the blend keeps counts as an overflow of every image pixel that reads it
        *stats = ctx->grid_stats;
        stats->trace_points = c[CNT_TRACE_POINTS];
        stats->pixels_hit = c[CNT_PIXELS_HIT];
        stats->step_overflows = ctx->grid_stats.step_overflows + c[CNT_OVERFLOWS];
        stats->kernel_launches = ctx->launches;
        return 0;
    }
    const bool rect = S.generator == ATMRT_GENERATOR_RECTILINEAR;  // no path cache
    if (!rect) CUDA_TRY(ctx, cudaMemcpy(pn.data(), ctx->d_pn.p, sizeof(int) * S.height, cudaMemcpyDeviceToHost));
    memset(stats, 0, sizeof(*stats));
    stats->ray_steps = c[CNT_RAY_STEPS];
    stats->trace_points = c[CNT_TRACE_POINTS];
    stats->pixels_hit = c[CNT_PIXELS_HIT];
    stats->step_overflows = c[CNT_OVERFLOWS];
    stats->path_steps = c[CNT_PATH_STEPS];
    stats->terrain_samples = rect ? c[CNT_PATH_STEPS] : (uint64_t)(S.x1 - S.x0) * (uint64_t)S.n_t;
    stats->n_terrain = rect ? 0 : S.n_t;
    int mx = 0;
    for (int v : pn) mx = std::max(mx, v);
    stats->n_path_max = mx;
    if (ctx->ring_used > 0) {
        const atmrt_ctx::StageEvents& E = ctx->ring[(ctx->ring_next - 1) % atmrt_ctx::RING];
        cudaEventElapsedTime(&stats->ms_terrain, E.a0, E.a1);
        cudaEventElapsedTime(&stats->ms_paths, E.b0, E.b1);
        cudaEventElapsedTime(&stats->ms_march, E.c0, E.c1);
        cudaEventElapsedTime(&stats->ms_total, E.t0, E.t1);
    }
    stats->kernel_launches = ctx->launches;
    return 0;
}

}  // namespace

// ---- multi-GPU group (atmrt_group_*): state and helpers ------------------------------------------
struct atmrt_group {
    std::vector<atmrt_ctx*> ctx;
    std::string err;
    atmrt_params params{};
    bool has_params = false;
    std::vector<void*> packed;  // per GPU: the packed terrain (owned)
    size_t packed_cap = 0;
    std::vector<cudaEvent_t> ev_slice;  // per GPU: its slice is retiled
};

static int gfail(atmrt_group* g, int code, const std::string& msg) {
    if (g) g->err = msg;
    else g_create_error = msg;
    return code;
}

// run fn(i) for every context on its own host thread; the first failure wins
template <class F>
static int group_parallel(atmrt_group* g, F fn) {
    const int n = (int)g->ctx.size();
    std::vector<int> rc(n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back([&, i] { rc[i] = fn(i); });
    rc[0] = fn(0);
    for (auto& t : th) t.join();
    for (int i = 0; i < n; ++i)
        if (rc[i]) return gfail(g, rc[i], "GPU " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err);
    return 0;
}

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int atmrt_abi_version(void) { return ATMRT_ABI_VERSION; }

// sizes of the ABI structs, for the ctypes mirror's self-check
int atmrt_abi_sizes(size_t* out, int n) {
    const size_t v[] = {sizeof(atmrt_altitude), sizeof(atmrt_atmosphere_def), sizeof(atmrt_params), sizeof(atmrt_tile_desc),
                        sizeof(atmrt_object),   sizeof(atmrt_meta),           sizeof(atmrt_trace_point), sizeof(atmrt_stats),
                        sizeof(atmrt_stage_ms), sizeof(atmrt_kernel_ms)};
    const int m = (int)(sizeof(v) / sizeof(v[0]));
    for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
    return m;
}

const char* atmrt_last_error(const atmrt_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int atmrt_create(int device, atmrt_ctx** out) {
    if (!out) return fail(nullptr, ATMRT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, ATMRT_ERR_NO_DEVICE,
                    std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, ATMRT_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, ATMRT_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major < 10) return fail(nullptr, ATMRT_ERR_NO_DEVICE, "device is not sm_100 (Blackwell): the library is built for sm_100a only");
    atmrt_ctx* ctx = new atmrt_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, ATMRT_ERR_CUDA, "cudaSetDevice failed");
    }
    // Stage B is a handful of latency-bound warps, stage A saturates the issue slots: give B's stream
    // the higher priority so that its blocks are placed first.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    bool ok = cudaStreamCreateWithPriority(&ctx->s_a, cudaStreamNonBlocking, prio_lo) == cudaSuccess &&
              cudaStreamCreateWithPriority(&ctx->s_b, cudaStreamNonBlocking, prio_hi) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_main, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->s_t, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->s_r, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&ctx->s_b2, cudaStreamNonBlocking, prio_hi) == cudaSuccess && cudaStreamCreateWithFlags(&ctx->s_c, cudaStreamNonBlocking) == cudaSuccess;
    cudaEvent_t* evs[] = {&ctx->ev_prep, &ctx->ev_a, &ctx->ev_b, &ctx->ev_terrain, &ctx->ev_tile, &ctx->ev_b1, &ctx->ev_s1, &ctx->ev_last};
    for (cudaEvent_t* ev : evs) ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
    for (cudaEvent_t& ev : ctx->ev_band) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
    cudaEvent_t* tevs[] = {&ctx->t_0, &ctx->t_1};
    for (cudaEvent_t* ev : tevs) ok = ok && cudaEventCreate(ev) == cudaSuccess;
    if (!ok) {
        delete ctx;
        return fail(nullptr, ATMRT_ERR_CUDA, "stream/event creation failed");
    }
    *out = ctx;
    return 0;
}

void atmrt_destroy(atmrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    DevBuf* bufs[] = {&ctx->d_objects_in, &ctx->d_objects, &ctx->d_dist, &ctx->d_colcalc, &ctx->d_tlat, &ctx->d_tlon, &ctx->d_telev,
                      &ctx->d_tclose, &ctx->d_pdist, &ctx->d_pelev, &ctx->d_plen, &ctx->d_pn,
                      &ctx->d_tmin1, &ctx->d_tmax1, &ctx->d_tmin2, &ctx->d_tmax2, &ctx->d_tmin3, &ctx->d_tmax3, &ctx->d_close1, &ctx->d_close2, &ctx->d_close3, &ctx->d_rmin1, &ctx->d_rmin3, &ctx->d_rmax3,
                      &ctx->d_rmax1, &ctx->d_rmin2, &ctx->d_rmax2, &ctx->d_obs, &ctx->d_counters, &ctx->d_atm_cells, &ctx->d_sweep_flags, &ctx->d_sweep_col, &ctx->d_sweep_hit, &ctx->d_cross, &ctx->d_cross_trig, &ctx->d_list, &ctx->d_count, &ctx->d_normals, &ctx->d_anchor, &ctx->d_stage, &ctx->d_atm_aux, &ctx->d_rgb, &ctx->d_meta,
                      &ctx->d_steps, &ctx->d_points, &ctx->d_counts, &ctx->d_interp, &ctx->d_grid_angles, &ctx->d_grid_points, &ctx->d_grid_counts, &ctx->d_probe_a, &ctx->d_probe_b, &ctx->d_probe_c, &ctx->d_probe_d};
    for (DevBuf* b : bufs) release(*b);
    for (DevBuf& b : ctx->textures) release(b);
    if (ctx->terrain_owned) cudaFree(ctx->terrain_owned);
    cudaEvent_t evs[] = {ctx->ev_prep, ctx->ev_a, ctx->ev_b, ctx->t_0, ctx->t_1, ctx->ev_last};
    for (cudaEvent_t ev : evs)
        if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev_band)
        if (ev) cudaEventDestroy(ev);
    for (atmrt_ctx::StageEvents& e : ctx->ring) {
        cudaEvent_t all[] = {e.a0, e.a1, e.b0, e.b1, e.c0, e.c1, e.t0, e.t1};
        for (cudaEvent_t ev : all)
            if (ev) cudaEventDestroy(ev);
        for (int i = 0; i < ATMRT_KERNEL_COUNT; ++i) {
            if (e.k0[i]) cudaEventDestroy(e.k0[i]);
            if (e.k1[i]) cudaEventDestroy(e.k1[i]);
        }
    }
    if (ctx->s_a) cudaStreamDestroy(ctx->s_a);
    if (ctx->s_b) cudaStreamDestroy(ctx->s_b);
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    if (ctx->s_b2) cudaStreamDestroy(ctx->s_b2);
    if (ctx->s_c) cudaStreamDestroy(ctx->s_c);
    if (ctx->ev_b1) cudaEventDestroy(ctx->ev_b1);
    if (ctx->ev_s1) cudaEventDestroy(ctx->ev_s1);
    if (ctx->s_t) cudaStreamDestroy(ctx->s_t);
    if (ctx->s_r) cudaStreamDestroy(ctx->s_r);
    if (ctx->ev_terrain) cudaEventDestroy(ctx->ev_terrain);
    if (ctx->ev_tile) cudaEventDestroy(ctx->ev_tile);
    delete ctx;
}

int atmrt_terrain_packed_bytes(const atmrt_tile_desc* tiles, int ntiles, size_t* bytes) {
    if (!bytes) return ATMRT_ERR_INVALID;
    TerrainLayout L;
    int rc = make_layout(nullptr, tiles, ntiles, &L);
    if (rc) return rc;
    *bytes = L.total;
    return 0;
}

int atmrt_pack_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, void* dev_dst) {
    if (!ctx || !dev_dst || (ntiles > 0 && !posts)) return fail(ctx, ATMRT_ERR_INVALID, "pack_terrain: NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TerrainLayout L;
    int rc = make_layout(ctx, tiles, ntiles, &L);
    if (rc) return rc;
    char* base = (char*)dev_dst;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemsetAsync(base, 0, L.total, s));
    if (ntiles > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(base + L.off_tiles, L.tiles.data(), sizeof(DevTile) * ntiles, cudaMemcpyHostToDevice, s));
    if (!L.lookup.empty())
        CUDA_TRY(ctx, cudaMemcpyAsync(base + L.off_lookup, L.lookup.data(), sizeof(int) * L.lookup.size(), cudaMemcpyHostToDevice, s));
    // All tiles travel back to back into one persistent staging area (the copy engine never waits for a
    // kernel), then the retile kernels run; one synchronisation at the end.
    std::vector<size_t> offs(ntiles + 1, 0);
    for (int i = 0; i < ntiles; ++i) offs[i + 1] = offs[i] + align_up(sizeof(int16_t) * (size_t)tiles[i].nlon * tiles[i].nlat, 256);
    rc = ensure(ctx, ctx->d_stage, std::max<size_t>(offs[ntiles], 256));
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < ntiles && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync((char*)ctx->d_stage.p + offs[i], posts[i], sizeof(int16_t) * (size_t)tiles[i].nlon * tiles[i].nlat, cudaMemcpyHostToDevice, s);
    for (int i = 0; i < ntiles && e == cudaSuccess; ++i) {
        const size_t n = (size_t)tiles[i].nlon * tiles[i].nlat;
        k_retile<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((const int16_t*)((const char*)ctx->d_stage.p + offs[i]), (int16_t*)(base + L.off_posts), L.tiles[i]);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(ctx, ATMRT_ERR_CUDA, std::string("pack_terrain: ") + cudaGetErrorString(e));
    return 0;
}

int atmrt_bind_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles, const void* dev_packed) {
    if (!ctx || !dev_packed) return fail(ctx, ATMRT_ERR_INVALID, "bind_terrain: NULL argument");
    TerrainLayout L;
    int rc = make_layout(ctx, tiles, ntiles, &L);
    if (rc) return rc;
    const char* base = (const char*)dev_packed;
    ctx->tile_descs.assign(tiles, tiles + ntiles);
    ctx->tiles_host = L.tiles;
    ctx->terrain.tiles = (const DevTile*)(base + L.off_tiles);
    ctx->terrain.lookup = (const int*)(base + L.off_lookup);
    ctx->terrain.posts = (const int16_t*)(base + L.off_posts);
    ctx->terrain.lat_min = L.lat_min, ctx->terrain.lon_min = L.lon_min;
    ctx->terrain.nlat_tiles = L.nlat_tiles, ctx->terrain.nlon_tiles = L.nlon_tiles;
    ctx->terrain.ntiles = ntiles;
    {  // a regular terrain (device_math.cuh): every tile is exactly its one-degree cell on one common grid
        DevTerrain& T = ctx->terrain;
        T.regular = ntiles > 0 ? 1 : 0;
        for (int i = 0; i < ntiles && T.regular; ++i) {
            const atmrt_tile_desc& d = tiles[i];
            const DevTile& t = L.tiles[i];
            if (d.nlat != tiles[0].nlat || d.nlon != tiles[0].nlon || d.lat_interval != tiles[0].lat_interval || d.lon_interval != tiles[0].lon_interval ||
                d.min_lat != (double)d.lat0 || d.min_lon != (double)d.lon0 || t.max_lat != d.min_lat + 1.0 || t.max_lon != d.min_lon + 1.0 ||
                d.lat0 < -32768 || d.lat0 > 32767 || d.lon0 < -32768 || d.lon0 > 32767)
                T.regular = 0;
        }
        if (T.regular) {
            T.r_nlat = L.tiles[0].nlat, T.r_nlon = L.tiles[0].nlon, T.r_mt_lat = L.tiles[0].mt_lat;
            T.r_lat_interval = L.tiles[0].lat_interval, T.r_lon_interval = L.tiles[0].lon_interval;
            T.r_inv_lat_interval = L.tiles[0].inv_lat_interval, T.r_inv_lon_interval = L.tiles[0].inv_lon_interval;
        }
    }
    ctx->has_terrain = true;
    ctx->rendered = false;
    return 0;
}

int atmrt_set_terrain(atmrt_ctx* ctx, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts) {
    if (!ctx) return ATMRT_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    size_t bytes = 0;
    TerrainLayout L;
    int rc = make_layout(ctx, tiles, ntiles, &L);
    if (rc) return rc;
    bytes = L.total;
    if (ctx->terrain_owned) {
        cudaFree(ctx->terrain_owned);
        ctx->terrain_owned = nullptr;
        ctx->has_terrain = false;
    }
    CUDA_TRY(ctx, cudaMalloc(&ctx->terrain_owned, bytes));
    rc = atmrt_pack_terrain(ctx, tiles, ntiles, posts, ctx->terrain_owned);
    if (rc) return rc;
    return atmrt_bind_terrain(ctx, tiles, ntiles, ctx->terrain_owned);
}

int atmrt_get_elev(atmrt_ctx* ctx, const double* lat, const double* lon, int n, double* elev) {
    if (!ctx || !lat || !lon || !elev || n < 0) return fail(ctx, ATMRT_ERR_INVALID, "get_elev: bad argument");
    if (!ctx->has_terrain) return fail(ctx, ATMRT_ERR_STATE, "get_elev before set_terrain");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    int rc = ensure(ctx, ctx->d_probe_a, 8 * (size_t)n) | ensure(ctx, ctx->d_probe_b, 8 * (size_t)n) | ensure(ctx, ctx->d_probe_c, 8 * (size_t)n);
    if (rc) return ATMRT_ERR_CUDA;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_a.p, lat, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_b.p, lon, 8 * (size_t)n, cudaMemcpyHostToDevice, s));
    k_get_elev<<<(n + 255) / 256, 256, 0, s>>>(ctx->terrain, (const double*)ctx->d_probe_a.p, (const double*)ctx->d_probe_b.p, n, (double*)ctx->d_probe_c.p);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(elev, ctx->d_probe_c.p, 8 * (size_t)n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_read_tile(atmrt_ctx* ctx, int tile_index, int16_t* posts) {
    if (!ctx || !posts) return fail(ctx, ATMRT_ERR_INVALID, "read_tile: NULL argument");
    if (!ctx->has_terrain) return fail(ctx, ATMRT_ERR_STATE, "read_tile before set_terrain");
    if (tile_index < 0 || tile_index >= (int)ctx->tiles_host.size()) return fail(ctx, ATMRT_ERR_INVALID, "tile index out of range");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const DevTile& t = ctx->tiles_host[tile_index];
    size_t n = (size_t)t.nlon * t.nlat;
    DevBuf raw;
    int rc = ensure(ctx, raw, sizeof(int16_t) * n);
    if (rc) return rc;
    cudaStream_t s = ctx->s_main;
    k_untile<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ctx->terrain.posts, (int16_t*)raw.p, t);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(posts, raw.p, sizeof(int16_t) * n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    release(raw);
    if (e != cudaSuccess) return fail(ctx, ATMRT_ERR_CUDA, std::string("read_tile: ") + cudaGetErrorString(e));
    return 0;
}

int atmrt_set_params(atmrt_ctx* ctx, const atmrt_params* params) {
    if (!ctx || !params) return fail(ctx, ATMRT_ERR_INVALID, "set_params: NULL argument");
    int rc = validate_params(ctx, *params);
    if (rc) return rc;
    DevAtmosphere tmp;
    rc = lower_atmosphere(ctx, params->atmosphere, params->wavelength, &tmp);
    if (rc) return rc;
    ctx->params = *params;
    ctx->has_params = true;
    ctx->rendered = false;
    return 0;
}

int atmrt_set_objects(atmrt_ctx* ctx, const atmrt_object* objects, int nobjects, const uint8_t* const* rgba_textures) {
    if (!ctx || nobjects < 0 || (nobjects > 0 && !objects)) return fail(ctx, ATMRT_ERR_INVALID, "set_objects: bad argument");
    if (nobjects > ATMRT_MAX_OBJECTS) return fail(ctx, ATMRT_ERR_INVALID, "too many objects (objects_close is a 64-bit mask)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < nobjects; ++i) {
        const atmrt_object& o = objects[i];
        if (o.kind == ATMRT_OBJECT_BILLBOARD) {
            if (!rgba_textures || !rgba_textures[i] || o.texture_width < 2 || o.texture_height < 2)
                return fail(ctx, ATMRT_ERR_INVALID, "billboard needs an RGBA texture of at least 2x2 texels");
        } else if (o.kind != ATMRT_OBJECT_FRUSTUM) {
            return fail(ctx, ATMRT_ERR_INVALID, "unknown object kind");
        }
    }
    for (DevBuf& b : ctx->textures) release(b);
    ctx->textures.assign(nobjects, DevBuf{});
    ctx->objects.assign(objects, objects + nobjects);
    for (int i = 0; i < nobjects; ++i) {
        if (objects[i].kind != ATMRT_OBJECT_BILLBOARD) continue;
        size_t bytes = (size_t)objects[i].texture_width * objects[i].texture_height * 4;
        int rc = ensure(ctx, ctx->textures[i], bytes);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaMemcpy(ctx->textures[i].p, rgba_textures[i], bytes, cudaMemcpyHostToDevice));
    }
    ctx->rendered = false;
    return 0;
}

int atmrt_set_march_mode(atmrt_ctx* ctx, int mode) {
    if (!ctx || mode < 0 || mode > 2) return fail(ctx, ATMRT_ERR_INVALID, "march mode must be 0, 1 or 2");
    ctx->march_mode = mode == 1 ? 1 : 0;
    ctx->sweep_enabled = mode == 0;
    return 0;
}

int atmrt_set_sweep_bands(atmrt_ctx* ctx, int bands) {
    if (!ctx || bands < -1) return fail(ctx, ATMRT_ERR_INVALID, "sweep bands must be >= -1 (0: automatic, -1: split below the horizon)");
    ctx->sweep_bands = bands;
    return 0;
}

int atmrt_set_path_mode(atmrt_ctx* ctx, int mode) {
    if (!ctx) return ATMRT_ERR_INVALID;
    ctx->path_mode = mode == 1 || mode == 2 ? mode : 0;
    return 0;
}

int atmrt_render_device(atmrt_ctx* ctx, void* rgb_dev, void* meta_dev, void* steps_dev, atmrt_stats* stats, void* stream) {
    if (!ctx) return ATMRT_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = prepare_render(ctx);
    if (rc) return rc;
    cudaStream_t main = stream ? (cudaStream_t)stream : ctx->s_main;
    // every render of a context works in the context's caches, flags and counters: a render on another stream than the one
    // before it starts behind that one
    if (ctx->last_recorded) CUDA_TRY(ctx, cudaStreamWaitEvent(main, ctx->ev_last, 0));
    RenderTargets rt;
    rt.rgb = (unsigned char*)rgb_dev, rt.meta = (atmrt_meta*)meta_dev, rt.steps = (int*)steps_dev;
    rc = launch_render(ctx, rt, main);
    if (rc) return rc;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_last, main));
    ctx->last_recorded = true;
    if (stats) {
        CUDA_TRY(ctx, cudaStreamSynchronize(main));
        return collect_stats(ctx, stats);
    }
    return 0;
}

// Render this context's column block into host memory; the block may be a sub-rectangle of a wider host image
// (host_width pixels per row, starting at column host_x0). Synchronous.
static int render_to_host(atmrt_ctx* ctx, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, int host_width, int host_x0, atmrt_stats* stats) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = prepare_render(ctx);
    if (rc) return rc;
    const DevScene& S = ctx->scene;
    const int wl = S.x1 - S.x0;
    const size_t npix = (size_t)wl * S.height;
    if (rgb && (rc = ensure(ctx, ctx->d_rgb, npix * 3))) return rc;
    if (meta && (rc = ensure(ctx, ctx->d_meta, npix * sizeof(atmrt_meta)))) return rc;
    if (steps && (rc = ensure(ctx, ctx->d_steps, npix * sizeof(int)))) return rc;
    RenderTargets rt;
    rt.rgb = rgb ? (unsigned char*)ctx->d_rgb.p : nullptr;
    rt.meta = meta ? (atmrt_meta*)ctx->d_meta.p : nullptr;
    rt.steps = steps ? (int*)ctx->d_steps.p : nullptr;
    rt.host_rgb = rgb, rt.host_meta = meta, rt.host_steps = steps;
    rt.host_width = host_width, rt.host_x0 = host_x0;
    cudaStream_t main = ctx->s_main;
    rc = launch_render(ctx, rt, main);
    if (rc) return rc;
    bool copy_all = !rt.host_copied;
    if (rt.host_copied) {  // bands went out behind the horizon sweep: valid unless a device-side fallback repainted pixels afterwards
        unsigned flags[2] = {0, 0};
        CUDA_TRY(ctx, cudaMemcpyAsync(flags, ctx->d_sweep_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, main));
        CUDA_TRY(ctx, cudaStreamSynchronize(main));
        copy_all = flags[0] != 0 || flags[1] != 0;
    }
    if (copy_all) {
        if ((rc = copy_rows_to_host(ctx, rt, wl, 0, S.height, main))) return rc;
        CUDA_TRY(ctx, cudaStreamSynchronize(main));
    }
    return collect_stats(ctx, stats);
}

int atmrt_render(atmrt_ctx* ctx, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, atmrt_stats* stats) {
    if (!ctx) return ATMRT_ERR_INVALID;
    return render_to_host(ctx, rgb, meta, steps, 0, 0, stats);
}

int atmrt_render_trace(atmrt_ctx* ctx, atmrt_trace_point* points, int32_t* counts, int max_points) {
    if (!ctx || !counts || max_points < 0 || (max_points > 0 && !points)) return fail(ctx, ATMRT_ERR_INVALID, "render_trace: bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = prepare_render(ctx);
    if (rc) return rc;
    const DevScene& S = ctx->scene;
    const size_t npix = (size_t)(S.x1 - S.x0) * S.height;
    if ((rc = ensure(ctx, ctx->d_points, npix * sizeof(atmrt_trace_point) * (size_t)std::max(max_points, 1)))) return rc;
    if ((rc = ensure(ctx, ctx->d_counts, npix * sizeof(int)))) return rc;
    RenderTargets rt;
    rt.points = (atmrt_trace_point*)ctx->d_points.p;
    rt.counts = (int*)ctx->d_counts.p;
    rt.max_points = max_points;
    cudaStream_t main = ctx->s_main;
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_points.p, 0, npix * sizeof(atmrt_trace_point) * (size_t)std::max(max_points, 1), main));
    rc = launch_render(ctx, rt, main);
    if (rc) return rc;
    if (max_points > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(points, ctx->d_points.p, npix * sizeof(atmrt_trace_point) * (size_t)max_points, cudaMemcpyDeviceToHost, main));
    CUDA_TRY(ctx, cudaMemcpyAsync(counts, ctx->d_counts.p, npix * sizeof(int), cudaMemcpyDeviceToHost, main));
    CUDA_TRY(ctx, cudaStreamSynchronize(main));
    return 0;
}

int atmrt_get_terrain_profile(atmrt_ctx* ctx, int x, int capacity, double* lat, double* lon, double* elev, double* normal,
                              uint64_t* objects_close, int* n) {
    if (!ctx || !n) return fail(ctx, ATMRT_ERR_INVALID, "get_terrain_profile: NULL argument");
    if (!ctx->rendered) return fail(ctx, ATMRT_ERR_STATE, "get_terrain_profile before a render");
    if (ctx->scene.generator != ATMRT_GENERATOR_FAST) return fail(ctx, ATMRT_ERR_STATE, "get_terrain_profile: only the Fast generator keeps a terrain cache");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const DevScene& S = ctx->scene;
    if (x < 0 || x >= S.x1 - S.x0) return fail(ctx, ATMRT_ERR_INVALID, "column out of range");
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    *n = S.n_t;
    int m = std::min(capacity, S.n_t);
    if (m <= 0) return 0;
    size_t off = (size_t)x * S.n_pad;
    const DevBuffers& B = ctx->buf;
    if (lat) CUDA_TRY(ctx, cudaMemcpy(lat, B.t_lat + off, 8 * (size_t)m, cudaMemcpyDeviceToHost));
    if (lon) CUDA_TRY(ctx, cudaMemcpy(lon, B.t_lon + off, 8 * (size_t)m, cudaMemcpyDeviceToHost));
    if (elev) CUDA_TRY(ctx, cudaMemcpy(elev, B.t_elev + off, 8 * (size_t)m, cudaMemcpyDeviceToHost));
    if (normal) {  // TerrainData::normal is not cached (kernels.cuh: sample_normal): evaluate it for this column
        if (ensure(ctx, ctx->d_probe_a, 24 * (size_t)S.n_t)) return ATMRT_ERR_CUDA;
        k_profile_normals<<<(S.n_t + 127) / 128, 128, 0, ctx->s_main>>>(S, ctx->terrain, B, x, (double*)ctx->d_probe_a.p);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_main));
        CUDA_TRY(ctx, cudaMemcpy(normal, ctx->d_probe_a.p, 24 * (size_t)m, cudaMemcpyDeviceToHost));
    }
    if (objects_close) {
        if (B.t_close)
            CUDA_TRY(ctx, cudaMemcpy(objects_close, B.t_close + off, 8 * (size_t)m, cudaMemcpyDeviceToHost));
        else
            memset(objects_close, 0, 8 * (size_t)m);
    }
    return 0;
}

int atmrt_get_path(atmrt_ctx* ctx, int y, int capacity, double* dist, double* elev, double* path_length, int* n) {
    if (!ctx || !n) return fail(ctx, ATMRT_ERR_INVALID, "get_path: NULL argument");
    if (!ctx->rendered) return fail(ctx, ATMRT_ERR_STATE, "get_path before a render");
    if (ctx->scene.generator != ATMRT_GENERATOR_FAST) return fail(ctx, ATMRT_ERR_STATE, "get_path: only the Fast generator keeps a path cache");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const DevScene& S = ctx->scene;
    if (y < 0 || y >= S.height) return fail(ctx, ATMRT_ERR_INVALID, "row out of range");
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    int len = 0;
    CUDA_TRY(ctx, cudaMemcpy(&len, ctx->buf.p_n + y, sizeof(int), cudaMemcpyDeviceToHost));
    *n = len;
    int m = std::min(capacity, len);
    if (m <= 0) return 0;
    // the cache is [row / 4][k][row % 4]: gather row y with a strided 2-D copy (pitch = one row group)
    const size_t pitch = (size_t)PATH_ROWS * sizeof(double), off = path_index(S.n_t, 0, y);
    if (dist) memcpy(dist, ctx->path_x.data(), sizeof(double) * (size_t)m);  // row-independent (prepare_render)
    if (elev) CUDA_TRY(ctx, cudaMemcpy2D(elev, sizeof(double), ctx->buf.p_elev + off, pitch, sizeof(double), (size_t)m, cudaMemcpyDeviceToHost));
    if (path_length)
        CUDA_TRY(ctx, cudaMemcpy2D(path_length, sizeof(double), ctx->buf.p_len + off, pitch, sizeof(double), (size_t)m, cudaMemcpyDeviceToHost));
    return 0;
}

int atmrt_stage_times(atmrt_ctx* ctx, atmrt_stage_ms* out) {
    if (!ctx || !out) return fail(ctx, ATMRT_ERR_INVALID, "stage_times: NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    double a = 0, b = 0, c = 0, t = 0;
    for (size_t i = 0; i < ctx->ring_used; ++i) {  // the most recent renders (at most RING of them)
        const atmrt_ctx::StageEvents& E = ctx->ring[(ctx->ring_next - 1 - i) % atmrt_ctx::RING];
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, E.a0, E.a1));
        a += ms;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, E.b0, E.b1));
        b += ms;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, E.c0, E.c1));
        c += ms;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, E.t0, E.t1));
        t += ms;
    }
    out->renders = (int32_t)ctx->ring_used;
    if (ctx->ring_used > 0) {
        const double n = (double)ctx->ring_used;
        out->ms_terrain = a / n, out->ms_paths = b / n, out->ms_march = c / n, out->ms_total = t / n;
    }
    ctx->ring_used = 0;
    return 0;
}

const char* atmrt_kernel_name(int i) {
    static const char* names[ATMRT_KERNEL_COUNT] = {"k_terrain_profile", "k_ray_paths_macro", "k_sweep_bits", "k_hit_normals",
                                                    "k_shade_tiles",     "k_march",           "k_rectilinear"};
    return i >= 0 && i < ATMRT_KERNEL_COUNT ? names[i] : "";
}

int atmrt_kernel_times(atmrt_ctx* ctx, atmrt_kernel_ms* out) {
    if (!ctx || !out) return fail(ctx, ATMRT_ERR_INVALID, "kernel_times: NULL argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    memset(out, 0, sizeof(*out));
    for (size_t i = 0; i < ctx->ring_used; ++i) {
        const atmrt_ctx::StageEvents& E = ctx->ring[(ctx->ring_next - 1 - i) % atmrt_ctx::RING];
        for (int k = 0; k < ATMRT_KERNEL_COUNT; ++k) {
            if (!(E.kmask >> k & 1u)) continue;
            float ms = 0.f;
            CUDA_TRY(ctx, cudaEventElapsedTime(&ms, E.k0[k], E.k1[k]));
            out->ms[k] += ms;
            out->renders_with[k] += 1;
        }
    }
    for (int k = 0; k < ATMRT_KERNEL_COUNT; ++k)
        if (out->renders_with[k] > 0) out->ms[k] /= (double)out->renders_with[k];
    out->renders = (int32_t)ctx->ring_used;
    return 0;
}

int atmrt_atmosphere_probe(atmrt_ctx* ctx, const double* h, int n, double* temperature, double* pressure, double* refractive_index) {
    if (!ctx || !h || n < 0) return fail(ctx, ATMRT_ERR_INVALID, "atmosphere_probe: bad argument");
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "atmosphere_probe before set_params");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevScene S{};
    int rc = lower_atmosphere(ctx, ctx->params.atmosphere, ctx->params.wavelength, &S.atm);
    if (rc) return rc;
    size_t bytes = sizeof(double) * (size_t)std::max(n, 1);
    if (ensure(ctx, ctx->d_probe_a, bytes) || ensure(ctx, ctx->d_probe_b, bytes) || ensure(ctx, ctx->d_probe_c, bytes) || ensure(ctx, ctx->d_probe_d, bytes))
        return ATMRT_ERR_CUDA;
    if (n == 0) return 0;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_a.p, h, bytes, cudaMemcpyHostToDevice, s));
    k_atm_probe<<<(n + 127) / 128, 128, 0, s>>>(S, (const double*)ctx->d_probe_a.p, n, (double*)ctx->d_probe_b.p, (double*)ctx->d_probe_c.p,
                                                  (double*)ctx->d_probe_d.p);
    CUDA_TRY(ctx, cudaGetLastError());
    if (temperature) CUDA_TRY(ctx, cudaMemcpyAsync(temperature, ctx->d_probe_b.p, bytes, cudaMemcpyDeviceToHost, s));
    if (pressure) CUDA_TRY(ctx, cudaMemcpyAsync(pressure, ctx->d_probe_c.p, bytes, cudaMemcpyDeviceToHost, s));
    if (refractive_index) CUDA_TRY(ctx, cudaMemcpyAsync(refractive_index, ctx->d_probe_d.p, bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_refraction_table(const atmrt_atmosphere_def* def, double wavelength, double* cells, int capacity, int* ncells, int* ncoef,
                           double* base, double* cell_height, int* cells_served, int* npieces) {
    if (!def) return ATMRT_ERR_INVALID;
    if (ncells) *ncells = ATM_CELLS;
    if (ncoef) *ncoef = ATM_FIELDS;
    if (base) *base = ATM_BASE;
    if (cell_height) *cell_height = ATM_CELL;
    DevAtmosphere a;
    int rc = lower_atmosphere(nullptr, *def, wavelength, &a);
    if (rc) return rc;
    std::vector<double> t;
    std::vector<DevGPiece> pieces;
    int served = 0;
    build_g_table(a, wavelength, t, pieces, &served);
    if (npieces) *npieces = (int)pieces.size();
    if (cells_served) *cells_served = served;
    if (cells) {
        if (capacity < (int)t.size()) return ATMRT_ERR_INVALID;
        memcpy(cells, t.data(), sizeof(double) * t.size());
    }
    return 0;
}

int atmrt_refraction_probe(atmrt_ctx* ctx, const double* h, int n, int with_pieces, double* g_table, double* g_libm, int* cells_served) {
    if (!ctx || !h || n < 0) return fail(ctx, ATMRT_ERR_INVALID, "refraction_probe: bad argument");
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "refraction_probe before set_params");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevScene S{};
    int rc = lower_atmosphere(ctx, ctx->params.atmosphere, ctx->params.wavelength, &S.atm);
    if (rc) return rc;
    std::vector<double> cells;
    std::vector<DevGPiece> pieces;
    int served = 0;
    build_g_table(S.atm, ctx->params.wavelength, cells, pieces, &served);
    if (cells_served) *cells_served = served;
    const size_t cell_bytes = sizeof(double) * cells.size();
    size_t bytes = sizeof(double) * (size_t)std::max(n, 1);
    if (ensure(ctx, ctx->d_probe_a, bytes) || ensure(ctx, ctx->d_probe_b, bytes) || ensure(ctx, ctx->d_probe_c, bytes) ||
        ensure(ctx, ctx->d_probe_d, cell_bytes + sizeof(DevGPiece) * ATM_MAX_PIECES))
        return ATMRT_ERR_CUDA;
    if (n == 0) return 0;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_a.p, h, bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_d.p, cells.data(), cell_bytes, cudaMemcpyHostToDevice, s));
    if (!pieces.empty())
        CUDA_TRY(ctx, cudaMemcpyAsync((char*)ctx->d_probe_d.p + cell_bytes, pieces.data(), sizeof(DevGPiece) * pieces.size(), cudaMemcpyHostToDevice, s));
    k_refraction_probe<<<(n + 127) / 128, 128, 0, s>>>(S, (const double*)ctx->d_probe_d.p, (const double*)ctx->d_probe_a.p, n, (double*)ctx->d_probe_b.p,
                                                         (double*)ctx->d_probe_c.p, with_pieces ? (const DevGPiece*)((const char*)ctx->d_probe_d.p + cell_bytes) : nullptr,
                                                         (int)pieces.size());
    CUDA_TRY(ctx, cudaGetLastError());
    if (g_table) CUDA_TRY(ctx, cudaMemcpyAsync(g_table, ctx->d_probe_b.p, bytes, cudaMemcpyDeviceToHost, s));
    if (g_libm) CUDA_TRY(ctx, cudaMemcpyAsync(g_libm, ctx->d_probe_c.p, bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_pixel_angles(atmrt_ctx* ctx, double* elevation_angle, double* azimuth) {
    if (!ctx) return ATMRT_ERR_INVALID;
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "pixel_angles before set_params");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const atmrt_params& p = ctx->params;
    DevScene S{};  // only the frame and the image geometry are read
    S.direction = p.direction, S.tilt = p.tilt, S.fov = p.fov;
    S.width = p.width, S.height = p.height, S.x0 = p.x0, S.x1 = p.x1, S.generator = p.generator;
    const size_t wl = (size_t)(p.x1 - p.x0), bytes = sizeof(double) * wl * (size_t)p.height;
    if ((elevation_angle && ensure(ctx, ctx->d_probe_a, bytes)) || (azimuth && ensure(ctx, ctx->d_probe_b, bytes))) return ATMRT_ERR_CUDA;
    cudaStream_t s = ctx->s_main;
    if (p.generator == ATMRT_GENERATOR_INTERPOLATING_RECTILINEAR) {
        double elev_step, dir_step;
        int rc = interpolating_steps(ctx, S, s, &elev_step, &dir_step);
        if (rc) return rc;
        k_interp_angles<<<dim3((unsigned)((wl + 127) / 128), (unsigned)p.height), 128, 0, s>>>(S, elev_step, dir_step, elevation_angle ? (double*)ctx->d_probe_a.p : nullptr,
                                                                                             azimuth ? (double*)ctx->d_probe_b.p : nullptr);
    } else
    k_pixel_angles<<<dim3((unsigned)((wl + 127) / 128), (unsigned)p.height), 128, 0, s>>>(S, elevation_angle ? (double*)ctx->d_probe_a.p : nullptr,
                                                                                        azimuth ? (double*)ctx->d_probe_b.p : nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    if (elevation_angle) CUDA_TRY(ctx, cudaMemcpyAsync(elevation_angle, ctx->d_probe_a.p, bytes, cudaMemcpyDeviceToHost, s));
    if (azimuth) CUDA_TRY(ctx, cudaMemcpyAsync(azimuth, ctx->d_probe_b.p, bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_ray_paths(atmrt_ctx* ctx, double start_h, const double* angles_deg, int n, double ray_step, int nsteps, double* x, double* h) {
    if (!ctx || !angles_deg || n < 0 || nsteps < 0 || !(ray_step > 0.0)) return fail(ctx, ATMRT_ERR_INVALID, "ray_paths: bad argument");
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "ray_paths before set_params");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const atmrt_params& p = ctx->params;
    DevScene S{};
    int rc = lower_atmosphere(ctx, p.atmosphere, p.wavelength, &S.atm);
    if (rc) return rc;
    // EarthModel::to_shape (earth_model/mod.rs:95-112), as prepare_render lowers it
    S.flat = p.earth_model == ATMRT_EARTH_FLAT_DISTORTED || p.earth_model == ATMRT_EARTH_AZIMUTHAL_EQUIDISTANT || p.earth_model == ATMRT_EARTH_OBSERVER_AE;
    S.radius = p.earth_model == ATMRT_EARTH_ELLIPSOID ? (2.0 * p.radius + p.ellipsoid_b) / 3.0 : p.radius;
    // RayState::x: the stepper's own running sum (phi += step / R, x = phi * R; flat: x += step), the same for every ray
    if (x) {
        const double d = S.flat ? ray_step : ray_step / S.radius;
        double t = 0.0;
        for (int k = 0; k < nsteps; ++k) {
            t += d;
            x[k] = S.flat ? t : t * S.radius;
        }
    }
    if (n == 0 || nsteps == 0 || !h) return 0;
    std::vector<double> cells;
    std::vector<DevGPiece> pieces;
    build_g_table(S.atm, p.wavelength, cells, pieces, nullptr);
    const size_t cell_bytes = sizeof(double) * cells.size(), out_bytes = sizeof(double) * (size_t)n * nsteps;
    if (ensure(ctx, ctx->d_probe_a, sizeof(double) * (size_t)n) || ensure(ctx, ctx->d_probe_b, out_bytes) ||
        ensure(ctx, ctx->d_probe_d, cell_bytes + sizeof(DevGPiece) * ATM_MAX_PIECES))
        return ATMRT_ERR_CUDA;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_a.p, angles_deg, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_d.p, cells.data(), cell_bytes, cudaMemcpyHostToDevice, s));
    if (!pieces.empty())
        CUDA_TRY(ctx, cudaMemcpyAsync((char*)ctx->d_probe_d.p + cell_bytes, pieces.data(), sizeof(DevGPiece) * pieces.size(), cudaMemcpyHostToDevice, s));
    DevBuffers B{};
    B.atm_cells = (const double*)ctx->d_probe_d.p;
    B.atm_pieces = (const DevGPiece*)((const char*)ctx->d_probe_d.p + cell_bytes);
    B.n_atm_pieces = (int)pieces.size();
    const int libm = ctx->path_mode == 1 ? 1 : 0;
    if (S.flat)
        k_ray_path_probe<true><<<(n + 127) / 128, 128, 0, s>>>(S, B, start_h, (const double*)ctx->d_probe_a.p, n, ray_step, nsteps, libm, (double*)ctx->d_probe_b.p);
    else
        k_ray_path_probe<false><<<(n + 127) / 128, 128, 0, s>>>(S, B, start_h, (const double*)ctx->d_probe_a.p, n, ray_step, nsteps, libm, (double*)ctx->d_probe_b.p);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(h, ctx->d_probe_b.p, out_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_elev_profile(atmrt_ctx* ctx, double azimuth, const double* dist, int n, double* lat, double* lon, double* elev) {
    if (!ctx || !dist || !elev || n < 0) return fail(ctx, ATMRT_ERR_INVALID, "elev_profile: bad argument");
    if (!ctx->has_terrain) return fail(ctx, ATMRT_ERR_STATE, "elev_profile before set_terrain");
    if (!ctx->has_params) return fail(ctx, ATMRT_ERR_STATE, "elev_profile before set_params");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    int rc = prepare_render(ctx);  // lowers the earth model into ctx->scene (and sizes the render buffers)
    if (rc) return rc;
    const size_t bytes = sizeof(double) * (size_t)n;
    if (ensure(ctx, ctx->d_probe_a, bytes) || ensure(ctx, ctx->d_probe_b, bytes) || ensure(ctx, ctx->d_probe_c, bytes) || ensure(ctx, ctx->d_probe_d, bytes))
        return ATMRT_ERR_CUDA;
    cudaStream_t s = ctx->s_main;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_probe_a.p, dist, bytes, cudaMemcpyHostToDevice, s));
    k_elev_profile<<<(n + 127) / 128, 128, 0, s>>>(ctx->scene, ctx->terrain, azimuth, (const double*)ctx->d_probe_a.p, n, (double*)ctx->d_probe_b.p,
                                                   (double*)ctx->d_probe_c.p, (double*)ctx->d_probe_d.p);
    CUDA_TRY(ctx, cudaGetLastError());
    if (lat) CUDA_TRY(ctx, cudaMemcpyAsync(lat, ctx->d_probe_b.p, bytes, cudaMemcpyDeviceToHost, s));
    if (lon) CUDA_TRY(ctx, cudaMemcpyAsync(lon, ctx->d_probe_c.p, bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaMemcpyAsync(elev, ctx->d_probe_d.p, bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ctx, cudaStreamSynchronize(s));
    return 0;
}

int atmrt_observer_altitude(atmrt_ctx* ctx, double* alt) {
    if (!ctx || !alt) return fail(ctx, ATMRT_ERR_INVALID, "observer_altitude: NULL argument");
    if (!ctx->rendered) return fail(ctx, ATMRT_ERR_STATE, "observer_altitude before a render");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    CUDA_TRY(ctx, cudaMemcpy(alt, ctx->buf.obs_alt, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int atmrt_fp64_peak(atmrt_ctx* ctx, double* gflops, double* dadd_ginstr) {
    if (!ctx) return ATMRT_ERR_INVALID;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->d_probe_a, 64);
    if (rc) return rc;
    cudaStream_t s = ctx->s_main;
    const int iters = 4096, blocks = ctx->num_sms * 8, threads = 256;
    cudaEvent_t e0 = ctx->t_0, e1 = ctx->t_1;
    double best[2] = {0.0, 0.0};
    for (int variant = 0; variant < 2; ++variant) {
        for (int rep = 0; rep < 5; ++rep) {
            CUDA_TRY(ctx, cudaEventRecord(e0, s));
            if (variant == 0)
                k_fp64_peak<true><<<blocks, threads, 0, s>>>((double*)ctx->d_probe_a.p, iters, 1.0000001, 1e-9);
            else
                k_fp64_peak<false><<<blocks, threads, 0, s>>>((double*)ctx->d_probe_a.p, iters, 1.0000001, 1e-9);
            CUDA_TRY(ctx, cudaEventRecord(e1, s));
            CUDA_TRY(ctx, cudaStreamSynchronize(s));
            float ms = 0.f;
            CUDA_TRY(ctx, cudaEventElapsedTime(&ms, e0, e1));
            double instr = (double)blocks * threads * (double)iters * 8.0;
            double rate = instr / (ms * 1e-3) / 1e9;  // G thread-instr/s
            if (rep > 0 && rate > best[variant]) best[variant] = rate;
        }
    }
    if (gflops) *gflops = best[0] * 2.0;
    if (dadd_ginstr) *dadd_ginstr = best[1];
    return 0;
}


// ---------------------------------------------------------------------------------------------
// Multi-GPU: one panorama over the GPUs of a box, in ONE process. The panorama shards by contiguous column blocks
// (pixels are independent: fast.rs:52-92), one context per GPU, one host thread per context. Terrain: every GPU
// uploads and retiles a contiguous SLICE of the tiles over its own PCIe link and pulls the other slices from its peers
// over NVLink (cudaMemcpyPeerAsync), so the host-to-device traffic of the frame is one copy of the terrain, not one per
// GPU. Image: every GPU copies its finished row bands straight into its columns of the host's row-major image (strided
// 2-D copies over its own PCIe link); nothing funnels through one GPU. No collective in the march.
// ---------------------------------------------------------------------------------------------
int atmrt_group_create(const int* devices, int n, atmrt_group** out) {
    if (!out || n < 1 || n > 64) return gfail(nullptr, ATMRT_ERR_INVALID, "group_create: bad argument");
    *out = nullptr;
    atmrt_group* g = new atmrt_group();
    for (int i = 0; i < n; ++i) {
        atmrt_ctx* c = nullptr;
        const int rc = atmrt_create(devices ? devices[i] : i, &c);
        if (rc) {
            for (atmrt_ctx* x : g->ctx) atmrt_destroy(x);
            delete g;
            return rc;  // (g_create_error is set)
        }
        g->ctx.push_back(c);
    }
    // peer access both ways between every pair (NVLink / NVSwitch on a B200 box); without it the peer copies are staged
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(g->ctx[i]->device);
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g->ctx[i]->device, g->ctx[j]->device);
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(g->ctx[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            }
        }
        cudaEvent_t ev = nullptr;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        g->ev_slice.push_back(ev);
    }
    g->packed.assign(n, nullptr);
    *out = g;
    return 0;
}

void atmrt_group_destroy(atmrt_group* g) {
    if (!g) return;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        cudaDeviceSynchronize();
        if (g->packed[i]) cudaFree(g->packed[i]);
        if (g->ev_slice[i]) cudaEventDestroy(g->ev_slice[i]);
        atmrt_destroy(g->ctx[i]);
    }
    delete g;
}

const char* atmrt_group_last_error(const atmrt_group* g) { return g ? g->err.c_str() : g_create_error.c_str(); }
atmrt_ctx* atmrt_group_context(atmrt_group* g, int i) { return g && i >= 0 && i < (int)g->ctx.size() ? g->ctx[(size_t)i] : nullptr; }

int atmrt_group_size(const atmrt_group* g) { return g ? (int)g->ctx.size() : 0; }

int atmrt_group_column_block(const atmrt_group* g, int width, int i, int* x0, int* x1) {
    if (!g || i < 0 || i >= (int)g->ctx.size() || !x0 || !x1) return ATMRT_ERR_INVALID;
    const long long n = (long long)g->ctx.size();
    *x0 = (int)((long long)i * width / n), *x1 = (int)((long long)(i + 1) * width / n);
    return 0;
}

// The terrain of a group: every GPU uploads a contiguous slice of the tiles over its own link (s_t), retiles every tile as it
// lands (s_r, behind the tile's copy: the next tile's copy runs meanwhile) and then pulls the other slices from its peers.
// `wait`: return when the terrain is in place on every GPU; otherwise ev_terrain of every context says when.
static int group_upload_terrain(atmrt_group* g, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, bool wait) {
    const int n = (int)g->ctx.size();
    TerrainLayout L;
    int rc = make_layout(g->ctx[0], tiles, ntiles, &L);
    if (rc) return gfail(g, rc, g->ctx[0]->err);
    // contiguous slices of tiles per GPU: slice i = tiles [i T / n, (i + 1) T / n), one contiguous range of packed posts
    auto first_tile = [&](int i) { return (int)((long long)i * ntiles / n); };
    auto post_byte = [&](int t) { return L.off_posts + sizeof(int16_t) * (size_t)(t < ntiles ? L.tiles[t].post_offset : (long long)((L.total - L.off_posts) / sizeof(int16_t))); };
    rc = group_parallel(g, [&](int i) -> int {
        atmrt_ctx* ctx = g->ctx[i];
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        if (!g->packed[i] || g->packed_cap < L.total) {
            if (g->packed[i]) CUDA_TRY(ctx, cudaFree(g->packed[i]));
            g->packed[i] = nullptr;
            CUDA_TRY(ctx, cudaMalloc(&g->packed[i], L.total));
        }
        char* base = (char*)g->packed[i];
        cudaStream_t s = ctx->s_t, r = ctx->s_r;
        const int t0 = first_tile(i), t1 = first_tile(i + 1);
        if (ntiles > 0) CUDA_TRY(ctx, cudaMemcpyAsync(base + L.off_tiles, L.tiles.data(), sizeof(DevTile) * ntiles, cudaMemcpyHostToDevice, s));
        if (!L.lookup.empty()) CUDA_TRY(ctx, cudaMemcpyAsync(base + L.off_lookup, L.lookup.data(), sizeof(int) * L.lookup.size(), cudaMemcpyHostToDevice, s));
        std::vector<size_t> offs(t1 - t0 + 1, 0);
        for (int t = t0; t < t1; ++t) offs[t - t0 + 1] = offs[t - t0] + align_up(sizeof(int16_t) * (size_t)tiles[t].nlon * tiles[t].nlat, 256);
        int e = ensure(ctx, ctx->d_stage, std::max<size_t>(offs[t1 - t0], 256));
        if (e) return e;
        if (t1 > t0) CUDA_TRY(ctx, cudaMemsetAsync(base + post_byte(t0), 0, post_byte(t1) - post_byte(t0), s));  // the padding of the edge micro-tiles
        for (int t = t0; t < t1; ++t) {
            const size_t np = (size_t)tiles[t].nlon * tiles[t].nlat;
            CUDA_TRY(ctx, cudaMemcpyAsync((char*)ctx->d_stage.p + offs[t - t0], posts[t], sizeof(int16_t) * np, cudaMemcpyHostToDevice, s));
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_tile, s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(r, ctx->ev_tile, 0));  // (the wait captures this record: the event is reused tile after tile)
            k_retile<<<(unsigned)((np + 255) / 256), 256, 0, r>>>((const int16_t*)((const char*)ctx->d_stage.p + offs[t - t0]), (int16_t*)(base + L.off_posts), L.tiles[t]);
        }
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_tile, s));  // the descriptors and the lookup table too, when the slice is empty
        CUDA_TRY(ctx, cudaStreamWaitEvent(r, ctx->ev_tile, 0));
        CUDA_TRY(ctx, cudaEventRecord(g->ev_slice[i], r));
        return 0;
    });
    if (rc) return rc;
    g->packed_cap = std::max(g->packed_cap, L.total);
    // all-gather of the slices over the peer links: every GPU pulls the slices it does not own
    return group_parallel(g, [&](int i) -> int {
        atmrt_ctx* ctx = g->ctx[i];
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        cudaStream_t r = ctx->s_r;
        for (int k = 1; k < n; ++k) {
            const int o = (i + k) % n;  // staggered: no two GPUs start on the same peer
            const size_t b0 = post_byte(first_tile(o)), b1 = post_byte(first_tile(o + 1));
            if (b1 <= b0) continue;
            CUDA_TRY(ctx, cudaStreamWaitEvent(r, g->ev_slice[o], 0));
            CUDA_TRY(ctx, cudaMemcpyPeerAsync((char*)g->packed[i] + b0, ctx->device, (const char*)g->packed[o] + b0, g->ctx[o]->device, b1 - b0, r));
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_terrain, r));
        if (wait) CUDA_TRY(ctx, cudaStreamSynchronize(r));
        ctx->terrain_in_flight = !wait;
        return atmrt_bind_terrain(ctx, tiles, ntiles, g->packed[i]);
    });
}

int atmrt_group_set_terrain(atmrt_group* g, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts) {
    if (!g || ntiles < 0 || (ntiles > 0 && (!tiles || !posts))) return gfail(g, ATMRT_ERR_INVALID, "group_set_terrain: bad argument");
    return group_upload_terrain(g, tiles, ntiles, posts, true);
}

int atmrt_group_set_params(atmrt_group* g, const atmrt_params* params) {
    if (!g || !params) return gfail(g, ATMRT_ERR_INVALID, "group_set_params: NULL argument");
    if (params->width < (int)g->ctx.size()) return gfail(g, ATMRT_ERR_INVALID, "group_set_params: fewer columns than GPUs");
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        atmrt_params p = *params;  // the group shards [0, width): x0 / x1 of the argument are ignored
        atmrt_group_column_block(g, params->width, (int)i, &p.x0, &p.x1);
        const int rc = atmrt_set_params(g->ctx[i], &p);
        if (rc) return gfail(g, rc, g->ctx[i]->err);
    }
    g->params = *params;
    g->has_params = true;
    return 0;
}

int atmrt_group_set_objects(atmrt_group* g, const atmrt_object* objects, int nobjects, const uint8_t* const* rgba_textures) {
    if (!g) return ATMRT_ERR_INVALID;
    for (atmrt_ctx* c : g->ctx) {
        const int rc = atmrt_set_objects(c, objects, nobjects, rgba_textures);
        if (rc) return gfail(g, rc, c->err);
    }
    return 0;
}

int atmrt_group_render(atmrt_group* g, uint8_t* rgb, atmrt_meta* meta, int32_t* steps, atmrt_stats* stats) {
    if (!g) return ATMRT_ERR_INVALID;
    if (!g->has_params) return gfail(g, ATMRT_ERR_STATE, "group_render before group_set_params");
    const int n = (int)g->ctx.size(), W = g->params.width;
    std::vector<atmrt_stats> st(n);
    const int rc = group_parallel(g, [&](int i) -> int {
        int x0 = 0, x1 = 0;
        atmrt_group_column_block(g, W, i, &x0, &x1);
        return render_to_host(g->ctx[i], rgb, meta, steps, W, x0, &st[i]);
    });
    if (rc) return rc;
    if (stats) {
        *stats = st[0];
        for (int i = 1; i < n; ++i) {
            stats->ray_steps += st[i].ray_steps, stats->trace_points += st[i].trace_points, stats->pixels_hit += st[i].pixels_hit;
            stats->step_overflows += st[i].step_overflows, stats->terrain_samples += st[i].terrain_samples;
            stats->kernel_launches += st[i].kernel_launches;
            stats->n_path_max = std::max(stats->n_path_max, st[i].n_path_max);
            stats->ms_terrain = std::max(stats->ms_terrain, st[i].ms_terrain), stats->ms_paths = std::max(stats->ms_paths, st[i].ms_paths);
            stats->ms_march = std::max(stats->ms_march, st[i].ms_march), stats->ms_total = std::max(stats->ms_total, st[i].ms_total);
        }
    }
    return 0;
}

int atmrt_group_render_tiles(atmrt_group* g, const atmrt_tile_desc* tiles, int ntiles, const int16_t* const* posts, uint8_t* rgb, atmrt_meta* meta,
                             int32_t* steps, atmrt_stats* stats) {
    if (!g || ntiles < 0 || (ntiles > 0 && (!tiles || !posts))) return gfail(g, ATMRT_ERR_INVALID, "group_render_tiles: bad argument");
    if (!g->has_params) return gfail(g, ATMRT_ERR_STATE, "group_render_tiles before group_set_params");
    int rc = group_upload_terrain(g, tiles, ntiles, posts, false);
    if (!rc) rc = atmrt_group_render(g, rgb, meta, steps, stats);  // (returns when the image is in host memory: the tiles have landed)
    for (atmrt_ctx* ctx : g->ctx) {
        if (rc && ctx->terrain_in_flight) {  // a failed render may not have waited for the copies
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(ctx->s_r);
        }
        ctx->terrain_in_flight = false;
    }
    return rc;
}

int atmrt_group_render_trace(atmrt_group* g, atmrt_trace_point* points, int32_t* counts, int max_points) {
    if (!g || !counts || max_points < 0 || (max_points > 0 && !points)) return gfail(g, ATMRT_ERR_INVALID, "group_render_trace: bad argument");
    if (!g->has_params) return gfail(g, ATMRT_ERR_STATE, "group_render_trace before group_set_params");
    const int W = g->params.width, H = g->params.height;
    return group_parallel(g, [&](int i) -> int {
        int x0 = 0, x1 = 0;
        atmrt_group_column_block(g, W, i, &x0, &x1);
        const size_t wl = (size_t)(x1 - x0), P = (size_t)max_points;
        std::vector<atmrt_trace_point> pts(wl * H * std::max<size_t>(P, 1));
        std::vector<int32_t> cnt(wl * H);
        const int rc = atmrt_render_trace(g->ctx[i], max_points > 0 ? pts.data() : nullptr, cnt.data(), max_points);
        if (rc) return gfail(g, rc, atmrt_last_error(g->ctx[i]));
        for (int y = 0; y < H; ++y) {  // the block's rows into the image's
            memcpy(counts + (size_t)y * W + x0, cnt.data() + (size_t)y * wl, wl * sizeof(int32_t));
            if (P) memcpy(points + ((size_t)y * W + x0) * P, pts.data() + (size_t)y * wl * P, wl * P * sizeof(atmrt_trace_point));
        }
        return 0;
    });
}

int atmrt_group_pixel_angles(atmrt_group* g, double* elevation_angle, double* azimuth) {
    if (!g) return ATMRT_ERR_INVALID;
    if (!g->has_params) return gfail(g, ATMRT_ERR_STATE, "group_pixel_angles before group_set_params");
    const int W = g->params.width, H = g->params.height;
    for (size_t i = 0; i < g->ctx.size(); ++i) {
        int x0 = 0, x1 = 0;
        atmrt_group_column_block(g, W, (int)i, &x0, &x1);
        const size_t wl = (size_t)(x1 - x0);
        std::vector<double> el(elevation_angle ? wl * H : 0), az(azimuth ? wl * H : 0);
        const int rc = atmrt_pixel_angles(g->ctx[i], elevation_angle ? el.data() : nullptr, azimuth ? az.data() : nullptr);
        if (rc) return gfail(g, rc, g->ctx[i]->err);
        for (int y = 0; y < H; ++y) {
            if (elevation_angle) memcpy(elevation_angle + (size_t)y * W + x0, el.data() + (size_t)y * wl, sizeof(double) * wl);
            if (azimuth) memcpy(azimuth + (size_t)y * W + x0, az.data() + (size_t)y * wl, sizeof(double) * wl);
        }
    }
    return 0;
}

// Page-locked host memory visible to every GPU (cudaHostAllocPortable): the copies of a render into it are asynchronous
// and run at the full rate of each GPU's own PCIe link.
void* atmrt_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, std::max<size_t>(bytes, 1), cudaHostAllocPortable) == cudaSuccess) return p;
    cudaGetLastError();
    // no device (or no page-locked memory left): ordinary memory -- the copies then run staged and synchronous
    p = malloc(std::max<size_t>(bytes, 1));
    if (p) {
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        g_pageable.insert(p);
    }
    return p;
}
void atmrt_host_free(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_pageable_mutex);
        auto it = g_pageable.find(p);
        if (it != g_pageable.end()) {
            g_pageable.erase(it);
            free(p);
            return;
        }
    }
    cudaFreeHost(p);
}

}  // extern "C"
