// device_paths.cuh -- Stage B: the ray-path cache (gen_path_cache, generators/utils.rs:136-174), one
// serial RK4 chain per image row.
//
// The stage is bound by the dependent-issue latency of ONE chain (N_t steps x the critical path of a
// step; a dependent f64 op costs 8.2 cycles on B200), not by throughput: all H chains are resident at
// once and the kernel ends when the longest one does -- and because every rank needs every row, this
// latency is the floor of the multi-GPU frame time. Everything here shortens the critical path of a
// step while the integration itself stays the reference's classical RK4 in f64:
//
//  1. The ray equation needs the atmosphere only through ONE function of altitude,
//         g(h) = dn(h) / n(h),   dn(h) = (n(h + eps) - n(h - eps)) / (2 eps),  eps = 0.01 m
//     (spherical: r'' = g (r'^2 + r^2) + 2 r'^2 / r + r; flat: h'' = g (1 + h'^2)). The reference
//     evaluates it with three Ciddor indices per RK4 stage (pow/exp + divisions: ~500 dependent
//     cycles even when the three run on separate lanes) and, because n = 1 + 2.8e-4 is rounded to
//     1.1e-16 before the difference is taken, every evaluation carries a relative noise of 2e-7.
//     The host lowers the atmosphere once per set_params into a table of g: 250 m cells, a degree-6
//     polynomial per cell fitted at Chebyshev nodes to the reference's own definition (same layers,
//     hydrostatic law, Ciddor terms and central difference) evaluated in x87 extended precision, and
//     checked at 33 points per cell to 1e-12 relative (atmrt_lib.cu:build_g_table). A table lookup is
//     one magic-number cell index, seven shared-memory loads and an Estrin evaluation: ~80 dependent
//     cycles. The table is 5 orders of magnitude closer to the real-number value of the reference's
//     formula than the reference's own f64 evaluation is (tests/test_noise_floor.py).
//  2. Cells the table cannot serve hold NaN coefficients, and a NaN result sends that lane's step to the
//     fallback: a cell with the start of a temperature function inside (g has a kink there) is served
//     by one polynomial per function (`pieces`), everything else (T <= 1 K, fit check failed, outside
//     -2.25 .. 189.75 km) by g_libm(), which is the oracle's arithmetic op for op (device_atm.cuh).
//     atmrt_set_path_mode(1) forces every evaluation through g_libm (validation, tests).
//  3. Two lookups per step instead of four, predicted one step ahead (rk4_step_shared): the stage
//     altitudes lie within 1e-4 m of the points a + d/2 b and a + d b, also when those are extrapolated
//     from the previous state; g and 1/r at the stages are the values at those points with a first-order
//     correction (neglected term < 1e-12 relative). The lookups leave the critical path; what remains on
//     it is the chain of the four slope updates.
//  4. Where g is smooth sixteen steps are taken as ONE RK4 step and the states in between come from the
//     cubic Hermite interpolant of its ends, on sixteen sub-lanes (kernels.cuh: k_ray_paths_macro); across
//     the starts of the temperature functions, where g jumps and the reference's result depends on how its
//     steps straddle the jump, the kernel takes the reference's single steps (rk4_step_shared). PathElem::dist
//     and calc_dist's dx / R do not depend on the row and come from two host tables.
#pragma once

#include "device_atm.cuh"

namespace atmrt {

// 250 m cells centred on ATM_BASE + j * 250 m, j = 0 .. 767 (-2.25 .. 189.75 km). The cell edges are the
// multiples of 250 m, so that the temperature functions of US-76 (and of any definition written in round
// numbers) start on an edge and no cell holds a kink of g.
constexpr int ATM_CELLS = 768;
constexpr double ATM_CELL = 250.0;
constexpr double ATM_BASE = -2125.0;
constexpr int ATM_FIELDS = 7;   // per cell: coefficients c0 .. c6 of g(h_j + 128 u), u in [-1, 1]; coefficient-major

// 1/x for normal x: MUFU.RCP64H seed (>= 20 bits) and two Newton steps (error ~1 ulp).
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// sqrt(x) for normal positive x (NaN in, NaN out): MUFU.RSQ64H seed, two coupled Newton steps and a
// final residual correction -- the libm sequence without its special-case branch, so that it can be
// scheduled inside the straight-line step.
__device__ __forceinline__ double sqrt_nr(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g), h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}

// predicated 8-byte global store (no branch)
__device__ __forceinline__ void stg_if(double* p, double v, bool ok) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %0, 0;\n\t@q st.global.f64 [%1], %2;\n\t}" ::"r"((unsigned)ok), "l"(p), "d"(v) : "memory");
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// g(h) = dn/n exactly as the reference forms it (Environment::n three times, central difference): the
// libm path, out of line so that it stays out of the hot loop's instruction footprint.
__device__ __noinline__ double g_libm(const DevAtmosphere& a, double h) {
    double n, dn;
    env_n_dn(a, h, &n, &dn);
    return dn / n;
}

// g(h) from the table. `tab` is the shared-space address of [ATM_FIELDS][ATM_CELLS]; `hs` = h - ATM_BASE,
// the altitude above the centre of cell 0 (the callers fold ATM_BASE into the subtraction that turns r
// into h). NaN when the cell does not serve the altitude (also for NaN, negative-overflow and too-high
// altitudes: they select the first or the last cell, which are never served).
__device__ __forceinline__ double g_table(unsigned tab, double hs) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52: the low word of hs / 250 + magic is the nearest cell index
    const double hx = fma(hs, 1.0 / ATM_CELL, magic);
    const unsigned j = min((unsigned)__double2loint(hx), (unsigned)(ATM_CELLS - 1));
    const unsigned cj = tab + j * 8u;
    const double c0 = lds_f64(cj), c1 = lds_f64(cj + 8u * ATM_CELLS), c2 = lds_f64(cj + 16u * ATM_CELLS),
                 c3 = lds_f64(cj + 24u * ATM_CELLS), c4 = lds_f64(cj + 32u * ATM_CELLS), c5 = lds_f64(cj + 40u * ATM_CELLS),
                 c6 = lds_f64(cj + 48u * ATM_CELLS);
    const double u = fma(magic - hx, ATM_CELL, hs) * (2.0 / ATM_CELL);  // (hs - 250 j) / 125; 250 j is exact
    const double u2 = u * u;
    const double p01 = fma(c1, u, c0), p23 = fma(c3, u, c2), p45 = fma(c5, u, c4);
    const double u4 = u2 * u2;
    return fma(fma(c6, u2, p45), u4, fma(p23, u2, p01));
}

// g and dg/dh at a base altitude: the value as g_table, the slope from the first two derivative terms
// (c1 + 2 c2 u) / 125 -- relative error x^2 / 2 with x = 125 m / H < 0.07, H the scale height of g, which
// is ample for the first-order corrections below (they move g by a relative 1e-8).
struct GBase {
    double g, t;
};
__device__ __forceinline__ GBase g_table2(unsigned tab, double hs) {
    const double magic = 6755399441055744.0;
    const double hx = fma(hs, 1.0 / ATM_CELL, magic);
    const unsigned j = min((unsigned)__double2loint(hx), (unsigned)(ATM_CELLS - 1));
    const unsigned cj = tab + j * 8u;
    const double c0 = lds_f64(cj), c1 = lds_f64(cj + 8u * ATM_CELLS), c2 = lds_f64(cj + 16u * ATM_CELLS),
                 c3 = lds_f64(cj + 24u * ATM_CELLS), c4 = lds_f64(cj + 32u * ATM_CELLS), c5 = lds_f64(cj + 40u * ATM_CELLS),
                 c6 = lds_f64(cj + 48u * ATM_CELLS);
    const double u = fma(magic - hx, ATM_CELL, hs) * (2.0 / ATM_CELL);
    const double u2 = u * u;
    const double p01 = fma(c1, u, c0), p23 = fma(c3, u, c2), p45 = fma(c5, u, c4);
    const double u4 = u2 * u2;
    GBase r;
    r.g = fma(fma(c6, u2, p45), u4, fma(p23, u2, p01));
    r.t = fma(c2 + c2, u, c1) * (2.0 / ATM_CELL);
    return r;
}

// A base point of the lookups: the altitude variable P (r or h), g and dg/dh there, and 1/P.
struct PathBase {
    double P, g, t, r;
};
template <bool FLAT>
__device__ __forceinline__ PathBase path_base(unsigned tab, double shift, double P) {
    const GBase v = g_table2(tab, P - shift);
    return PathBase{P, v.g, v.t, FLAT ? 0.0 : rcp_nr(P)};
}

// One classical RK4 step with two table lookups instead of four, both OFF the critical path. The four
// stage altitudes of step i are
//   a,   a + d/2 b,   a + d/2 b + (d/2)^2 kb1,   a + d b + d (d/2) kb2
// and step i+1 starts at a + d b + O(d^2 kb). With M_i ~ a + d/2 b and N_i ~ a + d b, every altitude lies
// within a few d^2 |kb| of one of M_i, N_i, N_{i-1} -- 1e-4 m for a 25 m step, 1e-2 m at 80 degrees of
// elevation -- even when M_i, N_i are PREDICTED one step ahead from the previous state
// (a_{i-1} + 3/2 d b_{i-1}, a_{i-1} + 2 d b_{i-1}). g is smooth on the scale of kilometres, so g at the
// stage altitudes is g at the base plus a first-order correction; the neglected g'' e^2 / 2 is below
// 1e-12 relative for |e| < 0.01 m (the reference's own g carries 3e-7 of rounding noise), and 1/r for the
// geometric term likewise (error (e/r)^2). `ok` is false -- the caller redoes the step with per-stage
// evaluations -- when a base is not served by the table or a correction distance exceeds 0.25 m
// (kilometre-long steps). The lookups for step i+1 depend only on the state at the start of step i, so
// the scheduler interleaves them with the chain of the four slope updates, which is all that is left on
// the critical path.
template <bool FLAT>
__device__ __forceinline__ bool rk4_step_shared(double d, double hd, double d6, double a, double b, const PathBase& E, const PathBase& M,
                                                const PathBase& N, double* a_out, double* b_out) {
    // stage 1 at a, from the previous step's N
    const double e1 = a - E.P;
    const double g1 = fma(E.t, e1, E.g);
    const double i1 = FLAT ? 0.0 : fma(-(e1 * E.r), E.r, E.r);
    const double bb1 = b * b;
    const double kb1 = fma(g1, FLAT ? 1.0 + bb1 : fma(a, a, bb1), FLAT ? 0.0 : fma(bb1 + bb1, i1, a));
    const double b2 = fma(hd, kb1, b), a2 = fma(hd, b, a);
    // stage 2 at a2 ~ M
    const double e2 = a2 - M.P;
    const double g2 = fma(M.t, e2, M.g);
    const double i2 = FLAT ? 0.0 : fma(-(e2 * M.r), M.r, M.r);
    const double bb2 = b2 * b2;
    const double kb2 = fma(g2, FLAT ? 1.0 + bb2 : fma(a2, a2, bb2), FLAT ? 0.0 : fma(bb2 + bb2, i2, a2));
    const double b3 = fma(hd, kb2, b), a3 = fma(hd, b2, a);
    // stage 3 at a3 = a2 + (d/2)^2 kb1
    const double e3 = a3 - M.P;
    const double g3 = fma(M.t, e3, M.g);
    const double i3 = FLAT ? 0.0 : fma(-(e3 * M.r), M.r, M.r);
    const double bb3 = b3 * b3;
    const double kb3 = fma(g3, FLAT ? 1.0 + bb3 : fma(a3, a3, bb3), FLAT ? 0.0 : fma(bb3 + bb3, i3, a3));
    const double b4 = fma(d, kb3, b), a4 = fma(d, b3, a);
    // stage 4 at a4 = a + d b + d (d/2) kb2 ~ N
    const double e4 = a4 - N.P;
    const double g4 = fma(N.t, e4, N.g);
    const double i4 = FLAT ? 0.0 : fma(-(e4 * N.r), N.r, N.r);
    const double bb4 = b4 * b4;
    const double kb4 = fma(g4, FLAT ? 1.0 + bb4 : fma(a4, a4, bb4), FLAT ? 0.0 : fma(bb4 + bb4, i4, a4));
    *a_out = fma((b + 2.0 * b2) + (2.0 * b3 + b4), d6, a);
    *b_out = fma((kb1 + 2.0 * kb2) + (2.0 * kb3 + kb4), d6, b);
    // served bases (a NaN g is an unserved cell -- or a NaN state, which the caller accepts) and short corrections
    const double sum = (E.g + M.g) + N.g;
    const double emax = fmax(fmax(fabs(e1), fabs(e2)), fmax(fabs(e3), fabs(e4)));
    return sum == sum && emax < 0.25;
}

// A cell that holds the start of a temperature function is not served by the table (g has a kink
// there); it is served piecewise instead: one polynomial per (cell, temperature function) over the part
// of the cell that function owns. A cell whose fit fails for another reason (the law degenerates where a
// temperature function approaches 0 K) is bisected down to 1 m pieces. Few, scanned linearly.
struct DevGPiece {
    double h_lo, h_hi;        // the piece serves h_lo <= h < h_hi
    double centre, inv_half;  // u = (h - centre) * inv_half
    double c[ATM_FIELDS];
};
constexpr int ATM_MAX_PIECES = 96;

// The fallback of a failed table lookup: the pieces, then libm. Out of line (rare).
__device__ __noinline__ double g_fallback(unsigned tab, const DevGPiece* __restrict__ pieces, int npieces, const DevAtmosphere& a, double h) {
    const double g = g_table(tab, h - ATM_BASE);
    if (g == g || h != h) return g;
    for (int k = 0; k < npieces; ++k) {
        const DevGPiece& p = pieces[k];
        if (h >= p.h_lo && h < p.h_hi) {
            const double u = (h - p.centre) * p.inv_half;
            const double u2 = u * u, u4 = u2 * u2;
            return fma(fma(p.c[6], u2, fma(p.c[5], u, p.c[4])), u4, fma(fma(p.c[3], u, p.c[2]), u2, fma(p.c[1], u, p.c[0])));
        }
    }
    return g_libm(a, h);
}

// One classical RK4 step of the ray equation (PathStepper::next, restated in device_atm.cuh:
// stepper_next) on the state (a, b) = (r, dr/dphi) or (h, dh/dx). MODE 0: g from the table only; returns
// false when a lookup could not serve a finite altitude (the caller redoes the step in mode 1).
// MODE 1: table, then pieces, then libm, per evaluation. MODE 2: every g through g_libm.
struct GSource {
    unsigned tab;
    const DevGPiece* pieces;
    int npieces;
};
template <bool FLAT, int MODE>
__device__ __forceinline__ bool rk4_step(const DevAtmosphere& atm, const GSource& gs, double radius, double d, double hd, double d6, double a,
                                         double b, double* a_out, double* b_out) {
    auto G = [&](double alt) -> double {
        if (MODE == 0) return g_table(gs.tab, alt - (FLAT ? ATM_BASE : radius + ATM_BASE));
        const double h = FLAT ? alt : alt - radius;
        if (MODE == 1) return g_fallback(gs.tab, gs.pieces, gs.npieces, atm, h);
        return g_libm(atm, h);
    };
    const double a2 = fma(hd, b, a);  // ka1 = b
    const double g1 = G(a), g2 = G(a2);
    const double bb1 = b * b;
    const double s1 = FLAT ? 1.0 + bb1 : fma(a, a, bb1);
    const double c1 = FLAT ? 0.0 : fma(2.0 * bb1, rcp_nr(a), a);
    const double inv_a2 = FLAT ? 0.0 : rcp_nr(a2);
    const double kb1 = fma(g1, s1, c1);
    const double b2 = fma(hd, kb1, b);
    const double a3 = fma(hd, b2, a);
    const double g3 = G(a3);
    const double inv_a3 = FLAT ? 0.0 : rcp_nr(a3);
    const double bb2 = b2 * b2;
    const double s2 = FLAT ? 1.0 + bb2 : fma(a2, a2, bb2);
    const double c2 = FLAT ? 0.0 : fma(2.0 * bb2, inv_a2, a2);
    const double kb2 = fma(g2, s2, c2);
    const double b3 = fma(hd, kb2, b);
    const double a4 = fma(d, b3, a);
    const double g4 = G(a4);
    const double inv_a4 = FLAT ? 0.0 : rcp_nr(a4);
    const double bb3 = b3 * b3;
    const double s3 = FLAT ? 1.0 + bb3 : fma(a3, a3, bb3);
    const double c3 = FLAT ? 0.0 : fma(2.0 * bb3, inv_a3, a3);
    const double kb3 = fma(g3, s3, c3);
    const double b4 = fma(d, kb3, b);
    const double bb4 = b4 * b4;
    const double s4 = FLAT ? 1.0 + bb4 : fma(a4, a4, bb4);
    const double c4 = FLAT ? 0.0 : fma(2.0 * bb4, inv_a4, a4);
    const double kb4 = fma(g4, s4, c4);
    // y += (k1 + 2 k2 + 2 k3 + k4) d / 6
    *a_out = fma((b + 2.0 * b2) + (2.0 * b3 + b4), d6, a);
    *b_out = fma((kb1 + 2.0 * kb2) + (2.0 * kb3 + kb4), d6, b);
    // a NaN g with finite altitudes = an unserved cell; with a NaN state every g is NaN and stays so.
    const double sum = (g1 + g2) + (g3 + g4);
    return MODE != 0 || sum == sum || a != a || b != b;
}

}  // namespace atmrt
