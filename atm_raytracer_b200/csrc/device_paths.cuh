// device_paths.cuh -- Stage B: the ray-path cache (gen_path_cache, generators/utils.rs:136-174), one
// serial RK4 chain per image row.
//
// The stage is bound by the dependent-issue latency of ONE chain (N_t steps x the critical path of a
// step; a dependent f64 op costs 8.2 cycles on B200), not by throughput: all H chains are resident at
// once and the kernel ends when the longest one does. Everything here shortens the critical path of a
// step while keeping the reference's arithmetic in f64:
//
//  1. RK4 structure. The refractive index is a function of altitude only, and the altitude input of
//     stage s+1 is a0 + w*d*ka_s with ka_s = b_s (the slope input of stage s). Stage 2's altitude
//     a0 + d/2*b0 is therefore known when the step starts, and stage 4's altitude needs kb_2 only.
//     The four index evaluations collapse into two dependent rounds: {stage 1, stage 2} then
//     {stage 3, stage 4}. No arithmetic changes, only the schedule.
//  2. Six lanes per row. Each round evaluates n at 2 altitudes x {h-eps, h, h+eps} (the central
//     difference dn/dh) on six lanes; the values are exchanged with shuffles and every lane applies
//     the identical RK4 update, so the six copies of the state never diverge.
//  3. Re-anchored hydrostatic pressure. Inside one temperature function T is linear in h, so for any
//     anchor altitude h_j of that function  p(h) = p(h_j) (T(h)/T(h_j))^alpha = p_j exp(alpha log1p(w))
//     with w = g (h - h_j) / T_j  (isothermal: p_j exp(k (h - h_j))) -- the same identity the reference
//     uses with the function's own reference point, moved to the centre of a 256 m cell. |w| <= 8e-3
//     and |alpha w| <= 0.034 * 128 / T, so log1p and exp are two short series (remainders < 1e-17)
//     with no range reduction, evaluated in Estrin form. The anchors p_j, T_j come from the
//     reference's own expressions (host libm pow / exp) and are staged in shared memory.
//  4. The remaining divisions use reciprocals: MUFU.RCP64H + two Newton steps for 1/T and 1/r (issued
//     early, they overlap the series), geometric series for 1/Z and 1/n (both within 1e-2 of 1).
//  5. The fast path is one straight-line block: cells it cannot serve (a temperature-function
//     boundary inside the cell, series not accurate enough, outside the table, NaN altitude) hold NaN
//     anchors, and a NaN result sends that one evaluation to the libm path of device_atm.cuh, which
//     is the arithmetic the oracle restates op for op. A cell is served when the truncation errors of
//     the two series, as an absolute error of n, stay below 3e-20 (atmrt_lib.cu:build_atm_table).
//
// Accuracy: p and (n - 1) carry ~1e-16 relative error, i.e. 3e-20 absolute in n -- four orders below
// the rounding of `1.0 + x` (1.1e-16) that the reference's own finite difference carries as noise (a
// relative 2e-7 of dn/dh per evaluation, ~2e-6 m of path altitude at 200 km; tests/test_noise_floor).
#pragma once

#include "device_atm.cuh"

namespace atmrt {

constexpr int ATM_CELLS = 768;  // 256 m cells centred on ATM_BASE + j * 256 m, j = 0 .. 767 (up to 194 km)
constexpr double ATM_CELL = 256.0;
constexpr double ATM_BASE = -2048.0;
constexpr int ATM_FIELDS = 5;   // per cell: p_j, T_j, g (K/m), w scale, alpha
// An isothermal function p_j exp(k dh) is served by the same formula as a linear one,
// p_j (1 + w)^alpha with w = k dh 2^-50 and alpha = 2^50 (relative error |k dh| 2^-51 < 1e-17).
constexpr double ATM_ISO_SCALE = 1125899906842624.0;  // 2^50

// Series coefficients live in constant memory so that DFMA reads them as constant-bank operands (a
// 64-bit literal costs two MOVs per use otherwise).
__constant__ double K_SER[16] = {
    -1.0 / 2.0, 1.0 / 3.0, -1.0 / 4.0, 1.0 / 5.0, -1.0 / 6.0, 1.0 / 7.0,                        // log1p: [0..5]
    1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,  // exp:   [6..13]
    1.0 / 362880.0,
    0.0, 0.0};
__constant__ double K_AIR[4] = {1.58123e-6, -2.9331e-8, 1.1043e-10, 1.83e-11};  // Ciddor compressibility a0, a1, a2, d

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

// 1/(1 + x) for |x| < 1e-2: (1 - x)(1 + x^2)(1 + x^4) = 1 - x + ... - x^7 (remainder x^8 < 1e-16).
__device__ __forceinline__ double rcp_1p(double x) {
    const double x2 = x * x, m = 1.0 - x;
    const double q = fma(x2, m, m);
    return fma(x2 * x2, q, q);
}

// The libm path, out of line so that it stays out of the hot loop's instruction footprint.
template <bool DRY>
__device__ __noinline__ double env_n_slow(const DevAtmosphere& a, double h) {
    return env_n_t<DRY>(a, h);
}

__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// Environment::n(h) on the short-chain path. `cells` is the shared-space address of the anchor table,
// [ATM_FIELDS][ATM_CELLS]; `sel` is the lane's stage altitude before the eps offset (r or h) and `hx` the
// same altitude in cell units plus 1.5 * 2^52 (an imprecise copy of the chain that is only used to pick
// the cell: its low word is the nearest cell index). *bad is set when the cell cannot serve the
// altitude (NaN anchors); the caller then takes env_n_slow.
template <bool DRY>
__device__ __forceinline__ double env_n_fast(const DevAtmosphere& a, unsigned cells, double h, double hx, bool* bad) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    const unsigned j = min((unsigned)__double2loint(hx), (unsigned)(ATM_CELLS - 1));  // negative, NaN -> an edge cell (NaN anchors)
    const unsigned cj = cells + j * 8u;
    const double pj = lds_f64(cj), tj = lds_f64(cj + 8u * ATM_CELLS), gj = lds_f64(cj + 16u * ATM_CELLS),
                 sj = lds_f64(cj + 24u * ATM_CELLS), aj = lds_f64(cj + 32u * ATM_CELLS);
    *bad = pj != pj;
    const double dh = h - fma(hx - magic, ATM_CELL, ATM_BASE);
    const double t = fma(gj, dh, tj);
    const double rt = rcp_nr(t);  // independent of the pressure chain: overlaps it
    const double w = dh * sj;
    // log1p(w) = w (1 - w/2 + w^2/3 - ... + w^6/7)
    const double w2 = w * w;
    const double l01 = fma(w, K_SER[0], 1.0), l23 = fma(w, K_SER[2], K_SER[1]), l45 = fma(w, K_SER[4], K_SER[3]);
    const double lq = fma(w2 * w2, fma(w2, K_SER[5], l45), fma(w2, l23, l01));
    const double v = (aj * w) * lq;  // alpha log1p(w)
    // exp(v), |v| < 0.08: Taylor to v^9
    const double v2 = v * v, v4 = v2 * v2;
    const double e01 = 1.0 + v, e23 = fma(v, K_SER[7], K_SER[6]), e45 = fma(v, K_SER[9], K_SER[8]),
                 e67 = fma(v, K_SER[11], K_SER[10]), e89 = fma(v, K_SER[13], K_SER[12]);
    const double ev = fma(v4 * v4, e89, fma(v4, fma(v2, e67, e45), fma(v2, e23, e01)));
    const double p = pj * ev;
    if (!DRY) return air_index_t<false>(a, p, t);
    // air_index_t<true>: n = 1 + (rho_a / rho_axs) r_axs, rho_a = p m_a / (Z R T),
    // Z = 1 - (p/T)(a0 + a1 tc + a2 tc^2) + (p/T)^2 d = 1 - e
    const double t_c = t - 273.15;
    const double poly = fma(fma(K_AIR[2], t_c, K_AIR[1]), t_c, K_AIR[0]);
    const double pt = p * rt;
    const double e = pt * fma(-pt, K_AIR[3], poly);
    return fma(pt * a.k_dry, rcp_1p(-e), 1.0);
}

}  // namespace atmrt
