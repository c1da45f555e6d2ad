// device_paths.cuh -- Stage B: the ray-path cache (gen_path_cache, generators/utils.rs:136-174), one
// serial RK4 chain per image row.
//
// The stage is bound by the dependent-issue latency of ONE chain (N_t steps x the critical path of a
// step), not by throughput: all H chains are resident at once and the kernel ends when the longest
// one does. Everything here therefore shortens the critical path of a step while keeping the
// reference's arithmetic in f64:
//
//  1. RK4 structure. The refractive index is a function of altitude only, and the altitude input of
//     stage s+1 is a0 + w*d*ka_s with ka_s = b_s (the slope input of stage s). Stage 2's altitude
//     a0 + d/2*b0 is therefore known when the step starts, and stage 4's altitude needs kb_2 only.
//     The four index evaluations collapse into two dependent rounds: {stage 1, stage 2} then
//     {stage 3, stage 4}. No arithmetic changes, only the schedule.
//  2. Six lanes per row. Each round evaluates n at 2 altitudes x {h-eps, h, h+eps} (the central
//     difference dn/dh) on six lanes; the six values are exchanged with shuffles and every lane
//     applies the identical RK4 update, so the six copies of the state never diverge.
//  3. Short dependency chains inside n(h): log and exp with Estrin-evaluated polynomials (depth
//     ~log2(degree) instead of degree), divisions by per-layer or per-render constants replaced by
//     multiplications with reciprocals computed once, the remaining reciprocals (1/T, 1/Z, 1/n, 1/r)
//     by MUFU.RCP64H + two Newton steps, issued early so they overlap the log/exp chain. All of it is
//     f64 and accurate to ~1e-16 relative in p and in (n - 1), i.e. 3e-20 absolute in n -- four
//     orders below the rounding of `1.0 + x` (1.1e-16) that the reference's own finite difference
//     carries as noise (a relative 2e-7 of dn/dh per evaluation). Anything outside the validated
//     argument ranges (T/T_ref outside [1/16, 16], |arg| > 690, NaN) takes the libm path of
//     device_atm.cuh, which is the arithmetic the oracle restates op for op.
#pragma once

#include "device_atm.cuh"

namespace atmrt {

// 1/x for normal x: MUFU.RCP64H seed (>= 20 bits) and two Newton steps (error ~1 ulp).
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// log(u) for u in [2^-4, 2^4]: u = 2^e * m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1),
// |s| <= 0.1716; the odd series is cut after s^23 (remainder < 1e-18 relative) and evaluated with
// Estrin's scheme in z = s^2.
__device__ __forceinline__ double log_fast(double u) {
    int hi = __double2hiint(u), lo = __double2loint(u);
    int e = (hi >> 20) - 1023;
    int mh = (hi & 0x000fffff) | 0x3ff00000;
    if (mh >= 0x3ff6a09f) {  // m >= ~sqrt(2): halve it
        mh -= 0x00100000;
        e += 1;
    }
    const double m = __hiloint2double(mh, lo);
    const double f = m - 1.0;
    const double y = rcp_nr(m + 1.0);
    double s = f * y;
    s = fma(fma(-(m + 1.0), s, f), y, s);  // one residual correction: s = f/(m+1) to < 1 ulp
    const double z = s * s;
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    // 2/3, 2/5, ..., 2/23
    const double p01 = fma(z, 2.0 / 5.0, 2.0 / 3.0), p23 = fma(z, 2.0 / 9.0, 2.0 / 7.0), p45 = fma(z, 2.0 / 13.0, 2.0 / 11.0);
    const double p67 = fma(z, 2.0 / 17.0, 2.0 / 15.0), p89 = fma(z, 2.0 / 21.0, 2.0 / 19.0);
    const double q0 = fma(z2, p23, p01), q1 = fma(z2, p67, p45), q2 = fma(z2, 2.0 / 23.0, p89);
    const double poly = fma(z8, q2, fma(z4, q1, q0));
    const double ed = (double)e;
    // e*ln2_hi is exact for |e| <= 4 (ln2_hi has 21 trailing zero bits)
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    return fma(ed, ln2_hi, fma(2.0, s, fma(s * z, poly, ed * ln2_lo)));
}

// exp(x) for |x| <= 690: x = k ln2 + r, |r| <= 0.3466, Taylor to r^13 (remainder 4e-18) in Estrin form,
// scaled by 2^k through the exponent field (the result is a normal number for |x| <= 690).
__device__ __forceinline__ double exp_fast(double x) {
    const double log2e = 1.44269504088896338700e+00, magic = 6755399441055744.0;  // 1.5 * 2^52
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double t = fma(x, log2e, magic);
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(kd, -ln2_hi, x);
    r = fma(kd, -ln2_lo, r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double c2 = 1.0 / 2, c3 = 1.0 / 6, c4 = 1.0 / 24, c5 = 1.0 / 120, c6 = 1.0 / 720, c7 = 1.0 / 5040, c8 = 1.0 / 40320,
                 c9 = 1.0 / 362880, c10 = 1.0 / 3628800, c11 = 1.0 / 39916800, c12 = 1.0 / 479001600, c13 = 1.0 / 6227020800.0;
    const double p01 = 1.0 + r, p23 = fma(r, c3, c2), p45 = fma(r, c5, c4), p67 = fma(r, c7, c6), p89 = fma(r, c9, c8);
    const double pab = fma(r, c11, c10), pcd = fma(r, c13, c12);
    const double q0 = fma(r2, p23, p01), q1 = fma(r2, p67, p45), q2 = fma(r2, pab, p89);
    const double v = fma(r8, fma(r4, pcd, q2), fma(r4, q1, q0));
    return __hiloint2double(__double2hiint(v) + (k << 20), __double2loint(v));
}

// The temperature function a lane is currently inside, kept in registers between evaluations.
struct LayerRegs {
    double lo, hi;  // [lo, hi): altitude range of the function
    double h_ref, t_ref, p_ref, gradient, expo;
    double inv_t_ref;  // 1 / t_ref
    double k_iso;      // -g M / (R t_ref), isothermal functions
};

__device__ __forceinline__ void load_layer(const DevAtmosphere& a, double h, LayerRegs& L) {
    const int i = atm_layer_index(a, h);
    const DevAtmLayer& l = a.layer[i];
    L.lo = l.start;
    L.hi = i + 1 < a.n ? a.layer[i + 1].start : __longlong_as_double(0x7ff0000000000000LL);
    L.h_ref = l.h_ref, L.t_ref = l.t_ref, L.p_ref = l.p_ref, L.gradient = l.gradient, L.expo = l.expo;
    L.inv_t_ref = 1.0 / l.t_ref;
    L.k_iso = l.gm / l.rt;
}

// The libm path, out of line so that it stays out of the hot loop's instruction footprint.
template <bool DRY>
__device__ __noinline__ double env_n_slow(const DevAtmosphere& a, double h) {
    return env_n_t<DRY>(a, h);
}

// Environment::n(h) on the short-chain path; bit-for-bit the libm path of device_atm.cuh when the
// arguments leave the validated ranges.
template <bool DRY>
__device__ __forceinline__ double env_n_fast(const DevAtmosphere& a, LayerRegs& L, double h) {
    if (!(h >= L.lo && h < L.hi)) load_layer(a, h, L);  // also taken (every time) for a NaN altitude
    const double t = L.t_ref + L.gradient * (h - L.h_ref);
    const double rt = rcp_nr(t);  // independent of the pressure chain: overlaps it
    double arg = 0.0;
    bool slow = false;
    if (L.gradient != 0.0) {
        const double u = t * L.inv_t_ref;
        if (u >= 0.0625 && u <= 16.0)
            arg = L.expo * log_fast(u);
        else
            slow = true;
    } else {
        arg = (h - L.h_ref) * L.k_iso;
    }
    if (slow || !(fabs(arg) <= 690.0)) return env_n_slow<DRY>(a, h);
    const double p = L.p_ref * exp_fast(arg);
    if (!DRY) return air_index_t<false>(a, p, t);
    // air_index_t<true>: n = 1 + (rho_a / rho_axs) r_axs, rho_a = p m_a / (Z R T),
    // Z = 1 - (p/T)(a0 + a1 tc + a2 tc^2) + (p/T)^2 d
    const double a0 = 1.58123e-6, a1 = -2.9331e-8, a2 = 1.1043e-10, d = 1.83e-11;
    const double t_c = t - 273.15;
    const double poly = a0 + a1 * t_c + a2 * t_c * t_c;
    const double pt = p * rt;
    const double z_m = 1.0 - pt * poly + pt * pt * d;
    return 1.0 + pt * a.k_dry * rcp_nr(z_m);
}

// d^2/dt^2 of the ray equation from n(h) and the three-point difference, with 1/n and 1/r as
// reciprocals (utils of atm-refraction 0.6: flat h'' = n'/n (1 + h'^2); spherical
// r'' = (n'/n)(r'^2 + r^2) + 2 r'^2 / r + r).
template <bool FLAT>
__device__ __forceinline__ double ray_accel(double a, double b, double inv_a, double n0, double n_minus, double n_plus) {
    const double g = (n_plus - n_minus) * 50.0 * rcp_nr(n0);  // (n2 - n1) / (2 eps) / n, eps = 0.01
    if (FLAT) return g * (1.0 + b * b);
    const double bb = b * b;
    return bb * g + a * a * g + 2.0 * bb * inv_a + a;
}

}  // namespace atmrt
