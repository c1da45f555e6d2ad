// device_atm.cuh -- atmosphere profiles, Ciddor refractive index and the ray-path stepper in f64.
//
// Restates the external crate atm-refraction 0.6 (Cargo.toml:8) from its published algorithms and
// from the reference's call sites (generators/utils.rs:142-160, ray_path.rs:57-91,
// atm_printer.rs:33-46): US-1976-style hydrostatic layers with linear temperature, the Ciddor (1996)
// refractive index, dn/dh by central difference (eps = 0.01 m) and classical RK4 on the Fermat ray
// equation in polar (spherical Earth) or Cartesian (flat Earth) form. PARITY UNPINNED against the
// real crate (its source is not available offline); parity is against oracle/atmrt_oracle.cpp.
#pragma once

#include "device_math.cuh"

namespace atmrt {

// One law of the atmosphere after host lowering (Atmosphere::from_def): a Linear temperature function, or one
// segment of a Spline temperature function (its end segments continued to the function's boundaries), with a
// reference altitude of known temperature and pressure.
constexpr int ATM_MAX_LAYERS = 32;  // Linear functions + Spline segments of one atmosphere
struct DevAtmLayer {
    double start;     // lower boundary (-inf for layer 0)
    double h_ref, t_ref, p_ref;
    double gradient;  // K/m (Linear)
    double expo;      // Linear, gradient != 0: -g*M/(R*gradient); cubic: -g*M/R
    double gm;        // -g*M
    double rt;        // R*t_ref (isothermal layers)
    double x0, c0, c1, c2, c3;  // cubic: T = c0 + x (c1 + x (c2 + x c3)), x = h - x0
    int cubic;
    int _pad;
};

struct DevAtmosphere {
    int n;
    int _pad;
    double humidity;
    DevAtmLayer layer[ATM_MAX_LAYERS];
    // Ciddor terms that depend only on the wavelength (lowered on the host with + - * / only).
    double r_axs, r_vs, m_a, rho_axs;
    double k_dry;  // m_a r_axs / (R rho_axs)
};

__device__ __forceinline__ int atm_layer_index(const DevAtmosphere& a, double h) {
    int idx = 0;
    for (int i = 1; i < a.n; ++i)
        if (h >= a.layer[i].start) idx = i;
    return idx;
}

// Gauss-Legendre, 8 points on [-1, 1]: nodes +-GL8_X[i], weights GL8_W[i].
#define ATMRT_GL8_X {0.1834346424956498049394761, 0.5255324099163289858177390, 0.7966664774136267395915539, 0.9602898564975362316835609}
#define ATMRT_GL8_W {0.3626837833783619829651504, 0.3137066458778872873379622, 0.2223810344533744705443560, 0.1012285362903762591525314}

__device__ __forceinline__ double cubic_temperature(const DevAtmLayer& l, double h) {
    const double x = h - l.x0;
    return fma(fma(fma(l.c3, x, l.c2), x, l.c1), x, l.c0);
}

__device__ __forceinline__ double layer_temperature(const DevAtmLayer& l, double h) {
    if (l.cubic) return cubic_temperature(l, h);
    return l.t_ref + l.gradient * (h - l.h_ref);
}

// Integral of dh / T over [h_ref, h] inside one Spline segment: the 8-point rule, nodes taken in ascending order
// (oracle: cubic_inverse_integral).
__device__ __noinline__ double cubic_inverse_integral(const DevAtmLayer& l, double h) {
    const double gx[4] = ATMRT_GL8_X, gw[4] = ATMRT_GL8_W;
    const double half = 0.5 * (h - l.h_ref), mid = 0.5 * (h + l.h_ref);
    double acc = 0.0;
#pragma unroll
    for (int i = 3; i >= 0; --i) acc += gw[i] / cubic_temperature(l, mid - half * gx[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) acc += gw[i] / cubic_temperature(l, mid + half * gx[i]);
    return acc * half;
}

// Hydrostatic pressure inside one layer: p_ref (T/T_ref)^expo for a linear temperature function,
// p_ref exp(-g M (h - h_ref) / (R T_ref)) for an isothermal one, p_ref exp(-g M / R * integral of dh / T) for a
// Spline segment. pow() is evaluated as
// exp(expo * log(x)) because this sits on the serial critical path of the ray stepper (pow costs 740
// dependent cycles on B200, log + exp 430): for x in [0.5, 1.5] and |expo| < 40 that stays within
// ~3 ulp of the correctly rounded power -- far below the half-ulp-of-n noise (4e-7 relative in
// dn/dh) that the reference's finite-difference derivative carries anyway.
__device__ __forceinline__ double layer_pressure(const DevAtmLayer& l, double h, double t) {
    double arg;
    if (l.cubic)
        arg = l.expo * cubic_inverse_integral(l, h);
    else if (l.gradient != 0.0)
        arg = l.expo * log(t / l.t_ref);
    else
        arg = l.gm * (h - l.h_ref) / l.rt;
    return l.p_ref * exp(arg);
}

// Ciddor (1996); p in Pa, t in K. DRY = relative humidity 0: every water-vapour term is then an
// exact zero (x_v = 0 => rho_v = 0, (..)*x_v = 0, 1 - x_v = 1), so dropping them is bit-identical.
template <bool DRY>
__device__ __forceinline__ double air_index_t(const DevAtmosphere& a, double p, double t_kelvin) {
    const double a0 = 1.58123e-6, a1 = -2.9331e-8, a2 = 1.1043e-10;
    const double b0 = 5.707e-6, b1 = -2.051e-8;
    const double c0 = 1.9898e-4, c1 = -2.376e-6;
    const double d = 1.83e-11, e = -0.765e-8;
    const double rho_vs = 0.00985938;
    const double gas_r = 8.314510, m_v = 0.018015;
    const double alpha = 1.00062, beta = 3.14e-8, gamma = 5.6e-7;
    const double sa = 1.2378847e-5, sb = -1.9121316e-2, sc = 33.93711047, sd = -6.3431645e3;

    const double t_c = t_kelvin - 273.15;
    const double pt = p / t_kelvin;
    if (DRY) {
        const double z_m = 1.0 - pt * (a0 + a1 * t_c + a2 * t_c * t_c) + pt * pt * d;
        const double rho_a = p * a.m_a / (z_m * gas_r * t_kelvin);
        return 1.0 + (rho_a / a.rho_axs) * a.r_axs;
    }
    const double svp = exp(sa * t_kelvin * t_kelvin + sb * t_kelvin + sc + sd / t_kelvin);
    const double f = alpha + beta * p + gamma * t_c * t_c;
    const double x_v = a.humidity * f * svp / p;
    const double z_m = 1.0 - pt * (a0 + a1 * t_c + a2 * t_c * t_c + (b0 + b1 * t_c) * x_v + (c0 + c1 * t_c) * x_v * x_v) +
                       pt * pt * (d + e * x_v * x_v);
    const double rho_v = x_v * p * m_v / (z_m * gas_r * t_kelvin);
    const double rho_a = (1.0 - x_v) * p * a.m_a / (z_m * gas_r * t_kelvin);
    return 1.0 + (rho_a / a.rho_axs) * a.r_axs + (rho_v / rho_vs) * a.r_vs;
}
__device__ __forceinline__ double air_index(const DevAtmosphere& a, double p, double t_kelvin) {
    return a.humidity != 0.0 ? air_index_t<false>(a, p, t_kelvin) : air_index_t<true>(a, p, t_kelvin);
}

// Environment::n(h)
template <bool DRY>
__device__ __forceinline__ double env_n_t(const DevAtmosphere& a, double h) {
    const DevAtmLayer& l = a.layer[atm_layer_index(a, h)];
    double t = layer_temperature(l, h);
    double p = layer_pressure(l, h, t);
    return air_index_t<DRY>(a, p, t);
}
__device__ __forceinline__ double env_n(const DevAtmosphere& a, double h) {
    return a.humidity != 0.0 ? env_n_t<false>(a, h) : env_n_t<true>(a, h);
}

// n(h) and dn/dh = (n(h+eps) - n(h-eps)) / (2 eps): three independent evaluations.
__device__ __forceinline__ void env_n_dn(const DevAtmosphere& a, double h, double* n, double* dn) {
    const double eps = 0.01;
    double n0 = env_n(a, h);
    double n1 = env_n(a, h - eps);
    double n2 = env_n(a, h + eps);
    *n = n0;
    *dn = (n2 - n1) / (2.0 * eps);
}

struct RayState {
    double x, h;
};

// Ray stepper state for one row of the path cache.
struct Stepper {
    // RK4 state: spherical (r, dr/dphi, phi) or flat (h, dh/dx, x)
    double a, b, t;
    // straight-line parameters
    double h0, ang;
};

__device__ __forceinline__ void stepper_init(Stepper& s, int flat, double radius, double start_h, double ang_rad) {
    s.h0 = start_h;
    s.ang = ang_rad;
    s.t = 0.0;
    if (flat) {
        s.a = start_h;
        s.b = tan(ang_rad);
    } else {
        s.a = radius + start_h;
        s.b = s.a * tan(ang_rad);
    }
}

__device__ __forceinline__ void deriv_sph(const DevAtmosphere& atm, double radius, double r, double dr, double* o_r, double* o_dr) {
    double n, dn;
    env_n_dn(atm, r - radius, &n, &dn);
    *o_r = dr;
    *o_dr = dr * dr * dn / n + r * r * dn / n + 2.0 * dr * dr / r + r;
}
__device__ __forceinline__ void deriv_flat(const DevAtmosphere& atm, double h, double dh, double* o_h, double* o_dh) {
    double n, dn;
    env_n_dn(atm, h, &n, &dn);
    *o_h = dh;
    *o_dh = dn / n * (1.0 + dh * dh);
}

// PathStepper::next(): one step of size `step` metres along the ground.
__device__ __forceinline__ RayState stepper_next(Stepper& s, const DevAtmosphere& atm, int flat, int straight, double radius, double step) {
    if (straight) {
        s.t += step;
        if (flat) return {s.t, s.h0 + s.t * tan(s.ang)};
        double r0 = radius + s.h0;
        double ph = s.t / radius;
        double rr = r0 * cos(s.ang) / cos(ph + s.ang);
        return {s.t, rr - radius};
    }
    double k1a, k1b, k2a, k2b, k3a, k3b, k4a, k4b;
    if (flat) {
        double d = step;
        deriv_flat(atm, s.a, s.b, &k1a, &k1b);
        deriv_flat(atm, s.a + 0.5 * d * k1a, s.b + 0.5 * d * k1b, &k2a, &k2b);
        deriv_flat(atm, s.a + 0.5 * d * k2a, s.b + 0.5 * d * k2b, &k3a, &k3b);
        deriv_flat(atm, s.a + d * k3a, s.b + d * k3b, &k4a, &k4b);
        s.a = s.a + (k1a + 2.0 * k2a + 2.0 * k3a + k4a) * d / 6.0;
        s.b = s.b + (k1b + 2.0 * k2b + 2.0 * k3b + k4b) * d / 6.0;
        s.t += d;
        return {s.t, s.a};
    }
    double d = step / radius;
    deriv_sph(atm, radius, s.a, s.b, &k1a, &k1b);
    deriv_sph(atm, radius, s.a + 0.5 * d * k1a, s.b + 0.5 * d * k1b, &k2a, &k2b);
    deriv_sph(atm, radius, s.a + 0.5 * d * k2a, s.b + 0.5 * d * k2b, &k3a, &k3b);
    deriv_sph(atm, radius, s.a + d * k3a, s.b + d * k3b, &k4a, &k4b);
    s.a = s.a + (k1a + 2.0 * k2a + 2.0 * k3a + k4a) * d / 6.0;
    s.b = s.b + (k1b + 2.0 * k2b + 2.0 * k3b + k4b) * d / 6.0;
    s.t += d;
    return {s.t * radius, s.a - radius};
}

// calc_dist, generators/utils.rs:42-53
__device__ __forceinline__ double calc_dist(int flat, double radius, RayState o, RayState n) {
    double dx = n.x - o.x, dh = n.h - o.h;
    if (flat) return sqrt(dx * dx + dh * dh);
    double avg_h = (n.h + o.h) / 2.0;
    dx = dx / radius * (avg_h + radius);
    return sqrt(dx * dx + dh * dh);
}

}  // namespace atmrt
