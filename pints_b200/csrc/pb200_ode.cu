// Fused adaptive Dormand-Prince 5(4) solve + objective for the ODE toy models
// (FitzHugh-Nagumo, Lotka-Volterra, SIR, Goodwin, Hes1, Repressilator).  One
// lane integrates one parameter vector; the observed series sits in shared
// memory (bulk-copied once per block); every accepted step is interpolated at
// the observation times that fall inside it (4th-order dense output), the
// residuals are squared and accumulated in registers, and only the score
// leaves the SM -- to the caller's array, or straight into the gathered arrays
// of all GPUs of a job (ScatterDev).
//
// In this file, in order:
//   tableau, models (load / rhs / scale(h) / hrhs = h f / cost_key)
//   ode_kernel        one lane per vector; plain layout (64-thread blocks) or
//                     the migrating layout (one block per SM, MigDev)
//   ode_pair_kernel   two lanes per vector (option, two-state models)
//   sort_kernel       cost-coherent scheduling (cooperative counting sort);
//                     local_cost_sort: the block-by-block sort inside
//                     ode_kernel where the launch is one block per SM
//   launchers         layout selection, launch_ode
// DESIGN.md section 3 has the measurements each choice rests on.
//
// Replaces, per parameter vector (citations relative to the reference root):
//   ToyODEModel._simulate           pints/toy/_toy_classes.py:190-234
//   FitzhughNagumoModel._rhs        pints/toy/_fitzhugh_nagumo_model.py:111-117
//   LotkaVolterraModel._rhs         pints/toy/_lotka_volterra_model.py:91-97
//   SIRModel.simulate/_rhs          pints/toy/_sir_model.py:96-111
//   GoodwinOscillatorModel._rhs     pints/toy/_goodwin_oscillator_model.py:101-107
//   Hes1Model._rhs                  pints/toy/_hes1_michaelis_menten.py:129-140
//   RepressilatorModel._rhs/simulate  pints/toy/_repressilator_model.py:91-108
//   Multi/SingleOutputProblem.evaluate   pints/_core.py:147-153, :257-265
//   objective __call__              see objective_finish()
// The reference integrates with LSODA (scipy odeint, rtol = atol = 1.49012e-8);
// this kernel uses DP5(4) at the same tolerances, see DESIGN.md for the
// measured agreement.
#include "pb200_common.cuh"
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

// --------------------------------------------------------------- tableau
// Dormand-Prince 5(4): stage weights A, 5th-order weights B, error weights E
// (5th minus 4th order), dense-output weights D (Hairer, Norsett & Wanner,
// contd5).  Kept in __constant__ memory so every DFMA/DMUL takes its coefficient
// straight from the constant bank instead of re-materialising 64-bit immediates
// (two UMOVs each) inside the step loop.
enum { I_A21, I_A31, I_A32, I_A41, I_A42, I_A43, I_A51, I_A52, I_A53, I_A54, I_A61, I_A62, I_A63, I_A64, I_A65, I_B1, I_B3, I_B4, I_B5, I_B6, I_E1, I_E3, I_E4, I_E5, I_E6, I_E7, I_D1, I_D3, I_D4, I_D5, I_D6, I_D7, N_TAB };
__constant__ double TAB[N_TAB] = {
    1.0 / 5.0,
    3.0 / 40.0,
    9.0 / 40.0,
    44.0 / 45.0,
    -56.0 / 15.0,
    32.0 / 9.0,
    19372.0 / 6561.0,
    -25360.0 / 2187.0,
    64448.0 / 6561.0,
    -212.0 / 729.0,
    9017.0 / 3168.0,
    -355.0 / 33.0,
    46732.0 / 5247.0,
    49.0 / 176.0,
    -5103.0 / 18656.0,
    35.0 / 384.0,
    500.0 / 1113.0,
    125.0 / 192.0,
    -2187.0 / 6784.0,
    11.0 / 84.0,
    71.0 / 57600.0,
    -71.0 / 16695.0,
    71.0 / 1920.0,
    -17253.0 / 339200.0,
    22.0 / 525.0,
    -1.0 / 40.0,
    -12715105075.0 / 11282082432.0,
    87487479700.0 / 32700410799.0,
    -10690763975.0 / 1880347072.0,
    701980252875.0 / 199316789632.0,
    -1453857185.0 / 822651844.0,
    69997945.0 / 29380423.0,
};
#define A21 TAB[I_A21]
#define A31 TAB[I_A31]
#define A32 TAB[I_A32]
#define A41 TAB[I_A41]
#define A42 TAB[I_A42]
#define A43 TAB[I_A43]
#define A51 TAB[I_A51]
#define A52 TAB[I_A52]
#define A53 TAB[I_A53]
#define A54 TAB[I_A54]
#define A61 TAB[I_A61]
#define A62 TAB[I_A62]
#define A63 TAB[I_A63]
#define A64 TAB[I_A64]
#define A65 TAB[I_A65]
#define B1 TAB[I_B1]
#define B3 TAB[I_B3]
#define B4 TAB[I_B4]
#define B5 TAB[I_B5]
#define B6 TAB[I_B6]
#define E1 TAB[I_E1]
#define E3 TAB[I_E3]
#define E4 TAB[I_E4]
#define E5 TAB[I_E5]
#define E6 TAB[I_E6]
#define E7 TAB[I_E7]
#define D1 TAB[I_D1]
#define D3 TAB[I_D3]
#define D4 TAB[I_D4]
#define D5 TAB[I_D5]
#define D6 TAB[I_D6]
#define D7 TAB[I_D7]


// ----------------------------------------------------------------- models
// Each model: NS states, NO outputs, OUT0 = index of the first output state.
// load() turns a parameter row into per-thread constants and y0; rhs() is the
// plain (autonomous) right-hand side, used once for the first step size;
// scale(h) folds the step size into the constants so that hrhs() returns
// K = h f(y) directly: every stage, the error estimate and the dense output
// then work on h-scaled increments and need no further multiplication by h.
struct FhnModel {
    static constexpr bool MIGRATES = true;
    static constexpr bool TIMED = false;
    static constexpr int NS = 2, NO = 2, OUT0 = 0, NP = 3;
    // V' = c (V - V^3/3 + R) = (c - (c/3) V^2) V + c R
    // R' = -(V - a + b R) / c  = (-1/c) V + (a/c) + (-b/c) R
    // written with per-vector constants so one evaluation is 6 FP64 ops
    double c, c3, nic, aoc, nboc;
    double hc, hc3, hnic, haoc, hnboc;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double /*t_first*/,
                                         double *t0) {
        const double a = x[0], b = x[1];
        c = x[2];
        c3 = c * (-1.0 / 3.0);
        nic = -1.0 / c;
        aoc = -nic * a;
        nboc = nic * b;
        y[0] = mc.c[0]; y[1] = mc.c[1];
        *t0 = 0.0;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        const double V = y[0], R = y[1];
        const double u = fma(c3, V * V, c);
        f[0] = fma(u, V, c * R);
        f[1] = fma(nboc, R, fma(nic, V, aoc));
    }
    __device__ __forceinline__ void scale(double h) {
        hc = h * c; hc3 = h * c3; hnic = h * nic; haoc = h * aoc; hnboc = h * nboc;
    }
    // scheduling key (see sort_keys_kernel): c sets the time scale of the
    // relaxation oscillation, vectors with equal c stay in phase longest
    __device__ __forceinline__ double cost_key(const double *) const { return fabs(c); }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        const double V = y[0], R = y[1];
        const double u = fma(hc3, V * V, hc);
        K[0] = fma(u, V, hc * R);
        K[1] = fma(hnboc, R, fma(hnic, V, haoc));
    }
};

struct LvModel {
    static constexpr bool MIGRATES = true;
    static constexpr bool TIMED = false;
    static constexpr int NS = 2, NO = 2, OUT0 = 0, NP = 4;
    double a, b, c, d;
    double ha, nhb, nhc, hd;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double, double *t0) {
        a = x[0]; b = x[1]; c = x[2]; d = x[3];
        y[0] = mc.c[0]; y[1] = mc.c[1];
        *t0 = 0.0;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        const double u = y[0], v = y[1];
        f[0] = u * fma(-b, v, a);         // a x - b x y
        f[1] = v * fma(d, u, -c);         // -c y + d x y
    }
    __device__ __forceinline__ void scale(double h) {
        ha = h * a; nhb = -h * b; nhc = -h * c; hd = h * d;
    }
    __device__ __forceinline__ double cost_key(const double *y) const {
        double f[NS];
        rhs(y, f);
        return fma(f[0], f[0], f[1] * f[1]);      // |f(y0)|^2, initial rate of change
    }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        const double u = y[0], v = y[1];
        K[0] = u * fma(nhb, v, ha);
        K[1] = v * fma(hd, u, nhc);
    }
};

struct SirModel {
    static constexpr bool MIGRATES = true;
    static constexpr bool TIMED = false;
    static constexpr int NS = 3, NO = 2, OUT0 = 1, NP = 3;
    double gamma, v;
    double nhg, hv;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double t_first,
                                         double *t0) {
        gamma = x[0]; v = x[1];
        double s0 = x[2];
        // default y0 of the reference is an integer array: S0 is truncated
        // toward zero on assignment (pints/toy/_sir_model.py:80,108-109)
        if (mc.c[3] != 0.0) s0 = trunc(s0);
        y[0] = s0; y[1] = mc.c[1]; y[2] = mc.c[2];
        *t0 = t_first;                    // initial condition holds at times[0]
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        const double ngsi = (-gamma * y[0]) * y[1];
        const double vi = v * y[1];
        f[0] = ngsi;
        f[1] = -ngsi - vi;
        f[2] = vi;
    }
    __device__ __forceinline__ void scale(double h) {
        nhg = -h * gamma; hv = h * v;
    }
    __device__ __forceinline__ double cost_key(const double *y) const {
        double f[NS];
        rhs(y, f);
        return fma(f[0], f[0], fma(f[1], f[1], f[2] * f[2]));
    }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        const double ngsi = (nhg * y[0]) * y[1];
        const double vi = hv * y[1];
        K[0] = ngsi;
        K[1] = -ngsi - vi;
        K[2] = vi;
    }
};

// MUFU.RCP64H: >= 20 good bits, on the XU pipe (off the FP64 pipe).
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// Reciprocal of a normal-range x: seed + Newton steps (20 -> 40 -> 80 bits);
// no slow path, two DFMAs per step.
template <int STEPS>
__device__ __forceinline__ double rcp_newton(double x) {
    double r = rcp_seed(x);
#pragma unroll
    for (int k = 0; k < STEPS; ++k) r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

// exp(x) for |x| <= 700 without the range branches of the library routine, so
// that independent exponentials interleave (the eight-state model evaluates
// sixteen per right-hand side): n = rint(x / ln 2) by the 1.5 * 2^52 shift,
// two-step Cody-Waite reduction to |r| <= 0.347, Taylor polynomial of degree
// 13 (remainder 4e-18), exponent added as an integer.  Within 1 ulp of exp().
__device__ __forceinline__ double exp_bounded(double x) {
    constexpr double SHIFT = 6755399441055744.0;
    const double t = fma(x, 1.4426950408889634, SHIFT);
    const int n = __double2loint(t);
    const double nf = t - SHIFT;
    double r = fma(nf, -6.93147180559945286e-01, x);
    r = fma(nf, -2.31904681384629956e-17, r);
    double p = 1.6059043836821613e-10;                // 1/13!
    p = fma(p, r, 2.08767569878681e-09);              // 1/12!
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}

// a / b for a normal-range b, no slow path: reciprocal by two Newton steps,
// quotient with one residual correction (last-place accurate).
__device__ __forceinline__ double div_normal(double a, double b) {
    const double r = rcp_newton<2>(b);
    const double q = a * r;
    return fma(r, fma(-b, q, a), q);
}

// log(x) for a positive normal x without branches: x = m 2^e with m in
// [sqrt(1/2), sqrt(2)), log m = 2 atanh((m - 1) / (m + 1)) as an odd series up
// to s^19 (|s| <= 0.1716, remainder 8e-18).  Within a few ulp.
__device__ __forceinline__ double log_positive(double x) {
    const int hi = __double2hiint(x);
    const int e = (hi - 0x3fe6a09e) >> 20;
    const double m = __hiloint2double(hi - (e << 20), __double2loint(x));
    const double s = div_normal(m - 1.0, m + 1.0);
    const double s2 = s * s;
    double p = 1.0 / 19.0;
    p = fma(p, s2, 1.0 / 17.0);
    p = fma(p, s2, 1.0 / 15.0);
    p = fma(p, s2, 1.0 / 13.0);
    p = fma(p, s2, 1.0 / 11.0);
    p = fma(p, s2, 1.0 / 9.0);
    p = fma(p, s2, 1.0 / 7.0);
    p = fma(p, s2, 1.0 / 5.0);
    p = fma(p, s2, 1.0 / 3.0);
    const double r = fma(p * s2, s, s);              // atanh(s)
    const double ef = static_cast<double>(e);
    return fma(ef, 6.93147180559945286e-01, fma(ef, 2.31904681384629956e-17, r + r));
}

// An exponent n and what numpy's power needs to know about it for a negative
// base (an integer-valued n keeps the sign rule, anything else gives NaN).
struct RealExponent {
    double n;
    bool integer, odd;
    __device__ __forceinline__ void set(double v) {
        n = v;
        integer = v == rint(v);
        odd = integer && (fmod(fabs(v), 2.0) == 1.0);
    }
};

// x^n = exp(n log |x|) with the branch-free pieces above (the library pow()
// is ~250 instructions with slow paths, and the Hill terms call it in every
// stage); the product n log|x| is clamped to exp_bounded's range (1e+-304).
// Relative error ~ |n log x| * 2e-16.  x = 0 gives 0 (n > 0), 1 (n = 0), inf.
__device__ __forceinline__ double pow_real(double x, const RealExponent &ex) {
    const double ax = fabs(x);
    const double a = fmin(fmax(ex.n * log_positive(ax), -700.0), 700.0);
    double r = exp_bounded(a);
    if (ax == 0.0) r = ex.n > 0.0 ? 0.0 : (ex.n == 0.0 ? 1.0 : INFINITY);
    if (x < 0.0) r = ex.integer ? (ex.odd ? -r : r) : NAN;
    if (!(ax == ax)) r = NAN;
    return r;
}

// Goodwin oscillator (pints/toy/_goodwin_oscillator_model.py:101-107):
//   x' = 1 / (1 + z^10) - m1 x,  y' = k2 x - m2 y,  z' = k3 y - m3 z
struct GoodwinModel {
    static constexpr int NS = 3, NO = 3, OUT0 = 0, NP = 5;
    static constexpr bool MIGRATES = false;     // needs 146 registers
    static constexpr bool TIMED = false;
    double k2, k3, m1, m2, m3;
    double h_, hk2, hk3, nhm1, nhm2, nhm3;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double, double *t0) {
        k2 = x[0]; k3 = x[1]; m1 = x[2]; m2 = x[3]; m3 = x[4];
        y[0] = mc.c[0]; y[1] = mc.c[1]; y[2] = mc.c[2];
        *t0 = 0.0;
    }
    __device__ __forceinline__ static double pow10(double z) {
        const double z2 = z * z, z4 = z2 * z2;
        return (z4 * z4) * z2;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        f[0] = 1.0 / (1.0 + pow10(y[2])) - m1 * y[0];
        f[1] = k2 * y[0] - m2 * y[1];
        f[2] = k3 * y[1] - m3 * y[2];
    }
    __device__ __forceinline__ void scale(double h) {
        h_ = h; hk2 = h * k2; hk3 = h * k3; nhm1 = -h * m1; nhm2 = -h * m2; nhm3 = -h * m3;
    }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        K[0] = fma(nhm1, y[0], h_ / (1.0 + pow10(y[2])));
        K[1] = fma(hk2, y[0], nhm2 * y[1]);
        K[2] = fma(hk3, y[1], nhm3 * y[2]);
    }
    __device__ __forceinline__ double cost_key(const double *y) const {
        double f[NS];
        rhs(y, f);
        return fma(f[0], f[0], fma(f[1], f[1], f[2] * f[2]));
    }
};

// Hes1 Michaelis-Menten model (pints/toy/_hes1_michaelis_menten.py:129-140),
// output = m; constants [m0, p1_0, p2_0, k_deg]:
//   m'  = -k_deg m + 1 / (1 + (p2 / P0)^h)
//   p1' = -k_deg p1 + v m - k1 p1,   p2' = -k_deg p2 + k1 p1
struct Hes1Model {
    static constexpr int NS = 3, NO = 1, OUT0 = 0, NP = 4;
    static constexpr bool MIGRATES = false;     // pow() needs > 128 registers
    static constexpr bool TIMED = false;
    double iP0, v, k1, kdeg;
    RealExponent hh;
    double h_, hv, hk1, nhkd;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double, double *t0) {
        iP0 = 1.0 / x[0]; v = x[1]; k1 = x[2]; hh.set(x[3]);
        y[0] = mc.c[0]; y[1] = mc.c[1]; y[2] = mc.c[2];
        kdeg = mc.c[3];
        *t0 = 0.0;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        f[0] = -kdeg * y[0] + 1.0 / (1.0 + pow_real(y[2] * iP0, hh));
        f[1] = -kdeg * y[1] + v * y[0] - k1 * y[1];
        f[2] = -kdeg * y[2] + k1 * y[1];
    }
    __device__ __forceinline__ void scale(double h) {
        h_ = h; hv = h * v; hk1 = h * k1; nhkd = -h * kdeg;
    }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        K[0] = fma(nhkd, y[0], div_normal(h_, 1.0 + pow_real(y[2] * iP0, hh)));
        K[1] = fma(hv, y[0], (nhkd - hk1) * y[1]);
        K[2] = fma(hk1, y[1], nhkd * y[2]);
    }
    __device__ __forceinline__ double cost_key(const double *y) const {
        double f[NS];
        rhs(y, f);
        return fma(f[0], f[0], fma(f[1], f[1], f[2] * f[2]));
    }
};

// Repressilator (pints/toy/_repressilator_model.py:91-108), six states, the
// three mRNA levels are the outputs; like SIR it is integrated from times[0]:
//   m_i' = -m_i + alpha / (1 + p_j^n) + alpha_0,   p_i' = -beta (p_i - m_i)
// with (i, j) = (0, 2), (1, 0), (2, 1) (protein j represses gene i).
struct RepressilatorModel {
    static constexpr int NS = 6, NO = 3, OUT0 = 0, NP = 4;
    static constexpr bool MIGRATES = false;     // 6 states need > 128 registers
    static constexpr bool TIMED = false;
    double a0, al, be;
    RealExponent nn;
    double h_, ha0, hal, hbe;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double t_first, double *t0) {
        a0 = x[0]; al = x[1]; be = x[2]; nn.set(x[3]);
#pragma unroll
        for (int k = 0; k < 6; ++k) y[k] = mc.c[k];
        *t0 = t_first;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        f[0] = -y[0] + al / (1.0 + pow_real(y[5], nn)) + a0;
        f[1] = -y[1] + al / (1.0 + pow_real(y[3], nn)) + a0;
        f[2] = -y[2] + al / (1.0 + pow_real(y[4], nn)) + a0;
        f[3] = -be * (y[3] - y[0]);
        f[4] = -be * (y[4] - y[1]);
        f[5] = -be * (y[5] - y[2]);
    }
    __device__ __forceinline__ void scale(double h) {
        h_ = h; ha0 = h * a0; hal = h * al; hbe = h * be;
    }
    __device__ __forceinline__ void hrhs(const double *y, double *K) const {
        K[0] = fma(-h_, y[0], div_normal(hal, 1.0 + pow_real(y[5], nn))) + ha0;
        K[1] = fma(-h_, y[1], div_normal(hal, 1.0 + pow_real(y[3], nn))) + ha0;
        K[2] = fma(-h_, y[2], div_normal(hal, 1.0 + pow_real(y[4], nn))) + ha0;
        K[3] = hbe * (y[0] - y[3]);
        K[4] = hbe * (y[1] - y[4]);
        K[5] = hbe * (y[2] - y[5]);
    }
    __device__ __forceinline__ double cost_key(const double *y) const {
        double f[NS], s2 = 0.0;
        rhs(y, f);
#pragma unroll
        for (int k = 0; k < NS; ++k) s2 = fma(f[k], f[k], s2);
        return s2;
    }
};

// Beeler-Reuter 1977 ventricular action potential
// (pints/toy/_beeler_reuter_model.py:105-183): eight states
// (V, Ca_i, m, h, j, d, f, x1), the first two are the outputs; the parameters
// are the natural logarithms of the five maximum conductances
// (gNa, gNaC, gCa, gK1, gx1).  The stimulus makes the right-hand side depend
// on time (TIMED): 25 uA/cm^2 while t mod 1000 < 2.  The m gate relaxes at
// 20 - 90 per ms over the whole action potential, so the explicit pair is
// stability-bound at h ~ 0.04 ms (7 - 8 thousand steps for 400 ms whatever
// the tolerance); the reference's LSODA switches to BDF there.
// Exponentials with the same rate share one exp() (14 instead of 27).
// consts: [V0, Ca0, m0, h0, j0, d0, f0, x10, C_m, E_Na, I_amp, I_period, I_length]
struct BeelerReuterModel {
    static constexpr int NS = 8, NO = 2, OUT0 = 0, NP = 5;
    static constexpr bool MIGRATES = false;
    static constexpr bool TIMED = true;
    double gNa, gNaC, gCa, gK1, gx1;
    double inv_cm, e_na, amp, period, inv_period, length, h_;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double t_first, double *t0) {
        gNa = exp(x[0]); gNaC = exp(x[1]); gCa = exp(x[2]); gK1 = exp(x[3]); gx1 = exp(x[4]);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = mc.c[k];
        inv_cm = 1.0 / mc.c[8]; e_na = mc.c[9]; amp = mc.c[10]; period = mc.c[11];
        length = mc.c[12];
        inv_period = 1.0 / period;
        h_ = 1.0;
        *t0 = t_first;
    }
    // stimulus current at `time`: (time % period) < length, times are not negative
    __device__ __forceinline__ double stimulus(double time) const {
        double ph = fma(-period, floor(time * inv_period), time);
        if (ph < 0.0) ph += period;
        if (ph >= period) ph -= period;
        return ph < length ? amp : 0.0;
    }
    // What the arrow-shaped Jacobian of the stiff integrator needs besides one
    // difference quotient in V: the gates' relaxation rates and three factors.
    struct Aux {
        double rate[6];       // alpha + beta of m, h, j, d, f, x1
        double dECa;          // V - E_Ca
        double m3, xf;        // m^3; Ix1 = gx1 x1 xf
        double cv[8];         // d f / d V (analytic: the exponentials are shared with f)
    };
    template <bool AUX>
    __device__ __forceinline__ void rhs_core(const double *y, double *f, double IStim, Aux *aux) const {
        const double V = y[0], Cai = y[1], m = y[2], hh = y[3], j = y[4], d = y[5],
                     ff = y[6], x1 = y[7];
        // one exponential per distinct rate; the shifted ones are constant
        // multiples.  Inside the exponentials V is clamped to +-800 mV (the
        // physiological range is -90 .. +50), which keeps every argument
        // within exp_bounded's range.
        const double Vc = fmin(fmax(V, -800.0), 800.0);
#define PB200_BR_EXP(rate, shift) exp_bounded((rate) * (Vc + (shift)))
        const double e_m01 = PB200_BR_EXP(-0.1, 47.0);            // also -0.1 (V + 32)
        const double e_m056 = PB200_BR_EXP(-0.056, 72.0);
        const double e_m25 = PB200_BR_EXP(-0.25, 77.0);           // also -0.25 (V + 78)
        const double e_m082 = PB200_BR_EXP(-0.082, 22.5);
        const double e_m2 = PB200_BR_EXP(-0.2, 78.0);             // also -0.2 (V + 30)
        const double e_m01b = PB200_BR_EXP(-0.01, -5.0);
        const double e_m072 = PB200_BR_EXP(-0.072, -5.0);
        const double e_m017 = PB200_BR_EXP(-0.017, 44.0);
        const double e_p05 = PB200_BR_EXP(0.05, 44.0);
        const double e_m008 = PB200_BR_EXP(-0.008, 28.0);
        const double e_p15 = PB200_BR_EXP(0.15, 28.0);
        const double e_m02 = PB200_BR_EXP(-0.02, 30.0);
        const double e_p04 = PB200_BR_EXP(0.04, 53.0);            // all the +-0.04 and 0.08 terms
        const double e_p083 = PB200_BR_EXP(0.083, 50.0);
        const double e_p057 = PB200_BR_EXP(0.057, 50.0);
        const double e_m06 = PB200_BR_EXP(-0.06, 20.0);
        const double inv_e04 = rcp_newton<2>(e_p04);
#undef PB200_BR_EXP
        // INa and its gates
        const double m3 = m * m * m;
        const double INa = (gNa * m3 * hh * j + gNaC) * (V - e_na);
        double al = div_normal(V + 47.0, 1.0 - e_m01);
        double be = 40.0 * e_m056;
        f[2] = al * (1.0 - m) - be * m;
        if (AUX) {
            aux->rate[0] = al + be;
            const double dal = (1.0 - 0.1 * e_m01 * al) * rcp_newton<1>(1.0 - e_m01);
            aux->cv[2] = dal * (1.0 - m) + 0.056 * be * m;
        }
        al = 0.126 * e_m25;
        be = div_normal(1.7, 1.0 + e_m082);
        f[3] = al * (1.0 - hh) - be * hh;
        if (AUX) {
            aux->rate[1] = al + be;
            const double dbe = be * 0.082 * e_m082 * rcp_newton<1>(1.0 + e_m082);
            aux->cv[3] = -0.25 * al * (1.0 - hh) - dbe * hh;
        }
        al = div_normal(0.055 * (e_m25 * 0.7788007830714049), 1.0 + e_m2);           // exp(-0.25)
        be = div_normal(0.3, 1.0 + e_m01 * 4.4816890703380645);                       // exp(1.5)
        f[4] = al * (1.0 - j) - be * j;
        if (AUX) {
            aux->rate[2] = al + be;
            const double e7 = e_m01 * 4.4816890703380645;
            const double dal = al * fma(0.2 * e_m2, rcp_newton<1>(1.0 + e_m2), -0.25);
            const double dbe = be * 0.1 * e7 * rcp_newton<1>(1.0 + e7);
            aux->cv[4] = dal * (1.0 - j) - dbe * j;
        }
        // ICa and its gates
        const double E_Ca = -82.3 - 13.0287 * log(Cai);
        const double ICa = gCa * d * ff * (V - E_Ca);
        al = div_normal(0.095 * e_m01b, e_m072 + 1.0);
        be = div_normal(0.07 * e_m017, e_p05 + 1.0);
        f[5] = al * (1.0 - d) - be * d;
        if (AUX) {
            aux->rate[3] = al + be;
            const double dal = al * fma(0.072 * e_m072, rcp_newton<1>(1.0 + e_m072), -0.01);
            const double dbe = be * fma(-0.05 * e_p05, rcp_newton<1>(1.0 + e_p05), -0.017);
            aux->cv[5] = dal * (1.0 - d) - dbe * d;
        }
        al = div_normal(0.012 * e_m008, e_p15 + 1.0);
        be = div_normal(0.0065 * e_m02, e_m2 * 14764.781565577267 + 1.0);             // exp(9.6)
        f[6] = al * (1.0 - ff) - be * ff;
        if (AUX) {
            aux->rate[4] = al + be;
            const double e15 = e_m2 * 14764.781565577267;
            const double dal = al * fma(-0.15 * e_p15, rcp_newton<1>(1.0 + e_p15), -0.008);
            const double dbe = be * fma(0.2 * e15, rcp_newton<1>(1.0 + e15), -0.02);
            aux->cv[6] = dal * (1.0 - ff) - dbe * ff;
        }
        f[1] = -1e-7 * ICa + 0.07 * (1e-7 - Cai);
        // IK1: exp(0.04 (V+85)) = e04 exp(1.28), exp(0.08 (V+53)) = e04^2,
        // exp(-0.04 (V+23)) = exp(1.2) / e04
        const double ik1_a = div_normal(4.0 * (e_p04 * 3.5966397255692817 - 1.0), e_p04 * e_p04 + e_p04);
        const double ik1_b = div_normal(0.2 * (V + 23.0), 1.0 - 3.3201169227365472 * inv_e04);
        const double IK1 = gK1 * (ik1_a + ik1_b);
        // Ix1: exp(0.04 (V+77)) = e04 exp(0.96), exp(0.04 (V+35)) = e04 exp(-0.72)
        const double xf = (e_p04 * 2.611696473423118 - 1.0) * (inv_e04 * 2.0544332106438876);   // exp(0.72)
        const double Ix1 = gx1 * x1 * xf;
        al = div_normal(0.0005 * e_p083, e_p057 + 1.0);
        be = div_normal(0.0013 * e_m06, 1.3674196065680964e-05 * inv_e04 + 1.0);      // exp(-11.2)
        f[7] = al * (1.0 - x1) - be * x1;
        if (AUX) {
            aux->rate[5] = al + be; aux->dECa = V - E_Ca; aux->m3 = m3; aux->xf = xf;
            const double e19 = 1.3674196065680964e-05 * inv_e04;
            const double dal = al * fma(-0.057 * e_p057, rcp_newton<1>(1.0 + e_p057), 0.083);
            const double dbe = be * fma(0.04 * e19, rcp_newton<1>(1.0 + e19), -0.06);
            aux->cv[7] = dal * (1.0 - x1) - dbe * x1;
            // currents: d IK1 / dV from the two quotients, d Ix1 / dV = gx1 x1 0.04 exp(-0.04 (V + 35))
            const double e08 = e_p04 * e_p04;
            const double d_a = (0.16 * 3.5966397255692817 * e_p04 - ik1_a * fma(0.08, e08, 0.04 * e_p04)) *
                               rcp_newton<1>(e08 + e_p04);
            const double ed = 3.3201169227365472 * inv_e04;
            const double d_b = (0.2 - ik1_b * 0.04 * ed) * rcp_newton<1>(1.0 - ed);
            const double dxf = 0.04 * 2.0544332106438876 * inv_e04;
            aux->cv[0] = -inv_cm * (gK1 * (d_a + d_b) + gx1 * x1 * dxf + gNa * m3 * hh * j + gNaC +
                                    gCa * d * ff);
            aux->cv[1] = -1e-7 * gCa * d * ff;
        }
        f[0] = -inv_cm * (IK1 + Ix1 + INa + ICa - IStim);
    }
    __device__ __forceinline__ void rhs_t(const double *y, double *f, double time) const {
        rhs_core<false>(y, f, stimulus(time), nullptr);
    }
    __device__ __forceinline__ void scale(double h) { h_ = h; }
    __device__ __forceinline__ void hrhs_t(const double *y, double *K, double time) const {
        rhs_t(y, K, time);
#pragma unroll
        for (int k = 0; k < NS; ++k) K[k] *= h_;
    }
    // every vector is stability-bound by the same gate: nothing to sort on
    __device__ __forceinline__ double cost_key(const double *) const { return gx1; }
};

// ss += r * r only where `valid`.  The compiler turns a predicated DFMA (or a
// branch) into a DFMA plus a 64-bit select, two FSELs per output and pass.
// Instead the HIGH word of r is cleared where the slot is not valid (one
// instruction): what is left is a denormal below 2^-1042 -- also when r was
// inf or NaN from a sentinel row -- whose square underflows to exactly +0, so
// the DFMA adds nothing.  (A valid r is never touched.)
__device__ __forceinline__ void acc_sq_if(double &ss, double r, bool valid) {
    const int hi = valid ? __double2hiint(r) : 0;
    const double rm = __hiloint2double(hi, __double2loint(r));
    ss = fma(rm, rm, ss);
}

// Order of two doubles from their bit patterns (integer pipe instead of a
// DSETP on the FP64 pipe).  Exact for non-negative a and any b whose sign bit
// marks "never" (-inf): times are non-negative, +inf is the largest pattern.
__device__ __forceinline__ bool le_bits(double a, double b) {
    return __double_as_longlong(a) <= __double_as_longlong(b);
}

// Step-size factor 0.9 * err^(-1/5), clamped to [0.2, 10], from the SUM of
// squared scaled errors (err^2 = sum / NS).  It runs on the SFU in fp32
// (lg2/ex2 approximations): the controller needs a few digits only, and
// accept/reject is decided on the fp64 sum, never on this value.
// sum = 0 -> 10, sum = inf or NaN -> 0.2.
template <int NS>
__device__ __forceinline__ float step_factor(double sum2) {
    const float n2 = static_cast<float>(sum2);
    float l, f;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(n2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(-0.1f * l));
    // 0.9 * NS^(1/10)
    const float k = NS == 1 ? 0.9f : NS == 2 ? 0.96459370f : NS == 3 ? 1.00450979f
                  : NS == 4 ? 1.03382570f : NS == 5 ? 1.05715438f : NS == 6 ? 1.07660593f
                                                             : 0.9f * exp2f(0.1f * log2f((float)NS));
    return fminf(fmaxf(k * f, 0.2f), 10.0f);
}

// The same factor without the SFU and without fp64 <-> fp32 conversions: the
// high word of a positive double is a fixed-point log2, so
//   hi(fac) = MAGIC - hi(sum2) / 10
// gives 0.9 (sum2 / NS)^(-1/10) to +-3.3 % (the piecewise-linear mantissa
// error, centred by the magic constant) in four integer instructions, and the
// clamps to [0.2, 10] (and to <= 1 after a rejection) are integer min / max on
// the same word.  A 3 % wobble in the safety factor leaves the number of
// attempted steps unchanged (measured: 384.6 against 385.1 per vector on the
// headline population).  sum2 = 0 -> 10, inf / NaN -> 0.2.
template <int NS>
__device__ __forceinline__ double step_factor_bits(double sum2, bool cap_at_one) {
    // 0x3FF00000 + 2^20 log2(0.9 NS^0.1) - 50000 (centring) + 0x3FF00000 / 10
    constexpr int32_t magic =
        NS == 1 ? 1179753186 : NS == 2 ? 1179858044 : NS == 3 ? 1179919381
        : NS == 4 ? 1179962902 : NS == 5 ? 1179996658 : NS == 6 ? 1180024239 : 1180067762;
    static_assert((NS >= 1 && NS <= 6) || NS == 8, "magic constant tabulated for 1..6 and 8 states");
    const uint32_t hi = static_cast<uint32_t>(__double2hiint(sum2));
    int32_t hy = magic - static_cast<int32_t>(__umulhi(hi, 0x1999999Au));
    hy = max(hy, 0x3FC99999);                        // 0.2
    hy = min(hy, cap_at_one ? 0x3FF00000 : 0x40240000);   // 1 or 10
    return __hiloint2double(hy, 0);
}

// Observed series access.  In shared memory the rows are read with explicit
// ld.shared through a 32-bit byte address kept in a register (the generic
// pointer form makes the compiler rebuild the shared-window base on every
// access); otherwise through the read-only global path.
// Opaque copy: keeps the compiler from recomputing the shared-window address
// (S2UR SR_CgaCtaId + three uniform ops) at every use inside the step loop.
__device__ __forceinline__ uint32_t launder_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

template <bool SMEM>
struct Series {
    uint32_t sbase;
    const double *gbase;
    __device__ __forceinline__ double at(uint32_t byte_off) const {
        double v;
        if (SMEM) {
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(sbase + byte_off));
        } else {
            v = __ldg(reinterpret_cast<const double *>(
                reinterpret_cast<const char *>(gbase) + byte_off));
        }
        return v;
    }
};

#ifndef PB200_TH1_FROM_TH
#define PB200_TH1_FROM_TH 0     // 1 - theta by subtraction instead of its own affine map
#endif
#ifndef PB200_SCALE_FROM_Y
#define PB200_SCALE_FROM_Y 0    // error weight rtol |y| + atol instead of rtol (|y| + |y_new|) / 2 + atol
#endif
#ifndef PB200_MIG_STRIDED
#define PB200_MIG_STRIDED 1     // migrating layout: warps dealt to the blocks round-robin
#endif
#ifndef PB200_STAGGER
#define PB200_STAGGER 0         // (development) start offset in cycles between a scheduler's warps
#endif
#ifndef PB200_LSORT_STRIDED
#define PB200_LSORT_STRIDED 2   // block-local sort: 2 = runs of 32 rows dealt round-robin, 1 = rows b, b + blocks, ..., 0 = one contiguous run
#endif
#ifndef PB200_TIMELINE
#define PB200_TIMELINE 0        // (development) per-warp start / end clocks
#endif
#if PB200_TIMELINE
__device__ long long pb200_timeline[160 * 16 * 4];
#endif
#ifndef PB200_PIPE_U
#define PB200_PIPE_U 2          // in-line interpolation slots per step attempt
#endif
#ifndef PB200_MAXTHREADS
#define PB200_MAXTHREADS 128
#endif
#ifndef PB200_CLOOP_WIDE
#define PB200_CLOOP_WIDE 4      // observations per trip of the wide (C) loop (1 = off)
#endif
#ifndef PB200_CLOOP_WIDTH
#define PB200_CLOOP_WIDTH 1     // observations per trip of the (C) loop
#endif
#ifndef PB200_LOOP_UNROLL
#define PB200_LOOP_UNROLL 2
#endif
#ifndef PB200_CTRL_BITS
#define PB200_CTRL_BITS 1       // integer step-size factor (no SFU)
#endif
#ifndef PB200_RCP_STEPS
#define PB200_RCP_STEPS 1
#endif
static_assert(PB200_PIPE_U + 1 <= PB200_SENTINEL_ROWS, "not enough sentinel rows");
static_assert(PB200_CLOOP_WIDE <= PB200_SENTINEL_ROWS, "not enough sentinel rows");

// Control flow is warp-uniform: every lane of a warp runs the same number of
// loop iterations (votes decide), finished lanes idle on harmless arithmetic.
// One iteration = one attempted step of every lane, a single basic block:
//   (C) observations of the pending interval beyond the U in-line slots
//       (uniform loop, only while some lane still has more than U pending),
//   (A) U predicated interpolation slots of the pending interval,
//   (B) the six stages on h-scaled increments, error estimate, controller,
//   (D) accept/reject by selects, dense-output coefficients parked for the
//       NEXT iteration's (C)/(A).
// (A) is independent of (B), so it fills the issue slots that the serial
// RHS -> stage -> RHS chain (9-cycle DFMA latency) leaves empty.
// ---------------------------------------------------------------------------
// Warp migration between the four schedulers of an SM (MIG = true).
// A warp stays on the scheduler (SM sub-partition) it was created on, and a
// single-wave launch takes as long as the BUSIEST scheduler: a population of
// 65 536 is 13.84 warps per SM, i.e. 4+4+3+3 per scheduler, and runs exactly
// as long as 16 warps per SM would (measured: 0.327 ms for 9..12 warps per SM,
// 0.425 ms for 13..16).  With one block per SM the placement is known (warp w
// sits on scheduler w mod 4), so the block is laid out as
//     warps 0 .. q4-1        workers                  (q4 a multiple of 4)
//     warps q4 .. q4+r-1     the r extra workers, on schedulers 0 .. r-1
//     warps q4+2 .. q4+1+r   receivers, parked on a named barrier on
//                            schedulers 2 .. 1+r    (r = 1 or 2)
// and about half way through the series each extra worker writes its state
// (y, h f(y), t, h, sums of squares, counters: 2.7 KB) to shared memory and
// retires; the receiver picks it up.  The busy schedulers drop from q+1 to q
// warps when the idle ones go from q to q+1: every scheduler carries about
// q + r/4 + ... warps on average instead of q+1 on the critical ones.
// Results are bit-identical to the plain layout (the state moves verbatim).
// ---------------------------------------------------------------------------
struct MigDev {
    int32_t q4;          // workers that never move
    int32_t r;           // extra workers that hand over to a receiver (0 = no migration)
    uint32_t off_mid;    // series byte offset at which they do
    uint32_t buf_off;    // byte offset of the hand-over area in dynamic shared memory
};

// Block shape the plain layout is compiled for.  The eight-state model would
// take 232 registers (8 warps per SM: two waves for a population of 65 536,
// 0.363 s).  Compiled for one block of up to 448 threads per SM it gets 128
// registers, part of the K arrays spills to local memory (touched once per
// stage, against ~700 instructions of right-hand side) and the 13.8 warps per
// SM of that population are resident at once; a single copy of the series per
// SM leaves the rest of the 256 KB for the spilled state in L1.  Measured at
// populations 16 384 / 65 536 / 262 144 (branch-free exp and division):
//   448 x 1, 128 registers   0.128 / 0.180 / 0.710 s      <- default
//   64 x 7,  128 registers   0.129 / 0.216 / 0.798 s
//   384 x 1, 168 registers   0.122 / 0.255 / 0.641 s      (two waves at 65 536)
// and with the library exp() and division (branches keep the sixteen
// exponentials of a right-hand side from interleaving): 0.192 / 0.205 / - s.
// Not unrolling the step loop for this model (halves 300 KB of code) brings
// the default to 0.087 / 0.159 / 0.630 s; a non-inlined right-hand side
// (25 KB of code, arguments through local memory) is slower: 0.112 / 0.193 /
// 0.80 s.
#ifndef PB200_BR_BLOCKS
#define PB200_BR_BLOCKS 1
#endif
#ifndef PB200_BR_THREADS
#define PB200_BR_THREADS 448
#endif
template <class M> struct LaunchShape {
    static constexpr int THREADS = PB200_MAXTHREADS, BLOCKS = 1;
};
template <> struct LaunchShape<BeelerReuterModel> {
    static constexpr int THREADS = PB200_BR_THREADS, BLOCKS = PB200_BR_BLOCKS;
};
#ifdef PB200_REP_THREADS
template <> struct LaunchShape<RepressilatorModel> {
    static constexpr int THREADS = PB200_REP_THREADS, BLOCKS = 1;
};
#endif
#ifndef PB200_BIG_UNROLL
#define PB200_BIG_UNROLL 2
#endif

// DENSE: register class of the instance.  The loop needs 124 registers to keep
// everything live, i.e. 16 warps per SM; populations of 16 < warps per SM <= 24
// (75 776 < N <= 113 664) would then take two waves, the second one nearly
// empty (48 % of the FP64 peak at N = 100 000 against 55 % at 65 536).  The
// same code compiled for 640-thread blocks gets 96 registers (12 bytes of
// spills), for 768-thread blocks 80 registers (114 bytes, outside the stage
// chain), and holds 20 / 24 warps per SM in ONE wave.
constexpr int dense_threads(int dense) { return dense == 2 ? 768 : dense == 1 ? 640 : 0; }

// ---------------------------------------------------------------------------
// Cost sort inside the evaluation kernel (one block per SM: migrating and dense
// layouts).  What the lock step of a warp needs is 32 vectors of similar cost
// in every warp; it does not need a GLOBAL order.  Block b takes the vectors
// b, b + blocks, b + 2 blocks, ... (a systematic sample of the population,
// whatever order the caller's rows are in) and sorts THOSE by the cost key in
// shared memory (counting
// sort over 1024 bins between the block's own smallest and largest key, no
// order inside a bin), under the bulk copy of the series: no launch of its own,
// no grid-wide barrier, no permutation in global memory.  Emulated on the
// headline population (tools/step_emulation.py): 3.71 interpolation passes per
// attempt for blocks of 448 vectors against 3.59 for the global sort and 4.60
// unsorted, i.e. 1.1 % more issue cycles than the global order -- against the
// 13.7 us (3.7 % of the step) that sort_kernel takes in front of the solve.
// A block's warps are in increasing order of cost, as the strided deal of the
// globally sorted order gave them, and every block holds an even sample of
// the population, so the SMs stay balanced as long as the costs of a
// population differ by a few per cent (an optimiser's or sampler's population).
// On a wide prior box (costs 1 : 5) the SMs' sums differ by a few per cent and
// the global order wins by 2 %: PB200_LOCAL_SORT=0 keeps sort_kernel.
// ---------------------------------------------------------------------------
constexpr int SORT_BINS = 1024;
struct LocalSortDev {
    int32_t on;                    // 0: launch order / perm as given
    uint32_t scratch_off;          // byte offset of LOCAL_SORT_WORDS words in dynamic shared memory
    int32_t *hint;                 // pb200_problem::sort_hint (or nullptr)
};

// Which sort pays is decided from what the previous launch measured (block 0 of
// every one-block-per-SM launch, at its end): the block-local sort costs about
// 1.2 x (relative mean absolute deviation of the attempted steps) x (kernel
// time) in lock-step waste and SM imbalance, the sort kernel 13.7 us in front
// of the solve.  Fitted on FitzHugh-Nagumo populations of relative width 2 /
// 5 / 10 / 20 % (deviation 0.8 / 1.6 / 3.0 / 6.8 % of 385 steps; local sort
// 9 us faster / 6 us faster / even / 14 us slower at 0.37-0.40 ms), and it
// predicts the other models: Lotka-Volterra and SIR populations of width 5 %
// (3.7 / 4.4 % of their steps, but kernels of 0.21 / 0.17 ms) run 3.5 % faster
// with the block-local sort (profiles/r02ac_local_sort.log, r02aq_*).  In SM
// cycles: local while deviation x elapsed cycles < PB200_LOCAL_SORT_CYCLES.
#ifndef PB200_LOCAL_SORT_CYCLES
#define PB200_LOCAL_SORT_CYCLES 22000.0f
#endif
constexpr int LOCAL_SORT_MAX = 768;          // vectors per block (the largest block shape)
constexpr int LOCAL_SORT_WORDS = SORT_BINS + LOCAL_SORT_MAX + 64;

// Returns the block-local index of the vector that takes local slot j; m is
// the number of vectors of this block (local index t is vector
// blockIdx.x + gridDim.x * t).
// row of the population that is local vector t of this block
__device__ __forceinline__ int64_t local_row(int t, int per_block) {
#if PB200_LSORT_STRIDED == 2
    // runs of 32 rows dealt to the blocks round-robin: coalesced, and still a
    // sample from all over the population
    return ((static_cast<int64_t>(t >> 5) * gridDim.x + blockIdx.x) << 5) + (t & 31);
#elif PB200_LSORT_STRIDED
    return blockIdx.x + static_cast<int64_t>(gridDim.x) * t;
#else
    return static_cast<int64_t>(blockIdx.x) * per_block + t;
#endif
}

template <class M>
__device__ __forceinline__ int local_cost_sort(const ProblemDev &P, const ModelConsts &mc,
                                               const double *__restrict__ params,
                                               int m, int m_full, int64_t ld, uint32_t *sh, int j) {
    uint32_t *hist = sh, *sorted = sh + SORT_BINS, *red = sh + SORT_BINS + LOCAL_SORT_MAX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    const int t = threadIdx.x;
    const bool have = t < m;
    for (int b = t; b < SORT_BINS; b += blockDim.x) hist[b] = 0u;
    float key = 0.0f;
    if (have) {
        double x[PB200_MAX_PARAMS];
#pragma unroll
        for (int k = 0; k < M::NP + M::NO; ++k)
            x[k] = (k < P.n_params) ? params[local_row(t, m_full) * ld + k] : 0.0;
        double early;
        if (!objective_precheck(P, x, &early)) {      // decided vectors keep key 0
            M model;
            double y[M::NS], t0;
            model.load(x, mc, y, P.n_times > 0 ? P.series[0] : 0.0, &t0);
            key = static_cast<float>(model.cost_key(y));
            if (!(key >= 0.0f)) key = 0.0f;            // NaN
            key = fminf(key, 3.0e38f);
        }
    }
    // keys are >= 0: their bit patterns order like the values
    uint32_t hi = have ? __float_as_uint(key) : 0u;
    uint32_t lo = have ? ~__float_as_uint(key) : 0u;
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_max_sync(0xffffffffu, lo);
    if (lane == 0) { red[warp] = hi; red[32 + warp] = lo; }
    __syncthreads();
    hi = lane < nwarps ? red[lane] : 0u;
    lo = lane < nwarps ? red[32 + lane] : 0u;
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_max_sync(0xffffffffu, lo);
    const float kmin = __uint_as_float(~lo), kmax = __uint_as_float(hi);
    const float scale = kmax > kmin ? static_cast<float>(SORT_BINS) / (kmax - kmin) : 0.0f;
    int bin = static_cast<int>((key - kmin) * scale);
    bin = min(max(bin, 0), SORT_BINS - 1);
    uint32_t rank = 0;
    if (have) rank = atomicAdd(&hist[bin], 1u);
    __syncthreads();
    // exclusive scan of the bins in place: threads 0..127 take eight bins each
    constexpr int PER = SORT_BINS / 128;
    uint32_t cnt[PER], run = 0, inc = 0;
    if (t < 128) {
        const uint4 a = reinterpret_cast<const uint4 *>(hist)[t * 2];
        const uint4 b = reinterpret_cast<const uint4 *>(hist)[t * 2 + 1];
        cnt[0] = a.x; cnt[1] = a.y; cnt[2] = a.z; cnt[3] = a.w;
        cnt[4] = b.x; cnt[5] = b.y; cnt[6] = b.z; cnt[7] = b.w;
#pragma unroll
        for (int k = 0; k < PER; ++k) run += cnt[k];
        inc = run;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) red[warp] = inc;
    }
    __syncthreads();
    if (t < 128) {
        uint32_t before = inc - run;
        for (int w = 0; w < warp; ++w) before += red[w];
#pragma unroll
        for (int k = 0; k < PER; ++k) { hist[t * PER + k] = before; before += cnt[k]; }
    }
    __syncthreads();
    const uint32_t pos = hist[bin] + rank;
    if (have) sorted[pos] = static_cast<uint32_t>(t);
    __syncthreads();
    return static_cast<int>(sorted[min(j, m - 1)]);
}

template <class M, bool SMEM, bool TRAJ, bool MIG, bool WIDE, int DENSE = 0>
__global__ void __launch_bounds__(DENSE ? dense_threads(DENSE)
                                        : (MIG ? 512 : LaunchShape<M>::THREADS),
                                  (MIG || DENSE) ? 1 : LaunchShape<M>::BLOCKS)
ode_kernel(const __grid_constant__ ProblemDev P,
           const __grid_constant__ ModelConsts mc,
           const double *__restrict__ params, int64_t n, int64_t ld,
           double *__restrict__ scores, double *__restrict__ traj,
           int32_t *__restrict__ steps_out, const int32_t *__restrict__ perm,
           const SolverDev S, const MigDev mig, const ScatterDev sc,
           const LocalSortDev ls) {
    constexpr int NS = M::NS, NO = M::NO, OUT0 = M::OUT0;
    constexpr uint32_t ROWB = (1 + NO) * 8;          // bytes per series row
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int U = PB200_PIPE_U;
    constexpr bool LSORT = (MIG || DENSE != 0) && !TRAJ;     // one-block-per-SM scoring instances
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ long long hint_t0;                 // LSORT: clock at the start (block 0)
    __shared__ float hint_sum[3];                 // LSORT: attempts, vectors, |deviations|
    if constexpr (LSORT) {
        if (threadIdx.x == 0) {
            hint_t0 = clock64();
            hint_sum[0] = hint_sum[1] = hint_sum[2] = 0.0f;
        }
    }
    // role of this warp in the migrating layout
    enum { WORKER = 0, SOURCE = 1, RECEIVER = 2 };
    int role = WORKER, mig_j = 0;

    Series<SMEM> series;
    series.gbase = P.series;
    series.sbase = 0;
    if (SMEM) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(P.n_times) + PB200_SENTINEL_ROWS) * ROWB + 15) & ~15ll);
        stage_series_begin(smem, P.series, bytes, &bar);     // waited for below
    }

    // lanes past the end of the population shadow the last vector and write
    // nothing: whole warps stay convergent for the votes below
    int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int lj = static_cast<int>(threadIdx.x);           // block-local slot (block-local sort)
    bool filler = false;
    if (MIG) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        int vw = warp;                                // warp whose vectors I hold
        if (warp >= mig.q4 + 2 && warp < mig.q4 + 2 + mig.r) {
            role = RECEIVER; mig_j = warp - mig.q4 - 2; vw = mig.q4 + mig_j;
        } else if (warp >= mig.q4 + mig.r) {
            filler = true; vw = 0;                    // filler warp (r = 1): idles
        } else if (warp >= mig.q4) {
            role = SOURCE; mig_j = warp - mig.q4;
        }
#if PB200_MIG_STRIDED
        // warp v of block b takes warp (v * blocks + b) of the launch order: after
        // the cost sort a block would otherwise hold 12-14 NEIGHBOURING warps, i.e.
        // SMs with uniformly cheap or uniformly expensive vectors
        gi = (static_cast<int64_t>(vw) * gridDim.x + blockIdx.x) * 32 + lane;
#else
        gi = (static_cast<int64_t>(blockIdx.x) * (mig.q4 + mig.r) + vw) * 32 + lane;
#endif
        lj = vw * 32 + lane;
    }
    // the block's own vectors in order of cost (see local_cost_sort)
    int li = -1, l_full = 0;
    if constexpr (LSORT) {
        if (ls.on) {
            l_full = MIG ? (mig.q4 + mig.r) * 32 : static_cast<int>(blockDim.x);
#if PB200_LSORT_STRIDED == 2
            // valid local vectors are a prefix (rows grow with the local index)
            const int m = __syncthreads_count(static_cast<int>(threadIdx.x) < l_full &&
                                              local_row(threadIdx.x, l_full) < n);
#elif PB200_LSORT_STRIDED
            // vectors b, b + G, ...: n / G of them, one more in the first n % G blocks
            const int64_t q = n / gridDim.x, rem = n % gridDim.x, b = blockIdx.x;
            const int m = static_cast<int>(q + (b < rem ? 1 : 0));
            const int64_t first_slot = b * q + (b < rem ? b : rem);
#else
            const int64_t first_slot = static_cast<int64_t>(blockIdx.x) * l_full;
            const int m = n - first_slot < l_full ? static_cast<int>(n - first_slot) : l_full;
#endif
            li = local_cost_sort<M>(P, mc, params, m, l_full, ld,
                                    reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(smem) + ls.scratch_off),
                                    lj);
            // position in the launch order: 32 consecutive ones per warp (put_score)
#if PB200_LSORT_STRIDED == 2
            gi = lj < m ? local_row(lj, l_full) : n;
#else
            gi = lj < m ? first_slot + lj : n;
#endif
        }
    }
    if (SMEM) {
        stage_series_wait(&bar);
        series.sbase = launder_u32(smem_u32(smem));
    }
#if PB200_TIMELINE
    const long long tl_start = clock64();
#endif
#if PB200_STAGGER
    if (MIG) {
        // warps of one scheduler start a quarter of an attempt apart
        const long long until = clock64() + static_cast<long long>((threadIdx.x >> 7) & 3) * PB200_STAGGER;
        while (clock64() < until) { }
    }
#endif
    // the role is the same for a whole warp; taking it from a vote tells the
    // compiler so (branches on it stay uniform, constants stay in uniform
    // registers)
    const bool is_source = MIG && __all_sync(FULL, role == SOURCE);
    const bool is_receiver = MIG && __all_sync(FULL, role == RECEIVER);
    const bool live = gi < n && !filler;
    // `perm` (optional) lists the vectors in order of a cost key, so that the
    // lanes of a warp integrate similar dynamics (see sort_* below)
    const int64_t slot = gi < n ? gi : n - 1;
    const int64_t i = (LSORT && li >= 0) ? local_row(li, l_full) : (perm ? perm[slot] : slot);

    const int nt = P.n_times;
    const int n_read = TRAJ ? P.n_model_params : P.n_params;
    double x[PB200_MAX_PARAMS];
#pragma unroll
    for (int k = 0; k < M::NP + NO; ++k)
        x[k] = (k < n_read) ? params[i * ld + k] : 0.0;

    bool idle = !live;
    if (!TRAJ) {
        double early;
        if (objective_precheck(P, x, &early)) {
            if (live && !is_receiver) {
                put_score(sc, scores, slot, i, early);
                if (steps_out) steps_out[i] = 0;
            }
            idle = true;
        }
    }

    M model;
    double y[NS], t;
    model.load(x, mc, y, nt > 0 ? series.at(0) : 0.0, &t);

    double ss[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) ss[o] = 0.0;

    // `off` is the byte offset of the next observation row.  The packed series
    // ends with sentinel rows whose time is +inf, so reading a few rows ahead
    // needs no bounds check.
    const uint32_t end = static_cast<uint32_t>(nt) * ROWB;
    uint32_t off = 0;
    double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * NO : nullptr;
    const bool traj_vec2 = TRAJ && (reinterpret_cast<uintptr_t>(traj) & 15) == 0;

    // observations at (or before) the initial time see the initial state
    if (!idle) {
        double t_out = series.at(0);
        while (t_out <= t) {
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (TRAJ) my_traj[off / ROWB * NO + o] = y[OUT0 + o];
                const double r = series.at(off + 8 + 8 * o) - y[OUT0 + o];
                ss[o] = fma(r, r, ss[o]);
            }
            off += ROWB;
            t_out = series.at(off);
        }
    } else {
        off = end;
    }

    const double half_rtol = S.half_rtol, atol = S.atol;
    double K1[NS], K2[NS], K3[NS], K4[NS], K5[NS], K6[NS], K7[NS];
    double yn[NS], w[NS];
    if constexpr (M::TIMED) model.rhs_t(y, K1, t); else model.rhs(y, K1);

    // ---- first step (Hairer's hinit, as scipy's select_initial_step) ----
    double h = S.h0;
    if (!(h > 0.0)) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double sc = fma(S.rtol, fabs(y[j]), atol);
            const double a0 = y[j] / sc, a1 = K1[j] / sc;
            d0 = fma(a0, a0, d0);
            d1 = fma(a1, a1, d1);
        }
        d0 = sqrt(d0 * (1.0 / NS));
        d1 = sqrt(d1 * (1.0 / NS));
        double h_a = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
#pragma unroll
        for (int j = 0; j < NS; ++j) w[j] = fma(h_a, K1[j], y[j]);
        if constexpr (M::TIMED) model.rhs_t(w, K2, t + h_a); else model.rhs(w, K2);
        double d2 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double sc = fma(S.rtol, fabs(y[j]), atol);
            const double a2 = (K2[j] - K1[j]) / sc;
            d2 = fma(a2, a2, d2);
        }
        d2 = sqrt(d2 * (1.0 / NS)) / h_a;
        const double dm = fmax(d1, d2);
        const double h_b = (dm <= 1e-15) ? fmax(1e-6, h_a * 1e-3)
                                         : pow(0.01 / dm, 0.2);
        h = fmin(100.0 * h_a, h_b);
    }
    // from here on K1 = h f(y)
#pragma unroll
    for (int j = 0; j < NS; ++j) K1[j] *= h;
    model.scale(h);

    int attempts = 0;
    bool failed = false;
    bool after_reject = false;
    const int max_steps = S.max_steps;
    const double t_last = nt > 0 ? series.at(end - ROWB) : -INFINITY;
    // below this step size the vector is given up (score NaN)
    const double h_min = 1e-14 * fmax(fabs(t_last), 1.0);

    // pending interval: dense-output coefficients of the last accepted step,
    // theta = tq * p_invh + p_c0, 1 - theta = -tq * p_invh + p_c1
    // (the pending interval always ends at the current time t: a rejected
    // step leaves t where every observation up to it has been consumed)
    double p_invh = 0.0, p_c0 = 0.0, p_c1 = 0.0;
    double p_y[NO], p_r2[NO], p_r3[NO], p_r4[NO], p_r5[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) p_y[o] = p_r2[o] = p_r3[o] = p_r4[o] = p_r5[o] = 0.0;

    // one interpolation of the pending interval at time tq against row `row`
    auto interpolate = [&](double tq, uint32_t row, bool valid) {
        const double th = fma(tq, p_invh, p_c0);
#if PB200_TH1_FROM_TH
        const double th1 = 1.0 - th;
#else
        const double th1 = fma(-tq, p_invh, p_c1);
#endif
        double v[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            // y + th (r2 + th1 (r3 + th (r4 + th1 r5)))
            v[o] = fma(th, fma(th1, fma(th, fma(th1, p_r5[o], p_r4[o]), p_r3[o]), p_r2[o]), p_y[o]);
            const double r = series.at(row + 8 + 8 * o) - v[o];
            acc_sq_if(ss[o], r, valid);
        }
        if (TRAJ) {
            if (valid) {
                double *dst = my_traj + row / ROWB * NO;
                if (NO == 2 && traj_vec2) {
                    // one 16-byte store per time point (a lane's trajectory is
                    // its own 16 KB region: stores never coalesce across lanes)
                    *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[NO - 1]);
                } else {
#pragma unroll
                    for (int o = 0; o < NO; ++o) dst[o] = v[o];
                }
            }
        }
    };

    // hand-over area of extra worker j: MIG_ND rows of 32 doubles, then 4 rows
    // of 32 words (off, attempts, flags, spare)
    constexpr int MIG_ND = 2 * NS + 2 + NO;
    double *mig_d = nullptr;
    uint32_t *mig_w = nullptr;
    if (is_source || is_receiver) {
        mig_d = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + mig.buf_off) +
                mig_j * (MIG_ND + 2) * 32 + (threadIdx.x & 31);
        mig_w = reinterpret_cast<uint32_t *>(mig_d - (threadIdx.x & 31) + MIG_ND * 32) +
                (threadIdx.x & 31);
    }
    if (is_receiver) {
        // parked until the extra worker has written its state
        asm volatile("bar.sync %0, 64;" ::"r"(1 + mig_j) : "memory");
#pragma unroll
        for (int j = 0; j < NS; ++j) { y[j] = mig_d[j * 32]; K1[j] = mig_d[(NS + j) * 32]; }
        t = mig_d[2 * NS * 32];
        h = mig_d[(2 * NS + 1) * 32];
#pragma unroll
        for (int o = 0; o < NO; ++o) ss[o] = mig_d[(2 * NS + 2 + o) * 32];
        off = mig_w[0];
        attempts = static_cast<int>(mig_w[32]);
        failed = (mig_w[64] & 1u) != 0;
        after_reject = (mig_w[64] & 2u) != 0;
        model.scale(h);
    }

    // K = h f(w) of one stage; models with a time-dependent right-hand side
    // (TIMED) also get the stage time t + c h
    auto stage = [&](const double *w_, double *K_, double c) {
        if constexpr (M::TIMED) model.hrhs_t(w_, K_, fma(c, h, t));
        else model.hrhs(w_, K_);
    };

    // the eight-state model's step is 150 KB of code: unrolling it would only
    // thrash the instruction cache (0.180 -> 0.159 s at population 65 536)
    constexpr int kLoopUnroll = M::NS >= 8 ? 1 : (M::NS >= 6 ? PB200_BIG_UNROLL : PB200_LOOP_UNROLL);
#pragma unroll kLoopUnroll
    for (;;) {
        const bool fin = (off >= end) || failed;
        if (__all_sync(FULL, fin)) break;
        // (lanes without a vector sit at off = end from the start: they do not count)
        if (is_source && __any_sync(FULL, !idle && off >= mig.off_mid)) break;
        // an attempt counts (and may fail the vector) only while the solution
        // has not reached the last observation; a lane that has consumed all its
        // observations is frozen (its attempts are forced to "rejected", so t, y
        // and h stay put while the slowest lane of the warp finishes)
        const bool working = !fin && t < t_last;
        if (working) {
            if (attempts >= max_steps) failed = true;
            else ++attempts;
        }

        // ---- (A) U interpolation slots (predicated) ----
        {
            double tq[U > 0 ? U : 1];
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < U; ++k) tq[k] = series.at(off + k * ROWB);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool valid = le_bits(tq[k], t);     // monotone in k
                interpolate(tq[k], off + k * ROWB, valid);
                cnt += valid ? 1u : 0u;
            }
            off += cnt * ROWB;
        }
        // ---- (B) six stages on K = h f (K1 is h f(y): first-same-as-last) ----
#pragma unroll
        for (int j = 0; j < NS; ++j) w[j] = fma(A21, K1[j], y[j]);
        stage(w, K2, 0.2);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(A32, K2[j], fma(A31, K1[j], y[j]));
        stage(w, K3, 0.3);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(A43, K3[j], fma(A42, K2[j], fma(A41, K1[j], y[j])));
        stage(w, K4, 0.8);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(A54, K4[j], fma(A53, K3[j], fma(A52, K2[j], fma(A51, K1[j], y[j]))));
        stage(w, K5, 8.0 / 9.0);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(A65, K5[j], fma(A64, K4[j], fma(A63, K3[j], fma(A62, K2[j], fma(A61, K1[j], y[j])))));
        stage(w, K6, 1.0);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            yn[j] = fma(B6, K6[j], fma(B5, K5[j], fma(B4, K4[j], fma(B3, K3[j], fma(B1, K1[j], y[j])))));
        stage(yn, K7, 1.0);
        // ---- embedded error estimate, sum of squared scaled errors ----
        // scale atol + rtol (|y| + |ynew|) / 2: one DADD with |.| modifiers
        // instead of max (DSETP + 64-bit select); the 1e-6-accurate reciprocal
        // seed is enough for a value that is only compared with NS and fed
        // to the controller.
        double sum2 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double e = fma(E7, K7[j], fma(E6, K6[j], fma(E5, K5[j], fma(E4, K4[j], fma(E3, K3[j], E1 * K1[j])))));
#if PB200_SCALE_FROM_Y
            // error weight from the state at the start of the step, as LSODA's
            // EWT = rtol |y| + atol (the reference's integrator)
            const double sc = fma(S.rtol, fabs(y[j]), atol);
#else
            const double sc = fma(half_rtol, fabs(y[j]) + fabs(yn[j]), atol);
#endif
            const double q = e * rcp_seed(sc);
            sum2 = j == 0 ? q * q : fma(q, q, sum2);
        }
        // accepted iff mean square <= 1; NaN compares as "rejected"
        const bool ok = !fin && static_cast<uint64_t>(__double_as_longlong(sum2)) <=
                                    static_cast<uint64_t>(__double_as_longlong(static_cast<double>(NS)));
        // no growth on the step right after a rejection, none on a rejection
#if PB200_CTRL_BITS
        const double facd = step_factor_bits<NS>(sum2, !ok || after_reject);
#else
        float fac = step_factor<NS>(sum2);
        if (!ok || after_reject) fac = fminf(fac, 1.0f);
        const double facd = static_cast<double>(fac);
#endif
        after_reject = !ok;
        // ---- (C) pending observations beyond the U slots ----
        // Smooth problems take few, long steps (Lotka-Volterra: 9 observations
        // per step): while at least half of the lanes still have WIDE or more
        // pending, take them WIDE at a time (independent chains, one vote per
        // trip); the rest, and bursty problems, go one per trip below.
        if constexpr (WIDE) {
            constexpr int NW = PB200_CLOOP_WIDE;
            double t_far = series.at(off + (NW - 1) * ROWB);
            while (__popc(__ballot_sync(FULL, le_bits(t_far, t))) >= 16) {
                uint32_t cnt = 0;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    const double tk = series.at(off + k * ROWB);
                    const bool more = le_bits(tk, t);
                    interpolate(tk, off + k * ROWB, more);
                    cnt += more ? 1u : 0u;
                }
                off += cnt * ROWB;
                t_far = series.at(off + (NW - 1) * ROWB);
            }
        }
        {
            constexpr int W = PB200_CLOOP_WIDTH;
            double tn[W];
#pragma unroll
            for (int k = 0; k < W; ++k) tn[k] = series.at(off + k * ROWB);
            while (__any_sync(FULL, le_bits(tn[0], t))) {
                uint32_t cnt = 0;
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const bool more = le_bits(tn[k], t);
                    interpolate(tn[k], off + k * ROWB, more);
                    cnt += more ? 1u : 0u;
                }
                off += cnt * ROWB;
#pragma unroll
                for (int k = 0; k < W; ++k) tn[k] = series.at(off + k * ROWB);
            }
        }

        // ---- (D) dense-output coefficients of this step (Hairer's contd5
        // on h-scaled increments); junk when rejected, never read (t stays) ----
        const double inv_h = rcp_newton<PB200_RCP_STEPS>(h);
        const double t_new = t + h;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const int j = OUT0 + o;
            const double d = yn[j] - y[j];
            const double b = K1[j] - d;
            p_y[o] = y[j];
            p_r2[o] = d;
            p_r3[o] = b;
            p_r4[o] = (d - K7[j]) - b;
            p_r5[o] = fma(D7, K7[j], fma(D6, K6[j], fma(D5, K5[j], fma(D4, K4[j], fma(D3, K3[j], D1 * K1[j])))));
        }
        p_invh = inv_h;
        p_c0 = -t * inv_h;
#if !PB200_TH1_FROM_TH
        p_c1 = t_new * inv_h;
#endif
        t = ok ? t_new : t;
        h *= facd;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            y[j] = ok ? yn[j] : y[j];
            K1[j] = (ok ? K7[j] : K1[j]) * facd;     // h_new f(y)
        }
        model.scale(h);
        // step size underflow: give up on this vector
        if (working && !ok && __double2hiint(h) < __double2hiint(h_min)) failed = true;
    }

    if (is_source) {
        // empty the pending interval (its coefficients do not travel), then
        // hand the state to the receiver and retire
        double t_next = series.at(off);
        while (__any_sync(FULL, le_bits(t_next, t))) {
            const bool more = le_bits(t_next, t);
            interpolate(t_next, off, more);
            off += more ? ROWB : 0u;
            t_next = series.at(off);
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) { mig_d[j * 32] = y[j]; mig_d[(NS + j) * 32] = K1[j]; }
        mig_d[2 * NS * 32] = t;
        mig_d[(2 * NS + 1) * 32] = h;
#pragma unroll
        for (int o = 0; o < NO; ++o) mig_d[(2 * NS + 2 + o) * 32] = ss[o];
        mig_w[0] = off;
        mig_w[32] = static_cast<uint32_t>(attempts);
        mig_w[64] = (failed ? 1u : 0u) | (after_reject ? 2u : 0u);
        __threadfence_block();
        asm volatile("bar.arrive %0, 64;" ::"r"(1 + mig_j) : "memory");
    } else if (!idle) {
        if (steps_out) steps_out[i] = attempts;
        if (TRAJ) {
            if (failed) {
                const int64_t first = static_cast<int64_t>(off / ROWB) * NO;
                const int64_t total = static_cast<int64_t>(nt) * NO;
                for (int64_t q = first; q < total; ++q) my_traj[q] = NAN;
            }
        } else {
            put_score(sc, scores, slot, i, failed ? NAN : objective_finish(P, ss, x));
        }
    }
#if PB200_TIMELINE
    if (MIG && (threadIdx.x & 31) == 0 && blockIdx.x < 160) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *row = pb200_timeline + (blockIdx.x * 16 + (threadIdx.x >> 5)) * 4;
        row[0] = tl_start; row[1] = clock64(); row[2] = smid; row[3] = role * 1000000 + attempts;
    }
#endif
    // what the next launch's choice of cost sort goes by (see PB200_LOCAL_SORT_CYCLES):
    // block 0 leaves the spread of its vectors' attempted steps, weighted with the
    // kernel's duration, in the problem's hint word
    if constexpr (LSORT) {
        if (ls.hint != nullptr && blockIdx.x == 0) {        // uniform
            const bool counted = !is_source && live && !idle && !failed;
            const float a = counted ? static_cast<float>(attempts) : 0.0f;
            float sa = a, sc_ = counted ? 1.0f : 0.0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sa += __shfl_xor_sync(FULL, sa, o);
                sc_ += __shfl_xor_sync(FULL, sc_, o);
            }
            if ((threadIdx.x & 31) == 0) { atomicAdd(&hint_sum[0], sa); atomicAdd(&hint_sum[1], sc_); }
            __syncthreads();
            const float cnt = fmaxf(hint_sum[1], 1.0f), mean = hint_sum[0] / cnt;
            float dv = counted ? fabsf(a - mean) : 0.0f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dv += __shfl_xor_sync(FULL, dv, o);
            if ((threadIdx.x & 31) == 0) atomicAdd(&hint_sum[2], dv);
            __syncthreads();
            if (threadIdx.x == 0) {
                const float rel = hint_sum[2] / (cnt * fmaxf(mean, 1.0f));
                const float cycles = static_cast<float>(clock64() - hint_t0);
                *ls.hint = rel * cycles < PB200_LOCAL_SORT_CYCLES ? 1 : 2;
            }
        }
    }
    // multi-GPU launches: the last block to get here raises this rank's flags
    if (!TRAJ) exchange_signal(sc);
}

// ======================================================================
// Stiff integrator for the Beeler-Reuter model: the 4-stage Rosenbrock method
// of Kaps & Rentrop in Shampine's parameterisation (order 4, embedded order 3;
// gamma = 1/2), replacing for this model the explicit DP5(4) template above,
// which is stability-bound by the sodium activation gate (7 700 steps for
// 400 ms at any tolerance).  The reference's LSODA switches to BDF on this
// model (pints/toy/_beeler_reuter_model.py:205-222).
//
// * Jacobian.  The system is an arrow: the six gates relax towards
//   voltage-dependent targets (dg/dt = alpha(V)(1-g) - beta(V) g: own diagonal
//   -(alpha+beta) and a V column), the V row is full and polynomial in the
//   gates, the Ca row couples to V, Ca, d, f.  All of it is analytic from
//   factors the right-hand side computes anyway (BeelerReuterModel::Aux): the
//   V column differentiates the twelve rate functions and two currents with
//   their exponentials already at hand (~90 FP64 instructions against ~550
//   for a difference quotient, i.e. another right-hand side).
// * Linear algebra.  (I/(gamma h) - J) x = b: the gates are eliminated by
//   substitution, leaving a 2 x 2 system in (V, Ca); the elimination factors
//   are prepared once per attempt, each of the four solves is ~40 FMAs.
// * Stimulus.  Steps are clipped to land exactly on the edges of the
//   stimulus (k period, k period + length); inside a step it is constant, so
//   the system is autonomous there (no df/dt term).
// * Dense output.  Cubic Hermite on (y_n, f_n, y_n+1, f_n+1); f_n+1 is the
//   next step's first right-hand side, so an attempt costs three right-hand
//   sides (two stages and f_n+1).
// * Error weights atol w_i + rtol |y_i| with w = 1 except 1e-6 for the
//   calcium concentration (its values are 2e-7 .. 7e-6).
// Prototype and convergence table: tools/br_rosenbrock_proto.py.
// ======================================================================
#ifndef PB200_BRR_THREADS
#define PB200_BRR_THREADS 128
#endif
#ifndef PB200_BRR_BLOCKS
#define PB200_BRR_BLOCKS 3
#endif

struct BrArrow {
    // (a I - J) x = b with the gates eliminated: gate_i = pg_i (b_i + cv_i xV)
    double pg[6], cv[6];          // 1 / (a - J_ii), column V of the gate rows
    double rv[6], rc[2];          // J_V,gate_i pg_i ; J_Ca,{d,f} pg_i
    double i11, i12, i21, i22;    // inverse of the reduced 2 x 2 matrix
    __device__ __forceinline__ void solve(const double *b, double *x) const {
        double bV = b[0], bC = b[1];
#pragma unroll
        for (int i = 0; i < 6; ++i) bV = fma(rv[i], b[2 + i], bV);
        bC = fma(rc[0], b[5], bC);
        bC = fma(rc[1], b[6], bC);
        const double xV = fma(i11, bV, i12 * bC);
        x[0] = xV;
        x[1] = fma(i21, bV, i22 * bC);
#pragma unroll
        for (int i = 0; i < 6; ++i) x[2 + i] = pg[i] * fma(cv[i], xV, b[2 + i]);
    }
};

template <bool SMEM, bool TRAJ>
__global__ void __launch_bounds__(PB200_BRR_THREADS, PB200_BRR_BLOCKS)
br_rosenbrock_kernel(const __grid_constant__ ProblemDev P,
                     const __grid_constant__ ModelConsts mc,
                     const double *__restrict__ params, int64_t n, int64_t ld,
                     double *__restrict__ scores, double *__restrict__ traj,
                     int32_t *__restrict__ steps_out, const int32_t *__restrict__ perm,
                     const SolverDev S, const ScatterDev sc) {
    using M = BeelerReuterModel;
    constexpr int NS = 8, NO = 2;
    constexpr uint32_t ROWB = (1 + NO) * 8;
    constexpr uint32_t FULL = 0xffffffffu;
    // Kaps-Rentrop / Shampine
    constexpr double R_GAM = 0.5, R_A21 = 2.0, R_A31 = 48.0 / 25.0, R_A32 = 6.0 / 25.0;
    constexpr double R_C21 = -8.0, R_C31 = 372.0 / 25.0, R_C32 = 12.0 / 5.0;
    constexpr double R_C41 = -112.0 / 125.0, R_C42 = -54.0 / 125.0, R_C43 = -2.0 / 5.0;
    constexpr double R_B1 = 19.0 / 9.0, R_B2 = 0.5, R_B3 = 25.0 / 108.0, R_B4 = 125.0 / 108.0;
    constexpr double R_E1 = 17.0 / 54.0, R_E2 = 7.0 / 36.0, R_E4 = 125.0 / 108.0;      // E3 = 0
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;

    Series<SMEM> series;
    series.gbase = P.series;
    series.sbase = 0;
    if (SMEM) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(P.n_times) + PB200_SENTINEL_ROWS) * ROWB + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series.sbase = launder_u32(smem_u32(smem));
    }
    const int64_t gi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool live = gi < n;
    const int64_t slot = live ? gi : n - 1;
    const int64_t i = perm ? perm[slot] : slot;
    const int nt = P.n_times;
    const int n_read = TRAJ ? P.n_model_params : P.n_params;
    double x[PB200_MAX_PARAMS];
#pragma unroll
    for (int k = 0; k < M::NP + NO; ++k)
        x[k] = (k < n_read) ? params[i * ld + k] : 0.0;

    bool idle = !live;
    if (!TRAJ) {
        double early;
        if (objective_precheck(P, x, &early)) {
            if (live) {
                put_score(sc, scores, slot, i, early);
                if (steps_out) steps_out[i] = 0;
            }
            idle = true;
        }
    }
    M model;
    double y[NS], t;
    model.load(x, mc, y, nt > 0 ? series.at(0) : 0.0, &t);
    double ss[NO] = {0.0, 0.0};
    const uint32_t end = static_cast<uint32_t>(nt) * ROWB;
    uint32_t off = 0;
    double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * NO : nullptr;
    if (!idle) {
        double t_out = series.at(0);
        while (t_out <= t) {               // observations at the initial time
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (TRAJ) my_traj[off / ROWB * NO + o] = y[o];
                const double r = series.at(off + 8 + 8 * o) - y[o];
                ss[o] = fma(r, r, ss[o]);
            }
            off += ROWB;
            t_out = series.at(off);
        }
    } else {
        off = end;
    }
    const double rtol = S.rtol, atol = S.atol;
    const double t_last = nt > 0 ? series.at(end - ROWB) : -INFINITY;
    const double h_min = 1e-13 * fmax(fabs(t_last), 1.0);
    double h = S.h0 > 0.0 ? S.h0 : 1e-3;
    double stim = model.stimulus(t);
    double f0[NS];
    M::Aux aux;
    bool have_f0 = false;
    int attempts = 0;
    bool failed = false;
    const int max_steps = S.max_steps;

    for (;;) {
        const bool fin = (off >= end) || failed;
        if (__all_sync(FULL, fin)) break;
        const bool working = !fin;
        if (working) {
            if (attempts >= max_steps) failed = true;
            else ++attempts;
        }
        // ---- clip the step at the next edge of the stimulus ----
        double ph = fma(-model.period, floor(t * model.inv_period), t);
        if (ph < 0.0) ph += model.period;
        if (ph >= model.period) ph -= model.period;
        const double edge = ph < model.length ? (t - ph) + model.length : (t - ph) + model.period;
        const bool clipped = edge - t < h;
        const double hs = clipped ? edge - t : h;
        // ---- f(y) and the pieces of the Jacobian ----
        if (!__all_sync(FULL, have_f0)) {
            // (first attempt, and lanes whose last step ended on an edge: all
            // lanes recompute; a lane that had f0 gets the same value again)
            stim = model.stimulus(fma(0.5, hs, t));
            model.template rhs_core<true>(y, f0, stim, &aux);
            have_f0 = true;
        }
        BrArrow ar;
        const double cvV = aux.cv[0], cvC = aux.cv[1];
#pragma unroll
        for (int k = 0; k < 6; ++k) ar.cv[k] = aux.cv[2 + k];
        const double ih = rcp_newton<2>(hs);
        const double a = ih * (1.0 / R_GAM);
        {
            const double V = y[0], Cai = y[1], m = y[2], hh = y[3], j = y[4], d = y[5], ff = y[6];
            const double dn = V - model.e_na, icm = model.inv_cm;
            const double dlog = 13.0287 * rcp_newton<2>(Cai);     // d(V - E_Ca) / dCa
#pragma unroll
            for (int k = 0; k < 6; ++k) ar.pg[k] = rcp_newton<2>(a + aux.rate[k]);
            const double jVm = -icm * model.gNa * 3.0 * m * m * hh * j * dn;
            const double jVh = -icm * model.gNa * aux.m3 * j * dn;
            const double jVj = -icm * model.gNa * aux.m3 * hh * dn;
            const double jVd = -icm * model.gCa * ff * aux.dECa;
            const double jVf = -icm * model.gCa * d * aux.dECa;
            const double jVx = -icm * model.gx1 * aux.xf;
            const double jVC = -icm * model.gCa * d * ff * dlog;
            const double jCC = -1e-7 * model.gCa * d * ff * dlog - 0.07;
            const double jCd = -1e-7 * model.gCa * ff * aux.dECa;
            const double jCf = -1e-7 * model.gCa * d * aux.dECa;
            ar.rv[0] = jVm * ar.pg[0]; ar.rv[1] = jVh * ar.pg[1]; ar.rv[2] = jVj * ar.pg[2];
            ar.rv[3] = jVd * ar.pg[3]; ar.rv[4] = jVf * ar.pg[4]; ar.rv[5] = jVx * ar.pg[5];
            ar.rc[0] = jCd * ar.pg[3]; ar.rc[1] = jCf * ar.pg[4];
            double aVV = a - cvV, aCV = -cvC;
#pragma unroll
            for (int k = 0; k < 6; ++k) aVV = fma(-ar.rv[k], ar.cv[k], aVV);
            aCV = fma(-ar.rc[0], ar.cv[3], aCV);
            aCV = fma(-ar.rc[1], ar.cv[4], aCV);
            const double aVC = -jVC, aCC = a - jCC;
            const double idet = rcp_newton<2>(aVV * aCC - aVC * aCV);
            ar.i11 = aCC * idet; ar.i12 = -aVC * idet;
            ar.i21 = -aCV * idet; ar.i22 = aVV * idet;
        }
        // ---- four stages ----
        double g1[NS], g2[NS], g3[NS], g4[NS], w[NS], fw[NS], b[NS];
        ar.solve(f0, g1);
#pragma unroll
        for (int k = 0; k < NS; ++k) w[k] = fma(R_A21, g1[k], y[k]);
        model.template rhs_core<false>(w, fw, stim, nullptr);
#pragma unroll
        for (int k = 0; k < NS; ++k) b[k] = fma(R_C21 * ih, g1[k], fw[k]);
        ar.solve(b, g2);
#pragma unroll
        for (int k = 0; k < NS; ++k) w[k] = fma(R_A32, g2[k], fma(R_A31, g1[k], y[k]));
        model.template rhs_core<false>(w, fw, stim, nullptr);
#pragma unroll
        for (int k = 0; k < NS; ++k) b[k] = fma(R_C32 * ih, g2[k], fma(R_C31 * ih, g1[k], fw[k]));
        ar.solve(b, g3);
#pragma unroll
        for (int k = 0; k < NS; ++k)
            b[k] = fma(R_C43 * ih, g3[k], fma(R_C42 * ih, g2[k], fma(R_C41 * ih, g1[k], fw[k])));
        ar.solve(b, g4);
        double yn[NS], sum2 = 0.0;
        constexpr double wt[NS] = {1.0, 1e-6, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            yn[k] = fma(R_B4, g4[k], fma(R_B3, g3[k], fma(R_B2, g2[k], fma(R_B1, g1[k], y[k]))));
            const double e = fma(R_E4, g4[k], fma(R_E2, g2[k], R_E1 * g1[k]));
            const double q = e * rcp_newton<1>(fma(rtol, fabs(y[k]), atol * wt[k]));
            sum2 = fma(q, q, sum2);
        }
        // accepted iff the mean square is <= 1 (NaN: rejected); a negative
        // calcium concentration is rejected as well (log in the next f)
        const bool ok = working && sum2 <= static_cast<double>(NS) && yn[1] > 0.0;
        // 0.9 err^(-1/4) in [0.2, 5]: err^2 = sum2 / NS
        float fac;
        {
            const float n2 = static_cast<float>(sum2 * (1.0 / NS));
            float l, f;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(n2));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(-0.125f * l));
            fac = fminf(fmaxf(0.9f * f, 0.2f), 5.0f);
            if (!(sum2 == sum2) || !(yn[1] > 0.0)) fac = 0.2f;
            if (!ok) fac = fminf(fac, 1.0f);
        }
        const double t_new = clipped ? edge : t + hs;
        // ---- f(y_new): dense output now, first right-hand side of the next step ----
        // (junk for a rejected lane, never used)
        double fn[NS];
        M::Aux aux_n;
        model.template rhs_core<true>(yn, fn, stim, &aux_n);
        // cubic Hermite on [t, t_new]: y + th (d0 + th (c2 + th c3))
        double c1[NO], c2[NO], c3[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            const double d = yn[o] - y[o], d0 = hs * f0[o], d1 = hs * fn[o];
            c1[o] = d0;
            c2[o] = fma(3.0, d, fma(-2.0, d0, -d1));
            c3[o] = fma(-2.0, d, d0 + d1);
        }
        double t_obs = series.at(off);
        while (__any_sync(FULL, ok && le_bits(t_obs, t_new))) {
            const bool valid = ok && le_bits(t_obs, t_new);
            const double th = (t_obs - t) * ih;
            double v[NO];
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                v[o] = fma(th, fma(th, fma(th, c3[o], c2[o]), c1[o]), y[o]);
                const double r = series.at(off + 8 + 8 * o) - v[o];
                acc_sq_if(ss[o], r, valid);
            }
            if (TRAJ) {
                if (valid) {
                    double *dst = my_traj + off / ROWB * NO;
#pragma unroll
                    for (int o = 0; o < NO; ++o) dst[o] = v[o];
                }
            }
            off += valid ? ROWB : 0u;
            t_obs = series.at(off);
        }
        // ---- accept / reject ----
        if (ok) {
#pragma unroll
            for (int k = 0; k < NS; ++k) { y[k] = yn[k]; f0[k] = fn[k]; }
            aux = aux_n;
            t = t_new;
            // the stimulus changes at an edge: the next step starts from a new f
            have_f0 = !clipped;
        }
        const double h_try = hs * static_cast<double>(fac);
        // an accepted, clipped step says nothing against the step it was cut from
        h = (ok && clipped) ? fmax(h, h_try) : h_try;
        if (working && !ok && hs < h_min) failed = true;
    }

    if (!idle) {
        if (steps_out) steps_out[i] = attempts;
        if (TRAJ) {
            if (failed) {
                const int64_t first = static_cast<int64_t>(off / ROWB) * NO;
                const int64_t total = static_cast<int64_t>(nt) * NO;
                for (int64_t q = first; q < total; ++q) my_traj[q] = NAN;
            }
        } else {
            put_score(sc, scores, slot, i, failed ? NAN : objective_finish(P, ss, x));
        }
    }
    if (!TRAJ) exchange_signal(sc);
}

// ======================================================================
// Pair kernel: TWO lanes per parameter vector, one per state component
// (FitzHugh-Nagumo, Lotka-Volterra: NS = NO = 2).  A population of 65 536 on
// one B200 is only 2048 one-lane warps = 3.5 per scheduler, too few to cover
// the 9-cycle dependent-DFMA latency of the serial stage chain.  Splitting a
// vector over a lane pair doubles the warps (6.9 per scheduler, one wave of
// 7 blocks x 128 threads per SM) and halves every warp's instruction stream:
// stage sums, error term, dense-output coefficients and interpolation are
// per component anyway; only the right-hand side needs the partner's stage
// value (one 64-bit lane exchange per stage) and the error norm one more.
// Controller arithmetic is duplicated in both lanes and bitwise identical.
// ======================================================================
__device__ __forceinline__ double pair_other(double v) {
    return __shfl_xor_sync(0xffffffffu, v, 1);
}

// Lane view of a two-state model: lane s integrates component s with
// K_s = hrhs(w_s, w_other); the constants hold the step size and are updated
// multiplicatively (hC *= h_new / h) together with the FSAL slope.
struct FhnLane {
    static constexpr int NP = 3, NC = 4;
    // s = 0 (V): f = (c3 V^2 + c) V + c R        C = {c3, c, c, 0}
    // s = 1 (R): f = (-b/c) R + ((-1/c) V + a/c)  C = {0, -b/c, -1/c, a/c}
    __device__ __forceinline__ static void load(const double *x, int s, double *C) {
        const double a = x[0], b = x[1], c = x[2];
        const double nic = -1.0 / c;
        C[0] = s ? 0.0 : c * (-1.0 / 3.0);
        C[1] = s ? nic * b : c;
        C[2] = s ? nic : c;
        C[3] = s ? -nic * a : 0.0;
    }
    __device__ __forceinline__ static double hrhs(const double *C, double ws, double wo) {
        return fma(fma(C[0], ws * ws, C[1]), ws, fma(C[2], wo, C[3]));
    }
};

struct LvLane {
    static constexpr int NP = 4, NC = 2;
    // s = 0 (x): f = x (a - b y)     C = {-b, a}
    // s = 1 (y): f = y (d x - c)     C = {d, -c}
    __device__ __forceinline__ static void load(const double *x, int s, double *C) {
        C[0] = s ? x[3] : -x[1];
        C[1] = s ? -x[2] : x[0];
    }
    __device__ __forceinline__ static double hrhs(const double *C, double ws, double wo) {
        return ws * fma(C[0], wo, C[1]);
    }
};

#ifndef PB200_PAIR_U
#define PB200_PAIR_U 3
#endif
#ifndef PB200_PAIR_THREADS
#define PB200_PAIR_THREADS 128
#endif
#ifndef PB200_PAIR_MINBLOCKS
#define PB200_PAIR_MINBLOCKS 7
#endif
static_assert(PB200_PAIR_U + 1 <= PB200_SENTINEL_ROWS, "not enough sentinel rows");

template <class L, bool SMEM, bool TRAJ>
__global__ void __launch_bounds__(PB200_PAIR_THREADS, PB200_PAIR_MINBLOCKS)
ode_pair_kernel(const __grid_constant__ ProblemDev P,
                const __grid_constant__ ModelConsts mc,
                const double *__restrict__ params, int64_t n, int64_t ld,
                double *__restrict__ scores, double *__restrict__ traj,
                int32_t *__restrict__ steps_out, const int32_t *__restrict__ perm,
                const SolverDev S, const ScatterDev sc) {
    constexpr int NO = 2, NC = L::NC;
    constexpr uint32_t ROWB = (1 + NO) * 8;
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int U = PB200_PAIR_U;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;

    Series<SMEM> series;
    series.gbase = P.series;
    series.sbase = 0;
    if (SMEM) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(P.n_times) + PB200_SENTINEL_ROWS) * ROWB + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series.sbase = launder_u32(smem_u32(smem));
    }

    const int64_t gt = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int s = static_cast<int>(gt & 1);          // my state component
    const bool live = (gt >> 1) < n;
    const int64_t slot = live ? (gt >> 1) : n - 1;
    const int64_t i = perm ? perm[slot] : slot;

    const int nt = P.n_times;
    const int n_read = TRAJ ? P.n_model_params : P.n_params;
    double x[PB200_MAX_PARAMS];
#pragma unroll
    for (int k = 0; k < L::NP + NO; ++k)
        x[k] = (k < n_read) ? params[i * ld + k] : 0.0;

    bool idle = !live;
    if (!TRAJ) {
        double early;
        if (objective_precheck(P, x, &early)) {
            if (live && s == 0) {
                put_score(sc, scores, slot, i, early);
                if (steps_out) steps_out[i] = 0;
            }
            idle = true;
        }
    }

    double C[NC];                       // step-scaled constants (h = 1 for now)
    L::load(x, s, C);
    double y = s ? mc.c[1] : mc.c[0];
    double t = 0.0;
    double ss = 0.0;

    const uint32_t vofs = 8 + 8 * s;    // my observed value inside a row
    const uint32_t end = static_cast<uint32_t>(nt) * ROWB;
    uint32_t off = 0;
    double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * NO + s : nullptr;

    if (!idle) {
        double t_out = series.at(0);
        while (t_out <= t) {
            if (TRAJ) my_traj[off / ROWB * NO] = y;
            const double r = series.at(off + vofs) - y;
            ss = fma(r, r, ss);
            off += ROWB;
            t_out = series.at(off);
        }
    } else {
        off = end;
    }

    const double half_rtol = S.half_rtol, atol = S.atol;
    double K1 = L::hrhs(C, y, pair_other(y));            // f(y), h = 1

    // ---- first step (Hairer's hinit, as scipy's select_initial_step) ----
    double h = S.h0;
    if (!(h > 0.0)) {
        const double sc = fma(S.rtol, fabs(y), atol);
        const double a0 = y / sc, a1 = K1 / sc;
        double d0 = a0 * a0, d1 = a1 * a1;
        d0 = sqrt((d0 + pair_other(d0)) * 0.5);
        d1 = sqrt((d1 + pair_other(d1)) * 0.5);
        const double h_a = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        const double w = fma(h_a, K1, y);
        const double k2 = L::hrhs(C, w, pair_other(w));
        const double a2 = (k2 - K1) / sc;
        double d2 = a2 * a2;
        d2 = sqrt((d2 + pair_other(d2)) * 0.5) / h_a;
        const double dm = fmax(d1, d2);
        const double h_b = (dm <= 1e-15) ? fmax(1e-6, h_a * 1e-3)
                                         : pow(0.01 / dm, 0.2);
        h = fmin(100.0 * h_a, h_b);
    }
    K1 *= h;
#pragma unroll
    for (int k = 0; k < NC; ++k) C[k] *= h;

    int attempts = 0;
    bool failed = false;
    bool after_reject = false;
    const int max_steps = S.max_steps;
    const double t_last = nt > 0 ? series.at(end - ROWB) : -INFINITY;
    const double h_min = 1e-14 * fmax(fabs(t_last), 1.0);

    double p_invh = 0.0, p_c0 = 0.0, p_c1 = 0.0;
    double p_y = 0.0, p_r2 = 0.0, p_r3 = 0.0, p_r4 = 0.0, p_r5 = 0.0;

    auto interpolate = [&](double tq, uint32_t row, bool valid) {
        const double th = fma(tq, p_invh, p_c0);
#if PB200_TH1_FROM_TH
        const double th1 = 1.0 - th;
#else
        const double th1 = fma(-tq, p_invh, p_c1);
#endif
        const double v = fma(th, fma(th1, fma(th, fma(th1, p_r5, p_r4), p_r3), p_r2), p_y);
        const double r = series.at(row + vofs) - v;
        if (TRAJ) { if (valid) my_traj[row / ROWB * NO] = v; }
        acc_sq_if(ss, r, valid);
    };

    constexpr int kLoopUnroll = PB200_LOOP_UNROLL;
#pragma unroll kLoopUnroll
    for (;;) {
        const bool fin = (off >= end) || failed;
        if (__all_sync(FULL, fin)) break;
        // an attempt counts (and may fail the vector) only while the solution
        // has not reached the last observation; a lane that has consumed all its
        // observations is frozen (its attempts are forced to "rejected", so t, y
        // and h stay put while the slowest lane of the warp finishes)
        const bool working = !fin && t < t_last;
        if (working) {
            if (attempts >= max_steps) failed = true;
            else ++attempts;
        }
        // ---- (C) pending observations beyond the U slots ----
        {
            double t_ahead = series.at(off + U * ROWB);
            while (__any_sync(FULL, le_bits(t_ahead, t))) {
                const bool more = le_bits(t_ahead, t);
                interpolate(series.at(off), off, more);
                off += more ? ROWB : 0u;
                t_ahead = series.at(off + U * ROWB);
            }
        }
        // ---- (A) U interpolation slots (predicated) ----
        {
            double tq[U > 0 ? U : 1];
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < U; ++k) tq[k] = series.at(off + k * ROWB);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool valid = le_bits(tq[k], t);
                interpolate(tq[k], off + k * ROWB, valid);
                cnt += valid ? 1u : 0u;
            }
            off += cnt * ROWB;
        }
        // ---- (B) six stages, my component; partner's stage value by shuffle ----
        double w = fma(A21, K1, y);
        const double K2 = L::hrhs(C, w, pair_other(w));
        w = fma(A32, K2, fma(A31, K1, y));
        const double K3 = L::hrhs(C, w, pair_other(w));
        w = fma(A43, K3, fma(A42, K2, fma(A41, K1, y)));
        const double K4 = L::hrhs(C, w, pair_other(w));
        w = fma(A54, K4, fma(A53, K3, fma(A52, K2, fma(A51, K1, y))));
        const double K5 = L::hrhs(C, w, pair_other(w));
        w = fma(A65, K5, fma(A64, K4, fma(A63, K3, fma(A62, K2, fma(A61, K1, y)))));
        const double K6 = L::hrhs(C, w, pair_other(w));
        const double yn = fma(B6, K6, fma(B5, K5, fma(B4, K4, fma(B3, K3, fma(B1, K1, y)))));
        const double K7 = L::hrhs(C, yn, pair_other(yn));
        // ---- error: my squared scaled component + the partner's ----
        const double e = fma(E7, K7, fma(E6, K6, fma(E5, K5, fma(E4, K4, fma(E3, K3, E1 * K1)))));
        const double sc = fma(half_rtol, fabs(y) + fabs(yn), atol);
        const double q = e * rcp_seed(sc);
        const double q2 = q * q;
        const double sum2 = q2 + pair_other(q2);          // commutative: same bits in both lanes
        const bool ok = !fin && static_cast<uint64_t>(__double_as_longlong(sum2)) <=
                                    static_cast<uint64_t>(__double_as_longlong(2.0));
#if PB200_CTRL_BITS
        const double facd = step_factor_bits<2>(sum2, !ok || after_reject);
#else
        float fac = step_factor<2>(sum2);
        if (!ok || after_reject) fac = fminf(fac, 1.0f);
        const double facd = static_cast<double>(fac);
#endif
        after_reject = !ok;
        // ---- (D) dense-output coefficients, accept / reject by selects ----
        const double inv_h = rcp_newton<PB200_RCP_STEPS>(h);
        const double t_new = t + h;
        {
            const double d = yn - y;
            const double b = K1 - d;
            p_y = y;
            p_r2 = d;
            p_r3 = b;
            p_r4 = (d - K7) - b;
            p_r5 = fma(D7, K7, fma(D6, K6, fma(D5, K5, fma(D4, K4, fma(D3, K3, D1 * K1)))));
        }
        p_invh = inv_h;
        p_c0 = -t * inv_h;
#if !PB200_TH1_FROM_TH
        p_c1 = t_new * inv_h;
#endif
        t = ok ? t_new : t;
        h *= facd;
        y = ok ? yn : y;
        K1 = (ok ? K7 : K1) * facd;
#pragma unroll
        for (int k = 0; k < NC; ++k) C[k] *= facd;
        if (working && !ok && __double2hiint(h) < __double2hiint(h_min)) failed = true;
    }

    const double ss_other = pair_other(ss);
    if (idle) {
        // decided before the solve, or past the end of the population
    } else if (TRAJ) {
        if (failed) {
            for (uint32_t row = off; row < end; row += ROWB) my_traj[row / ROWB * NO] = NAN;
        }
        if (s == 0 && steps_out) steps_out[i] = attempts;
    } else if (s == 0) {
        if (steps_out) steps_out[i] = attempts;
        const double both[2] = {ss, ss_other};
        put_score(sc, scores, slot, i, failed ? NAN : objective_finish(P, both, x));
    }
    if (!TRAJ) exchange_signal(sc);
}

// ======================================================================
// Cost-coherent scheduling.  Lanes of one warp run in lock step, so a warp
// pays for the slowest of its 32 vectors in every respect: attempted steps,
// and interpolation passes per step (the lanes of a warp sit at different
// phases of the oscillation, so in every step SOME lane has many pending
// observations).  Grouping vectors with similar dynamics removes most of that:
// on the headline population the interpolation passes per attempt drop from
// 4.6 to 3.5, on a wide prior box the whole kernel gets 30 % faster.  The key
// comes from the model (cost_key): the time-scale parameter c for
// FitzHugh-Nagumo, |f(y0)|^2 (the initial rate of change) otherwise.  A counting
// sort over SORT_BINS linear bins between the smallest and the largest key is
// enough (no order is needed inside a bin) and is one cooperative launch.
// ======================================================================
static int sm_count();
constexpr int SORT_THREADS = 256;
constexpr int SORT_MAX_BLOCKS = 1024;

struct SortHeader {
    uint32_t hist[SORT_BINS];
    uint32_t cursor[SORT_BINS];
    uint32_t block_max[SORT_MAX_BLOCKS];      // per block: max of bits(key)   (keys >= 0)
    uint32_t block_negmax[SORT_MAX_BLOCKS];   // per block: max of ~bits(key), i.e. ~min
};

// One cooperative launch, three phases separated by grid-wide barriers:
//   A  keys (and each block's key range); clear the histogram
//   B  bin index per vector, histogram (shared-memory atomics, then global)
//   C  every block scans the 1024-bin histogram for itself, reserves its range
//      in every bin it touches and scatters its vectors' indices
template <class M>
__global__ void __launch_bounds__(SORT_THREADS)
sort_kernel(const __grid_constant__ ProblemDev P,
            const __grid_constant__ ModelConsts mc,
            const double *__restrict__ params, int64_t n, int64_t ld,
            float *__restrict__ keys, SortHeader *__restrict__ hdr,
            int32_t *__restrict__ perm) {
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t sh_a[SORT_BINS];      // B: histogram      C: scan of the histogram
    __shared__ uint32_t sh_b[SORT_BINS];      // C: this block's count per bin, then its start
    __shared__ uint32_t sh_w[2 * SORT_THREADS / 32];
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- A ----
    uint32_t hi = 0u, lo = 0u;
    for (int64_t i = first; i < n; i += stride) {
        double x[PB200_MAX_PARAMS];
#pragma unroll
        for (int k = 0; k < M::NP + M::NO; ++k)
            x[k] = (k < P.n_params) ? params[i * ld + k] : 0.0;
        float key = 0.0f;
        double early;
        if (!objective_precheck(P, x, &early)) {      // decided vectors keep key 0
            M model;
            double y[M::NS], t0;
            model.load(x, mc, y, P.n_times > 0 ? P.series[0] : 0.0, &t0);
            key = static_cast<float>(model.cost_key(y));
            if (!(key >= 0.0f)) key = 0.0f;            // NaN
            key = fminf(key, 3.0e38f);
        }
        keys[i] = key;
        hi = max(hi, __float_as_uint(key));
        lo = max(lo, ~__float_as_uint(key));
    }
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_max_sync(0xffffffffu, lo);
    if (lane == 0) { sh_w[warp] = hi; sh_w[SORT_THREADS / 32 + warp] = lo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < SORT_THREADS / 32; ++w) {
            hi = max(hi, sh_w[w]);
            lo = max(lo, sh_w[SORT_THREADS / 32 + w]);
        }
        hdr->block_max[blockIdx.x] = hi;
        hdr->block_negmax[blockIdx.x] = lo;
    }
    for (int64_t b = first; b < 2 * SORT_BINS; b += stride) hdr->hist[b] = 0;   // hist + cursor
    grid.sync();

    // ---- B ----
    hi = 0u; lo = 0u;
    for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += blockDim.x) {
        hi = max(hi, hdr->block_max[b]);
        lo = max(lo, hdr->block_negmax[b]);
    }
    hi = __reduce_max_sync(0xffffffffu, hi);
    lo = __reduce_max_sync(0xffffffffu, lo);
    if (lane == 0) { sh_w[warp] = hi; sh_w[SORT_THREADS / 32 + warp] = lo; }
    for (int b = threadIdx.x; b < SORT_BINS; b += blockDim.x) sh_a[b] = 0;
    __syncthreads();
    for (int w = 0; w < SORT_THREADS / 32; ++w) {
        hi = max(hi, sh_w[w]);
        lo = max(lo, sh_w[SORT_THREADS / 32 + w]);
    }
    const float kmin = __uint_as_float(~lo), kmax = __uint_as_float(hi);
    const float scale = kmax > kmin ? static_cast<float>(SORT_BINS) / (kmax - kmin) : 0.0f;
    uint32_t *bins = reinterpret_cast<uint32_t *>(keys);
    for (int64_t i = first; i < n; i += stride) {
        int bin = static_cast<int>((keys[i] - kmin) * scale);
        bin = min(max(bin, 0), SORT_BINS - 1);
        atomicAdd(&sh_a[bin], 1u);
        bins[i] = static_cast<uint32_t>(bin);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < SORT_BINS; b += blockDim.x)
        if (sh_a[b]) atomicAdd(&hdr->hist[b], sh_a[b]);
    grid.sync();

    // ---- C ----
    constexpr int PER = SORT_BINS / SORT_THREADS;
    uint32_t v[PER], run = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) { v[k] = hdr->hist[threadIdx.x * PER + k]; run += v[k]; }
    uint32_t inc = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();                      // sh_w and sh_a are reused
    if (lane == 31) sh_w[warp] = inc;
    for (int b = threadIdx.x; b < SORT_BINS; b += blockDim.x) sh_b[b] = 0;
    __syncthreads();
    uint32_t before = inc - run;
    for (int w = 0; w < warp; ++w) before += sh_w[w];
#pragma unroll
    for (int k = 0; k < PER; ++k) { sh_a[threadIdx.x * PER + k] = before; before += v[k]; }
    // ranks inside this block (one round of the grid-stride loop at a time)
    for (int64_t base_i = static_cast<int64_t>(blockIdx.x) * blockDim.x; base_i < n; base_i += stride) {
        const int64_t i = base_i + threadIdx.x;
        uint32_t bin = 0, rank = 0;
        if (i < n) {
            bin = bins[i];
            rank = atomicAdd(&sh_b[bin], 1u);
        }
        __syncthreads();
        for (int b = threadIdx.x; b < SORT_BINS; b += blockDim.x)
            if (sh_b[b]) sh_b[b] = atomicAdd(&hdr->cursor[b], sh_b[b]);
        __syncthreads();
        if (i < n) perm[sh_a[bin] + sh_b[bin] + rank] = static_cast<int32_t>(i);
        __syncthreads();
        for (int b = threadIdx.x; b < SORT_BINS; b += blockDim.x) sh_b[b] = 0;
        __syncthreads();
    }
}

// Builds the permutation in `ws` (pb200_sort_workspace_bytes(n) bytes);
// returns the device pointer to it through *perm.
template <class M>
int build_permutation(const pb200_problem *p, const double *d_params, int64_t n,
                      int64_t ld, void *ws, const int32_t **perm, cudaStream_t st) {
    SortHeader *hdr = static_cast<SortHeader *>(ws);
    float *keys = reinterpret_cast<float *>(hdr + 1);
    int32_t *out = reinterpret_cast<int32_t *>(keys + n);
    // co-resident blocks only (grid-wide barriers): what the runtime says fits
    // (registers, 9 KB of static shared memory), at most 4 per SM
    int64_t grid = (n + SORT_THREADS - 1) / SORT_THREADS;
    static int per_sm[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = (dev >= 0 && dev < 64) ? dev : 0;
    if (!per_sm[dev]) {
        int fit = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&fit, sort_kernel<M>, SORT_THREADS, 0) !=
                cudaSuccess || fit < 1) {
            cudaGetLastError();
            fit = 1;
        }
        per_sm[dev] = fit > 4 ? 4 : fit;
    }
    const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm[dev];
    if (grid > cap) grid = cap;
    if (grid > SORT_MAX_BLOCKS) grid = SORT_MAX_BLOCKS;
    void *args[] = {const_cast<ProblemDev *>(&p->dev), const_cast<ModelConsts *>(&p->mc),
                    &d_params, &n, &ld, &keys, &hdr, &out};
    int rc = pb200_check_cuda(
        cudaLaunchCooperativeKernel(reinterpret_cast<void *>(sort_kernel<M>),
                                    dim3(static_cast<unsigned>(grid)), dim3(SORT_THREADS),
                                    args, 0, st),
        "cooperative launch of the sort kernel");
    *perm = out;
    return rc;
}

// device facts and tuning knobs of the migrating layout
static int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}
static bool mig_enabled() {
    static const int on = [] {
        const char *e = getenv("PB200_MIGRATE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on != 0;
}
// PB200_LOCAL_SORT: 0 = always sort_kernel, 1 = always the block-local sort,
// anything else = by the spread of the previous population (sort_hint)
static int local_sort_mode() {
    static const int mode = [] {
        const char *e = getenv("PB200_LOCAL_SORT");
        return (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : 2;
    }();
    return mode;
}
// The problem's hint word (allocated on first use, never while the stream is
// being captured); nullptr when there is none.
static int32_t *sort_hint_of(const pb200_problem *p, cudaStream_t st) {
    if (p->sort_hint) return p->sort_hint;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (p->sort_hint) return p->sort_hint;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        return nullptr;
    }
    int32_t *w = nullptr;
    if (cudaHostAlloc(reinterpret_cast<void **>(&w), 64, cudaHostAllocMapped | cudaHostAllocPortable) !=
        cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    w[0] = 0;
    p->sort_hint = w;
    return w;
}
static bool local_plain_enabled() {
    static const int on = [] {
        const char *e = getenv("PB200_LOCAL_PLAIN");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on != 0;
}
static bool dense_enabled() {
    static const int on = [] {
        const char *e = getenv("PB200_DENSE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on != 0;
}
// Hand-over point of the extra workers as a fraction of the observations, by
// the number of warps a scheduler carries without them: 0.55 for three or
// four, 0.60 for two, 0.65 for one (swept at populations 65 536 / 58 000,
// 47 000 / 43 000 and 24 000 / 20 000: profiles/r02ap_migrate_at*.log; the
// differences are below 1 %).  PB200_MIGRATE_AT overrides.
static double mig_alpha(int warps_per_scheduler) {
    static const double forced = [] {
        const char *e = getenv("PB200_MIGRATE_AT");
        const double v = e ? atof(e) : 0.0;
        return (v > 0.0 && v < 1.0) ? v : 0.0;
    }();
    if (forced > 0.0) return forced;
    return warps_per_scheduler >= 3 ? 0.55 : warps_per_scheduler == 2 ? 0.60 : 0.65;
}

// series rows + the +inf sentinel rows, in whole 16-byte units
static int64_t staged_bytes(const pb200_problem *p, int no) {
    return ((static_cast<int64_t>(p->dev.n_times) + PB200_SENTINEL_ROWS) * (1 + no) * 8 + 15) & ~15ll;
}

#define PB200_LAUNCH(KERN, GRID, BLOCK, DYN)                                             \
    do {                                                                                 \
        auto kern = KERN;                                                                \
        const size_t dyn_ = (DYN);                                                       \
        if (dyn_ > 48 * 1024) {                                                          \
            int rc = pb200_check_cuda(                                                   \
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     static_cast<int>(dyn_)),                            \
                "cudaFuncSetAttribute");                                                 \
            if (rc) return rc;                                                           \
        }                                                                                \
        kern<<<static_cast<unsigned>(GRID), BLOCK, dyn_, st>>>(                          \
            p->dev, p->mc, d_params, n, ld, d_scores, d_traj, d_steps, perm, s PB200_XARG); \
    } while (0)

// one lane per parameter vector (any model)
template <class M>
int launch_model(const pb200_problem *p, const double *d_params, int64_t n,
                 int64_t ld, double *d_scores, double *d_traj,
                 int32_t *d_steps, const SolverDev &s, int block,
                 void *sort_ws, const ScatterDev &sc, cudaStream_t st) {
    const int32_t *perm = nullptr;
    const int64_t bytes = staged_bytes(p, M::NO);
#ifdef PB200_NO_SMEM
    const bool use_smem = false;
#else
    const bool use_smem = bytes <= 96 * 1024;   // keeps >= 2 blocks per SM
#endif
    const bool traj = d_traj != nullptr;
    const bool auto_layout = block <= 0;      // an explicit block size means the plain layout
    if (block <= 0) {
        block = 64;
        if (LaunchShape<M>::THREADS > 128) {
            // one block per SM as long as the population fits a single wave
            const int64_t per_sm = ((n + 31) / 32 + sm_count() - 1) / sm_count() * 32;
            block = per_sm > LaunchShape<M>::THREADS ? LaunchShape<M>::THREADS
                                                     : static_cast<int>(per_sm);
        }
    }
    if (block > LaunchShape<M>::THREADS) block = LaunchShape<M>::THREADS;
    block = (block + 31) & ~31;                 // whole warps: the kernel votes
    const int64_t grid = (n + block - 1) / block;
    if (grid <= 0) return PB200_OK;
    if (grid > 2147483647ll) {
        pb200_set_error("too many parameter vectors for one launch: %lld", (long long)n);
        return PB200_EINVAL;
    }
    const size_t dyn = use_smem ? static_cast<size_t>(bytes) : 0;
    // single-wave populations with one or two warps per SM more than a
    // multiple of four: migrating layout, one block per SM (see MigDev);
    // 17 .. 24 warps per SM: the dense register classes (see ode_kernel)
    MigDev mig = {0, 0, 0, 0};
    LocalSortDev ls = {0, 0, nullptr};
    bool use_mig = false, want_local = false;
    int dense = 0;
    if (M::MIGRATES && auto_layout && use_smem && !traj && mig_enabled()) {
        const int64_t warps = (n + 31) / 32;
        const int sms = sm_count();
        const int64_t per_sm = (warps + sms - 1) / sms;
        const int r = static_cast<int>(per_sm % 4);
        // the dense instances exist for the model's default interpolation loop
        const bool default_wide = (p->dev.model != PB200_MODEL_FHN) == (s.wide != 0);
        // (23 or 24 warps per SM without migration are slower at 80 registers
        // than two waves at 124: 0.694 against 0.656 ms, profiles/r02i_dense_probe.log)
        if (per_sm > 16 && per_sm <= 24 && default_wide && dense_enabled() &&
            (per_sm <= 20 || r == 1 || r == 2))
            dense = per_sm <= 20 ? 1 : 2;
        // Which cost sort (if any).  The block-local sort (local_cost_sort) is
        // 6-9 us faster on populations whose costs differ by a few per cent
        // (an optimiser's or sampler's population), the global order of
        // sort_kernel is up to 14 us faster on wide ones; every one-block-per-SM
        // launch leaves what it measured in the problem's hint word (see
        // PB200_LOCAL_SORT_CYCLES) and the next launch goes by it (the first one
        // takes the global sort).
        const bool single_wave = per_sm >= 5 && (per_sm <= 16 || dense);
        if (sort_ws && single_wave && local_sort_mode() != 0) {
            ls.hint = sort_hint_of(p, st);
            const int seen = ls.hint ? *static_cast<volatile int32_t *>(ls.hint) : 0;
            want_local = local_sort_mode() == 1 || seen == 1;
        }
        if (per_sm >= 5 && (per_sm <= 14 || dense) && (r == 1 || r == 2)) {
            use_mig = true;
            mig.q4 = static_cast<int32_t>(per_sm - r);
            mig.r = r;
            const int64_t row = (1 + M::NO) * 8;
            int64_t mid = static_cast<int64_t>(mig_alpha(mig.q4 / 4) * p->dev.n_times);
            mid = mid < 1 ? 1 : mid;
            mig.off_mid = static_cast<uint32_t>(mid * row);
            mig.buf_off = static_cast<uint32_t>(bytes);
        } else if (single_wave && !dense && local_plain_enabled()) {
            // nothing to migrate (the schedulers carry equal numbers of warps, or
            // three of them one more): the same one-block-per-SM kernel without
            // extra workers.  Its round-robin deal of the sorted warps keeps the
            // SMs level where 64-thread blocks land wherever a slot frees up
            // (population 73 000, 16 warps per SM: 0.441 -> 0.422 ms, wide prior
            // box 0.702 -> 0.643 ms), and the sort can run inside it (0.416 ms;
            // profiles/r02af_local_plain*.log)
            use_mig = true;
            mig.q4 = static_cast<int32_t>(per_sm);
            mig.r = 0;
            mig.off_mid = 0xffffffffu;
            mig.buf_off = static_cast<uint32_t>(bytes);
        }
        if (dense && !use_mig) {
            // plain layout, one block of per_sm warps per SM
            block = static_cast<int>(per_sm) * 32;
        }
    }
    const bool one_block_per_sm = use_mig || (dense != 0);
    size_t sort_dyn = 0;
    if (want_local && one_block_per_sm) {
        ls.on = 1;
        sort_dyn = static_cast<size_t>(LOCAL_SORT_WORDS) * 4 + 16;
    }
    if (sort_ws && !ls.on) {
        int rc = build_permutation<M>(p, d_params, n, ld, sort_ws, &perm, st);
        if (rc) return rc;
    }
#define PB200_XARG , mig, sc, ls
    constexpr bool W = !std::is_same<M, FhnModel>::value;      // the model's default loop
    if (use_mig) {
        if constexpr (M::MIGRATES) {
        const int mblock = (mig.q4 + (mig.r ? 2 + mig.r : 0)) * 32;
        const int64_t per_block = static_cast<int64_t>(mig.q4 + mig.r) * 32;
        const int64_t mgrid = (n + per_block - 1) / per_block;
        size_t mdyn = dyn + static_cast<size_t>(mig.r) * (2 * M::NS + 2 + M::NO + 2) * 32 * 8;
        if (ls.on) {
            ls.scratch_off = static_cast<uint32_t>((mdyn + 15) & ~static_cast<size_t>(15));
            mdyn = ls.scratch_off + sort_dyn;
        }
        if (dense == 2) PB200_LAUNCH((ode_kernel<M, true, false, true, W, 2>), mgrid, mblock, mdyn);
        else if (dense == 1) PB200_LAUNCH((ode_kernel<M, true, false, true, W, 1>), mgrid, mblock, mdyn);
        else if (s.wide) PB200_LAUNCH((ode_kernel<M, true, false, true, true>), mgrid, mblock, mdyn);
        else PB200_LAUNCH((ode_kernel<M, true, false, true, false>), mgrid, mblock, mdyn);
        }
    } else if (dense) {
        if constexpr (M::MIGRATES) {
        const int64_t dgrid = (n + block - 1) / block;
        size_t ddyn = dyn;
        if (ls.on) {
            ls.scratch_off = static_cast<uint32_t>((ddyn + 15) & ~static_cast<size_t>(15));
            ddyn = ls.scratch_off + sort_dyn;
        }
        if (dense == 2) PB200_LAUNCH((ode_kernel<M, true, false, false, W, 2>), dgrid, block, ddyn);
        else PB200_LAUNCH((ode_kernel<M, true, false, false, W, 1>), dgrid, block, ddyn);
        }
    } else if (use_smem) {
        if (traj) PB200_LAUNCH((ode_kernel<M, true, true, false, false>), grid, block, dyn);
        else if (s.wide) PB200_LAUNCH((ode_kernel<M, true, false, false, true>), grid, block, dyn);
        else PB200_LAUNCH((ode_kernel<M, true, false, false, false>), grid, block, dyn);
    } else {
        if (traj) PB200_LAUNCH((ode_kernel<M, false, true, false, false>), grid, block, dyn);
        else PB200_LAUNCH((ode_kernel<M, false, false, false, false>), grid, block, dyn);
    }
#undef PB200_XARG
    return pb200_check_cuda(cudaGetLastError(), "ode_kernel launch");
}

// Beeler-Reuter on the stiff integrator (br_rosenbrock_kernel); PB200_BR_EXPLICIT=1
// keeps the explicit DP5(4) template for comparisons.  Every vector costs the
// same here (the step size follows the action potential, not the parameters):
// no cost sort.
static bool br_explicit() {
    static const int on = [] {
        const char *e = getenv("PB200_BR_EXPLICIT");
        return (e && e[0] == '1') ? 1 : 0;
    }();
    return on != 0;
}
int launch_br_rosenbrock(const pb200_problem *p, const double *d_params, int64_t n,
                         int64_t ld, double *d_scores, double *d_traj, int32_t *d_steps,
                         const SolverDev &s, const ScatterDev &sc, cudaStream_t st) {
    const int32_t *perm = nullptr;
    const int64_t bytes = staged_bytes(p, BeelerReuterModel::NO);
    const bool use_smem = bytes <= 64 * 1024;
    const int block = PB200_BRR_THREADS;
    const int64_t grid = (n + block - 1) / block;
    if (grid <= 0) return PB200_OK;
    if (grid > 2147483647ll) {
        pb200_set_error("too many parameter vectors for one launch: %lld", (long long)n);
        return PB200_EINVAL;
    }
    const size_t dyn = use_smem ? static_cast<size_t>(bytes) : 0;
#define PB200_XARG , sc
    if (use_smem) {
        if (d_traj) PB200_LAUNCH((br_rosenbrock_kernel<true, true>), grid, block, dyn);
        else PB200_LAUNCH((br_rosenbrock_kernel<true, false>), grid, block, dyn);
    } else {
        if (d_traj) PB200_LAUNCH((br_rosenbrock_kernel<false, true>), grid, block, dyn);
        else PB200_LAUNCH((br_rosenbrock_kernel<false, false>), grid, block, dyn);
    }
#undef PB200_XARG
    return pb200_check_cuda(cudaGetLastError(), "br_rosenbrock_kernel launch");
}

// two lanes per parameter vector (two-state models)
template <class L, class M>
int launch_pair(const pb200_problem *p, const double *d_params, int64_t n,
                int64_t ld, double *d_scores, double *d_traj,
                int32_t *d_steps, const SolverDev &s, int block,
                void *sort_ws, const ScatterDev &sc, cudaStream_t st) {
    const int32_t *perm = nullptr;
    if (sort_ws) {
        int rc = build_permutation<M>(p, d_params, n, ld, sort_ws, &perm, st);
        if (rc) return rc;
    }
    const int64_t bytes = staged_bytes(p, 2);
#ifdef PB200_NO_SMEM
    const bool use_smem = false;
#else
    // one wave of PB200_PAIR_MINBLOCKS blocks per SM must fit in 227 KB
    const bool use_smem = bytes <= (227 * 1024) / PB200_PAIR_MINBLOCKS - 1024;
#endif
    const bool traj = d_traj != nullptr;
    if (block <= 0) block = PB200_PAIR_THREADS;
    if (block > PB200_PAIR_THREADS) block = PB200_PAIR_THREADS;
    block = (block + 31) & ~31;
    const int64_t threads = 2 * n;
    const int64_t grid = (threads + block - 1) / block;
    if (grid <= 0) return PB200_OK;
    if (grid > 2147483647ll) {
        pb200_set_error("too many parameter vectors for one launch: %lld", (long long)n);
        return PB200_EINVAL;
    }
    const size_t dyn = use_smem ? static_cast<size_t>(bytes) : 0;
#define PB200_XARG , sc
    if (use_smem) {
        if (traj) PB200_LAUNCH((ode_pair_kernel<L, true, true>), grid, block, dyn);
        else PB200_LAUNCH((ode_pair_kernel<L, true, false>), grid, block, dyn);
    } else {
        if (traj) PB200_LAUNCH((ode_pair_kernel<L, false, true>), grid, block, dyn);
        else PB200_LAUNCH((ode_pair_kernel<L, false, false>), grid, block, dyn);
    }
#undef PB200_XARG
    return pb200_check_cuda(cudaGetLastError(), "ode_pair_kernel launch");
}
#undef PB200_LAUNCH

}  // namespace

#ifndef PB200_DEFAULT_LANES
#define PB200_DEFAULT_LANES 1
#endif

int launch_ode(const pb200_problem *p, const double *d_params, int64_t n,
               int64_t ld, double *d_scores, double *d_traj, int32_t *d_steps,
               const SolverDev &s_in, int block, int lanes, void *sort_ws,
               const ScatterDev &sc, cudaStream_t st) {
    SolverDev s = s_in;
    // automatic: the wide pending-observation loop pays when steps span many
    // observations (Lotka-Volterra, SIR, ...: 9+ per step, -35 % time); for
    // FitzHugh-Nagumo's relaxation oscillations (2.6 per step) its extra
    // ballot and load per attempt cost 3 %
    if (s.wide < 0)
        s.wide = (p->dev.model == PB200_MODEL_FHN || p->dev.model == PB200_MODEL_BEELER_REUTER) ? 0 : 1;
    if (lanes <= 0) lanes = PB200_DEFAULT_LANES;
    if (lanes != 1 && lanes != 2) {
        pb200_set_error("lanes per vector must be 0 (automatic), 1 or 2, got %d", lanes);
        return PB200_EINVAL;
    }
    switch (p->dev.model) {
        case PB200_MODEL_FHN:
            if (lanes == 2)
                return launch_pair<FhnLane, FhnModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
            return launch_model<FhnModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_LOTKA_VOLTERRA:
            if (lanes == 2)
                return launch_pair<LvLane, LvModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
            return launch_model<LvModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_SIR:           // three states: always one lane per vector
            return launch_model<SirModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_GOODWIN:
            return launch_model<GoodwinModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_HES1:
            return launch_model<Hes1Model>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_REPRESSILATOR:
            return launch_model<RepressilatorModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
        case PB200_MODEL_BEELER_REUTER:
            if (br_explicit())
                return launch_model<BeelerReuterModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, sort_ws, sc, st);
            return launch_br_rosenbrock(p, d_params, n, ld, d_scores, d_traj, d_steps, s, sc, st);
    }
    pb200_set_error("model %d is not an ODE model", p->dev.model);
    return PB200_EUNSUPPORTED;
}

int64_t ode_sort_workspace_bytes(int64_t n) {
    return static_cast<int64_t>(sizeof(SortHeader)) + 8 * (n > 0 ? n : 0);
}

#if PB200_TIMELINE
extern "C" int pb200_debug_timeline(long long *host, int n_words) {
    return (int)cudaMemcpyFromSymbol(host, pb200_timeline, sizeof(long long) * n_words);
}
#endif
