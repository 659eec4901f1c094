// Fused adaptive Dormand-Prince 5(4) solve + objective for the ODE toy models
// (FitzHugh-Nagumo, Lotka-Volterra, SIR).  One thread integrates one parameter
// vector; the observed series sits in shared memory (bulk-copied once per
// block); every accepted step is interpolated at the observation times that
// fall inside it (4th-order dense output), the residuals are squared and
// accumulated in registers, and only the score leaves the SM.
//
// Replaces, per parameter vector (citations relative to the reference root):
//   ToyODEModel._simulate           pints/toy/_toy_classes.py:190-234
//   FitzhughNagumoModel._rhs        pints/toy/_fitzhugh_nagumo_model.py:111-117
//   LotkaVolterraModel._rhs         pints/toy/_lotka_volterra_model.py:91-97
//   SIRModel.simulate/_rhs          pints/toy/_sir_model.py:96-111
//   Multi/SingleOutputProblem.evaluate   pints/_core.py:147-153, :257-265
//   objective __call__              see objective_finish()
// The reference integrates with LSODA (scipy odeint, rtol = atol = 1.49012e-8);
// this kernel uses DP5(4) at the same tolerances, see DESIGN.md for the
// measured agreement.
#include "pb200_common.cuh"

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
// Each model: NS states, NO outputs, OUT0 = index of the first output state,
// load() turns a parameter row into per-thread constants and y0, rhs() is the
// (autonomous) right-hand side.
struct FhnModel {
    static constexpr int NS = 2, NO = 2, OUT0 = 0, NP = 3;
    // V' = c (V - V^3/3 + R) = (c - (c/3) V^2) V + c R
    // R' = -(V - a + b R) / c  = (-1/c) V + (a/c) + (-b/c) R
    // written with per-vector constants so one evaluation is 6 FP64 ops
    double c, c3, nic, aoc, nboc;
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
};

struct LvModel {
    static constexpr int NS = 2, NO = 2, OUT0 = 0, NP = 4;
    double a, b, c, d;
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
};

struct SirModel {
    static constexpr int NS = 3, NO = 2, OUT0 = 1, NP = 3;
    double gamma, v;
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
};

// MUFU.RCP64H: >= 20 good bits, on the XU pipe (off the FP64 pipe).
__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// Full-precision reciprocal for normal-range x: seed + two Newton steps
// (20 -> 40 -> 80 bits); no slow path, four DFMAs.
__device__ __forceinline__ double rcp_newton(double x) {
    double r = rcp_seed(x);
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// Mean square of err_i / (atol + rtol * (|y_i| + |ynew_i|) / 2).  The scale
// uses the mean of the two magnitudes (one DADD with |.| modifiers) instead of
// their max (a DSETP plus a 64-bit select); the 1e-6-accurate reciprocal seed
// is enough because the result is only compared with 1 and fed to the
// step-size controller.
template <int NS>
__device__ __forceinline__ double err_norm2(const double *esum, double h,
                                            const double *y, const double *yn,
                                            double half_rtol, double atol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const double sc = fma(half_rtol, fabs(y[i]) + fabs(yn[i]), atol);
        const double q = (esum[i] * h) * rcp_seed(sc);
        s = fma(q, q, s);
    }
    return s * (1.0 / NS);
}

// Step-size factor 0.9 * norm^(-1/5), clamped to [0.2, 10], from norm^2.  It
// runs on the SFU in fp32 (lg2/ex2 approximations): the controller needs a few
// digits only, and accept/reject is decided in fp64 on norm^2 <= 1, never on
// this value.  norm2 = 0 -> 10, norm2 = inf or NaN -> 0.2.
__device__ __forceinline__ float step_factor(double norm2) {
    const float n2 = static_cast<float>(norm2);
    float l, f;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(n2));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(-0.1f * l));
    return fminf(fmaxf(0.9f * f, 0.2f), 10.0f);
}

// Observed series access.  In shared memory the rows are read with explicit
// ld.shared through a 32-bit byte address kept in a register (the generic
// pointer form makes the compiler rebuild the shared-window base on every
// access); otherwise through the read-only global path.
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

template <class M, bool SMEM, bool TRAJ>
#ifndef PB200_PIPE_U
#define PB200_PIPE_U 2          // deferred interpolation slots per step attempt
#endif
#ifndef PB200_MAXTHREADS
#define PB200_MAXTHREADS 128
#endif
__global__ void __launch_bounds__(PB200_MAXTHREADS)
ode_kernel(const __grid_constant__ ProblemDev P,
           const __grid_constant__ ModelConsts mc,
           const double *__restrict__ params, int64_t n, int64_t ld,
           double *__restrict__ scores, double *__restrict__ traj,
           int32_t *__restrict__ steps_out, const SolverDev S) {
    constexpr int NS = M::NS, NO = M::NO, OUT0 = M::OUT0;
    constexpr uint32_t ROWB = (1 + NO) * 8;          // bytes per series row
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;

    Series<SMEM> series;
    series.gbase = P.series;
    series.sbase = 0;
    if (SMEM) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(P.n_times) + PB200_SENTINEL_ROWS) * ROWB + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series.sbase = smem_u32(smem);
    }

    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;

    const int nt = P.n_times;
    const int n_read = TRAJ ? P.n_model_params : P.n_params;
    double x[PB200_MAX_PARAMS];
#pragma unroll
    for (int k = 0; k < M::NP + NO; ++k)
        x[k] = (k < n_read) ? params[i * ld + k] : 0.0;

    if (!TRAJ) {
        double early;
        if (objective_precheck(P, x, &early)) {
            scores[i] = early;
            if (steps_out) steps_out[i] = 0;
            return;
        }
    }

    M model;
    double y[NS], t;
    model.load(x, mc, y, nt > 0 ? series.at(0) : 0.0, &t);
    if (nt == 0) t = 0.0;

    double ss[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) ss[o] = 0.0;

    // `off` is the byte offset of the next observation row, `t_out` its time.
    // The packed series ends with a sentinel row whose time is +inf, so the
    // interpolation loop needs no bounds check and one register compare per
    // step decides whether any interpolation is needed.
    const uint32_t end = static_cast<uint32_t>(nt) * ROWB;
    uint32_t off = 0;
    double t_out = series.at(0);
    double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * NO : nullptr;

    // observations at (or before) the initial time see the initial state
    while (t_out <= t) {
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            if (TRAJ) my_traj[o] = y[OUT0 + o];
            const double r = series.at(off + 8 + 8 * o) - y[OUT0 + o];
            ss[o] = fma(r, r, ss[o]);
        }
        if (TRAJ) my_traj += NO;
        off += ROWB;
        t_out = series.at(off);
    }

    const double half_rtol = S.half_rtol, atol = S.atol;
    double k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS];
    double yn[NS], w[NS], esum[NS];
    model.rhs(y, k1);

    // ---- first step (Hairer's hinit, as scipy's select_initial_step) ----
    double h = S.h0;
    if (!(h > 0.0)) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double sc = fma(S.rtol, fabs(y[j]), atol);
            const double a0 = y[j] / sc, a1 = k1[j] / sc;
            d0 = fma(a0, a0, d0);
            d1 = fma(a1, a1, d1);
        }
        d0 = sqrt(d0 * (1.0 / NS));
        d1 = sqrt(d1 * (1.0 / NS));
        double h_a = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
#pragma unroll
        for (int j = 0; j < NS; ++j) w[j] = fma(h_a, k1[j], y[j]);
        model.rhs(w, k2);
        double d2 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double sc = fma(S.rtol, fabs(y[j]), atol);
            const double a2 = (k2[j] - k1[j]) / sc;
            d2 = fma(a2, a2, d2);
        }
        d2 = sqrt(d2 * (1.0 / NS)) / h_a;
        const double dm = fmax(d1, d2);
        const double h_b = (dm <= 1e-15) ? fmax(1e-6, h_a * 1e-3)
                                         : pow(0.01 / dm, 0.2);
        h = fmin(100.0 * h_a, h_b);
    }

    int attempts = 0;
    bool failed = false;
    bool after_reject = false;
    const int max_steps = S.max_steps;

    // The observations inside an accepted step are NOT interpolated right away:
    // the step's dense-output coefficients are parked ("pending interval") and
    // consumed during the NEXT attempt, where U branch-free interpolation slots
    // sit in the same basic block as the six stages.  The slots are independent
    // of the stage arithmetic, so they fill the issue slots that the serial
    // RHS -> stage -> RHS dependency chain (9-cycle DFMA latency) leaves empty;
    // with only ~3.5 warps per scheduler at population 65 536 that is where the
    // FP64 pipe was idling.  Steps that contain more than U observations drain
    // the rest in a (divergent) loop before the coefficients are overwritten.
    constexpr int U = PB200_PIPE_U;
    double p_t = 0.0, p_end = -INFINITY, p_invh = 0.0;
    double p_y[NO], p_r2[NO], p_r3[NO], p_r4[NO], p_r5[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) p_y[o] = p_r2[o] = p_r3[o] = p_r4[o] = p_r5[o] = 0.0;
    const double t_last = nt > 0 ? series.at(end - ROWB) : -INFINITY;

    // one interpolation of the pending interval at time tq against row `row`
    auto interpolate = [&](double tq, uint32_t row, bool valid) {
        const double th = (tq - p_t) * p_invh;
        const double th1 = 1.0 - th;
        const double thq = th * th1;
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            // y + th ((r2 + th1 r3) + (th th1)(r4 + th1 r5))
            const double v = fma(th, fma(thq, fma(th1, p_r5[o], p_r4[o]),
                                         fma(th1, p_r3[o], p_r2[o])), p_y[o]);
            if (TRAJ) { if (valid) my_traj[(row - off) / ROWB * NO + o] = v; }
            const double r = valid ? series.at(row + 8 + 8 * o) - v : 0.0;
            ss[o] = fma(r, r, ss[o]);
        }
    };
    auto drain = [&]() {
        while (t_out <= p_end) {
            interpolate(t_out, off, true);
            if (TRAJ) my_traj += NO;
            off += ROWB;
            t_out = series.at(off);
        }
    };

    while (off < end) {
        if (p_end >= t_last) {          // every remaining observation is pending
            drain();
            break;
        }
        if (attempts >= max_steps) {
            failed = true;
            break;
        }
        ++attempts;
        const double inv_h = rcp_newton(h);
        // ---- U deferred interpolation slots (branch free) ----
        if (U > 0) {
            double tq[U > 0 ? U : 1];
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < U; ++k) tq[k] = series.at(off + k * ROWB);
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const bool valid = tq[k] <= p_end;     // monotone in k
                interpolate(tq[k], off + k * ROWB, valid);
                cnt += valid ? 1u : 0u;
            }
            if (TRAJ) my_traj += cnt * NO;
            off += cnt * ROWB;
            t_out = series.at(off);
        }
        // ---- six stages (k1 is f(y): first-same-as-last) ----
#pragma unroll
        for (int j = 0; j < NS; ++j) w[j] = fma(h, A21 * k1[j], y[j]);
        model.rhs(w, k2);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(h, fma(A32, k2[j], A31 * k1[j]), y[j]);
        model.rhs(w, k3);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(h, fma(A43, k3[j], fma(A42, k2[j], A41 * k1[j])), y[j]);
        model.rhs(w, k4);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(h, fma(A54, k4[j], fma(A53, k3[j], fma(A52, k2[j], A51 * k1[j]))), y[j]);
        model.rhs(w, k5);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            w[j] = fma(h, fma(A65, k5[j], fma(A64, k4[j], fma(A63, k3[j], fma(A62, k2[j], A61 * k1[j])))), y[j]);
        model.rhs(w, k6);
#pragma unroll
        for (int j = 0; j < NS; ++j)
            yn[j] = fma(h, fma(B6, k6[j], fma(B5, k5[j], fma(B4, k4[j], fma(B3, k3[j], B1 * k1[j])))), y[j]);
        model.rhs(yn, k7);
        // ---- embedded error estimate (to be scaled by h) ----
#pragma unroll
        for (int j = 0; j < NS; ++j)
            esum[j] = fma(E7, k7[j], fma(E6, k6[j], fma(E5, k5[j], fma(E4, k4[j], fma(E3, k3[j], E1 * k1[j])))));
        const double n2 = err_norm2<NS>(esum, h, y, yn, half_rtol, atol);
        float fac = step_factor(n2);
        // observations of the pending interval beyond the U slots
        drain();

        if (n2 <= 1.0) {
            // ---- accepted: park the dense-output coefficients ----
            const double t_new = t + h;
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                const int j = OUT0 + o;
                const double d = yn[j] - y[j];
                const double b = fma(h, k1[j], -d);
                p_y[o] = y[j];
                p_r2[o] = d;
                p_r3[o] = b;
                p_r4[o] = (d - h * k7[j]) - b;
                p_r5[o] = h * fma(D7, k7[j], fma(D6, k6[j], fma(D5, k5[j], fma(D4, k4[j], fma(D3, k3[j], D1 * k1[j])))));
            }
            p_t = t;
            p_end = t_new;
            p_invh = inv_h;
            t = t_new;
#pragma unroll
            for (int j = 0; j < NS; ++j) { y[j] = yn[j]; k1[j] = k7[j]; }
            // no growth on the step right after a rejection
            if (after_reject) fac = fminf(fac, 1.0f);
            after_reject = false;
        } else {
            // rejected (also when the norm is NaN, which gives fac = 0.2)
            fac = fminf(fac, 1.0f);
            after_reject = true;
            // step size underflow: give up on this vector (score NaN)
            if (!(fabs(h) * 0.2 > 1e-14 * fmax(fabs(t), 1.0))) {
                failed = true;
                break;
            }
        }
        h *= static_cast<double>(fac);
    }

    if (steps_out) steps_out[i] = attempts;
    if (TRAJ) {
        if (failed) {
            const int64_t rest = (static_cast<int64_t>(end) - off) / ROWB * NO;
            for (int64_t q = 0; q < rest; ++q) my_traj[q] = NAN;
        }
    } else {
        scores[i] = failed ? NAN : objective_finish(P, ss, x);
    }
}

template <class M>
int launch_model(const pb200_problem *p, const double *d_params, int64_t n,
                 int64_t ld, double *d_scores, double *d_traj,
                 int32_t *d_steps, const SolverDev &s, int block,
                 cudaStream_t st) {
    // series rows + the +inf sentinel rows, in whole 16-byte units
    const int64_t bytes = ((static_cast<int64_t>(p->dev.n_times) + PB200_SENTINEL_ROWS) * (1 + M::NO) * 8 + 15) & ~15ll;
#ifdef PB200_NO_SMEM
    const bool use_smem = false;
#else
    const bool use_smem = bytes <= 96 * 1024;   // keeps >= 2 blocks per SM
#endif
    const bool traj = d_traj != nullptr;
    if (block <= 0) block = 64;
    const int64_t grid = (n + block - 1) / block;
    if (grid <= 0) return PB200_OK;
    if (grid > 2147483647ll) {
        pb200_set_error("too many parameter vectors for one launch: %lld", (long long)n);
        return PB200_EINVAL;
    }
#define PB200_LAUNCH(SM, TR)                                                             \
    do {                                                                                 \
        auto kern = ode_kernel<M, SM, TR>;                                               \
        size_t dyn = SM ? static_cast<size_t>(bytes) : 0;                                \
        if (dyn > 48 * 1024) {                                                           \
            int rc = pb200_check_cuda(                                                   \
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     static_cast<int>(dyn)),                             \
                "cudaFuncSetAttribute");                                                 \
            if (rc) return rc;                                                           \
        }                                                                                \
        kern<<<static_cast<unsigned>(grid), block, dyn, st>>>(                           \
            p->dev, p->mc, d_params, n, ld, d_scores, d_traj, d_steps, s);               \
    } while (0)
    if (use_smem) {
        if (traj) PB200_LAUNCH(true, true); else PB200_LAUNCH(true, false);
    } else {
        if (traj) PB200_LAUNCH(false, true); else PB200_LAUNCH(false, false);
    }
#undef PB200_LAUNCH
    return pb200_check_cuda(cudaGetLastError(), "ode_kernel launch");
}

}  // namespace

int launch_ode(const pb200_problem *p, const double *d_params, int64_t n,
               int64_t ld, double *d_scores, double *d_traj, int32_t *d_steps,
               const SolverDev &s, int block, cudaStream_t st) {
    switch (p->dev.model) {
        case PB200_MODEL_FHN:
            return launch_model<FhnModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, st);
        case PB200_MODEL_LOTKA_VOLTERRA:
            return launch_model<LvModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, st);
        case PB200_MODEL_SIR:
            return launch_model<SirModel>(p, d_params, n, ld, d_scores, d_traj, d_steps, s, block, st);
    }
    pb200_set_error("model %d is not an ODE model", p->dev.model);
    return PB200_EUNSUPPORTED;
}
