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
constexpr double A21 = 1.0 / 5.0;
constexpr double A31 = 3.0 / 40.0, A32 = 9.0 / 40.0;
constexpr double A41 = 44.0 / 45.0, A42 = -56.0 / 15.0, A43 = 32.0 / 9.0;
constexpr double A51 = 19372.0 / 6561.0, A52 = -25360.0 / 2187.0,
                 A53 = 64448.0 / 6561.0, A54 = -212.0 / 729.0;
constexpr double A61 = 9017.0 / 3168.0, A62 = -355.0 / 33.0,
                 A63 = 46732.0 / 5247.0, A64 = 49.0 / 176.0,
                 A65 = -5103.0 / 18656.0;
constexpr double B1 = 35.0 / 384.0, B3 = 500.0 / 1113.0, B4 = 125.0 / 192.0,
                 B5 = -2187.0 / 6784.0, B6 = 11.0 / 84.0;
// error estimate = h * sum E_i k_i  (5th minus 4th order weights)
constexpr double E1 = 71.0 / 57600.0, E3 = -71.0 / 16695.0, E4 = 71.0 / 1920.0,
                 E5 = -17253.0 / 339200.0, E6 = 22.0 / 525.0, E7 = -1.0 / 40.0;
// dense output (Hairer, Norsett & Wanner, contd5)
constexpr double D1 = -12715105075.0 / 11282082432.0,
                 D3 = 87487479700.0 / 32700410799.0,
                 D4 = -10690763975.0 / 1880347072.0,
                 D5 = 701980252875.0 / 199316789632.0,
                 D6 = -1453857185.0 / 822651844.0,
                 D7 = 69997945.0 / 29380423.0;

constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;

// ----------------------------------------------------------------- models
// Each model: NS states, NO outputs, OUT0 = index of the first output state,
// load() turns a parameter row into per-thread constants and y0, rhs() is the
// (autonomous) right-hand side.
struct FhnModel {
    static constexpr int NS = 2, NO = 2, OUT0 = 0, NP = 3;
    double a, b, c, neg_inv_c;
    __device__ __forceinline__ void load(const double *x, const ModelConsts &mc,
                                         double *y, double /*t_first*/,
                                         double *t0) {
        a = x[0]; b = x[1]; c = x[2];
        neg_inv_c = -1.0 / c;
        y[0] = mc.c[0]; y[1] = mc.c[1];
        *t0 = 0.0;
    }
    __device__ __forceinline__ void rhs(const double *y, double *f) const {
        const double V = y[0], R = y[1];
        // (V - V^3/3 + R) * c ;  (V - a + b R) / -c
        const double v2 = V * V;
        const double cubic = fma(v2 * (-1.0 / 3.0), V, V);
        f[0] = (cubic + R) * c;
        f[1] = fma(b, R, V - a) * neg_inv_c;
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
        const double uv = u * v;
        f[0] = fma(a, u, -(b * uv));      // a x - b x y
        f[1] = fma(d, uv, -(c * v));      // -c y + d x y
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
        const double gsi = gamma * y[0] * y[1];
        const double vi = v * y[1];
        f[0] = -gsi;
        f[1] = gsi - vi;
        f[2] = vi;
    }
};

__device__ __forceinline__ double rcp_approx(double x) {
    // MUFU.RCP64H seed (>= 20 bits) + one Newton step: ~1e-12 relative, plenty
    // for the step-size controller (accept/reject never depends on it).
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return fma(r, fma(-x, r, 1.0), r);
}

template <int NS>
__device__ __forceinline__ double err_norm2(const double *err, const double *y,
                                            const double *yn, double rtol,
                                            double atol) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const double sc = fma(rtol, fmax(fabs(y[i]), fabs(yn[i])), atol);
        const double q = err[i] * rcp_approx(sc);
        s = fma(q, q, s);
    }
    return s * (1.0 / NS);
}

// factor = SAFETY * norm^(-1/5) from norm^2, through the SFU in fp32: the
// controller only needs a few digits.
__device__ __forceinline__ double step_factor(double norm2) {
    const float n2 = static_cast<float>(norm2);
    // norm^(-0.2) = (norm2)^(-0.1)
    const float f = exp2f(-0.1f * __log2f(n2));
    return static_cast<double>(0.9f * f);
}

template <class M, bool SMEM, bool TRAJ>
__global__ void __launch_bounds__(128)
ode_kernel(const __grid_constant__ ProblemDev P,
           const __grid_constant__ ModelConsts mc,
           const double *__restrict__ params, int64_t n, int64_t ld,
           double *__restrict__ scores, double *__restrict__ traj,
           int32_t *__restrict__ steps_out, const SolverDev S) {
    constexpr int NS = M::NS, NO = M::NO, OUT0 = M::OUT0;
    constexpr int ROW = 1 + NO;
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;

    const double *series;
    if (SMEM) {
        const uint32_t bytes =
            static_cast<uint32_t>((static_cast<int64_t>(P.n_times) * ROW * 8 + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series = smem;
    } else {
        series = P.series;
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
    model.load(x, mc, y, nt > 0 ? series[0] : 0.0, &t);

    double ss[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) ss[o] = 0.0;

    int idx = 0;
    double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * NO : nullptr;

    // observations at (or before) the initial time see the initial state
    while (idx < nt && series[idx * ROW] <= t) {
#pragma unroll
        for (int o = 0; o < NO; ++o) {
            if (TRAJ) my_traj[idx * NO + o] = y[OUT0 + o];
            const double r = series[idx * ROW + 1 + o] - y[OUT0 + o];
            ss[o] = fma(r, r, ss[o]);
        }
        ++idx;
    }

    const double rtol = S.rtol, atol = S.atol;
    double k1[NS], k2[NS], k3[NS], k4[NS], k5[NS], k6[NS], k7[NS];
    double yn[NS], w[NS], err[NS];
    model.rhs(y, k1);

    // ---- first step (Hairer's hinit, as scipy's select_initial_step) ----
    double h = S.h0;
    if (!(h > 0.0)) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double sc = fma(rtol, fabs(y[j]), atol);
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
            const double sc = fma(rtol, fabs(y[j]), atol);
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

    while (idx < nt) {
        if (attempts >= max_steps || !(fabs(h) > 1e-14 * fmax(fabs(t), 1.0))) {
            failed = true;
            break;
        }
        ++attempts;
        // ---- six stages ----
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
        // ---- embedded error ----
#pragma unroll
        for (int j = 0; j < NS; ++j)
            err[j] = h * fma(E7, k7[j], fma(E6, k6[j], fma(E5, k5[j], fma(E4, k4[j], fma(E3, k3[j], E1 * k1[j])))));
        const double n2 = err_norm2<NS>(err, y, yn, rtol, atol);

        if (n2 <= 1.0) {
            // ---- accepted: interpolate at the observation times inside ----
            const double t_new = t + h;
            if (series[idx * ROW] <= t_new) {
                const double inv_h = 1.0 / h;
                double r2[NS], r3[NS], r4[NS], r5[NS];
#pragma unroll
                for (int o = 0; o < NO; ++o) {
                    const int j = OUT0 + o;
                    r2[j] = yn[j] - y[j];
                    r3[j] = fma(h, k1[j], -r2[j]);
                    r4[j] = (r2[j] - h * k7[j]) - r3[j];
                    r5[j] = h * fma(D7, k7[j], fma(D6, k6[j], fma(D5, k5[j], fma(D4, k4[j], fma(D3, k3[j], D1 * k1[j])))));
                }
                do {
                    const double th = (series[idx * ROW] - t) * inv_h;
                    const double th1 = 1.0 - th;
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        const int j = OUT0 + o;
                        const double v = fma(th, fma(th1, fma(th, fma(th1, r5[j], r4[j]), r3[j]), r2[j]), y[j]);
                        if (TRAJ) my_traj[idx * NO + o] = v;
                        const double r = series[idx * ROW + 1 + o] - v;
                        ss[o] = fma(r, r, ss[o]);
                    }
                    ++idx;
                } while (idx < nt && series[idx * ROW] <= t_new);
            }
            t = t_new;
#pragma unroll
            for (int j = 0; j < NS; ++j) { y[j] = yn[j]; k1[j] = k7[j]; }
            double fac = (n2 < 1e-20) ? MAX_FACTOR
                                      : fmin(MAX_FACTOR, step_factor(n2));
            if (after_reject) fac = fmin(fac, 1.0);
            h *= fac;
            after_reject = false;
        } else {
            // rejected (also when the norm is NaN): shrink and retry
            const double fac = (n2 > 1.0) ? fmax(MIN_FACTOR, step_factor(n2))
                                          : MIN_FACTOR;
            h *= fac;
            after_reject = true;
        }
    }

    if (steps_out) steps_out[i] = attempts;
    if (TRAJ) {
        if (failed)
            for (int q = idx * NO; q < nt * NO; ++q) my_traj[q] = NAN;
    } else {
        scores[i] = failed ? NAN : objective_finish(P, ss, x);
    }
}

template <class M>
int launch_model(const pb200_problem *p, const double *d_params, int64_t n,
                 int64_t ld, double *d_scores, double *d_traj,
                 int32_t *d_steps, const SolverDev &s, int block,
                 cudaStream_t st) {
    const int64_t bytes = (static_cast<int64_t>(p->dev.n_times) * (1 + M::NO) * 8 + 15) & ~15ll;
    const bool use_smem = bytes <= 96 * 1024;   // keeps >= 2 blocks per SM
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
