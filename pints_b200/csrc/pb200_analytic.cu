// Closed-form forward models fused with the objective: Logistic, Hodgkin-Huxley
// IK (piecewise-exponential gate under a voltage-step protocol) and Constant.
// The time points of one parameter vector are independent, so one WARP owns a
// vector: lanes stride over the observation times (conflict-free shared-memory
// reads of the staged series), accumulate residual squares per output, and a
// shuffle tree produces the sums.  Warps loop over vectors (persistent grid).
//
// Replaces (citations relative to the reference root):
//   LogisticModel._simulate            pints/toy/_logistic_model.py:62-86
//   HodgkinHuxleyIKModel.simulate      pints/toy/_hh_ik_model.py:175-214
//   ConstantModel.simulate             pints/toy/_constant_model.py:75-93
//   Single/MultiOutputProblem.evaluate pints/_core.py:147-153, :257-265
//   objective __call__                 see objective_finish()
#include <cstdlib>
#include "pb200_common.cuh"
#include "pb200_sampler_dev.cuh"

namespace {

constexpr int WARPS = 8;                 // warps per block
#ifndef PB200_CHAIN_T_ASK
#define PB200_CHAIN_T_ASK 1
#endif
#ifndef PB200_CHAIN_T_TELL
#define PB200_CHAIN_T_TELL 1
#endif
#ifndef PB200_POINT_W
#define PB200_POINT_W 8                  // points per lane and trip of the one-output loop
#endif
constexpr int SEG_STRIDE = 11;           // doubles per HH-IK segment record

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ---- per-vector model state (warp-uniform unless noted) -------------------
struct NoBlockAux {};

struct LogisticEval {
    static constexpr bool POINT_AUX = false;
    using BlockAux = NoBlockAux;
    double r, k, c;
    bool zero;
    __device__ __forceinline__ void prepare(const double *x, const ModelConsts &mc,
                                            double *, int, bool, const BlockAux &) {
        r = x[0]; k = x[1];
        const double p0 = mc.c[0];
        zero = (p0 == 0.0) || (k < 0.0);
        c = __dsub_rn(__ddiv_rn(k, p0), 1.0);
    }
    // k / (1 + c * exp(-r t)), products and sums rounded one by one as NumPy
    __device__ __forceinline__ double value(double t, int, const double *) const {
        if (zero) return 0.0;
        const double e = exp(__dmul_rn(-r, t));
        return __ddiv_rn(k, __dadd_rn(1.0, __dmul_rn(c, e)));
    }
};

struct ConstantEval {
    static constexpr bool POINT_AUX = false;
    using BlockAux = NoBlockAux;
    double p[PB200_MAX_OUTPUTS];
    __device__ __forceinline__ void prepare(const double *x, const ModelConsts &,
                                            double *, int no, bool, const BlockAux &) {
        for (int o = 0; o < PB200_MAX_OUTPUTS; ++o)
            p[o] = (o < no) ? __dmul_rn(x[o], static_cast<double>(o + 1)) : 0.0;
    }
    __device__ __forceinline__ double value(double, int o, const double *) const {
        return p[o];
    }
};

// HH-IK: the protocol has n_ev events -> n_ev + 1 constant-voltage segments.
// Lane s prepares segment s: rates, time constant, steady state, and the decay
// over the whole segment; one lane then chains the gate value across events
// (n_next = inf - (inf - n_start) * decay).  Records go to a per-warp table in
// shared memory: [inf, n_start, t_start, -1/tau].
//
// Which segment a time point lies in, and how far into it, depends on the
// protocol only, not on the parameters: when the series is staged in shared
// memory each block works that out ONCE (prepare_points: the row's time is
// replaced by t - t_start and the segment index goes to a byte array), so the
// per-point work of every vector is one multiply, one exp and the polynomial
// instead of a binary search over the events and an IEEE division as well.
//
// Inside a segment the decay is exp(-(t - t_start) / tau), and a lane visits
// points 32 apart: where those are equally spaced (checked bit for bit, once
// per block, per segment) the lane's next exponential is the previous one
// times F = exp(-(32 dt) / tau), one multiplication instead of an exp().  The
// point loop therefore runs segment by segment: one exp() for a lane's first
// point in the segment, the recurrence after that (chains of ~6 factors on the
// reference protocol, rounding <= 1e-15), plain exp() in segments whose
// spacing is not uniform.
struct HhBlockAux {
    int lo[PB200_MAX_EVENTS + 2];          // first point of segment s (lo[n_seg] = n_times)
    int uniform[PB200_MAX_EVENTS + 1];     // points 32 apart are equally spaced in s
    double d32[PB200_MAX_EVENTS + 1];      // that spacing
    int uniform1[PB200_MAX_EVENTS + 1];    // the first 32 points of s are equally spaced
    double d1[PB200_MAX_EVENTS + 1];       // that spacing
    double dt0[PB200_MAX_EVENTS + 1];      // time of the first point into the segment
};

struct HhEval {
    static constexpr bool POINT_AUX = true;
    using BlockAux = HhBlockAux;
    double g_max, e_k;
    int n_ev;
    const double *events;     // in the kernel parameter (constant bank)
    // record of segment s: [inf, n_start, t_start, tau], or for the staged
    // (fast) path [inf, inf - n_start, -1/tau, v - E_k, F, E_a, G, G^2, G^4,
    // G^8, G^16]: a lane's first decay factor in the segment is
    // E_a G^lane = exp(-(dt_a + lane d1) / tau), five selected factors
    // instead of an exp()
    __device__ __forceinline__ void prepare(const double *x, const ModelConsts &mc,
                                            double *table, int, bool fast,
                                            const BlockAux &aux) {
        const int lane = threadIdx.x & 31;
        const double n0 = mc.c[0];
        g_max = mc.c[1]; e_k = mc.c[2];
        const double v_hold = mc.c[3];
        n_ev = static_cast<int>(mc.c[4]);
        events = &mc.c[5];
        const double *volts = &mc.c[5 + n_ev];
        mc_vhold = v_hold;
        mc_volts = volts;
        const double p1 = x[0], p2 = x[1], p3 = x[2], p4 = x[3], p5 = x[4];
        for (int s = lane; s <= n_ev; s += 32) {
            const double v = (s == 0) ? v_hold : volts[s - 1];
            const double t_start = (s == 0) ? 0.0 : events[s - 1];
            // a = p1 * (-(v + 75) + p2) / (exp((-(v + 75) + p2) / p3) - 1)
            const double q = __dadd_rn(-__dadd_rn(v, 75.0), p2);
            const double a = __ddiv_rn(__dmul_rn(p1, q),
                                       __dsub_rn(exp(__ddiv_rn(q, p3)), 1.0));
            // b = p4 * exp((-v - 75) / p5)
            const double b = __dmul_rn(p4, exp(__ddiv_rn(__dsub_rn(-v, 75.0), p5)));
            const double tau = __ddiv_rn(1.0, __dadd_rn(a, b));
            const double inf = __dmul_rn(a, tau);
            double decay = 0.0;
            if (s < n_ev)
                decay = exp(__ddiv_rn(-__dsub_rn(events[s], t_start), tau));
            table[s * SEG_STRIDE + 0] = inf;
            table[s * SEG_STRIDE + 1] = decay;      // replaced by n_start below
            table[s * SEG_STRIDE + 2] = fast ? -1.0 / tau : t_start;
            table[s * SEG_STRIDE + 3] = fast ? __dsub_rn(v, e_k) : tau;
            if (fast && aux.uniform[s] && aux.lo[s + 1] - aux.lo[s] > 32)
                table[s * SEG_STRIDE + 4] = exp(__dmul_rn(aux.d32[s], -1.0 / tau));
            if (fast && aux.uniform1[s]) {
                double *rec = table + s * SEG_STRIDE;
                rec[5] = exp(__dmul_rn(aux.dt0[s], -1.0 / tau));
                const double g = exp(__dmul_rn(aux.d1[s], -1.0 / tau));
                const double g2 = __dmul_rn(g, g), g4 = __dmul_rn(g2, g2);
                const double g8 = __dmul_rn(g4, g4);
                rec[6] = g; rec[7] = g2; rec[8] = g4; rec[9] = g8;
                rec[10] = __dmul_rn(g8, g8);
            }
        }
        __syncwarp();
        if (lane == 0) {
            double n_start = n0;
            for (int s = 0; s <= n_ev; ++s) {
                const double inf = table[s * SEG_STRIDE + 0];
                const double decay = table[s * SEG_STRIDE + 1];
                table[s * SEG_STRIDE + 1] = fast ? __dsub_rn(inf, n_start) : n_start;
                n_start = __dsub_rn(inf, __dmul_rn(__dsub_rn(inf, n_start), decay));
            }
        }
        __syncwarp();
    }
    // segment of t: number of events <= t  (half-open [t_last, t_next))
    __device__ __forceinline__ double value(double t, int, const double *table) const {
        int lo = 0, hi = n_ev;            // answer in [0, n_ev]
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (events[mid] <= t) lo = mid + 1; else hi = mid;
        }
        const int s = lo;
        const double inf = table[s * SEG_STRIDE + 0];
        const double n_start = table[s * SEG_STRIDE + 1];
        const double t_start = table[s * SEG_STRIDE + 2];
        const double tau = table[s * SEG_STRIDE + 3];
        const double e = exp(__ddiv_rn(-__dsub_rn(t, t_start), tau));
        const double nn = __dsub_rn(inf, __dmul_rn(__dsub_rn(inf, n_start), e));
        // g_max * n**4 * (v - E_k)
        const double volt = (s == 0) ? mc_vhold : mc_volts[s - 1];
        const double n2 = __dmul_rn(nn, nn);
        const double n4 = __dmul_rn(n2, n2);
        return __dmul_rn(__dmul_rn(g_max, n4), __dsub_rn(volt, e_k));
    }
    // once per block: segment index and time into the segment of every point,
    // the point range of every segment and whether its spacing is uniform
    __device__ static void prepare_points(double *series, int nt, int row,
                                          const ModelConsts &mc, uint8_t *seg,
                                          BlockAux &aux) {
        const int n_ev = static_cast<int>(mc.c[4]);
        const double *events = &mc.c[5];
        for (int j = threadIdx.x; j < nt; j += blockDim.x) {
            const double t = series[j * row];
            int lo = 0, hi = n_ev;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (events[mid] <= t) lo = mid + 1; else hi = mid;
            }
            seg[j] = static_cast<uint8_t>(lo);
            series[j * row] = __dsub_rn(t, lo == 0 ? 0.0 : events[lo - 1]);
        }
        for (int s = threadIdx.x; s <= n_ev + 1; s += blockDim.x) {
            aux.lo[s] = s > n_ev ? nt : -1;
            if (s <= n_ev) {
                aux.uniform[s] = 1; aux.d32[s] = 0.0;
                aux.uniform1[s] = 1; aux.d1[s] = 0.0; aux.dt0[s] = 0.0;
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += blockDim.x)
            if (j == 0 || seg[j] != seg[j - 1]) aux.lo[seg[j]] = j;
        __syncthreads();
        if (threadIdx.x == 0)                       // segments without points
            for (int s = n_ev; s >= 0; --s)
                if (aux.lo[s] < 0) aux.lo[s] = aux.lo[s + 1];
        __syncthreads();
        for (int j = threadIdx.x; j < nt; j += blockDim.x) {
            const int s = seg[j], a = aux.lo[s];
            if (j >= a + 32) {
                const double d = __dsub_rn(series[j * row], series[(j - 32) * row]);
                const double d0 = __dsub_rn(series[(a + 32) * row], series[a * row]);
                if (j == a + 32) aux.d32[s] = d0;
                if (d != d0) aux.uniform[s] = 0;
            }
            if (j == a) aux.dt0[s] = series[j * row];
            if (j > a && j < a + 32) {
                const double d = __dsub_rn(series[j * row], series[(j - 1) * row]);
                const double d0 = __dsub_rn(series[(a + 1) * row], series[a * row]);
                // lane k's time must be dt_a + k d1 exactly
                const double want = __dadd_rn(series[a * row],
                                              __dmul_rn(static_cast<double>(j - a), d0));
                if (j == a + 1) aux.d1[s] = d0;
                if (d != d0 || want != series[j * row]) aux.uniform1[s] = 0;
            }
        }
        __syncthreads();
    }
    // n = inf - (inf - n_start) exp(-(t - t_start) / tau), with the division by
    // tau as a multiplication by -1/tau (argument error <= 2 ulp, i.e. a
    // relative 1e-14 on the exponential at most)
    __device__ __forceinline__ double value_at(double dt, int s, const double *table) const {
        const double *rec = table + s * SEG_STRIDE;
        const double e = exp(__dmul_rn(dt, rec[2]));
        const double nn = __dsub_rn(rec[0], __dmul_rn(rec[1], e));
        const double n2 = __dmul_rn(nn, nn);
        const double n4 = __dmul_rn(n2, n2);
        return __dmul_rn(__dmul_rn(g_max, n4), rec[3]);
    }
    double mc_vhold;
    const double *mc_volts;
};

// dispatch on the per-point auxiliary data (HH-IK with a staged series)
template <class E>
__device__ __forceinline__ void point_aux(double *series, int nt, int row,
                                          const ModelConsts &mc, uint8_t *seg,
                                          typename E::BlockAux &aux) {
    if constexpr (E::POINT_AUX) E::prepare_points(series, nt, row, mc, seg, aux);
}
template <class E, bool SMEM>
__device__ __forceinline__ double point_value(const E &ev, double t, int o,
                                              const double *table, const uint8_t *seg,
                                              int j) {
    if constexpr (E::POINT_AUX && SMEM) return ev.value_at(t, seg[j], table);
    else return ev.value(t, o, table);
}

// Dynamic shared memory: [staged series][segment index per point (HH-IK)]
// [one segment table per warp, `table_doubles` each]; the block size is chosen
// at launch (8 warps for short series, 32 when the series fills most of an SM's
// shared memory and only one block fits).
// CHAIN: the persistent MCMC loop.  Chains of the adaptive-covariance family
// are independent, so ONE WARP runs one chain through all its iterations
// without ever leaving the kernel: lane 0 proposes (ac_ask_one), the warp
// scores the proposal exactly as an evaluation launch would (same code, same
// reduction order), lane 0 accepts / adapts / stores (ac_tell_one).  No grid
// barrier, no launch per iteration: for a few chains on a closed-form model an
// iteration was four dependent launches of a few microseconds each.
struct ChainDev {
    double *cur, *cur_f, *prop, *fx, *mu, *sigma, *log_lambda, *samples;
    uint8_t *accepted;
    int64_t *acc_count;
    const int64_t *rows;          // planned iterations, 8 slots each (IterTable)
    int64_t first_row, n_iter, sample_stride;
    double target;
    uint64_t seed;
    int32_t require_finite, d;
    uint32_t state_off;           // byte offset of the per-warp chain state in shared memory
};

template <class E, bool SMEM, bool TRAJ, bool BIG, bool CHAIN = false>
__global__ void __launch_bounds__(BIG ? 1024 : 256)
analytic_kernel(const __grid_constant__ ProblemDev P,
                const __grid_constant__ ModelConsts mc,
                const double *__restrict__ params, int64_t n, int64_t ld,
                double *__restrict__ scores, double *__restrict__ traj,
                uint32_t table_off, uint32_t table_doubles, const ScatterDev sc,
                const ChainDev ch) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ typename E::BlockAux aux;
    const int warps = blockDim.x >> 5;

    const int row = P.row, no = P.n_outputs, nt = P.n_times;
    const double *series;
    uint8_t *seg_of = nullptr;
    if (SMEM) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(nt) + PB200_SENTINEL_ROWS) * row * 8 + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series = smem;
        if (E::POINT_AUX) {
            seg_of = reinterpret_cast<uint8_t *>(smem) + bytes;
            point_aux<E>(smem, nt, row, mc, seg_of, aux);
            __syncthreads();
        }
    } else {
        series = P.series;
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    double *table = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + table_off) +
                    static_cast<size_t>(warp) * table_doubles;
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * warps;
    const int n_read = TRAJ ? P.n_model_params : P.n_params;

    // score (or trajectory) of row i, whose parameters are at `prow`; the
    // score is returned in every lane
    auto eval_row = [&](int64_t i, const double *prow) -> double {
        double x[PB200_MAX_PARAMS];
#pragma unroll
        for (int k = 0; k < PB200_MAX_PARAMS; ++k)
            x[k] = (k < n_read) ? prow[k] : 0.0;

        if (!TRAJ) {
            double early;
            if (objective_precheck(P, x, &early)) return early;     // warp-uniform
        }
        E ev;
        ev.prepare(x, mc, table, no, SMEM && E::POINT_AUX, aux);

        double ss[PB200_MAX_OUTPUTS];
#pragma unroll
        for (int o = 0; o < PB200_MAX_OUTPUTS; ++o) ss[o] = 0.0;
        double *my_traj = TRAJ ? traj + i * static_cast<int64_t>(nt) * no : nullptr;

        if constexpr (SMEM && E::POINT_AUX) {
            // segment by segment (see HhEval): a lane's points are 32 apart.
            // The first point of a lane in a segment takes an exp(); the loop
            // over the rest is branch-free: one multiplication by F for the
            // decay, the polynomial, a residual that is zeroed for lanes past
            // the end of the segment, pointers instead of index arithmetic.
            const int n_seg = ev.n_ev + 1;
            const double g_max = ev.g_max;
            double acc = 0.0;
            for (int s = 0; s < n_seg; ++s) {
                const int a = aux.lo[s], b = aux.lo[s + 1];
                if (a >= b) continue;
                const double *rec = table + s * SEG_STRIDE;
                const double inf = rec[0], dn = rec[1], mrt = rec[2], drive = rec[3];
                const double F = rec[4];
                const bool uni = aux.uniform[s] != 0;
                const int last = b - 1;
                auto point = [&](int j, double e) {
                    const bool valid = j <= last;
                    const double nn = __dsub_rn(inf, __dmul_rn(dn, e));
                    const double n2 = __dmul_rn(nn, nn);
                    const double n4 = __dmul_rn(n2, n2);
                    const double v = __dmul_rn(__dmul_rn(g_max, n4), drive);
                    const int jj = valid ? j : last;
                    if (TRAJ) { if (valid) my_traj[j] = v; }
                    double r = __dsub_rn(series[jj * row + 1], v);
                    r = valid ? r : 0.0;
                    acc = __dadd_rn(acc, __dmul_rn(r, r));
                };
                int j = a + lane;
                double e;
                if (aux.uniform1[s] != 0) {             // block-uniform
                    e = rec[5];
                    e = (lane & 1) ? __dmul_rn(e, rec[6]) : e;
                    e = (lane & 2) ? __dmul_rn(e, rec[7]) : e;
                    e = (lane & 4) ? __dmul_rn(e, rec[8]) : e;
                    e = (lane & 8) ? __dmul_rn(e, rec[9]) : e;
                    e = (lane & 16) ? __dmul_rn(e, rec[10]) : e;
                } else {
                    e = exp(__dmul_rn(series[(j <= last ? j : last) * row], mrt));
                }
                point(j, e);
                if (uni) {
                    for (j += 32; j - lane <= last; j += 32) {
                        e = __dmul_rn(e, F);
                        point(j, e);
                    }
                } else {
                    for (j += 32; j - lane <= last; j += 32) {
                        e = exp(__dmul_rn(series[(j <= last ? j : last) * row], mrt));
                        point(j, e);
                    }
                }
            }
            ss[0] = acc;
        } else if (no == 1) {
            // One output (the usual closed-form problem): PB200_POINT_W of a
            // lane's points per trip, so that their exponentials and divisions
            // -- a few hundred cycles of dependent arithmetic each -- overlap.
            // Slots past the end read the last row and add +0 (which leaves any
            // sum as it is); the squares are added in the order of the points:
            // the score is bit for bit that of the one-point loop below.  (A
            // warp of the persistent chain loop has nothing else to hide that
            // latency with: 20 -> 15 us per MCMC iteration at four per trip.)
            constexpr int W = PB200_POINT_W;
            double acc = 0.0;
            for (int j0 = lane; j0 < nt; j0 += 32 * W) {
                double tt[W], dd[W], vv[W];
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int jj = j0 + 32 * u;
                    const int jc = jj < nt ? jj : nt - 1;
                    tt[u] = series[jc * row];
                    dd[u] = series[jc * row + 1];
                }
#pragma unroll
                for (int u = 0; u < W; ++u)
                    vv[u] = point_value<E, SMEM>(ev, tt[u], 0, table, seg_of, j0 + 32 * u);
#pragma unroll
                for (int u = 0; u < W; ++u) {
                    const int jj = j0 + 32 * u;
                    const bool valid = jj < nt;
                    if (TRAJ) { if (valid) my_traj[jj] = vv[u]; }
                    double r = __dsub_rn(dd[u], vv[u]);
                    r = valid ? r : 0.0;
                    acc = __dadd_rn(acc, __dmul_rn(r, r));
                }
            }
            ss[0] = acc;
        } else
        for (int j = lane; j < nt; j += 32) {
            const double t = series[j * row];
#pragma unroll
            for (int o = 0; o < PB200_MAX_OUTPUTS; ++o) {
                if (o < no) {
                    const double v = point_value<E, SMEM>(ev, t, o, table, seg_of, j);
                    if (TRAJ) my_traj[j * no + o] = v;
                    const double r = __dsub_rn(series[j * row + 1 + o], v);
                    ss[o] = __dadd_rn(ss[o], __dmul_rn(r, r));
                }
            }
        }
        if (TRAJ) return 0.0;
#pragma unroll
        for (int o = 0; o < PB200_MAX_OUTPUTS; ++o)
            if (o < no) ss[o] = warp_sum(ss[o]);
        return objective_finish(P, ss, x);
    };

    if constexpr (CHAIN) {
        // chain j = this warp, for all planned iterations.  Its state (current
        // point and score, proposal, mu, sigma, log lambda: 3 d + d^2 + 3
        // doubles) lives in shared memory for the duration of the loop, so an
        // iteration makes no dependent round trip to global memory; only the
        // chain sample is stored (fire and forget).  The d normals of a
        // proposal are drawn by d lanes at once.
        const int64_t j = static_cast<int64_t>(blockIdx.x) * warps + warp;
        if (j < n) {
            const int d = ch.d, dd = d * d;
            double *st = reinterpret_cast<double *>(reinterpret_cast<char *>(smem) + ch.state_off) +
                         static_cast<size_t>(warp) * (3 * d + dd + 3 + PB200_MAX_PARAMS);
            double *s_cur = st, *s_prop = st + d, *s_mu = st + 2 * d, *s_sig = st + 3 * d;
            double *s_curf = s_sig + dd, *s_ll = s_curf + 1, *s_z = s_ll + 2;
            for (int k = lane; k < d; k += 32) {
                s_cur[k] = ch.cur[j * d + k];
                s_mu[k] = ch.mu[j * d + k];
            }
            for (int k = lane; k < dd; k += 32) s_sig[k] = ch.sigma[j * dd + k];
            if (lane == 0) { s_curf[0] = ch.cur_f[j]; s_ll[0] = ch.log_lambda[j]; }
            __syncwarp();
            int64_t n_acc = 0;
            bool acc = false;
            double f_new = 0.0;
            for (int64_t it = 0; it < ch.n_iter; ++it) {
                const int64_t *r = ch.rows + 8 * (ch.first_row + it);
                const uint64_t off_ask = static_cast<uint64_t>(r[0]);
                if (lane < d) s_z[lane] = philox_normal_at(j * d + lane, ch.seed, off_ask);
                __syncwarp();
                // (the serial part, one lane: dimension known at compile time up to 6)
                if (lane == 0) {
                    switch (PB200_CHAIN_T_ASK ? d : 0) {
#define PB200_ASK(D) case D: ac_ask_row<D>(s_cur, s_sig, s_ll, s_prop, j, d, ch.seed, off_ask, s_z); break
                        PB200_ASK(1); PB200_ASK(2); PB200_ASK(3); PB200_ASK(4); PB200_ASK(5); PB200_ASK(6);
#undef PB200_ASK
                        default: ac_ask_row<0>(s_cur, s_sig, s_ll, s_prop, j, d, ch.seed, off_ask, s_z);
                    }
                }
                __syncwarp();
                f_new = eval_row(j, s_prop);
                if (lane == 0) {
                    double *sample = ch.samples ? ch.samples + j * ch.sample_stride + r[3] : nullptr;
                    const int variant = static_cast<int>(r[6]);
                    const double gamma = slot_f64(r[5]);
                    const uint64_t off_tell = static_cast<uint64_t>(r[2]);
                    switch (PB200_CHAIN_T_TELL ? d : 0) {
#define PB200_TELL(D) case D: acc = ac_tell_row<D>(variant, ch.require_finite, s_cur, s_curf, s_prop,  \
                                                   f_new, s_mu, s_sig, s_ll, sample, j, d, gamma,      \
                                                   ch.target, ch.seed, off_tell); break
                        PB200_TELL(1); PB200_TELL(2); PB200_TELL(3); PB200_TELL(4); PB200_TELL(5); PB200_TELL(6);
#undef PB200_TELL
                        default: acc = ac_tell_row<0>(variant, ch.require_finite, s_cur, s_curf, s_prop,
                                                      f_new, s_mu, s_sig, s_ll, sample, j, d, gamma,
                                                      ch.target, ch.seed, off_tell);
                    }
                    n_acc += acc ? 1 : 0;
                }
                __syncwarp();
            }
            // state back to the arrays the launch-per-step kernels use
            for (int k = lane; k < d; k += 32) {
                ch.cur[j * d + k] = s_cur[k];
                ch.mu[j * d + k] = s_mu[k];
                ch.prop[j * d + k] = s_prop[k];
            }
            for (int k = lane; k < dd; k += 32) ch.sigma[j * dd + k] = s_sig[k];
            if (lane == 0) {
                ch.cur_f[j] = s_curf[0];
                ch.log_lambda[j] = s_ll[0];
                ch.fx[j] = f_new;
                ch.accepted[j] = acc ? 1 : 0;
                if (ch.acc_count) ch.acc_count[j] += n_acc;
            }
        }
    } else {
        for (int64_t i = static_cast<int64_t>(blockIdx.x) * warps + warp; i < n;
             i += warps_total) {
            const double score = eval_row(i, params + i * ld);
            if (!TRAJ && lane == 0) put_score(sc, scores, i, i, score);
            __syncwarp();
        }
        if (!TRAJ) exchange_signal(sc);
    }
}

template <class E>
int launch_eval(const pb200_problem *p, const double *d_params, int64_t n,
                int64_t ld, double *d_scores, double *d_traj, const ScatterDev &sc,
                cudaStream_t st) {
    const int64_t bytes = ((static_cast<int64_t>(p->dev.n_times) + PB200_SENTINEL_ROWS) * p->dev.row * 8 + 15) & ~15ll;
    const size_t aux_bytes = E::POINT_AUX ? (static_cast<size_t>(p->dev.n_times) + 15) & ~size_t(15) : 0;
    // per-warp segment table (HH-IK only): n_events + 1 records
    const uint32_t table_doubles = E::POINT_AUX
        ? static_cast<uint32_t>((static_cast<int>(p->mc.c[4]) + 1) * SEG_STRIDE) : 0;
    const bool use_smem = bytes <= 100 * 1024;
    const bool traj = d_traj != nullptr;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // a long staged series leaves room for one block per SM: make it a big one
    const size_t staged = use_smem ? static_cast<size_t>(bytes) + aux_bytes : 0;
    const int warps = staged > 48 * 1024 ? 32 : 8;
    const int per_sm = staged > 48 * 1024 ? 1 : 4;
    const size_t dyn = staged + static_cast<size_t>(warps) * table_doubles * 8;
    int64_t grid = (n + warps - 1) / warps;
    const int64_t cap = static_cast<int64_t>(sms) * per_sm;      // persistent grid
    if (grid > cap) grid = cap;
    if (grid <= 0) return PB200_OK;
#define PB200_LAUNCH(SM, TR, BG)                                                         \
    do {                                                                                 \
        auto kern = analytic_kernel<E, SM, TR, BG>;                                      \
        if (dyn > 32 * 1024) {                                                           \
            int rc = pb200_check_cuda(                                                   \
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     static_cast<int>(dyn)),                             \
                "cudaFuncSetAttribute");                                                 \
            if (rc) return rc;                                                           \
        }                                                                                \
        kern<<<static_cast<unsigned>(grid), warps * 32, dyn, st>>>(                      \
            p->dev, p->mc, d_params, n, ld, d_scores, d_traj,                            \
            static_cast<uint32_t>(staged), table_doubles, sc, ChainDev{});               \
    } while (0)
    if (warps == 32) {                       // implies use_smem
        if (traj) PB200_LAUNCH(true, true, true); else PB200_LAUNCH(true, false, true);
    } else if (use_smem) {
        if (traj) PB200_LAUNCH(true, true, false); else PB200_LAUNCH(true, false, false);
    } else {
        if (traj) PB200_LAUNCH(false, true, false); else PB200_LAUNCH(false, false, false);
    }
#undef PB200_LAUNCH
    return pb200_check_cuda(cudaGetLastError(), "analytic_kernel launch");
}

// the persistent MCMC loop: one warp per chain, all chains resident at once
template <class E>
int launch_chain(const pb200_problem *p, int64_t n, const ChainDev &ch_in, cudaStream_t st) {
    const int64_t bytes = ((static_cast<int64_t>(p->dev.n_times) + PB200_SENTINEL_ROWS) * p->dev.row * 8 + 15) & ~15ll;
    const size_t aux_bytes = E::POINT_AUX ? (static_cast<size_t>(p->dev.n_times) + 15) & ~size_t(15) : 0;
    const uint32_t table_doubles = E::POINT_AUX
        ? static_cast<uint32_t>((static_cast<int>(p->mc.c[4]) + 1) * SEG_STRIDE) : 0;
    if (bytes > 100 * 1024) {
        pb200_set_error("pb200_mcmc_chain_run: the series does not fit shared memory");
        return PB200_EUNSUPPORTED;
    }
    const size_t staged = static_cast<size_t>(bytes) + aux_bytes;
    const bool big = staged > 48 * 1024;
    // few chains: spread them over the SMs (one 2-warp block each) rather than
    // packing eight warps onto one SM
    const int warps = big ? 32 : (n <= 2 * 148 ? 2 : 8);
    const size_t tables = static_cast<size_t>(warps) * table_doubles * 8;
    const size_t state = static_cast<size_t>(warps) *
                         (3 * ch_in.d + ch_in.d * ch_in.d + 3 + PB200_MAX_PARAMS) * 8;
    const size_t dyn = staged + tables + state;
    if (dyn > 200 * 1024) {
        pb200_set_error("pb200_mcmc_chain_run: series, tables and chain state need %zu bytes of "
                        "shared memory", dyn);
        return PB200_EUNSUPPORTED;
    }
    ChainDev ch = ch_in;
    ch.state_off = static_cast<uint32_t>(staged + tables);
    const int64_t grid = (n + warps - 1) / warps;
    const ScatterDev none = {};
#define PB200_LAUNCH(BG)                                                                 \
    do {                                                                                 \
        auto kern = analytic_kernel<E, true, false, BG, true>;                           \
        if (dyn > 32 * 1024) {                                                           \
            int rc = pb200_check_cuda(                                                   \
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                     static_cast<int>(dyn)),                             \
                "cudaFuncSetAttribute");                                                 \
            if (rc) return rc;                                                           \
        }                                                                                \
        kern<<<static_cast<unsigned>(grid), warps * 32, dyn, st>>>(                      \
            p->dev, p->mc, nullptr, n, ch.d, ch.fx, nullptr,                             \
            static_cast<uint32_t>(staged), table_doubles, none, ch);                     \
    } while (0)
    if (big) PB200_LAUNCH(true); else PB200_LAUNCH(false);
#undef PB200_LAUNCH
    return pb200_check_cuda(cudaGetLastError(), "analytic_kernel (chain loop) launch");
}

// ---- unfused baseline, second half: objective from stored trajectories ----
// A streaming kernel: it reads N * n_t * n_o doubles once and writes N scores,
// so HBM bandwidth bounds it.  One warp per trajectory; the observed series is
// staged in shared memory (bulk copy, once per persistent block); the lanes
// sweep the trajectory with 16-byte streaming loads, four in flight per lane
// (2 KB per warp and trip), so that a resident SM keeps > 100 KB on its way.
// VEC = 2: two outputs, a double2 is one time point; VEC = 1: one output, a
// double2 is two time points; VEC = 0: any layout, scalar loads.
constexpr int ST_WARPS = 8;

template <int VEC, int U = 4>
__global__ void __launch_bounds__(ST_WARPS * 32)
score_traj_kernel(const __grid_constant__ ProblemDev P, const double *__restrict__ traj,
                  const double *__restrict__ params, int64_t n, int64_t ld,
                  double *__restrict__ scores, int staged) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int row = P.row, no = P.n_outputs, nt = P.n_times;
    const double *series = P.series;
    if (staged) {
        const uint32_t bytes = static_cast<uint32_t>(
            ((static_cast<int64_t>(nt) + PB200_SENTINEL_ROWS) * row * 8 + 15) & ~15ll);
        stage_series(smem, P.series, bytes, &bar);
        series = smem;
    }
    const int64_t warps_total = static_cast<int64_t>(gridDim.x) * ST_WARPS;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * ST_WARPS + warp; i < n;
         i += warps_total) {
        double x[PB200_MAX_PARAMS];
#pragma unroll
        for (int k = 0; k < PB200_MAX_PARAMS; ++k)
            x[k] = (k < P.n_params) ? params[i * ld + k] : 0.0;
        double early;
        if (objective_precheck(P, x, &early)) {
            if (lane == 0) scores[i] = early;
            continue;
        }
        double ss[PB200_MAX_OUTPUTS];
#pragma unroll
        for (int o = 0; o < PB200_MAX_OUTPUTS; ++o) ss[o] = 0.0;
        const double *tr = traj + i * static_cast<int64_t>(nt) * no;
        const int total = nt * no;
        if (VEC == 2) {
            // pair q = time point q: (output 0, output 1)
            const double2 *tv = reinterpret_cast<const double2 *>(tr);
            int q = lane;
            for (; q + 32 * (U - 1) < nt; q += 32 * U) {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldcs(tv + q + 32 * u);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double *d = series + (q + 32 * u) * 3 + 1;
                    const double r0 = d[0] - v[u].x, r1 = d[1] - v[u].y;
                    ss[0] = fma(r0, r0, ss[0]);
                    ss[1] = fma(r1, r1, ss[1]);
                }
            }
            for (; q < nt; q += 32) {
                const double2 v = __ldcs(tv + q);
                const double *d = series + q * 3 + 1;
                const double r0 = d[0] - v.x, r1 = d[1] - v.y;
                ss[0] = fma(r0, r0, ss[0]);
                ss[1] = fma(r1, r1, ss[1]);
            }
        } else if (VEC == 1) {
            // pair q = time points 2q and 2q + 1 (nt even, checked at launch)
            const double2 *tv = reinterpret_cast<const double2 *>(tr);
            const int pairs = nt / 2;
            int q = lane;
            for (; q + 96 < pairs; q += 128) {
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldcs(tv + q + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double *d = series + (q + 32 * u) * 4 + 1;
                    const double r0 = d[0] - v[u].x, r1 = d[2] - v[u].y;
                    ss[0] = fma(r0, r0, ss[0]);
                    ss[0] = fma(r1, r1, ss[0]);
                }
            }
            for (; q < pairs; q += 32) {
                const double2 v = __ldcs(tv + q);
                const double *d = series + q * 4 + 1;
                const double r0 = d[0] - v.x, r1 = d[2] - v.y;
                ss[0] = fma(r0, r0, ss[0]);
                ss[0] = fma(r1, r1, ss[0]);
            }
        } else {
            // flat, coalesced sweep over the nt*no stored values
            for (int q = lane; q < total; q += 32) {
                const int j = q / no, o = q - j * no;
                const double r = series[j * row + 1 + o] - __ldcs(tr + q);
                const double sq = r * r;
#pragma unroll
                for (int oo = 0; oo < PB200_MAX_OUTPUTS; ++oo)
                    if (oo == o) ss[oo] += sq;
            }
        }
#pragma unroll
        for (int o = 0; o < PB200_MAX_OUTPUTS; ++o)
            if (o < no) ss[o] = warp_sum(ss[o]);
        if (lane == 0) scores[i] = objective_finish(P, ss, x);
    }
}

}  // namespace

int launch_analytic(const pb200_problem *p, const double *d_params, int64_t n,
                    int64_t ld, double *d_scores, double *d_traj,
                    const ScatterDev &sc, cudaStream_t st) {
    switch (p->dev.model) {
        case PB200_MODEL_LOGISTIC:
            return launch_eval<LogisticEval>(p, d_params, n, ld, d_scores, d_traj, sc, st);
        case PB200_MODEL_HH_IK:
            return launch_eval<HhEval>(p, d_params, n, ld, d_scores, d_traj, sc, st);
        case PB200_MODEL_CONSTANT:
            return launch_eval<ConstantEval>(p, d_params, n, ld, d_scores, d_traj, sc, st);
    }
    pb200_set_error("model %d is not a closed-form model", p->dev.model);
    return PB200_EUNSUPPORTED;
}

int pb200_chain_loop(const pb200_problem *p, int64_t n, int32_t d, int32_t require_finite,
                     double *cur, double *cur_f, double *prop, double *fx, uint8_t *accepted,
                     int64_t *acc_count, double *mu, double *sigma, double *log_lambda,
                     double *samples, int64_t sample_stride, double target, uint64_t seed,
                     const int64_t *rows, int64_t first_row, int64_t n_iter, cudaStream_t st) {
    ChainDev ch;
    ch.cur = cur; ch.cur_f = cur_f; ch.prop = prop; ch.fx = fx; ch.mu = mu; ch.sigma = sigma;
    ch.log_lambda = log_lambda; ch.samples = samples; ch.accepted = accepted;
    ch.acc_count = acc_count; ch.rows = rows; ch.first_row = first_row; ch.n_iter = n_iter;
    ch.sample_stride = sample_stride; ch.target = target; ch.seed = seed;
    ch.require_finite = require_finite; ch.d = d;
    switch (p->dev.model) {
        case PB200_MODEL_LOGISTIC: return launch_chain<LogisticEval>(p, n, ch, st);
        case PB200_MODEL_HH_IK: return launch_chain<HhEval>(p, n, ch, st);
        case PB200_MODEL_CONSTANT: return launch_chain<ConstantEval>(p, n, ch, st);
    }
    pb200_set_error("pb200_mcmc_chain_run: model %d is not a closed-form model", p->dev.model);
    return PB200_EUNSUPPORTED;
}

int launch_score_traj(const pb200_problem *p, const double *d_traj,
                      const double *d_params, int64_t n, int64_t ld,
                      double *d_scores, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int no = p->dev.n_outputs, nt = p->dev.n_times;
    const int64_t bytes = ((static_cast<int64_t>(nt) + PB200_SENTINEL_ROWS) * p->dev.row * 8 + 15) & ~15ll;
    const int staged = bytes <= 48 * 1024 ? 1 : 0;          // else read through L1
    // persistent grid: 8 blocks of 8 warps per SM (4 when the series is staged
    // and large), each warp takes every (grid * 8)-th trajectory
    static const int blocks_per_sm = [] {
        const char *e = getenv("PB200_ST_BLOCKS");
        return e ? atoi(e) : 8;
    }();
    int per_sm = blocks_per_sm;
    if (staged && bytes * per_sm > 200 * 1024) per_sm = static_cast<int>(200 * 1024 / bytes);
    int64_t grid = (n + ST_WARPS - 1) / ST_WARPS;
    if (grid > static_cast<int64_t>(sms) * per_sm) grid = static_cast<int64_t>(sms) * per_sm;
    if (grid <= 0) return PB200_OK;
    const size_t dyn = staged ? static_cast<size_t>(bytes) : 0;
    // 16-byte loads need 16-byte aligned trajectories: the buffer itself and,
    // for one output, an even number of time points
    const bool aligned = (reinterpret_cast<uintptr_t>(d_traj) & 15) == 0;
    const int vec = (aligned && no == 2) ? 2 : (aligned && no == 1 && nt % 2 == 0) ? 1 : 0;
#define PB200_ST(V)                                                                        \
    score_traj_kernel<V><<<static_cast<unsigned>(grid), ST_WARPS * 32, dyn, st>>>(         \
        p->dev, d_traj, d_params, n, ld, d_scores, staged)
    static const int unroll = [] {
        const char *e = getenv("PB200_ST_UNROLL");
        return e ? atoi(e) : 8;
    }();
    if (vec == 2 && unroll == 8)
        score_traj_kernel<2, 8><<<static_cast<unsigned>(grid), ST_WARPS * 32, dyn, st>>>(
            p->dev, d_traj, d_params, n, ld, d_scores, staged);
    else if (vec == 2) PB200_ST(2);
    else if (vec == 1) PB200_ST(1);
    else PB200_ST(0);
#undef PB200_ST
    return pb200_check_cuda(cudaGetLastError(), "score_traj_kernel launch");
}
