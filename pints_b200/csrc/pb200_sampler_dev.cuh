// Device-side pieces of the sampler kernels that more than one translation
// unit needs: the counter-based random numbers, the per-iteration table of a
// replayed loop, and one chain's ask() / tell() of the adaptive-covariance
// family (pints/_mcmc/_adaptive_covariance.py:109-131, :236-303;
// pints/_mcmc/_haario_bardenet_ac.py:59-73).  pb200_samplers.cu launches them
// one thread per chain; the chain kernel of pb200_analytic.cu runs them inside
// a persistent warp, between two evaluations of the closed-form model.  The
// same source compiled in both places gives the same bits.
#pragma once
#include "pb200_common.cuh"

// The library's exp / log / sincospi are inlined as sequences of separately
// rounded multiplications and additions that the assembler may or may not
// contract into FMAs depending on the code around them: inlined into the
// unrolled ask / tell of the chain loop, exp(log lambda) came out one ulp away
// from the copy in the one-thread-per-chain kernels (chains differing in the
// last bit from iteration 46 on).  The samplers therefore call them through
// these non-inlined functions: one body each, whatever the caller looks like
// (a call costs ~30 cycles, four times per iteration of a chain).
static __device__ __noinline__ double sampler_exp(double x) { return exp(x); }
static __device__ __noinline__ double sampler_log(double x) { return log(x); }
static __device__ __noinline__ void sampler_sincospi(double x, double *sn, double *cs) {
    sincospi(x, sn, cs);
}

// ------------------------------------------------------------------ Philox
struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint64_t counter, uint64_t stream,
                                            uint64_t seed) {
    uint32_t c0 = static_cast<uint32_t>(counter), c1 = static_cast<uint32_t>(counter >> 32);
    uint32_t c2 = static_cast<uint32_t>(stream), c3 = static_cast<uint32_t>(stream >> 32);
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

// 53-bit uniform in (0, 1)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    const uint64_t bits = (static_cast<uint64_t>(hi >> 5) << 26) | (lo >> 6);
    return (static_cast<double>(bits) + 0.5) * (1.0 / 9007199254740992.0);
}


// element `idx` of the stream pb200_philox_uniform / pb200_philox_normal would
// write for (seed, offset): the fused kernels below draw in place and stay
// bit-identical to the separate launches
__device__ __forceinline__ double philox_uniform_at(int64_t idx, uint64_t seed, uint64_t offset) {
    const U4 r = philox4x32_10(offset + (idx >> 1), 1, seed);
    return (idx & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
}
__device__ __forceinline__ double philox_normal_at(int64_t idx, uint64_t seed, uint64_t offset) {
    const U4 r = philox4x32_10(offset + (idx >> 1), 2, seed);
    const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * sampler_log(u1));
    double sn, cs;
    sampler_sincospi(2.0 * u2, &sn, &cs);
    return (idx & 1) ? rad * sn : rad * cs;
}

// Per-iteration scalars of a sampler loop read from DEVICE memory, so that one
// captured iteration (ask, solve, tell) can be replayed from a CUDA graph
// without the host in the loop: row `*counter` of a table planned in advance,
// eight 64-bit slots per iteration
//   0 ask offset (normals)   1 second ask offset (DE partner pairs)
//   2 tell offset (uniforms) 3 sample offset   4 gamma of ask (DE)
//   5 gamma of tell          6 variant of tell 7 unused
// and a counter that pb200_iter_advance increments at the end of the
// iteration.  rows == nullptr: the by-value arguments are used.
struct IterTable {
    const int64_t *rows;
    const int64_t *counter;
    __device__ __forceinline__ const int64_t *row() const { return rows + 8 * (*counter); }
};
__device__ __forceinline__ double slot_f64(int64_t v) { return __longlong_as_double(v); }


// ask() of one chain: x* = x + chol(sigma e^{log lambda}) z.  `cur`, `sigma`,
// `log_lambda`, `prop` point at THIS chain's rows (global memory, or the shared
// memory copy of the chain loop); z are the normals pb200_philox_normal would
// write at [j d, (j + 1) d) for (seed, offset), drawn here unless given.
// D > 0: the dimension at compile time (d == D): the loops unroll, the factor
// stays in registers and nothing branches -- the same expressions in the same
// order, so the same bits as D = 0 (the chain loop's serial part: one lane runs
// this between two evaluations, and every branch of the generic loops is a
// stall of the whole warp).
template <int D = 0>
__device__ __forceinline__ void ac_ask_row(const double *cur, const double *sg,
                                           const double *log_lambda, double *prop, int64_t j,
                                           int d_, uint64_t seed, uint64_t offset,
                                           const double *z_given = nullptr) {
    const int d = D > 0 ? D : d_;
    const double scale = sampler_exp(log_lambda[0]);
    double L[(D > 0 ? D * D : PB200_MAX_PARAMS * PB200_MAX_PARAMS)];
#pragma unroll
    for (int a = 0; a < d; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            // (explicit roundings: an unrolled `scale * sg - L * L` may contract
            // either product into the FMA, the loop form contracts the second)
            double s = __dmul_rn(scale, sg[a * d + b]);
#pragma unroll
            for (int k = 0; k < b; ++k) s = fma(-L[a * d + k], L[b * d + k], s);
            if (a == b) {
                L[a * d + a] = s > 0.0 ? sqrt(s) : 0.0;
            } else {
                const double piv = L[b * d + b];
                L[a * d + b] = piv > 0.0 ? s / piv : 0.0;
            }
        }
    }
    double z[D > 0 ? D : PB200_MAX_PARAMS];
#pragma unroll
    for (int k = 0; k < d; ++k)
        z[k] = z_given ? z_given[k] : philox_normal_at(j * d + k, seed, offset);
#pragma unroll
    for (int a = 0; a < d; ++a) {
        double s = cur[a];
#pragma unroll
        for (int k = 0; k <= a; ++k) s = fma(L[a * d + k], z[k], s);
        prop[a] = s;
    }
}

// the same with the arrays of all chains (row j is used)
__device__ __forceinline__ void ac_ask_one(const double *cur, const double *sigma,
                                           const double *log_lambda, double *prop, int64_t j,
                                           int d, uint64_t seed, uint64_t offset) {
    ac_ask_row(cur + j * d, sigma + j * static_cast<int64_t>(d) * d, log_lambda + j,
               prop + j * d, j, d, seed, offset);
}

// tell() of one chain given the score f_new of its proposal: log-uniform,
// accept/reject, chain sample, adaptation (variant < 0: none).  `cur`, `cur_f`,
// `prop`, `mu`, `sg` (sigma), `log_lambda` point at THIS chain's rows; `sample`
// at where its new current point goes (or NULL).  Returns the decision.
// D as in ac_ask_row.
template <int D = 0>
__device__ __forceinline__ bool ac_tell_row(int variant, int require_finite, double *cur,
                                            double *cur_f, const double *prop, double f_new,
                                            double *mu, double *sg, double *log_lambda,
                                            double *sample, int64_t j, int d_, double gamma,
                                            double target, uint64_t seed, uint64_t offset) {
    const int d = D > 0 ? D : d_;
    constexpr int DM = D > 0 ? D : PB200_MAX_PARAMS;
    const double log_u = sampler_log(philox_uniform_at(j, seed, offset));
    const double lr = __dsub_rn(f_new, cur_f[0]);
    const bool acc = (log_u < lr) && (!require_finite || isfinite(f_new));
    double x[DM], prev[DM], y[DM];
#pragma unroll
    for (int k = 0; k < d; ++k) {
        prev[k] = cur[k];
        y[k] = prop[k];
        x[k] = acc ? y[k] : prev[k];
    }
    if (acc) {
#pragma unroll
        for (int k = 0; k < d; ++k) cur[k] = x[k];
        cur_f[0] = f_new;
    }
    if (sample) {
#pragma unroll
        for (int k = 0; k < d; ++k) sample[k] = x[k];
    }
    if (variant < 0) return acc;
    // adaptation, as ac_adapt_kernel
    const double keep = __dsub_rn(1.0, gamma);
    double p = acc ? 1.0 : 0.0;
    if (variant != PB200_AC_HAARIO_BARDENET) p = lr < 0.0 ? sampler_exp(lr) : 1.0;
    double dev[DM];
#pragma unroll
    for (int k = 0; k < d; ++k) {
        const double m = __dadd_rn(__dmul_rn(keep, mu[k]), __dmul_rn(gamma, x[k]));
        mu[k] = m;
        double v = x[k];
        if (variant == PB200_AC_RAO_BLACKWELL)
            v = __dadd_rn(__dmul_rn(p, y[k]), __dmul_rn(__dsub_rn(1.0, p), prev[k]));
        dev[k] = __dsub_rn(v, m);
    }
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int b = 0; b < d; ++b) {
            const double outer = __dmul_rn(dev[a], dev[b]);
            sg[a * d + b] = __dadd_rn(__dmul_rn(keep, sg[a * d + b]),
                                      __dmul_rn(gamma, outer));
        }
    if (variant != PB200_AC_RAO_BLACKWELL)
        log_lambda[0] = __dadd_rn(log_lambda[0], __dmul_rn(gamma, __dsub_rn(p, target)));
    return acc;
}

// the same with the arrays of all chains (row j is used), plus the decision
// flag and the acceptance count
__device__ __forceinline__ void ac_tell_one(int variant, int require_finite, double *cur,
                                            double *cur_f, const double *prop, double f_new,
                                            uint8_t *accepted, int64_t *acc_count, double *mu,
                                            double *sigma, double *log_lambda, double *samples,
                                            int64_t sample_stride, int64_t sample_offset,
                                            int64_t j, int d, double gamma, double target,
                                            uint64_t seed, uint64_t offset) {
    const bool acc = ac_tell_row(
        variant, require_finite, cur + j * d, cur_f + j, prop + j * d, f_new, mu + j * d,
        sigma + j * static_cast<int64_t>(d) * d, log_lambda + j,
        samples ? samples + j * sample_stride + sample_offset : nullptr, j, d, gamma, target,
        seed, offset);
    accepted[j] = acc ? 1 : 0;
    if (acc_count) acc_count[j] += acc ? 1 : 0;
}
