// Companion sampler kernels: one thread per chain.
//
//   mh_accept_kernel    Metropolis accept/reject, in place
//                       AdaptiveCovarianceMC.tell
//                         pints/_mcmc/_adaptive_covariance.py:264-278
//                       DifferentialEvolutionMCMC.tell
//                         pints/_mcmc/_differential_evolution.py:314-326
//   hbac_adapt_kernel   Haario-Bardenet mu / sigma / log-lambda update
//                         pints/_mcmc/_adaptive_covariance.py:88-107, :286-300
//                         pints/_mcmc/_haario_bardenet_ac.py:65-68
//   hbac_propose_kernel x + chol(sigma e^{log lambda}) z   (device-RNG mode)
//   de_propose_kernel   x_j + gamma (x_r1 - x_r2) + eps_j
//                         pints/_mcmc/_differential_evolution.py:103-115
//   philox_*            counter-based normals / uniforms / index pairs
//
// Everything that must reproduce the reference bit for bit uses the explicit
// round-to-nearest intrinsics (__dmul_rn/__dadd_rn/__dsub_rn), which the
// compiler never contracts into FMAs: NumPy rounds every product and sum.
#include "pb200_common.cuh"
#include "pb200_sampler_dev.cuh"

namespace {

__global__ void mh_accept_kernel(double *__restrict__ cur, double *__restrict__ cur_f,
                                 const double *__restrict__ prop,
                                 const double *__restrict__ fx,
                                 const double *__restrict__ log_u,
                                 uint8_t *__restrict__ accepted,
                                 double *__restrict__ log_ratio, int64_t n, int d,
                                 int require_finite) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double f_new = fx[j];
    const double ratio = __dsub_rn(f_new, cur_f[j]);
    if (log_ratio) log_ratio[j] = ratio;
    bool acc = log_u[j] < ratio;                 // NaN compares false
    if (require_finite) acc = acc && isfinite(f_new);
    if (acc) {
        for (int k = 0; k < d; ++k) cur[j * d + k] = prop[j * d + k];
        cur_f[j] = f_new;
    }
    accepted[j] = acc ? 1 : 0;
}

__global__ void hbac_adapt_kernel(const double *__restrict__ cur,
                                  const uint8_t *__restrict__ accepted,
                                  double *__restrict__ mu, double *__restrict__ sigma,
                                  double *__restrict__ log_lambda, int64_t n, int d,
                                  double gamma, double target) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double keep = __dsub_rn(1.0, gamma);
    double dev[PB200_MAX_PARAMS];
    // mu <- (1 - gamma) mu + gamma x
    for (int k = 0; k < d; ++k) {
        const double x = cur[j * d + k];
        const double m = __dadd_rn(__dmul_rn(keep, mu[j * d + k]), __dmul_rn(gamma, x));
        mu[j * d + k] = m;
        dev[k] = __dsub_rn(x, m);
    }
    // sigma <- (1 - gamma) sigma + gamma (x - mu)(x - mu)^T   (the new mu)
    double *sg = sigma + j * static_cast<int64_t>(d) * d;
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            const double outer = __dmul_rn(dev[a], dev[b]);
            sg[a * d + b] = __dadd_rn(__dmul_rn(keep, sg[a * d + b]),
                                      __dmul_rn(gamma, outer));
        }
    // log lambda += gamma (accepted - target)
    const double p = accepted[j] ? 1.0 : 0.0;
    log_lambda[j] = __dadd_rn(log_lambda[j], __dmul_rn(gamma, __dsub_rn(p, target)));
}

// Adaptation step of the three adaptive-covariance samplers.  They share
//   mu    <- (1 - gamma) mu + gamma x                       (x = new current)
//   sigma <- (1 - gamma) sigma + gamma v v^T
// and differ in v and in the scale:
//   HAARIO_BARDENET  v = x - mu,  log lambda += gamma (accepted - target)
//                    pints/_mcmc/_haario_bardenet_ac.py:65-68
//   HAARIO           v = x - mu,  log lambda += gamma (p - target),
//                    p = exp(log_ratio) if log_ratio < 0 else 1
//                    pints/_mcmc/_haario_ac.py:63-66
//   RAO_BLACKWELL    v = (p Y + (1 - p) X) - mu, Y the proposal, X the
//                    previous point; fixed scale
//                    pints/_mcmc/_rao_blackwell_ac.py:58-74
// (mu is the updated one in every case).  `log_ratio < 0` is false for NaN, so a
// NaN score gives p = 1 exactly as in the reference.
__global__ void ac_adapt_kernel(int variant, const double *__restrict__ cur,
                                const double *__restrict__ prev,
                                const double *__restrict__ prop,
                                const uint8_t *__restrict__ accepted,
                                const double *__restrict__ log_ratio,
                                double *__restrict__ mu, double *__restrict__ sigma,
                                double *__restrict__ log_lambda, int64_t n, int d,
                                double gamma, double target) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double keep = __dsub_rn(1.0, gamma);
    double p = accepted[j] ? 1.0 : 0.0;
    if (variant != PB200_AC_HAARIO_BARDENET) {
        const double lr = log_ratio[j];
        p = lr < 0.0 ? sampler_exp(lr) : 1.0;
    }
    double dev[PB200_MAX_PARAMS];
    for (int k = 0; k < d; ++k) {
        const double x = cur[j * d + k];
        const double m = __dadd_rn(__dmul_rn(keep, mu[j * d + k]), __dmul_rn(gamma, x));
        mu[j * d + k] = m;
        double v = x;
        if (variant == PB200_AC_RAO_BLACKWELL)
            v = __dadd_rn(__dmul_rn(p, prop[j * d + k]),
                          __dmul_rn(__dsub_rn(1.0, p), prev[j * d + k]));
        dev[k] = __dsub_rn(v, m);
    }
    double *sg = sigma + j * static_cast<int64_t>(d) * d;
    for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) {
            const double outer = __dmul_rn(dev[a], dev[b]);
            sg[a * d + b] = __dadd_rn(__dmul_rn(keep, sg[a * d + b]),
                                      __dmul_rn(gamma, outer));
        }
    if (variant != PB200_AC_RAO_BLACKWELL)
        log_lambda[j] = __dadd_rn(log_lambda[j], __dmul_rn(gamma, __dsub_rn(p, target)));
}

__global__ void hbac_propose_kernel(const double *__restrict__ cur,
                                    const double *__restrict__ sigma,
                                    const double *__restrict__ log_lambda,
                                    const double *__restrict__ z,
                                    double *__restrict__ prop, int64_t n, int d) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double scale = sampler_exp(log_lambda[j]);
    const double *sg = sigma + j * static_cast<int64_t>(d) * d;
    double L[PB200_MAX_PARAMS * PB200_MAX_PARAMS];
    // in-thread Cholesky of scale * sigma (lower), semi-definite tolerant
    for (int a = 0; a < d; ++a) {
        for (int b = 0; b <= a; ++b) {
            double s = scale * sg[a * d + b];
            for (int k = 0; k < b; ++k) s -= L[a * d + k] * L[b * d + k];
            if (a == b) {
                L[a * d + a] = s > 0.0 ? sqrt(s) : 0.0;
            } else {
                const double piv = L[b * d + b];
                L[a * d + b] = piv > 0.0 ? s / piv : 0.0;
            }
        }
    }
    for (int a = 0; a < d; ++a) {
        double s = cur[j * d + a];
        for (int k = 0; k <= a; ++k) s = fma(L[a * d + k], z[j * d + k], s);
        prop[j * d + a] = s;
    }
}

__global__ void de_propose_kernel(const double *__restrict__ cur_all, int64_t first,
                                  const double *__restrict__ eps,
                                  const int32_t *__restrict__ r1,
                                  const int32_t *__restrict__ r2,
                                  double *__restrict__ prop, int64_t n, int d,
                                  double gamma) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double *xj = cur_all + (first + j) * d;
    const double *xa = cur_all + static_cast<int64_t>(r1[j]) * d;
    const double *xb = cur_all + static_cast<int64_t>(r2[j]) * d;
    for (int k = 0; k < d; ++k) {
        const double step = __dmul_rn(gamma, __dsub_rn(xa[k], xb[k]));
        prop[j * d + k] = __dadd_rn(__dadd_rn(xj[k], step), eps[j * d + k]);
    }
}

__global__ void philox_uniform_kernel(double *out, int64_t n, uint64_t seed,
                                      uint64_t offset) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (2 * t >= n) return;
    const U4 r = philox4x32_10(offset + t, 1, seed);
    out[2 * t] = u53(r.x, r.y);
    if (2 * t + 1 < n) out[2 * t + 1] = u53(r.z, r.w);
}

__global__ void philox_normal_kernel(double *out, int64_t n, uint64_t seed,
                                     uint64_t offset) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (2 * t >= n) return;
    const U4 r = philox4x32_10(offset + t, 2, seed);
    const double u1 = u53(r.x, r.y), u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * sampler_log(u1));
    double s, c;
    sampler_sincospi(2.0 * u2, &s, &c);
    out[2 * t] = rad * c;
    if (2 * t + 1 < n) out[2 * t + 1] = rad * s;
}

__global__ void philox_pairs_kernel(int32_t *r1, int32_t *r2, int64_t n, int64_t first,
                                    int64_t n_total, uint64_t seed, uint64_t offset) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t self = first + t;
    const U4 r = philox4x32_10(offset + t, 3, seed);
    // first index among the n_total - 1 others, second among the n_total - 2 left
    int64_t a = static_cast<int64_t>((static_cast<uint64_t>(r.x) * static_cast<uint64_t>(n_total - 1)) >> 32);
    int64_t b = static_cast<int64_t>((static_cast<uint64_t>(r.y) * static_cast<uint64_t>(n_total - 2)) >> 32);
    if (a >= self) ++a;
    const int64_t lo = a < self ? a : self, hi = a < self ? self : a;
    if (b >= lo) ++b;
    if (b >= hi) ++b;
    r1[t] = static_cast<int32_t>(a);
    r2[t] = static_cast<int32_t>(b);
}

__global__ void fp64_probe_kernel(double *sink, int64_t iters) {
    const double a = 1.0000000001, b = 1e-9;
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int64_t i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) sink[0] = s;       // never true; keeps the chains alive
}

inline unsigned blocks_for(int64_t n, int block) {
    return static_cast<unsigned>((n + block - 1) / block);
}

}  // namespace

extern "C" {

int pb200_mh_accept_ratio(double *d_current, double *d_current_f, const double *d_proposed,
                          const double *d_fx, const double *d_log_u, uint8_t *d_accepted,
                          double *d_log_ratio, int64_t n, int32_t d, int32_t require_finite,
                          void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_mh_accept: bad n=%lld or d=%d", (long long)n, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    mh_accept_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current, d_current_f, d_proposed, d_fx, d_log_u, d_accepted, d_log_ratio, n, d,
        require_finite);
    return pb200_check_cuda(cudaGetLastError(), "mh_accept_kernel launch");
}

int pb200_mh_accept(double *d_current, double *d_current_f, const double *d_proposed,
                    const double *d_fx, const double *d_log_u, uint8_t *d_accepted,
                    int64_t n, int32_t d, int32_t require_finite, void *stream) {
    return pb200_mh_accept_ratio(d_current, d_current_f, d_proposed, d_fx, d_log_u, d_accepted,
                                 nullptr, n, d, require_finite, stream);
}

int pb200_ac_adapt(int32_t variant, const double *d_current, const double *d_previous,
                   const double *d_proposed, const uint8_t *d_accepted,
                   const double *d_log_ratio, double *d_mu, double *d_sigma,
                   double *d_log_lambda, int64_t n, int32_t d, double gamma, double target,
                   void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_ac_adapt: bad n=%lld or d=%d", (long long)n, d);
        return PB200_EINVAL;
    }
    if (variant < PB200_AC_HAARIO_BARDENET || variant > PB200_AC_RAO_BLACKWELL) {
        pb200_set_error("pb200_ac_adapt: unknown variant %d", variant);
        return PB200_EUNSUPPORTED;
    }
    if (n == 0) return PB200_OK;
    if (!d_current || !d_accepted || !d_mu || !d_sigma ||
        (variant != PB200_AC_RAO_BLACKWELL && !d_log_lambda) ||
        (variant != PB200_AC_HAARIO_BARDENET && !d_log_ratio) ||
        (variant == PB200_AC_RAO_BLACKWELL && (!d_previous || !d_proposed))) {
        pb200_set_error("pb200_ac_adapt: a buffer this variant needs is NULL");
        return PB200_EINVAL;
    }
    ac_adapt_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        variant, d_current, d_previous, d_proposed, d_accepted, d_log_ratio, d_mu, d_sigma,
        d_log_lambda, n, d, gamma, target);
    return pb200_check_cuda(cudaGetLastError(), "ac_adapt_kernel launch");
}

int pb200_hbac_adapt(const double *d_current, const uint8_t *d_accepted, double *d_mu,
                     double *d_sigma, double *d_log_lambda, int64_t n, int32_t d,
                     double gamma, double target, void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_hbac_adapt: bad n=%lld or d=%d", (long long)n, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    hbac_adapt_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current, d_accepted, d_mu, d_sigma, d_log_lambda, n, d, gamma, target);
    return pb200_check_cuda(cudaGetLastError(), "hbac_adapt_kernel launch");
}

int pb200_hbac_propose(const double *d_current, const double *d_sigma,
                       const double *d_log_lambda, const double *d_z, double *d_proposed,
                       int64_t n, int32_t d, void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_hbac_propose: bad n=%lld or d=%d", (long long)n, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    hbac_propose_kernel<<<blocks_for(n, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current, d_sigma, d_log_lambda, d_z, d_proposed, n, d);
    return pb200_check_cuda(cudaGetLastError(), "hbac_propose_kernel launch");
}

int pb200_de_propose(const double *d_current_all, int64_t first, const double *d_eps,
                     const int32_t *d_r1, const int32_t *d_r2, double *d_proposed,
                     int64_t n, int32_t d, double gamma, void *stream) {
    if (n < 0 || first < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_de_propose: bad n=%lld, first=%lld or d=%d",
                        (long long)n, (long long)first, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    de_propose_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current_all, first, d_eps, d_r1, d_r2, d_proposed, n, d, gamma);
    return pb200_check_cuda(cudaGetLastError(), "de_propose_kernel launch");
}

__global__ void iter_advance_kernel(int64_t *counter) { *counter += 1; }

// ask of the adaptive-covariance family on the device: normals + Cholesky
// proposal in one launch (= pb200_philox_normal then pb200_hbac_propose)
__global__ void ac_ask_kernel(const double *__restrict__ cur, const double *__restrict__ sigma,
                              const double *__restrict__ log_lambda,
                              double *__restrict__ prop, int64_t n, int d, uint64_t seed,
                              uint64_t offset, const IterTable it) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (it.rows) offset = static_cast<uint64_t>(it.row()[0]);
    ac_ask_one(cur, sigma, log_lambda, prop, j, d, seed, offset);
}

// tell of the family on the device, one launch: log-uniform, accept/reject,
// acceptance count, adaptation (variant < 0: none) and the chain sample
// (= pb200_philox_uniform, log, pb200_mh_accept[_ratio], pb200_ac_adapt, copy)
__global__ void ac_tell_kernel(int variant, int require_finite, double *__restrict__ cur,
                               double *__restrict__ cur_f, const double *__restrict__ prop,
                               const double *__restrict__ fx, uint8_t *__restrict__ accepted,
                               int64_t *__restrict__ acc_count, double *__restrict__ mu,
                               double *__restrict__ sigma, double *__restrict__ log_lambda,
                               double *__restrict__ samples, int64_t sample_stride,
                               int64_t sample_offset, int64_t n, int d, double gamma,
                               double target, uint64_t seed, uint64_t offset,
                               const IterTable it) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (it.rows) {
        const int64_t *r = it.row();
        offset = static_cast<uint64_t>(r[2]);
        sample_offset = r[3];
        gamma = slot_f64(r[5]);
        variant = static_cast<int>(r[6]);
    }
    ac_tell_one(variant, require_finite, cur, cur_f, prop, fx[j], accepted, acc_count, mu, sigma,
                log_lambda, samples, sample_stride, sample_offset, j, d, gamma, target, seed,
                offset);
}

static int ac_ask_impl(const double *d_current, const double *d_sigma,
                       const double *d_log_lambda, double *d_proposed, int64_t n, int32_t d,
                       uint64_t seed, uint64_t offset, IterTable it, void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_ac_ask: bad n=%lld or d=%d", (long long)n, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    ac_ask_kernel<<<blocks_for(n, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current, d_sigma, d_log_lambda, d_proposed, n, d, seed, offset, it);
    return pb200_check_cuda(cudaGetLastError(), "ac_ask_kernel launch");
}

int pb200_ac_ask(const double *d_current, const double *d_sigma, const double *d_log_lambda,
                 double *d_proposed, int64_t n, int32_t d, uint64_t seed, uint64_t offset,
                 void *stream) {
    return ac_ask_impl(d_current, d_sigma, d_log_lambda, d_proposed, n, d, seed, offset,
                       IterTable{nullptr, nullptr}, stream);
}

static int iter_table_ok(const int64_t *d_rows, const int64_t *d_counter, const char *who) {
    if (d_rows && d_counter) return PB200_OK;
    pb200_set_error("%s: the iteration table and its counter are required", who);
    return PB200_EINVAL;
}

int pb200_ac_ask_it(const double *d_current, const double *d_sigma, const double *d_log_lambda,
                    double *d_proposed, int64_t n, int32_t d, uint64_t seed,
                    const int64_t *d_rows, const int64_t *d_counter, void *stream) {
    if (int rc = iter_table_ok(d_rows, d_counter, "pb200_ac_ask_it")) return rc;
    return ac_ask_impl(d_current, d_sigma, d_log_lambda, d_proposed, n, d, seed, 0,
                       IterTable{d_rows, d_counter}, stream);
}

int pb200_iter_advance(int64_t *d_counter, void *stream) {
    if (!d_counter) { pb200_set_error("pb200_iter_advance: NULL counter"); return PB200_EINVAL; }
    iter_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(d_counter);
    return pb200_check_cuda(cudaGetLastError(), "iter_advance_kernel launch");
}

// ask of the differential-evolution sampler on the device, one launch:
// error draws, partner indices and proposal (= pb200_philox_normal scaled by
// b*, pb200_philox_pairs, pb200_de_propose)
__global__ void de_ask_kernel(const double *__restrict__ cur_all, int64_t first,
                              int64_t n_total, const double *__restrict__ b_star,
                              double *__restrict__ prop, int64_t n, int d, double gamma,
                              uint64_t seed, uint64_t offset_eps, uint64_t offset_pairs,
                              const IterTable it) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (it.rows) {
        const int64_t *r = it.row();
        offset_eps = static_cast<uint64_t>(r[0]);
        offset_pairs = static_cast<uint64_t>(r[1]);
        gamma = slot_f64(r[4]);
    }
    const int64_t self = first + j;
    const U4 r = philox4x32_10(offset_pairs + j, 3, seed);
    int64_t a = static_cast<int64_t>((static_cast<uint64_t>(r.x) * static_cast<uint64_t>(n_total - 1)) >> 32);
    int64_t b = static_cast<int64_t>((static_cast<uint64_t>(r.y) * static_cast<uint64_t>(n_total - 2)) >> 32);
    if (a >= self) ++a;
    const int64_t lo = a < self ? a : self, hi = a < self ? self : a;
    if (b >= lo) ++b;
    if (b >= hi) ++b;
    const double *xj = cur_all + self * d;
    const double *xa = cur_all + a * d;
    const double *xb = cur_all + b * d;
    for (int k = 0; k < d; ++k) {
        const double eps = philox_normal_at(j * d + k, seed, offset_eps) * b_star[k];
        const double step = __dmul_rn(gamma, __dsub_rn(xa[k], xb[k]));
        prop[j * d + k] = __dadd_rn(__dadd_rn(xj[k], step), eps);
    }
}

static int de_ask_impl(const double *d_current_all, int64_t first, int64_t n_total,
                       const double *d_b_star, double *d_proposed, int64_t n, int32_t d,
                       double gamma, uint64_t seed, uint64_t offset_eps,
                       uint64_t offset_pairs, IterTable it, void *stream) {
    if (n < 0 || first < 0 || n_total < 3 || first + n > n_total || d < 1 ||
        d > PB200_MAX_PARAMS) {
        pb200_set_error("pb200_de_ask: bad n=%lld, first=%lld, n_total=%lld or d=%d",
                        (long long)n, (long long)first, (long long)n_total, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    de_ask_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_current_all, first, n_total, d_b_star, d_proposed, n, d, gamma, seed, offset_eps,
        offset_pairs, it);
    return pb200_check_cuda(cudaGetLastError(), "de_ask_kernel launch");
}

int pb200_de_ask(const double *d_current_all, int64_t first, int64_t n_total,
                 const double *d_b_star, double *d_proposed, int64_t n, int32_t d,
                 double gamma, uint64_t seed, uint64_t offset_eps, uint64_t offset_pairs,
                 void *stream) {
    return de_ask_impl(d_current_all, first, n_total, d_b_star, d_proposed, n, d, gamma, seed,
                       offset_eps, offset_pairs, IterTable{nullptr, nullptr}, stream);
}

int pb200_de_ask_it(const double *d_current_all, int64_t first, int64_t n_total,
                    const double *d_b_star, double *d_proposed, int64_t n, int32_t d,
                    uint64_t seed, const int64_t *d_rows, const int64_t *d_counter,
                    void *stream) {
    if (int rc = iter_table_ok(d_rows, d_counter, "pb200_de_ask_it")) return rc;
    return de_ask_impl(d_current_all, first, n_total, d_b_star, d_proposed, n, d, 0.0, seed, 0, 0,
                       IterTable{d_rows, d_counter}, stream);
}

static int ac_tell_impl(int32_t variant, int32_t require_finite, double *d_current,
                        double *d_current_f,
                        const double *d_proposed, const double *d_fx, uint8_t *d_accepted,
                        int64_t *d_acceptance_count, double *d_mu, double *d_sigma,
                        double *d_log_lambda, double *d_samples, int64_t sample_stride,
                        int64_t sample_offset, int64_t n, int32_t d, double gamma, double target,
                        uint64_t seed, uint64_t offset, IterTable it, void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS || variant > PB200_AC_RAO_BLACKWELL) {
        pb200_set_error("pb200_ac_tell: bad n=%lld, d=%d or variant=%d", (long long)n, d, variant);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    if (!d_current || !d_current_f || !d_proposed || !d_fx || !d_accepted ||
        (variant >= 0 && (!d_mu || !d_sigma || !d_log_lambda))) {
        pb200_set_error("pb200_ac_tell: a buffer this call needs is NULL");
        return PB200_EINVAL;
    }
    ac_tell_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        variant, require_finite, d_current, d_current_f, d_proposed, d_fx, d_accepted, d_acceptance_count,
        d_mu, d_sigma, d_log_lambda, d_samples, sample_stride, sample_offset, n, d, gamma,
        target, seed, offset, it);
    return pb200_check_cuda(cudaGetLastError(), "ac_tell_kernel launch");
}

int pb200_ac_tell(int32_t variant, int32_t require_finite, double *d_current,
                  double *d_current_f,
                  const double *d_proposed, const double *d_fx, uint8_t *d_accepted,
                  int64_t *d_acceptance_count, double *d_mu, double *d_sigma,
                  double *d_log_lambda, double *d_samples, int64_t sample_stride,
                  int64_t sample_offset, int64_t n, int32_t d, double gamma, double target,
                  uint64_t seed, uint64_t offset, void *stream) {
    return ac_tell_impl(variant, require_finite, d_current, d_current_f, d_proposed, d_fx,
                        d_accepted, d_acceptance_count, d_mu, d_sigma, d_log_lambda, d_samples,
                        sample_stride, sample_offset, n, d, gamma, target, seed, offset,
                        IterTable{nullptr, nullptr}, stream);
}

int pb200_ac_tell_it(int32_t variant, int32_t require_finite, double *d_current,
                     double *d_current_f, const double *d_proposed, const double *d_fx,
                     uint8_t *d_accepted, int64_t *d_acceptance_count, double *d_mu,
                     double *d_sigma, double *d_log_lambda, double *d_samples,
                     int64_t sample_stride, int64_t n, int32_t d, double target, uint64_t seed,
                     const int64_t *d_rows, const int64_t *d_counter, void *stream) {
    if (int rc = iter_table_ok(d_rows, d_counter, "pb200_ac_tell_it")) return rc;
    return ac_tell_impl(variant, require_finite, d_current, d_current_f, d_proposed, d_fx,
                        d_accepted, d_acceptance_count, d_mu, d_sigma, d_log_lambda, d_samples,
                        sample_stride, 0, n, d, 0.0, target, seed, 0,
                        IterTable{d_rows, d_counter}, stream);
}

// ---- device-resident xNES (pints/_optimisers/_xnes.py:63-91, :147-182) ----
// ask: z_i ~ N(0, I) from Philox (the numbers pb200_philox_normal(n d, seed,
// offset) gives), x_i = mu + A z_i, and the boundary test of :79-87.  A point
// outside the boundaries is never evaluated by the reference; here its row of
// x_eval holds mu instead (a harmless point, no stream compaction, no host
// round trip) and the caller overrides its score with +inf from `inside`.
__global__ void xnes_ask_kernel(const double *__restrict__ mu, const double *__restrict__ A,
                                const double *__restrict__ lower,
                                const double *__restrict__ upper, double *__restrict__ z,
                                double *__restrict__ x, double *__restrict__ x_eval,
                                uint8_t *__restrict__ inside, int64_t n, int d,
                                uint64_t seed, const uint64_t *__restrict__ d_offset) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t offset = *d_offset;
    double zi[PB200_MAX_PARAMS];
    for (int k = 0; k < d; ++k) {
        zi[k] = philox_normal_at(i * d + k, seed, offset);
        z[i * d + k] = zi[k];
    }
    bool in = true;
    double xi[PB200_MAX_PARAMS];
    for (int a = 0; a < d; ++a) {
        // mu + dot(A[a, :], z) with the products added left to right
        double acc = __dmul_rn(A[a * d], zi[0]);
        for (int k = 1; k < d; ++k) acc = __dadd_rn(acc, __dmul_rn(A[a * d + k], zi[k]));
        xi[a] = __dadd_rn(mu[a], acc);
        x[i * d + a] = xi[a];
        if (lower) in = in && !(xi[a] < lower[a]) && !(xi[a] >= upper[a]);
    }
    inside[i] = in ? 1 : 0;
    for (int a = 0; a < d; ++a) x_eval[i * d + a] = in ? xi[a] : mu[a];
}

// tell: the utility-weighted sums  Gd = sum_r u_r z_(r),  Gm = sum_r u_r z_(r) z_(r)^T
// over the population in rank order (order[r] = index of the r-th best).
// Deterministic: block b sums a contiguous range of ranks serially, one thread
// per output quantity; a second single-block pass adds the partial sums in
// block order.
constexpr int XNES_SPLIT = 16;     // threads per output quantity (fewer for large d)
__global__ void xnes_sums_kernel(const double *__restrict__ z, const int64_t *__restrict__ order,
                                 const double *__restrict__ us, double *__restrict__ partial,
                                 int64_t n, int d, int split) {
    // thread (q, s): quantity q over the ranks r0 + s, r0 + s + SPLIT, ... of
    // this block's range (independent gathers in flight), then the SPLIT
    // partial sums of a quantity are added in a fixed order
    extern __shared__ double part[];
    const int m = d + d * d;
    const int q = threadIdx.x / split, sub = threadIdx.x % split;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = per * blockIdx.x, r1 = r0 + per < n ? r0 + per : n;
    if (q < m) {
        const int a = q < d ? q : (q - d) / d, b = q < d ? -1 : (q - d) % d;
        double acc = 0.0;
        for (int64_t r = r0 + sub; r < r1; r += split) {
            const double *zr = z + order[r] * d;
            const double term = b < 0 ? zr[a] : zr[a] * zr[b];
            acc = fma(us[r], term, acc);
        }
        part[threadIdx.x] = acc;
    }
    __syncthreads();
    if (q < m && sub == 0) {
        double acc = 0.0;
        for (int k = 0; k < split; ++k) acc += part[q * split + k];
        partial[static_cast<int64_t>(blockIdx.x) * m + q] = acc;
    }
}

__global__ void xnes_sums_final_kernel(const double *__restrict__ partial, double *__restrict__ out,
                                       int blocks, int m, const int64_t *__restrict__ order,
                                       const double *__restrict__ x,
                                       const double *__restrict__ fx, int d) {
    const int q = threadIdx.x;
    if (x && q < d + 1) {
        // the best point of the generation and its score, for the same read-back
        const int64_t best = order[0];
        out[m + q] = q < d ? x[best * d + q] : fx[best];
    }
    if (q >= m) return;
    double acc = 0.0;
    for (int b = 0; b < blocks; ++b) acc += partial[static_cast<int64_t>(b) * m + q];
    out[q] = acc;
}

// The d x d part of tell() (pints/_optimisers/_xnes.py:170-182; SNES:
// pints/_optimisers/_snes.py:150-166) in one block, so that the host is out of
// the generation loop:
//   mu <- mu + eta_mu A Gd,   A <- A expm(eta_A / 2 (Gm - sum(u) I))
// expm by scaling and squaring of the Taylor series: the argument is divided by
// 2^s until its 1-norm is at most 1/2, the series is summed to degree 18 in
// Horner form (remainder below 2e-23), the result squared s times.
// `state` = [mu (d), A (d, d), Philox offset (u64)], `sums` = what
// pb200_xnes_sums wrote, `best` = [f_best, x_best (d), f_guessed,
// generations (as double)], `hist` (optional) row g mod hist_rows receives
// [f_best, x_best (d), f_guessed] after generation g.
constexpr int XNES_TAYLOR = 18;
__global__ void xnes_update_kernel(double *__restrict__ state, const double *__restrict__ sums,
                                   double *__restrict__ best, double *__restrict__ hist,
                                   int hist_rows, double eta_mu, double eta_A, double sum_us,
                                   int d, uint64_t offset_step, int separable) {
    __shared__ double M[PB200_MAX_PARAMS * PB200_MAX_PARAMS];
    __shared__ double E[PB200_MAX_PARAMS * PB200_MAX_PARAMS];
    __shared__ double T[PB200_MAX_PARAMS * PB200_MAX_PARAMS];
    __shared__ double A0[PB200_MAX_PARAMS * PB200_MAX_PARAMS];
    __shared__ double col[PB200_MAX_PARAMS];
    __shared__ int s_shift;
    const int q = threadIdx.x, dd = d * d, m = d + dd;
    const int a = q / d, b = q % d;
    double *mu = state, *A = state + d;
    if (q < dd) A0[q] = A[q];
    __syncthreads();
    if (separable) {
        // sigma <- sigma exp(eta / 2 (Gm_aa - sum(u))), mu <- mu + eta_mu sigma Gd
        if (q < d) {
            const double sigma = A0[q * d + q];
            mu[q] = mu[q] + eta_mu * sigma * sums[q];
            A[q * d + q] = sigma * exp(0.5 * eta_A * (sums[d + q * d + q] - sum_us));
        }
    } else {
        if (q < dd) M[q] = 0.5 * eta_A * (sums[d + q] - (a == b ? sum_us : 0.0));
        __syncthreads();
        if (q < d) {
            double c = 0.0;
            for (int r = 0; r < d; ++r) c += fabs(M[r * d + q]);
            col[q] = c;
        }
        __syncthreads();
        if (q == 0) {
            double norm = 0.0;
            for (int k = 0; k < d; ++k) norm = fmax(norm, col[k]);
            int sh = 0;
            // NaN or inf: no scaling, the NaNs travel on as they do on the host
            while (norm > 0.5 && sh < 60 && norm < INFINITY) { norm *= 0.5; ++sh; }
            s_shift = sh;
        }
        __syncthreads();
        const int sh = s_shift;
        if (q < dd) {
            M[q] = ldexp(M[q], -sh);
            E[q] = a == b ? 1.0 : 0.0;
        }
        __syncthreads();
        // E = I + M / k E, k = degree .. 1
        for (int k = XNES_TAYLOR; k >= 1; --k) {
            double acc = 0.0;
            if (q < dd) {
                for (int r = 0; r < d; ++r) acc = fma(M[a * d + r], E[r * d + b], acc);
                acc = acc / static_cast<double>(k) + (a == b ? 1.0 : 0.0);
            }
            __syncthreads();
            if (q < dd) E[q] = acc;
            __syncthreads();
        }
        for (int k = 0; k < sh; ++k) {
            double acc = 0.0;
            if (q < dd)
                for (int r = 0; r < d; ++r) acc = fma(E[a * d + r], E[r * d + b], acc);
            __syncthreads();
            if (q < dd) E[q] = acc;
            __syncthreads();
        }
        if (q < dd) {
            double acc = 0.0;
            for (int r = 0; r < d; ++r) acc = fma(A0[a * d + r], E[r * d + b], acc);
            T[q] = acc;
        }
        if (q < d) {
            double acc = 0.0;
            for (int r = 0; r < d; ++r) acc = fma(A0[q * d + r], sums[r], acc);
            mu[q] = mu[q] + eta_mu * acc;
        }
        __syncthreads();
        if (q < dd) A[q] = T[q];
    }
    // best point so far (pints/_optimisers/_xnes.py:160-165) and bookkeeping
    __syncthreads();
    const double f_gen = sums[m + d];
    const bool better = f_gen < best[0];
    const double generation = best[d + 2];
    if (q < d && better) best[1 + q] = sums[m + q];
    __syncthreads();
    if (q == 0) {
        if (better) best[0] = f_gen;
        best[d + 1] = f_gen;
        best[d + 2] = generation + 1.0;
        uint64_t *offset = reinterpret_cast<uint64_t *>(state + d + dd);
        *offset += offset_step;
    }
    __syncthreads();
    if (hist && q < d + 2) {
        const int row = static_cast<int>(static_cast<long long>(generation) % hist_rows);
        hist[row * (d + 2) + q] = best[q];
    }
}

int pb200_xnes_ask(const double *d_mu, const double *d_A, const double *d_lower,
                   const double *d_upper, double *d_z, double *d_x, double *d_x_eval,
                   uint8_t *d_inside, int64_t n, int32_t d, uint64_t seed,
                   const uint64_t *d_offset, void *stream) {
    if (n < 0 || d < 1 || d > PB200_MAX_PARAMS || !d_offset ||
        (d_lower == nullptr) != (d_upper == nullptr)) {
        pb200_set_error("pb200_xnes_ask: bad n=%lld, d=%d or half a boundary", (long long)n, d);
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    xnes_ask_kernel<<<blocks_for(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_mu, d_A, d_lower, d_upper, d_z, d_x, d_x_eval, d_inside, n, d, seed, d_offset);
    return pb200_check_cuda(cudaGetLastError(), "xnes_ask_kernel launch");
}

int pb200_xnes_sums(const double *d_z, const int64_t *d_order, const double *d_us,
                    double *d_partial, int32_t n_blocks, double *d_out, const double *d_x,
                    const double *d_fx, int64_t n, int32_t d, void *stream) {
    const int m = d + d * d;
    if (n < 1 || d < 1 || d > PB200_MAX_PARAMS || n_blocks < 1) {
        pb200_set_error("pb200_xnes_sums: bad n=%lld, d=%d or n_blocks=%d", (long long)n, d, n_blocks);
        return PB200_EINVAL;
    }
    const int threads = (m + 31) & ~31;
    const int split = 1024 / m < XNES_SPLIT ? 1024 / m : XNES_SPLIT;   // m <= 272
    const int wide = (m * split + 31) & ~31;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    xnes_sums_kernel<<<n_blocks, wide, wide * sizeof(double), st>>>(d_z, d_order, d_us, d_partial,
                                                                    n, d, split);
    xnes_sums_final_kernel<<<1, threads, 0, st>>>(d_partial, d_out, n_blocks, m, d_order,
                                                  (d_x && d_fx) ? d_x : nullptr, d_fx, d);
    return pb200_check_cuda(cudaGetLastError(), "xnes_sums_kernel launch");
}

int pb200_xnes_update(double *d_state, const double *d_sums, double *d_best, double *d_hist,
                      int32_t hist_rows, double eta_mu, double eta_A, double sum_us, int32_t d,
                      uint64_t offset_step, int32_t separable, void *stream) {
    if (d < 1 || d > PB200_MAX_PARAMS || !d_state || !d_sums || !d_best ||
        (d_hist && hist_rows < 1)) {
        pb200_set_error("pb200_xnes_update: bad d=%d, NULL buffer or hist_rows=%d", d, hist_rows);
        return PB200_EINVAL;
    }
    int threads = (d * d + 31) & ~31;
    if (threads < 32) threads = 32;
    xnes_update_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        d_state, d_sums, d_best, d_hist, hist_rows, eta_mu, eta_A, sum_us, d, offset_step,
        separable);
    return pb200_check_cuda(cudaGetLastError(), "xnes_update_kernel launch");
}

int pb200_philox_normal(double *d_out, int64_t n, uint64_t seed, uint64_t offset,
                        void *stream) {
    if (n < 0) { pb200_set_error("pb200_philox_normal: n < 0"); return PB200_EINVAL; }
    if (n == 0) return PB200_OK;
    philox_normal_kernel<<<blocks_for((n + 1) / 2, 256), 256, 0,
                           static_cast<cudaStream_t>(stream)>>>(d_out, n, seed, offset);
    return pb200_check_cuda(cudaGetLastError(), "philox_normal_kernel launch");
}

int pb200_philox_uniform(double *d_out, int64_t n, uint64_t seed, uint64_t offset,
                         void *stream) {
    if (n < 0) { pb200_set_error("pb200_philox_uniform: n < 0"); return PB200_EINVAL; }
    if (n == 0) return PB200_OK;
    philox_uniform_kernel<<<blocks_for((n + 1) / 2, 256), 256, 0,
                            static_cast<cudaStream_t>(stream)>>>(d_out, n, seed, offset);
    return pb200_check_cuda(cudaGetLastError(), "philox_uniform_kernel launch");
}

int pb200_philox_pairs(int32_t *d_r1, int32_t *d_r2, int64_t n, int64_t first,
                       int64_t n_total, uint64_t seed, uint64_t offset, void *stream) {
    if (n < 0 || n_total < 3 || first < 0 || first + n > n_total) {
        pb200_set_error("pb200_philox_pairs: need n_total >= 3 and rows inside it");
        return PB200_EINVAL;
    }
    if (n == 0) return PB200_OK;
    philox_pairs_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_r1, d_r2, n, first, n_total, seed, offset);
    return pb200_check_cuda(cudaGetLastError(), "philox_pairs_kernel launch");
}

int pb200_fp64_probe(double *d_sink, int32_t blocks, int32_t threads, int64_t iters,
                     void *stream) {
    if (blocks < 1 || threads < 1 || threads > 1024 || iters < 0) {
        pb200_set_error("pb200_fp64_probe: bad launch shape");
        return PB200_EINVAL;
    }
    fp64_probe_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(d_sink, iters);
    return pb200_check_cuda(cudaGetLastError(), "fp64_probe_kernel launch");
}

}  // extern "C"
