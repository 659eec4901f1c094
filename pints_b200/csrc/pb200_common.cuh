// Shared device-side definitions of the pints-b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/pints_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pints-b200 kernels are written for sm_100a (Blackwell B200) only"
#endif

// ---------------------------------------------------------------------------
// Problem description passed by value as a kernel parameter (constant bank,
// so every read is a uniform load).
// ---------------------------------------------------------------------------
#define PB200_MAX_MODEL_CONSTS (5 + 2 * PB200_MAX_EVENTS)

struct ProblemDev {
    int32_t model;
    int32_t objective;
    int32_t n_times;
    int32_t n_outputs;
    int32_t n_model_params;       // parameters consumed by the forward model
    int32_t n_params;             // row length (adds sigmas for GAUSSIAN)
    int32_t n_prior;              // 0 = no prior box
    int32_t row;                  // doubles per series row = 1 + n_outputs
    const double *series;         // device, packed [n_times][row]
    double sign;
    double log_prior;
    double oc[2 * PB200_MAX_OUTPUTS];      // objective constants
    double lo[PB200_MAX_PARAMS];
    double hi[PB200_MAX_PARAMS];
    // composed prior (pints/_log_priors.py:212-219): terms added in order;
    // kind 0 = constant c0, kind 1 = Gaussian c0 - c1 (x[idx] - c2)^2
    int32_t n_terms;              // 0 = the constant log_prior above
    int32_t term_kind[PB200_MAX_PARAMS];
    int32_t term_idx[PB200_MAX_PARAMS];
    double term_c[PB200_MAX_PARAMS][3];
};

// log prior of a point inside the box: the constant of a single
// UniformLogPrior, or the terms of a ComposedLogPrior summed in the reference's
// order (0 + p_1 + p_2 + ..., each sum rounded).
__device__ __forceinline__ double log_prior_of(const ProblemDev &P, const double *x) {
    if (P.n_terms == 0) return P.log_prior;
    double acc = 0.0;
    for (int k = 0; k < P.n_terms; ++k) {
        double term = P.term_c[k][0];
        if (P.term_kind[k] == 1) {
            // offset - factor * (x - mean)**2   (pints/_log_priors.py:467-468)
            const double d = __dsub_rn(x[P.term_idx[k]], P.term_c[k][2]);
            term = __dsub_rn(term, __dmul_rn(P.term_c[k][1], __dmul_rn(d, d)));
        }
        acc = __dadd_rn(acc, term);
    }
    return acc;
}

// Model constants live apart so the ODE kernels do not carry the HH-IK table.
struct ModelConsts {
    double c[PB200_MAX_MODEL_CONSTS];
};

struct pb200_problem {
    ProblemDev dev;
    ModelConsts mc;
    int32_t n_model_consts;
    int64_t series_bytes;         // bytes staged into shared memory
    // One word of page-locked, device-mapped host memory (allocated at the first
    // sorted launch): a one-block-per-SM launch leaves 1 = costs narrow, 2 = wide
    // here (from its vectors' attempted steps), and the NEXT launch picks its sort by it
    // (launch_model in pb200_ode.cu); no synchronisation, a stale value only
    // costs a per cent or two of speed, never a result.
    mutable int32_t *sort_hint = nullptr;
};

// Where a score goes.  n = 0: the caller's d_scores[i].  n > 0: the exchange
// step of a multi-GPU job is fused into the kernel.  Every rank owns a staging
// array of 16-byte pairs {score, row} (symmetric memory: each rank's copy is
// mapped into every process over NVLink) and a vector's score is stored into
// ALL copies at pair `offset + slot`, where `slot` is the vector's position in
// the launch order (after the cost sort) and `row` its original row: the lanes
// of a warp hold 32 consecutive slots, so a warp's stores to one rank are one
// contiguous 512-byte run (4 NVLink write packets) however the sort has
// permuted the rows; exchange_finish_kernel on the reading side puts the
// scores back in row order.  With a multicast mapping of the staging array
// (NVSwitch replicates the store) one multimem.st per lane replaces the n
// unicast stores.  A block that has stored all its scores adds its share of
// EXCHANGE_UNIT to this rank's counter on every rank (one system-scope fence,
// then relaxed remote atomics: the shares of a launch add up to exactly
// EXCHANGE_UNIT whatever the grid size); the reader spins (acquire) until the
// counters of all ranks have reached epoch * EXCHANGE_UNIT, so there is no
// barrier launch and no block waits for another one.
#define PB200_EXCHANGE_UNIT_LOG2 20
struct ScatterDev {
    void *stage[PB200_MAX_PEERS];        // rank q's staging array, as mapped in this process
    unsigned long long *flag[PB200_MAX_PEERS];   // this rank's counter in rank q's flag array
    void *stage_mc;                      // multicast address of the staging arrays (or NULL)
    int32_t n;                           // ranks (0 = no exchange)
    int32_t pad;
    int64_t offset;                      // first pair of this rank's block
};

__device__ __forceinline__ void put_score(const ScatterDev &sc, double *scores,
                                          int64_t slot, int64_t i, double v) {
    if (sc.n == 0) {
        scores[i] = v;
        return;
    }
    const int64_t at = (sc.offset + slot) * 16;
    if (sc.stage_mc) {
        const long long ib = static_cast<long long>(i);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(static_cast<char *>(sc.stage_mc) + at),
                     "f"(__int_as_float(__double2loint(v))), "f"(__int_as_float(__double2hiint(v))),
                     "f"(__int_as_float(static_cast<int>(ib & 0xffffffffll))),
                     "f"(__int_as_float(static_cast<int>(ib >> 32)))
                     : "memory");
    } else {
        const double2 pair = make_double2(v, __longlong_as_double(i));
        for (int r = 0; r < sc.n; ++r)
            *reinterpret_cast<double2 *>(static_cast<char *>(sc.stage[r]) + at) = pair;
    }
}

__device__ __forceinline__ void red_add_relaxed_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// End of a kernel that stored scores through put_score with sc.n > 0; every
// thread of the block must reach it (none may have exited).  The block
// barrier orders all stores of the block before thread 0's system-scope
// fence (one fence for all destinations; a st.release per destination would
// serialise one fence each), then the block's share goes to every rank.
__device__ __forceinline__ void exchange_signal(const ScatterDev &sc) {
    if (sc.n == 0) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned long long blocks = gridDim.x, b = blockIdx.x;
        const unsigned long long share = (((b + 1) << PB200_EXCHANGE_UNIT_LOG2) / blocks) -
                                         ((b << PB200_EXCHANGE_UNIT_LOG2) / blocks);
        for (int r = 0; r < sc.n; ++r) red_add_relaxed_sys(sc.flag[r], share);
    }
}

struct SolverDev {
    double rtol, atol, h0;
    double half_rtol;             // 0.5 * rtol, used every step
    int32_t max_steps;
    int32_t wide;                 // take pending observations four at a time first
};

// ---------------------------------------------------------------------------
// Shared-memory staging of the observed series with the bulk-copy engine
// (cp.async.bulk -> SASS UBLKCP), completion through an mbarrier.
// One elected thread issues the copy; everybody waits on the barrier phase.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Issue: barrier initialisation (one block-wide barrier) and the copies.
__device__ __forceinline__ void stage_series_begin(double *dst, const double *src,
                                                   uint32_t bytes, uint64_t *bar) {
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile(
            "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a),
            "r"(bytes)
            : "memory");
        const uint32_t chunk = 32768u;
        uint32_t done = 0;
        while (done < bytes) {
            uint32_t nb = bytes - done < chunk ? bytes - done : chunk;
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst) + done),
                "l"(reinterpret_cast<const char *>(src) + done), "r"(nb), "r"(bar_a)
                : "memory");
            done += nb;
        }
    }
}

// Wait for phase 0 of the barrier (every thread that reads the series).
__device__ __forceinline__ void stage_series_wait(uint64_t *bar) {
    const uint32_t bar_a = smem_u32(bar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_a)
            : "memory");
    }
}

__device__ __forceinline__ void stage_series(double *dst, const double *src,
                                             uint32_t bytes, uint64_t *bar) {
    stage_series_begin(dst, src, bytes, bar);
    stage_series_wait(bar);
}

// ---------------------------------------------------------------------------
// Objective epilogue.  Products and sums are rounded separately, in the order
// NumPy evaluates the reference expressions, so closed-form models agree with
// the reference to the last few ulps:
//   SSE    np.sum(np.sum((sim - data)**2, axis=0) * weights)
//          (pints/_error_measures.py:373-376)
//   KNOWN  np.sum(offset + multip * np.sum(error**2, axis=0))
//          (pints/_log_likelihoods.py:1039-1041)
//   GAUSS  np.sum(-logn - nt*log(sigma) - np.sum(error**2, axis=0)/(2 sigma**2))
//          (pints/_log_likelihoods.py:1123-1129)
//   MSE    (pints/_error_measures.py:111-113), RMSE / NRMSE (:159-161, :227-229)
// then LogPosterior adds the prior constant (pints/_log_pdfs.py:207-212) and
// ProbabilityBasedError flips the sign (pints/_error_measures.py:183-184).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double objective_finish(const ProblemDev &P,
                                                   const double *ss,
                                                   const double *x) {
    double acc = 0.0;
    const int no = P.n_outputs;
    if (P.objective == PB200_OBJ_SSE) {
        for (int o = 0; o < no; ++o) {
            double term = __dmul_rn(ss[o], P.oc[o]);
            acc = (o == 0) ? term : __dadd_rn(acc, term);
        }
    } else if (P.objective == PB200_OBJ_MSE) {
        //   ninv * np.sum(weights * np.sum((sim - data)**2, axis=0), axis=0)
        for (int o = 0; o < no; ++o) {
            double term = __dmul_rn(P.oc[o], ss[o]);
            acc = (o == 0) ? term : __dadd_rn(acc, term);
        }
        acc = __dmul_rn(P.oc[no], acc);
    } else if (P.objective == PB200_OBJ_RMSE) {
        //   norm * np.sqrt(ninv * np.sum((sim - data)**2))
        acc = __dmul_rn(P.oc[1], __dsqrt_rn(__dmul_rn(P.oc[0], ss[0])));
    } else if (P.objective == PB200_OBJ_GAUSSIAN_KNOWN_SIGMA) {
        for (int o = 0; o < no; ++o) {
            double term = __dadd_rn(P.oc[o], __dmul_rn(P.oc[no + o], ss[o]));
            acc = (o == 0) ? term : __dadd_rn(acc, term);
        }
    } else {
        const double logn = P.oc[0];
        const double nt = static_cast<double>(P.n_times);
        for (int o = 0; o < no; ++o) {
            const double sigma = x[P.n_model_params + o];
            double a = __dsub_rn(-logn, __dmul_rn(nt, log(sigma)));
            double b = __ddiv_rn(ss[o], __dmul_rn(2.0, __dmul_rn(sigma, sigma)));
            double term = __dsub_rn(a, b);
            acc = (o == 0) ? term : __dadd_rn(acc, term);
        }
    }
    if (P.n_prior > 0) acc = __dadd_rn(log_prior_of(P, x), acc);
    return P.sign * acc;
}

// Checks done before any simulation: the prior box (a point outside scores
// -inf and is never simulated) and sigma <= 0 for the unknown-sigma
// likelihood.  Returns true when the vector is decided without a solve.
__device__ __forceinline__ bool objective_precheck(const ProblemDev &P,
                                                   const double *x,
                                                   double *score) {
    if (P.n_prior > 0) {
        bool outside = false;
        for (int i = 0; i < P.n_prior; ++i)
            outside = outside || (x[i] < P.lo[i]) || (x[i] >= P.hi[i]);
        if (outside) {
            *score = P.sign * -INFINITY;
            return true;
        }
    }
    if (P.objective == PB200_OBJ_GAUSSIAN) {
        bool bad = false;
        for (int o = 0; o < P.n_outputs; ++o)
            bad = bad || (x[P.n_model_params + o] <= 0.0);
        if (bad) {
            double v = -INFINITY;
            if (P.n_prior > 0) v = log_prior_of(P, x) + v;
            *score = P.sign * v;
            return true;
        }
    }
    return false;
}

// host-side launchers implemented in the kernel translation units
int launch_ode(const pb200_problem *p, const double *d_params, int64_t n,
               int64_t ld, double *d_scores, double *d_traj, int32_t *d_steps,
               const SolverDev &s, int block, int lanes, void *sort_ws,
               const ScatterDev &sc, cudaStream_t st);
int64_t ode_sort_workspace_bytes(int64_t n);
int launch_analytic(const pb200_problem *p, const double *d_params, int64_t n,
                    int64_t ld, double *d_scores, double *d_traj,
                    const ScatterDev &sc, cudaStream_t st);
int launch_score_traj(const pb200_problem *p, const double *d_traj,
                      const double *d_params, int64_t n, int64_t ld,
                      double *d_scores, cudaStream_t st);
int pb200_chain_loop(const pb200_problem *p, int64_t n, int32_t d, int32_t require_finite,
                     double *cur, double *cur_f, double *prop, double *fx, uint8_t *accepted,
                     int64_t *acc_count, double *mu, double *sigma, double *log_lambda,
                     double *samples, int64_t sample_stride, double target, uint64_t seed,
                     const int64_t *rows, int64_t first_row, int64_t n_iter, cudaStream_t st);
int launch_exchange_signal(const ScatterDev &sc, cudaStream_t st);
int launch_exchange_finish(const void *stage, const void *flags, int n_ranks, int64_t per_rank,
                           int64_t rows_per_rank, unsigned long long epoch, int64_t n_total,
                           double *d_scores_all, cudaStream_t st);
void pb200_set_error(const char *fmt, ...);
int pb200_check_cuda(cudaError_t e, const char *what);
