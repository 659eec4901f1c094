// C-ABI entry points (include/pints_b200.h): argument checking, the host-side
// problem object, and dispatch to the kernel launchers.  No device memory is
// allocated here and nothing synchronises: work is enqueued on the caller's
// stream.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#include <pthread.h>
#include <sched.h>
#include "pb200_common.cuh"

static thread_local char g_error[512] = "";

void pb200_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int pb200_check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return PB200_OK;
    pb200_set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return PB200_ECUDA;
}

namespace {

struct ModelInfo {
    int n_params;      // forward-model parameters (-1: equals n_outputs)
    int n_outputs;     // -1: any (1..PB200_MAX_OUTPUTS)
    int min_consts;
    bool ode;
};

const ModelInfo kModels[PB200_N_MODELS] = {
    /* LOGISTIC       */ {2, 1, 1, false},
    /* HH_IK          */ {5, 1, 5, false},
    /* CONSTANT       */ {-1, -1, 0, false},
    /* FHN            */ {3, 2, 2, true},
    /* LOTKA_VOLTERRA */ {4, 2, 2, true},
    /* SIR            */ {3, 2, 4, true},
    /* GOODWIN        */ {5, 3, 3, true},
    /* HES1           */ {4, 1, 4, true},
    /* REPRESSILATOR  */ {4, 3, 6, true},
    /* BEELER_REUTER  */ {5, 2, 13, true},
};

// Default tolerance of a model.  The reference integrates everything with
// scipy odeint's rtol = atol = 1.49012e-8 (pints/toy/_toy_classes.py:225-233)
// and DP5(4) is run at the same value -- except Lotka-Volterra: its benchmark
// likelihood (sigma = 0.1, 1000 points) turns a trajectory error of 1e-7 into a
// log-likelihood error of 1.2e-6 at that tolerance (LSODA itself: 1.9e-5), and
// the error is proportional to the tolerance (sweep in
// profiles/r02_lv_tolerance.md), so its default is 5e-9: 3.8e-7, inside the
// 1e-6 bar with margin, for 23 % more steps.  Beeler-Reuter (the reference:
// LSODA at rtol 1e-4 / atol 1e-6, 1.7e-3 from the converged log-likelihood)
// runs on the Rosenbrock integrator at 1e-7: 8.5e-7 from the converged
// log-likelihood in 1 650 steps (profiles/r02x_br_rosenbrock.log).
double default_tolerance(int model) {
    if (model == PB200_MODEL_BEELER_REUTER) return 1e-7;
    return model == PB200_MODEL_LOTKA_VOLTERRA ? 5e-9 : 1.49012e-8;
}

SolverDev resolve_opts(const pb200_solver_opts *o, int model, int *block, int *lanes) {
    SolverDev s;
    s.rtol = (o && o->rtol > 0) ? o->rtol : default_tolerance(model);
    s.atol = (o && o->atol > 0) ? o->atol : default_tolerance(model);
    s.half_rtol = 0.5 * s.rtol;
    s.h0 = (o && o->h0 > 0) ? o->h0 : 0.0;
    s.max_steps = (o && o->max_steps > 0) ? o->max_steps : 1000000;
    s.wide = (o && o->wide_interpolation != 0) ? (o->wide_interpolation > 0 ? 1 : 0) : -1;
    *block = (o && o->block > 0) ? o->block : 0;
    *lanes = (o && o->lanes > 0) ? o->lanes : 0;
    return s;
}

int check_batch(const pb200_problem *p, const void *d_params, int64_t n, int64_t ld,
                const void *d_out, const char *who, bool model_only = false) {
    if (!p) { pb200_set_error("%s: problem is NULL", who); return PB200_EINVAL; }
    if (n < 0) { pb200_set_error("%s: n = %lld < 0", who, (long long)n); return PB200_EINVAL; }
    if (n == 0) return PB200_OK;
    if (!d_params || !d_out) { pb200_set_error("%s: NULL device pointer", who); return PB200_EINVAL; }
    if (!p->dev.series) {
        pb200_set_error("%s: the problem was created without a series buffer", who);
        return PB200_EINVAL;
    }
    const int need = model_only ? p->dev.n_model_params : p->dev.n_params;
    if (ld < need) {
        pb200_set_error("%s: row stride %lld is shorter than the %d parameters of this problem",
                        who, (long long)ld, need);
        return PB200_EINVAL;
    }
    return PB200_OK;
}

// ---------------------------------------------------------------------------
// Staging copies (pageable host array -> page-locked buffer) cut into slices
// and done by a few persistent threads of this library, so that their speed
// does not depend on the host program's OpenMP settings (torchrun starts every
// rank with OMP_NUM_THREADS=1).  Workers sleep on a condition variable after a
// short spin; the caller takes the first slice itself.
class CopyPool {
public:
    // never destroyed: the (detached) workers may outlive static destruction
    static CopyPool &instance() {
        static CopyPool *pool = [] {
            CopyPool *p = new CopyPool();
            pthread_atfork(nullptr, nullptr, [] { instance().forget_workers(); });
            return p;
        }();
        return *pool;
    }
    // in a forked child the worker threads do not exist: start over
    void forget_workers() {
        workers_ = 0;
        jobs_.clear();
        pending_.store(0);
    }
    void copy(char *dst, const char *src, size_t bytes, int threads) {
        int t = threads > 0 ? threads : default_threads_;
        if (t > kMaxThreads) t = kMaxThreads;
        const size_t min_slice = 128 * 1024;
        if (static_cast<size_t>(t) > bytes / min_slice) t = static_cast<int>(bytes / min_slice);
        if (t <= 1) { memcpy(dst, src, bytes); return; }
        std::lock_guard<std::mutex> serial(call_mutex_);       // one copy at a time
        grow(t - 1);
        const size_t slice = ((bytes + t - 1) / t + 63) & ~size_t(63);
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int w = 0; w < t - 1; ++w) {
                const size_t lo = static_cast<size_t>(w + 1) * slice;
                const size_t hi = lo + slice < bytes ? lo + slice : bytes;
                jobs_[w] = Job{dst + lo, src + lo, lo < bytes ? hi - lo : 0};
            }
            for (size_t w = t - 1; w < jobs_.size(); ++w) jobs_[w] = Job{nullptr, nullptr, 0};
            pending_.store(t - 1, std::memory_order_relaxed);
            generation_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        memcpy(dst, src, slice < bytes ? slice : bytes);
        while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
    }

private:
    static constexpr int kMaxThreads = 16;
    struct Job { char *dst; const char *src; size_t bytes; };
    CopyPool() {
        int cores = static_cast<int>(std::thread::hardware_concurrency());
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
        int share = 1;
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = atoi(e) > 0 ? atoi(e) : 1;
        int t = cores / share / 2;
        if (const char *e = getenv("PB200_COPY_THREADS")) t = atoi(e);
        default_threads_ = t < 1 ? 1 : (t > 8 ? 8 : t);
        jobs_.reserve(kMaxThreads);
    }
    void grow(int n) {
        std::lock_guard<std::mutex> lk(m_);
        while (workers_ < n) {
            const int id = workers_++;
            jobs_.push_back(Job{nullptr, nullptr, 0});
            const uint64_t seen = generation_.load(std::memory_order_relaxed);
            std::thread([this, id, seen] { run(id, seen); }).detach();
        }
    }
    void run(int id, uint64_t seen) {
        for (;;) {
            // spin for about a millisecond (an evaluate loop calls again within
            // that time), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            while (generation_.load(std::memory_order_acquire) == seen) {
                for (int k = 0; k < 64; ++k) __builtin_ia32_pause();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(1000)) break;
            }
            Job job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
                seen = generation_.load(std::memory_order_relaxed);
                job = jobs_[id];
                jobs_[id].dst = nullptr;
            }
            if (job.dst) {
                if (job.bytes) memcpy(job.dst, job.src, job.bytes);
                pending_.fetch_sub(1, std::memory_order_release);
            }
        }
    }
    std::mutex m_, call_mutex_;
    std::condition_variable cv_;
    int workers_ = 0;
    std::vector<Job> jobs_;
    std::atomic<uint64_t> generation_{0};
    std::atomic<int> pending_{0};
    int default_threads_ = 1;
};

}  // namespace

extern "C" {

int pb200_version(void) { return PB200_VERSION; }

const char *pb200_last_error(void) { return g_error; }

int64_t pb200_series_len(int32_t n_times, int32_t n_outputs) {
    if (n_times < 0 || n_outputs < 1) return -1;
    // n_times rows plus the sentinel rows (time = +inf), in whole 16-byte units
    const int64_t len = (static_cast<int64_t>(n_times) + PB200_SENTINEL_ROWS) * (1 + n_outputs);
    return (len + 1) & ~1ll;
}

int pb200_problem_create(int32_t model, int32_t objective, int32_t n_times,
                         int32_t n_outputs, const double *d_series,
                         const double *model_consts, int32_t n_model_consts,
                         const double *obj_consts, int32_t n_obj_consts, double sign,
                         const double *prior_lower, const double *prior_upper,
                         int32_t n_prior, double log_prior_value, pb200_problem **out) {
    if (!out) { pb200_set_error("pb200_problem_create: out is NULL"); return PB200_EINVAL; }
    *out = nullptr;
    if (model < 0 || model >= PB200_N_MODELS) {
        pb200_set_error("unknown model id %d", model);
        return PB200_EUNSUPPORTED;
    }
    if (objective < 0 || objective >= PB200_N_OBJECTIVES) {
        pb200_set_error("unknown objective id %d", objective);
        return PB200_EUNSUPPORTED;
    }
    const ModelInfo &mi = kModels[model];
    if (n_times < 0) { pb200_set_error("n_times = %d < 0", n_times); return PB200_EINVAL; }
    if (n_outputs < 1 || n_outputs > PB200_MAX_OUTPUTS ||
        (mi.n_outputs > 0 && n_outputs != mi.n_outputs)) {
        pb200_set_error("model %d does not have %d outputs", model, n_outputs);
        return PB200_EINVAL;
    }
    if (n_times > 0 && !d_series) { pb200_set_error("d_series is NULL"); return PB200_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(d_series) & 15) != 0) {
        pb200_set_error("d_series must be 16-byte aligned");
        return PB200_EINVAL;
    }
    if (n_model_consts < mi.min_consts || n_model_consts > PB200_MAX_MODEL_CONSTS ||
        (n_model_consts > 0 && !model_consts)) {
        pb200_set_error("model %d needs at least %d constants, got %d", model,
                        mi.min_consts, n_model_consts);
        return PB200_EINVAL;
    }
    if (model == PB200_MODEL_HH_IK) {
        const int n_ev = static_cast<int>(model_consts[4]);
        if (n_ev < 0 || n_ev > PB200_MAX_EVENTS || n_model_consts != 5 + 2 * n_ev) {
            pb200_set_error("HH-IK protocol: %d events needs %d constants, got %d", n_ev,
                            5 + 2 * n_ev, n_model_consts);
            return PB200_EINVAL;
        }
    }
    const int want_obj = objective == PB200_OBJ_SSE ? n_outputs
                         : objective == PB200_OBJ_GAUSSIAN_KNOWN_SIGMA ? 2 * n_outputs
                         : objective == PB200_OBJ_MSE ? n_outputs + 1
                         : objective == PB200_OBJ_RMSE ? 2 : 1;
    if (objective == PB200_OBJ_RMSE && n_outputs != 1) {
        pb200_set_error("the root-mean-squared errors are defined for single-output problems");
        return PB200_EINVAL;
    }
    if (n_obj_consts != want_obj || !obj_consts) {
        pb200_set_error("objective %d needs %d constants, got %d", objective, want_obj,
                        n_obj_consts);
        return PB200_EINVAL;
    }
    const int n_model_params = mi.n_params > 0 ? mi.n_params : n_outputs;
    const int n_params = n_model_params + (objective == PB200_OBJ_GAUSSIAN ? n_outputs : 0);
    if (n_params > PB200_MAX_PARAMS) {
        pb200_set_error("%d parameters exceed the supported %d", n_params, PB200_MAX_PARAMS);
        return PB200_EUNSUPPORTED;
    }
    if (n_prior != 0 && (n_prior != n_params || !prior_lower || !prior_upper)) {
        pb200_set_error("prior box must have %d entries, got %d", n_params, n_prior);
        return PB200_EINVAL;
    }
    if (!(sign == 1.0 || sign == -1.0)) {
        pb200_set_error("sign must be +1 or -1");
        return PB200_EINVAL;
    }
    pb200_problem *p = new (std::nothrow) pb200_problem();
    if (!p) { pb200_set_error("out of host memory"); return PB200_ENOMEM; }
    memset(p, 0, sizeof(*p));
    p->dev.model = model;
    p->dev.objective = objective;
    p->dev.n_times = n_times;
    p->dev.n_outputs = n_outputs;
    p->dev.n_model_params = n_model_params;
    p->dev.n_params = n_params;
    p->dev.n_prior = n_prior;
    p->dev.row = 1 + n_outputs;
    p->dev.series = d_series;
    p->dev.sign = sign;
    p->dev.log_prior = n_prior ? log_prior_value : 0.0;
    for (int i = 0; i < n_obj_consts; ++i) p->dev.oc[i] = obj_consts[i];
    for (int i = 0; i < n_prior; ++i) {
        p->dev.lo[i] = prior_lower[i];
        p->dev.hi[i] = prior_upper[i];
    }
    for (int i = 0; i < n_model_consts; ++i) p->mc.c[i] = model_consts[i];
    p->n_model_consts = n_model_consts;
    p->series_bytes = pb200_series_len(n_times, n_outputs) * 8;
    *out = p;
    return PB200_OK;
}

int pb200_problem_set_prior_terms(pb200_problem *p, int32_t n_terms, const int32_t *kinds,
                                  const int32_t *index, const double *consts) {
    if (!p) { pb200_set_error("pb200_problem_set_prior_terms: problem is NULL"); return PB200_EINVAL; }
    if (n_terms < 0 || n_terms > PB200_MAX_PARAMS || (n_terms > 0 && (!kinds || !index || !consts))) {
        pb200_set_error("pb200_problem_set_prior_terms: 0..%d terms", PB200_MAX_PARAMS);
        return PB200_EINVAL;
    }
    if (n_terms > 0 && p->dev.n_prior == 0) {
        pb200_set_error("pb200_problem_set_prior_terms: the problem was created without a prior box");
        return PB200_EINVAL;
    }
    for (int k = 0; k < n_terms; ++k) {
        if (kinds[k] < 0 || kinds[k] > 1 ||
            (kinds[k] == 1 && (index[k] < 0 || index[k] >= p->dev.n_params))) {
            pb200_set_error("pb200_problem_set_prior_terms: bad term %d", k);
            return PB200_EINVAL;
        }
    }
    p->dev.n_terms = n_terms;
    for (int k = 0; k < n_terms; ++k) {
        p->dev.term_kind[k] = kinds[k];
        p->dev.term_idx[k] = index[k];
        for (int c = 0; c < 3; ++c) p->dev.term_c[k][c] = consts[3 * k + c];
    }
    return PB200_OK;
}

void pb200_problem_destroy(pb200_problem *p) {
    if (p && p->sort_hint) cudaFreeHost(p->sort_hint);
    delete p;
}

int32_t pb200_problem_sort_hint(const pb200_problem *p) {
    return (p && p->sort_hint) ? *static_cast<volatile int32_t *>(p->sort_hint) : 0;
}

int32_t pb200_problem_n_parameters(const pb200_problem *p) {
    return p ? p->dev.n_params : -1;
}

int64_t pb200_sort_workspace_bytes(int64_t n) { return ode_sort_workspace_bytes(n); }

namespace {
int eval_impl(const pb200_problem *p, const double *d_params, int64_t n, int64_t ld,
              double *d_scores, int32_t *d_steps, const pb200_solver_opts *opts,
              void *d_sort_ws, int64_t sort_ws_bytes, const ScatterDev &sc, void *stream,
              const char *who) {
    int rc = check_batch(p, d_params, n, ld, sc.n ? static_cast<const void *>(sc.stage[0]) : d_scores,
                         who);
    if (rc || n == 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kModels[p->dev.model].ode) {
        int block, lanes;
        SolverDev s = resolve_opts(opts, p->dev.model, &block, &lanes);
        if (d_sort_ws) {
            if (sort_ws_bytes < ode_sort_workspace_bytes(n) ||
                (reinterpret_cast<uintptr_t>(d_sort_ws) & 15) != 0) {
                pb200_set_error("%s: workspace of %lld bytes (16-byte aligned) "
                                "needed, got %lld", who, (long long)ode_sort_workspace_bytes(n),
                                (long long)sort_ws_bytes);
                return PB200_EINVAL;
            }
            if (n > 2147483647ll) d_sort_ws = nullptr;       // int32 permutation
        }
        return launch_ode(p, d_params, n, ld, d_scores, nullptr, d_steps, s, block, lanes,
                          d_sort_ws, sc, st);
    }
    if (d_steps) {
        rc = pb200_check_cuda(cudaMemsetAsync(d_steps, 0, n * sizeof(int32_t), st),
                              "cudaMemsetAsync(d_steps)");
        if (rc) return rc;
    }
    return launch_analytic(p, d_params, n, ld, d_scores, nullptr, sc, st);
}
const ScatterDev kNoScatter = {};

// checks an exchange descriptor and turns it into the kernels' view of it
int exchange_scatter(const pb200_exchange_desc *ex, ScatterDev *sc, const char *who) {
    if (!ex) { pb200_set_error("%s: exchange descriptor is NULL", who); return PB200_EINVAL; }
    if (ex->n_ranks < 1 || ex->n_ranks > PB200_MAX_PEERS || ex->rank < 0 ||
        ex->rank >= ex->n_ranks || ex->per_rank < 1 || ex->epoch < 1) {
        pb200_set_error("%s: needs 1..%d ranks, a rank among them, per_rank >= 1 and "
                        "epoch >= 1", who, PB200_MAX_PEERS);
        return PB200_EINVAL;
    }
    *sc = ScatterDev{};
    for (int q = 0; q < ex->n_ranks; ++q) {
        if (!ex->stage[q] || !ex->flags[q] ||
            (reinterpret_cast<uintptr_t>(ex->stage[q]) & 15) != 0 ||
            (reinterpret_cast<uintptr_t>(ex->flags[q]) & 7) != 0) {
            pb200_set_error("%s: staging / flag array of rank %d is NULL or misaligned", who, q);
            return PB200_EINVAL;
        }
        sc->stage[q] = ex->stage[q];
        sc->flag[q] = static_cast<unsigned long long *>(ex->flags[q]) + ex->rank;
    }
    if ((reinterpret_cast<uintptr_t>(ex->stage_multicast) & 15) != 0) {
        pb200_set_error("%s: multicast address is misaligned", who);
        return PB200_EINVAL;
    }
    sc->stage_mc = ex->stage_multicast;
    sc->n = ex->n_ranks;
    sc->offset = static_cast<int64_t>(ex->rank) * ex->per_rank;
    return PB200_OK;
}
}  // namespace

int pb200_eval_sorted(const pb200_problem *p, const double *d_params, int64_t n, int64_t ld,
                      double *d_scores, int32_t *d_steps, const pb200_solver_opts *opts,
                      void *d_sort_ws, int64_t sort_ws_bytes, void *stream) {
    return eval_impl(p, d_params, n, ld, d_scores, d_steps, opts, d_sort_ws, sort_ws_bytes,
                     kNoScatter, stream, "pb200_eval");
}

int pb200_eval_exchange(const pb200_problem *p, const double *d_params, int64_t n, int64_t ld,
                        int32_t *d_steps, const pb200_solver_opts *opts, void *d_sort_ws,
                        int64_t sort_ws_bytes, const pb200_exchange_desc *ex, void *stream) {
    ScatterDev sc;
    int rc = exchange_scatter(ex, &sc, "pb200_eval_exchange");
    if (rc) return rc;
    if (n < 0 || n > ex->per_rank) {
        pb200_set_error("pb200_eval_exchange: n = %lld rows do not fit a block of %lld",
                        (long long)n, (long long)ex->per_rank);
        return PB200_EINVAL;
    }
    if (n == 0) return launch_exchange_signal(sc, static_cast<cudaStream_t>(stream));
    return eval_impl(p, d_params, n, ld, nullptr, d_steps, opts, d_sort_ws, sort_ws_bytes, sc,
                     stream, "pb200_eval_exchange");
}

int pb200_exchange_finish(const pb200_exchange_desc *ex, int64_t rows_per_rank, int64_t n_total,
                          double *d_scores_all, void *stream) {
    ScatterDev sc;
    int rc = exchange_scatter(ex, &sc, "pb200_exchange_finish");
    if (rc) return rc;
    if (rows_per_rank < 1 || rows_per_rank > ex->per_rank || n_total < 0 ||
        n_total > rows_per_rank * ex->n_ranks || (n_total > 0 && !d_scores_all)) {
        pb200_set_error("pb200_exchange_finish: %lld scores do not fit %d blocks of %lld rows "
                        "(staged in blocks of %lld)", (long long)n_total, ex->n_ranks,
                        (long long)rows_per_rank, (long long)ex->per_rank);
        return PB200_EINVAL;
    }
    return launch_exchange_finish(ex->stage[ex->rank], ex->flags[ex->rank], ex->n_ranks,
                                  ex->per_rank, rows_per_rank, ex->epoch, n_total, d_scores_all,
                                  static_cast<cudaStream_t>(stream));
}

int pb200_eval(const pb200_problem *p, const double *d_params, int64_t n, int64_t ld,
               double *d_scores, int32_t *d_steps, const pb200_solver_opts *opts,
               void *stream) {
    return pb200_eval_sorted(p, d_params, n, ld, d_scores, d_steps, opts, nullptr, 0, stream);
}

int pb200_eval_host(const pb200_problem *p, const double *h_params, int64_t n, int64_t ld,
                    double *h_scores, int32_t *h_steps, double *d_params_ws,
                    double *d_scores_ws, int32_t *d_steps_ws, void *d_sort_ws,
                    int64_t sort_ws_bytes, const pb200_solver_opts *opts, void *stream) {
    int rc = check_batch(p, h_params, n, ld, h_scores, "pb200_eval_host");
    if (rc || n == 0) return rc;
    if (!d_params_ws || !d_scores_ws || ((h_steps == nullptr) != (d_steps_ws == nullptr))) {
        pb200_set_error("pb200_eval_host: missing device workspace");
        return PB200_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = pb200_check_cuda(cudaMemcpyAsync(d_params_ws, h_params, static_cast<size_t>(n) * ld * 8,
                                          cudaMemcpyHostToDevice, st), "copy of parameters");
    if (rc) return rc;
    rc = pb200_eval_sorted(p, d_params_ws, n, ld, d_scores_ws, d_steps_ws, opts, d_sort_ws,
                           sort_ws_bytes, stream);
    if (rc) return rc;
    rc = pb200_check_cuda(cudaMemcpyAsync(h_scores, d_scores_ws, static_cast<size_t>(n) * 8,
                                          cudaMemcpyDeviceToHost, st), "copy of scores");
    if (rc) return rc;
    if (h_steps)
        rc = pb200_check_cuda(cudaMemcpyAsync(h_steps, d_steps_ws, static_cast<size_t>(n) * 4,
                                              cudaMemcpyDeviceToHost, st), "copy of steps");
    return rc;
}

int pb200_mcmc_chain_run(const pb200_problem *p, int32_t require_finite, double *d_current,
                         double *d_current_f, double *d_proposed, double *d_fx,
                         uint8_t *d_accepted, int64_t *d_acceptance_count, double *d_mu,
                         double *d_sigma, double *d_log_lambda, double *d_samples,
                         int64_t sample_stride, int64_t n, int32_t d, double target,
                         uint64_t seed, const int64_t *d_rows, int64_t first_row,
                         int64_t n_iterations, void *stream) {
    if (!p) { pb200_set_error("pb200_mcmc_chain_run: problem is NULL"); return PB200_EINVAL; }
    if (kModels[p->dev.model].ode) {
        pb200_set_error("pb200_mcmc_chain_run: closed-form models only (an ODE solve per "
                        "iteration is long enough for separate launches)");
        return PB200_EUNSUPPORTED;
    }
    if (n < 0 || n_iterations < 0 || first_row < 0 || d != p->dev.n_params) {
        pb200_set_error("pb200_mcmc_chain_run: bad n = %lld, iterations = %lld, first_row = %lld "
                        "or d = %d (the problem has %d parameters)", (long long)n,
                        (long long)n_iterations, (long long)first_row, d, p->dev.n_params);
        return PB200_EINVAL;
    }
    if (n == 0 || n_iterations == 0) return PB200_OK;
    if (!d_current || !d_current_f || !d_proposed || !d_fx || !d_accepted || !d_mu || !d_sigma ||
        !d_log_lambda || !d_rows || !p->dev.series) {
        pb200_set_error("pb200_mcmc_chain_run: NULL buffer");
        return PB200_EINVAL;
    }
    return pb200_chain_loop(p, n, d, require_finite, d_current, d_current_f, d_proposed, d_fx,
                            d_accepted, d_acceptance_count, d_mu, d_sigma, d_log_lambda,
                            d_samples, sample_stride, target, seed, d_rows, first_row,
                            n_iterations, static_cast<cudaStream_t>(stream));
}

int pb200_host_is_pinned(const void *ptr) {
    if (!ptr) return 0;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();                      // plain pageable memory on old drivers
        return 0;
    }
    return a.type == cudaMemoryTypeHost ? 1 : 0;
}

int pb200_copy_async(void *dst, const void *src, int64_t bytes, int32_t to_device,
                     void *stream) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) {
        pb200_set_error("pb200_copy_async: bad arguments");
        return PB200_EINVAL;
    }
    if (bytes == 0) return PB200_OK;
    return pb200_check_cuda(
        cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes),
                        to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                        static_cast<cudaStream_t>(stream)),
        "pb200_copy_async");
}

int pb200_host_copy(void *dst, const void *src, int64_t bytes, int32_t threads) {
    if (bytes < 0 || (bytes > 0 && (!dst || !src))) {
        pb200_set_error("pb200_host_copy: bad arguments");
        return PB200_EINVAL;
    }
    if (bytes == 0) return PB200_OK;
    CopyPool::instance().copy(static_cast<char *>(dst), static_cast<const char *>(src),
                              static_cast<size_t>(bytes), threads);
    return PB200_OK;
}

int pb200_simulate(const pb200_problem *p, const double *d_params, int64_t n, int64_t ld,
                   double *d_traj, int32_t *d_steps, const pb200_solver_opts *opts,
                   void *stream) {
    int rc = check_batch(p, d_params, n, ld, d_traj, "pb200_simulate", true);
    if (rc || n == 0) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kModels[p->dev.model].ode) {
        int block, lanes;
        SolverDev s = resolve_opts(opts, p->dev.model, &block, &lanes);
        return launch_ode(p, d_params, n, ld, nullptr, d_traj, d_steps, s, block, lanes, nullptr,
                          ScatterDev{}, st);
    }
    if (d_steps) {
        rc = pb200_check_cuda(cudaMemsetAsync(d_steps, 0, n * sizeof(int32_t), st),
                              "cudaMemsetAsync(d_steps)");
        if (rc) return rc;
    }
    return launch_analytic(p, d_params, n, ld, nullptr, d_traj, ScatterDev{}, st);
}

int pb200_score_trajectories(const pb200_problem *p, const double *d_traj,
                             const double *d_params, int64_t n, int64_t ld,
                             double *d_scores, void *stream) {
    int rc = check_batch(p, d_params, n, ld, d_scores, "pb200_score_trajectories");
    if (rc || n == 0) return rc;
    if (!d_traj) { pb200_set_error("pb200_score_trajectories: d_traj is NULL"); return PB200_EINVAL; }
    return launch_score_traj(p, d_traj, d_params, n, ld, d_scores,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
