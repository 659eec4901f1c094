/*
 * pints_b200.h -- C ABI of the B200-native batched evaluation engine for PINTS.
 *
 * This is the drop-in boundary of the one hot path this repository
 * accelerates: scoring N parameter vectors as
 *
 *     model.simulate(x, times) -> residual -> sum of squares -> objective
 *
 * in one fused launch, plus the per-iteration state update of two MCMC
 * samplers.  The reference (pints-team/pints v0.6.2) is pure Python and has no
 * FFI of its own; every entry point below names the reference function
 * (file:line, relative to the reference root) whose work it replaces.  The
 * binding a maintainer would add is the ctypes shim in
 * pints_b200/_lib.py (shown in INTEGRATION.md).
 *
 * Conventions
 *  - plain C: pointers, sizes, scalars.  No C++ or torch types.
 *  - every function returns 0 on success and a negative PB200_E* code on
 *    failure; pb200_last_error() returns a thread-local message.  Nothing is
 *    thrown across the boundary and nothing aborts.
 *  - pointers named d_* are DEVICE pointers owned by the caller (in this
 *    repository: torch tensors).  The library never allocates or frees device
 *    memory; a pb200_problem is a small host object holding constants and the
 *    caller's d_series pointer, which must stay valid while it is used.
 *  - `stream` is a cudaStream_t passed as void* (0 = default stream).  All work
 *    is enqueued asynchronously on it; the caller synchronises.
 *  - like the reference evaluators (pints/_evaluation.py:156-158) one problem
 *    handle must not be used from several host threads at once; distinct
 *    handles and streams are independent.
 *  - there is no CPU fallback anywhere behind this interface.
 */
#ifndef PINTS_B200_H
#define PINTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PB200_VERSION 100           /* 0.1.0 */

/* error codes */
#define PB200_OK            0
#define PB200_EINVAL       -1       /* bad argument                        */
#define PB200_EUNSUPPORTED -2       /* model/objective/size not supported   */
#define PB200_ECUDA        -3       /* CUDA runtime error (see last_error)  */
#define PB200_ENOMEM       -4       /* host allocation failed               */

/* forward models (pints/toy/...) */
#define PB200_MODEL_LOGISTIC        0   /* _logistic_model.py:62-86            */
#define PB200_MODEL_HH_IK           1   /* _hh_ik_model.py:135-214             */
#define PB200_MODEL_CONSTANT        2   /* _constant_model.py:75-93            */
#define PB200_MODEL_FHN             3   /* _fitzhugh_nagumo_model.py:111-117   */
#define PB200_MODEL_LOTKA_VOLTERRA  4   /* _lotka_volterra_model.py:91-97      */
#define PB200_MODEL_SIR             5   /* _sir_model.py:96-111                */
#define PB200_MODEL_GOODWIN         6   /* _goodwin_oscillator_model.py:101-107 */
#define PB200_MODEL_HES1            7   /* _hes1_michaelis_menten.py:129-140   */
#define PB200_MODEL_REPRESSILATOR   8   /* _repressilator_model.py:91-108      */
#define PB200_MODEL_BEELER_REUTER   9   /* _beeler_reuter_model.py:105-183     */
#define PB200_N_MODELS              10

/* objectives */
#define PB200_OBJ_SSE                   0  /* pints/_error_measures.py:373-376     */
#define PB200_OBJ_GAUSSIAN_KNOWN_SIGMA  1  /* pints/_log_likelihoods.py:1012-1041  */
#define PB200_OBJ_GAUSSIAN              2  /* pints/_log_likelihoods.py:1110-1129  */
#define PB200_OBJ_MSE                   3  /* pints/_error_measures.py:74-114      */
#define PB200_OBJ_RMSE                  4  /* pints/_error_measures.py:135-161 (normalised)
                                              and :201-229 (plain, norm = 1)       */
#define PB200_N_OBJECTIVES              5

#define PB200_MAX_OUTPUTS   8
#define PB200_MAX_PARAMS    16      /* model parameters + sigma parameters  */
#define PB200_MAX_EVENTS    64      /* HH-IK protocol events                */
#define PB200_MAX_PEERS     16      /* GPUs of one job (pb200_eval_scatter) */
#define PB200_SENTINEL_ROWS 8       /* +inf rows that end the packed series */

typedef struct pb200_problem pb200_problem;

/* Adaptive Dormand-Prince 5(4) controls for the ODE models.  The reference
 * integrates with scipy's LSODA at rtol = atol = 1.49012e-8
 * (pints/toy/_toy_classes.py:225-233); the defaults here match that, except
 * for Lotka-Volterra (5e-9), whose benchmark log-likelihood needs it to stay
 * within 1e-6 of a converged solve (pb200_api.cu: default_tolerance). */
typedef struct pb200_solver_opts {
    double  rtol;        /* <= 0 -> the model's default (1.49012e-8; LV 5e-9) */
    double  atol;        /* <= 0 -> the same                                  */
    double  h0;          /* first step; <= 0 -> automatic                    */
    int32_t max_steps;   /* attempted steps per vector; <= 0 -> 1000000.     */
                         /* A vector that exceeds it scores NaN.             */
    int32_t block;       /* threads per block; <= 0 -> automatic (which also */
                         /* allows the one-block-per-SM migrating layout)    */
    int32_t lanes;       /* GPU lanes per parameter vector: 1, or 2 (one per */
                         /* state, two-state models only); <= 0 -> automatic */
    int32_t wide_interpolation;  /* observations that pile up inside long     */
                         /* steps are taken four at a time: > 0 on, < 0 off, */
                         /* 0 -> automatic (on for the smooth models, off    */
                         /* for FitzHugh-Nagumo's relaxation oscillations)   */
} pb200_solver_opts;

int         pb200_version(void);
const char *pb200_last_error(void);

/* Number of doubles the packed series buffer must hold (see below). */
int64_t pb200_series_len(int32_t n_times, int32_t n_outputs);

/*
 * Describe one inference problem: a forward model, the observed series and an
 * objective.  Replaces the object graph
 *   ErrorMeasure/LogPDF -> Single/MultiOutputProblem -> ForwardModel
 * (pints/_core.py:102-327, pints/_log_pdfs.py:134-241).
 *
 * d_series   device buffer of pb200_series_len() doubles, 16-byte aligned:
 *            n_times rows of [t_j, v_j0 .. v_j(n_outputs-1)], then
 *            PB200_SENTINEL_ROWS rows whose time is +infinity (values 0), zero
 *            padded: the kernels read a few rows ahead without bounds checks.
 *            Times must be non-negative and non-decreasing
 *            (pints/_core.py:129-133, :238-242); the caller checks that.
 * model_consts (host)
 *   LOGISTIC        [p0]
 *   HH_IK           [n0, g_max, E_k, v_hold, n_events, events[n_events],
 *                    voltages[n_events]]   voltages[i] applies after events[i]
 *   CONSTANT        []           (n_outputs parameters, output o = p_o*(o+1))
 *   FHN             [y0_V, y0_R]
 *   LOTKA_VOLTERRA  [y0_x, y0_y]
 *   SIR             [y0_S (unused), y0_I, y0_R, truncate_S0 (0|1)]
 *   GOODWIN         [y0_x, y0_y, y0_z]
 *   HES1            [m0, p1_0, p2_0, k_deg]          (output: m)
 *   REPRESSILATOR   [y0 (6 values)]                  (outputs: m1, m2, m3)
 *   BEELER_REUTER   [V0, Ca0, m0, h0, j0, d0, f0, x10, C_m, E_Na, I_amp,
 *                    I_period, I_length]             (outputs: V, Ca_i;
 *                    parameters: log of gNa, gNaC, gCa, gK1, gx1)
 * obj_consts (host)
 *   SSE                   weights[n_outputs]
 *   GAUSSIAN_KNOWN_SIGMA  offset[n_outputs], multip[n_outputs]
 *   GAUSSIAN              [0.5 * n_times * log(2 pi)]
 *   MSE                   weights[n_outputs], 1 / (n_times * n_outputs)
 *   RMSE                  [1 / n_times, norm]   single output; norm = 1 for
 *                         RootMeanSquaredError, 1 / sqrt(mean(values^2)) for
 *                         NormalisedRootMeanSquaredError
 * sign       +1, or -1 for ProbabilityBasedError
 *            (pints/_error_measures.py:183-184)
 * prior_lower/upper (host, n_prior entries or NULL/0): box of a
 *            UniformLogPrior combined through LogPosterior
 *            (pints/_log_pdfs.py:207-212, pints/_log_priors.py:1319-1320);
 *            log_prior_value is the prior's constant inside the box.
 */
int pb200_problem_create(int32_t model, int32_t objective,
                         int32_t n_times, int32_t n_outputs,
                         const double *d_series,
                         const double *model_consts, int32_t n_model_consts,
                         const double *obj_consts, int32_t n_obj_consts,
                         double sign,
                         const double *prior_lower, const double *prior_upper,
                         int32_t n_prior, double log_prior_value,
                         pb200_problem **out);
void pb200_problem_destroy(pb200_problem *p);

/*
 * ComposedLogPrior (pints/_log_priors.py:170-219) inside LogPosterior: the box
 * given to pb200_problem_create holds the supports of all sub-priors (+-inf
 * where a sub-prior is unbounded) and decides "-inf without simulating"; the
 * value inside is the sum of n_terms terms added in the sub-priors' order, as
 * the reference adds them:
 *   kind 0  constant consts[3k]: a UniformLogPrior's -log(prod(range))
 *           (pints/_log_priors.py:1315-1320)
 *   kind 1  GaussianLogPrior on parameter index[k]:
 *           consts[3k] - consts[3k+1] * (x - consts[3k+2])^2  with
 *           offset = log(1/sqrt(2 pi sd^2)), factor = 1/(2 sd^2), mean
 *           (pints/_log_priors.py:463-468)
 * n_terms = 0 restores the single constant of pb200_problem_create.
 */
int pb200_problem_set_prior_terms(pb200_problem *p, int32_t n_terms,
                                  const int32_t *kinds, const int32_t *index,
                                  const double *consts);

/* Row length of a parameter vector for this problem
 * (model parameters + n_outputs sigmas for GAUSSIAN). */
int32_t pb200_problem_n_parameters(const pb200_problem *p);

/*
 * Score n parameter vectors: the fused replacement of
 *   SequentialEvaluator._evaluate / ParallelEvaluator._evaluate
 *   (pints/_evaluation.py:478-482, :284-372)
 * and everything they call per vector.  Trajectories never leave the SM.
 *
 * d_params  n rows of `ld` doubles (ld >= n_parameters), row major
 * d_scores  n doubles
 * d_steps   n int32 or NULL: attempted RK steps per vector (ODE models),
 *           0 for closed-form models; used for flop accounting
 */
int pb200_eval(const pb200_problem *p, const double *d_params, int64_t n,
               int64_t ld, double *d_scores, int32_t *d_steps,
               const pb200_solver_opts *opts, void *stream);

/*
 * The same with cost-coherent scheduling (ODE models; ignored for closed-form
 * models): a counting sort on |f(y0)|^2 groups vectors with similar dynamics
 * into the same warps before the fused launch, so that lock-stepped lanes take
 * similar step sequences.  Scores and step counts are written at the original
 * row positions and do not depend on the grouping.  d_sort_ws is a 16-byte
 * aligned device workspace of pb200_sort_workspace_bytes(n) bytes owned by the
 * caller (NULL = no sorting, i.e. pb200_eval).
 *
 * Launches with one block per SM (the single-wave populations of the
 * migrating and dense layouts) can also sort inside the evaluation kernel,
 * block by block in shared memory, which saves the sort kernel's launch; that
 * pays when the costs of a population differ by a few per cent only, so every
 * such launch leaves what it measured (the spread of its vectors' attempted
 * steps, weighted with the kernel's duration) in a word of the problem and the
 * next launch chooses by it.  pb200_problem_sort_hint reads that word:
 * 0 = nothing seen yet, 1 = narrow (next one-block-per-SM launch sorts in the
 * kernel), 2 = wide (sort kernel).  Environment PB200_LOCAL_SORT=0 / 1 forces
 * either.  Scores never depend on the choice.
 */
int64_t pb200_sort_workspace_bytes(int64_t n);
int32_t pb200_problem_sort_hint(const pb200_problem *p);
int pb200_eval_sorted(const pb200_problem *p, const double *d_params, int64_t n,
                      int64_t ld, double *d_scores, int32_t *d_steps,
                      const pb200_solver_opts *opts, void *d_sort_ws,
                      int64_t sort_ws_bytes, void *stream);

/*
 * Multi-GPU variant with the exchange step fused in (the
 * one collective of the path is the all-gather of the scores, which replaces
 * the result queue of ParallelEvaluator, pints/_evaluation.py:316-337).
 *
 * Every rank owns, in memory that all ranks of the job have mapped (e.g. torch
 * symmetric memory over NVLink / NVSwitch):
 *   a staging array of n_ranks * per_rank pairs of 16 bytes {double score,
 *   int64 row}: rank r's block starts at pair r * per_rank;
 *   a flag array of n_ranks uint64, zero before the first generation.
 * pb200_eval_exchange scores this rank's n (<= per_rank) rows and the kernel
 * stores each score, with its row number, into the staging array of EVERY rank
 * in launch order -- after the cost sort a warp's 32 vectors are 32 consecutive
 * pairs, i.e. one contiguous 512-byte run per destination, not 32 scattered
 * 8-byte stores -- or, when stage_multicast is given (a multicast mapping of
 * the staging arrays: NVSwitch replicates the store), with one multimem.st per
 * vector.  Every block of the launch, once its scores are stored, adds its
 * share of 2^20 to this rank's counter in every rank's flag array (one
 * system-scope fence, then relaxed remote atomics; the shares of one launch
 * sum to exactly 2^20 whatever the grid); with n = 0 a one-thread kernel adds
 * the whole 2^20.  There is no barrier launch.
 *
 * pb200_exchange_finish (the reading side, same stream, after the call above)
 * waits until all n_ranks local counters have reached epoch * 2^20 (acquire;
 * it traps after 30 s without progress rather than hang the GPU) and writes
 * the n_total scores of the whole population in row order to d_scores_all,
 * where rank r scored the rows [r * rows_per_rank, min(n_total, (r + 1) *
 * rows_per_rank)) (rows_per_rank <= per_rank: the blocks of this generation
 * may be shorter than the staging blocks).
 *
 * Use epoch = 1, 2, 3, ... -- every rank takes part in every generation -- and
 * alternate between TWO staging arrays (one flag array serves both): a rank
 * can then run at most one generation ahead of the slowest reader, which is
 * still on the other array.  d_sort_ws as in pb200_eval_sorted (NULL = no
 * sorting).
 */
typedef struct pb200_exchange_desc {
    int32_t  n_ranks;                    /* 1 .. PB200_MAX_PEERS                */
    int32_t  rank;                       /* this rank                           */
    int64_t  per_rank;                   /* rows per rank block                 */
    void    *stage[PB200_MAX_PEERS];     /* rank q's staging array, as mapped in
                                            this process (q = rank: its own)    */
    void    *flags[PB200_MAX_PEERS];     /* rank q's flag array, as mapped here */
    void    *stage_multicast;            /* multicast address of the staging
                                            arrays, or NULL                     */
    uint64_t epoch;                      /* generation number, >= 1, increasing */
} pb200_exchange_desc;

int pb200_eval_exchange(const pb200_problem *p, const double *d_params, int64_t n,
                        int64_t ld, int32_t *d_steps,
                        const pb200_solver_opts *opts, void *d_sort_ws,
                        int64_t sort_ws_bytes, const pb200_exchange_desc *ex,
                        void *stream);
int pb200_exchange_finish(const pb200_exchange_desc *ex, int64_t rows_per_rank,
                          int64_t n_total, double *d_scores_all, void *stream);

/*
 * The same through HOST buffers, for callers that hold their positions in host
 * memory (which is what the reference's evaluators receive): enqueues the
 * host->device copy of the parameters, the fused kernel and the device->host
 * copy of the scores on `stream`, then returns; the caller synchronises the
 * stream before reading h_scores.  h_params / h_scores should be page-locked
 * (pinned) for the copies to be asynchronous.  d_params_ws (n*ld doubles) and
 * d_scores_ws (n doubles) are device workspaces owned by the caller; h_steps /
 * d_steps_ws are optional (both NULL or both given); d_sort_ws as in
 * pb200_eval_sorted (NULL = no sorting).
 */
int pb200_eval_host(const pb200_problem *p, const double *h_params, int64_t n,
                    int64_t ld, double *h_scores, int32_t *h_steps,
                    double *d_params_ws, double *d_scores_ws,
                    int32_t *d_steps_ws, void *d_sort_ws, int64_t sort_ws_bytes,
                    const pb200_solver_opts *opts, void *stream);

/*
 * Building block of the pipelined host path: one asynchronous copy on `stream`
 * between a page-locked host buffer and a device buffer (to_device = 1: host to
 * device, 0: device to host).  CudaEvaluator stages a batch into pinned memory
 * in slices and sends each slice while it stages the next, then launches ONE
 * fused kernel over the whole batch.
 */
int pb200_copy_async(void *dst, const void *src, int64_t bytes,
                     int32_t to_device, void *stream);
/* Host-side staging copy (an ordinary, pageable array of positions into a
 * page-locked buffer that pb200_copy_async can send): cut into slices and done
 * by `threads` persistent threads of this library (<= 0: automatic, from the
 * CPUs this process may run on divided by LOCAL_WORLD_SIZE; PB200_COPY_THREADS
 * overrides), independent of the host program's OpenMP settings.  Synchronous. */
int pb200_host_copy(void *dst, const void *src, int64_t bytes, int32_t threads);
/* 1 if `ptr` points into page-locked host memory (then a batch can be sent
 * from where it lies, without staging), else 0. */
int pb200_host_is_pinned(const void *ptr);

/*
 * Batched ForwardModel.simulate / Problem.evaluate
 * (pints/_core.py:147-153, :257-265): writes the n trajectories
 * d_traj[n][n_times][n_outputs].  Also the first half of the unfused baseline.
 */
int pb200_simulate(const pb200_problem *p, const double *d_params, int64_t n,
                   int64_t ld, double *d_traj, int32_t *d_steps,
                   const pb200_solver_opts *opts, void *stream);

/* Second half of the unfused baseline: objective from stored trajectories. */
int pb200_score_trajectories(const pb200_problem *p, const double *d_traj,
                             const double *d_params, int64_t n, int64_t ld,
                             double *d_scores, void *stream);

/*
 * Metropolis accept/reject for n chains (bit-exact with the reference given
 * the same inputs):  accept_j = isfinite-or-not rule AND log_u_j < fx_j - cur_f_j
 *   require_finite = 1: AdaptiveCovarianceMC.tell
 *                       (pints/_mcmc/_adaptive_covariance.py:264-278)
 *   require_finite = 0: DifferentialEvolutionMCMC.tell
 *                       (pints/_mcmc/_differential_evolution.py:314-326)
 * Updates d_current[n][d] / d_current_f[n] in place from d_proposed / d_fx and
 * writes d_accepted[n] (uint8).
 */
int pb200_mh_accept(double *d_current, double *d_current_f,
                    const double *d_proposed, const double *d_fx,
                    const double *d_log_u, uint8_t *d_accepted,
                    int64_t n, int32_t d, int32_t require_finite,
                    void *stream);

/* The same, also writing log_ratio_j = fx_j - cur_f_j (taken BEFORE the update;
 * the adaptive samplers below use it whatever the decision was). */
int pb200_mh_accept_ratio(double *d_current, double *d_current_f,
                          const double *d_proposed, const double *d_fx,
                          const double *d_log_u, uint8_t *d_accepted,
                          double *d_log_ratio, int64_t n, int32_t d,
                          int32_t require_finite, void *stream);

/*
 * Adaptation step of the adaptive-covariance family
 * (pints/_mcmc/_adaptive_covariance.py:88-107, :286-300):
 *   HAARIO_BARDENET  pints/_mcmc/_haario_bardenet_ac.py:65-68 (= pb200_hbac_adapt)
 *   HAARIO           pints/_mcmc/_haario_ac.py:63-66: log lambda moves with the
 *                    acceptance PROBABILITY min(1, exp(log_ratio))
 *   RAO_BLACKWELL    pints/_mcmc/_rao_blackwell_ac.py:58-74: sigma is updated with
 *                    p * proposed + (1 - p) * previous instead of the new point;
 *                    needs d_previous (the points before the accept step) and
 *                    d_proposed; d_log_lambda is not touched
 * Products and sums are rounded one by one as NumPy does; exp() may differ from
 * NumPy's in the last place, so HAARIO / RAO_BLACKWELL states agree with the
 * reference to ~1e-15 relative (decisions are taken elsewhere and are exact).
 */
#define PB200_AC_HAARIO_BARDENET 0
#define PB200_AC_HAARIO          1
#define PB200_AC_RAO_BLACKWELL   2
int pb200_ac_adapt(int32_t variant, const double *d_current,
                   const double *d_previous, const double *d_proposed,
                   const uint8_t *d_accepted, const double *d_log_ratio,
                   double *d_mu, double *d_sigma, double *d_log_lambda,
                   int64_t n, int32_t d, double gamma, double target,
                   void *stream);

/*
 * Haario-Bardenet adaptation after the accept step
 * (pints/_mcmc/_adaptive_covariance.py:286-300, :88-107;
 *  pints/_mcmc/_haario_bardenet_ac.py:65-68), separately rounded products and
 * sums exactly as NumPy evaluates them.  gamma = (adaptations + 1) ** -eta is
 * chain independent and computed by the caller.
 */
int pb200_hbac_adapt(const double *d_current, const uint8_t *d_accepted,
                     double *d_mu, double *d_sigma, double *d_log_lambda,
                     int64_t n, int32_t d, double gamma, double target,
                     void *stream);

/*
 * Device-side proposals for the Haario-Bardenet sampler:
 *   x* = x + L z,  L L^T = sigma * exp(log_lambda)   (Cholesky, in-thread)
 * The reference factors with LAPACK's SVD inside numpy.random.multivariate_normal
 * (pints/_mcmc/_haario_bardenet_ac.py:70-73); Cholesky gives a different but
 * equally distributed proposal, so this entry point is statistical parity
 * only.  d_z holds n*d standard normals (see pb200_philox_normal).
 */
int pb200_hbac_propose(const double *d_current, const double *d_sigma,
                       const double *d_log_lambda, const double *d_z,
                       double *d_proposed, int64_t n, int32_t d,
                       void *stream);

/*
 * Differential-evolution proposals
 * (pints/_mcmc/_differential_evolution.py:103-115):
 *   x*_j = (x_j + gamma * (x_r1 - x_r2)) + eps_j      left to right, no FMA
 * d_current_all holds the positions of ALL chains (e.g. after an all-gather),
 * d_r1/d_r2 are int32 chain indices into it; this call proposes for the n
 * chains [first, first + n), and d_eps/d_r1/d_r2/d_proposed have n rows.
 */
int pb200_de_propose(const double *d_current_all, int64_t first,
                     const double *d_eps, const int32_t *d_r1,
                     const int32_t *d_r2, double *d_proposed, int64_t n,
                     int32_t d, double gamma, void *stream);

/*
 * One whole ask() / tell() of the adaptive-covariance family per launch, for
 * device-RNG runs (CudaMCMCController): the numbers drawn are exactly those
 * pb200_philox_normal / pb200_philox_uniform produce for (seed, offset), and
 * the arithmetic is that of pb200_hbac_propose, pb200_mh_accept_ratio and
 * pb200_ac_adapt, so fused and unfused runs give identical chains.
 *   ask   x* = x + chol(sigma e^{log lambda}) z
 *   tell  accept iff [isfinite(fx) and] log(u) < fx - cur_f (require_finite as
 *         in pb200_mh_accept); count; adapt with
 *         `variant` (PB200_AC_*, or < 0 for none: initial phase, Metropolis);
 *         optionally store the new current point of chain j at
 *         d_samples[j * sample_stride + sample_offset ..+d]
 */
int pb200_ac_ask(const double *d_current, const double *d_sigma,
                 const double *d_log_lambda, double *d_proposed, int64_t n,
                 int32_t d, uint64_t seed, uint64_t offset, void *stream);
int pb200_ac_tell(int32_t variant, int32_t require_finite, double *d_current,
                  double *d_current_f, const double *d_proposed,
                  const double *d_fx,
                  uint8_t *d_accepted, int64_t *d_acceptance_count,
                  double *d_mu, double *d_sigma, double *d_log_lambda,
                  double *d_samples, int64_t sample_stride,
                  int64_t sample_offset, int64_t n, int32_t d, double gamma,
                  double target, uint64_t seed, uint64_t offset, void *stream);

/* The same for DifferentialEvolutionMCMC.ask
 * (pints/_mcmc/_differential_evolution.py:103-115): eps_j = b* z_j with the
 * normals of (seed, offset_eps), partners from (seed, offset_pairs) as
 * pb200_philox_pairs draws them, then the proposal of pb200_de_propose; its
 * tell() is pb200_ac_tell with variant < 0 and require_finite = 0. */
int pb200_de_ask(const double *d_current_all, int64_t first, int64_t n_total,
                 const double *d_b_star, double *d_proposed, int64_t n,
                 int32_t d, double gamma, uint64_t seed, uint64_t offset_eps,
                 uint64_t offset_pairs, void *stream);

/* The same three launches with their per-iteration scalars read from DEVICE
 * memory, so that one captured MCMC iteration (ask, solve, tell, advance) can
 * be replayed from a CUDA graph with no host work in the loop.  d_rows: a
 * table planned in advance, eight 64-bit slots per iteration
 *   0 ask offset   1 second ask offset (DE partner pairs)   2 tell offset
 *   3 sample_offset   4 gamma of ask (DE, double bits)   5 gamma of tell
 *   (double bits)   6 variant of tell   7 unused;
 * row *d_counter is used, pb200_iter_advance adds one to the counter.  The
 * `variant` argument of pb200_ac_tell_it only says which buffers must be
 * present (the row's variant is the one applied). */
int pb200_ac_ask_it(const double *d_current, const double *d_sigma,
                    const double *d_log_lambda, double *d_proposed, int64_t n,
                    int32_t d, uint64_t seed, const int64_t *d_rows,
                    const int64_t *d_counter, void *stream);
int pb200_de_ask_it(const double *d_current_all, int64_t first, int64_t n_total,
                    const double *d_b_star, double *d_proposed, int64_t n,
                    int32_t d, uint64_t seed, const int64_t *d_rows,
                    const int64_t *d_counter, void *stream);
int pb200_ac_tell_it(int32_t variant, int32_t require_finite, double *d_current,
                     double *d_current_f, const double *d_proposed,
                     const double *d_fx, uint8_t *d_accepted,
                     int64_t *d_acceptance_count, double *d_mu, double *d_sigma,
                     double *d_log_lambda, double *d_samples,
                     int64_t sample_stride, int64_t n, int32_t d, double target,
                     uint64_t seed, const int64_t *d_rows,
                     const int64_t *d_counter, void *stream);
int pb200_iter_advance(int64_t *d_counter, void *stream);

/* The whole MCMC loop of the adaptive-covariance family on a closed-form model
 * in ONE launch (MCMCController.run's loop, pints/_mcmc/__init__.py:706-894, for
 * samplers whose chains are independent): one persistent warp per chain runs
 * n_iterations of  ask (pb200_ac_ask's arithmetic)  ->  score the proposal (the
 * arithmetic of pb200_eval, same reduction order)  ->  tell (pb200_ac_tell's),
 * reading iteration it's scalars from row first_row + it of the table d_rows
 * (layout above; the variant of row slot 6 is applied, the store offset of
 * slot 3 is used when d_samples is given).  Chains are bit-identical to the
 * launch-per-step loop.  Meant for few chains (up to ~10^3) on a cheap model,
 * where an iteration of separate launches is launch-latency bound; d must be
 * the problem's parameter count; all n warps must be resident at once. */
int pb200_mcmc_chain_run(const pb200_problem *p, int32_t require_finite,
                         double *d_current, double *d_current_f,
                         double *d_proposed, double *d_fx, uint8_t *d_accepted,
                         int64_t *d_acceptance_count, double *d_mu,
                         double *d_sigma, double *d_log_lambda,
                         double *d_samples, int64_t sample_stride, int64_t n,
                         int32_t d, double target, uint64_t seed,
                         const int64_t *d_rows, int64_t first_row,
                         int64_t n_iterations, void *stream);

/* Device-resident xNES (pints/_optimisers/_xnes.py:63-91 ask, :147-182 tell).
 * ask: z (n, d) standard normals = pb200_philox_normal(n d, seed, *d_offset)
 * (mean, A and the offset are read from device memory, so that a generation
 * captured in a CUDA graph can be replayed with new values),
 * x = mu + A z (A row-major (d, d)), inside[i] = boundary test of row i
 * (d_lower / d_upper both NULL: no boundaries); x_eval = x where inside, mu
 * elsewhere -- the reference never evaluates a point outside the boundaries,
 * the caller overrides those scores with +inf.
 * sums: out[0..d) = sum_r us[r] z[order[r]], out[d + a d + b] = sum_r us[r]
 * z[order[r]][a] z[order[r]][b]; d_partial holds n_blocks (d + d d) doubles;
 * deterministic (fixed summation order).  With d_x and d_fx (both optional)
 * out[d + d d ..] also receives the best point x[order[0]] (d values) and its
 * score fx[order[0]]. */
int pb200_xnes_ask(const double *d_mu, const double *d_A, const double *d_lower,
                   const double *d_upper, double *d_z, double *d_x,
                   double *d_x_eval, uint8_t *d_inside, int64_t n, int32_t d,
                   uint64_t seed, const uint64_t *d_offset, void *stream);
int pb200_xnes_sums(const double *d_z, const int64_t *d_order,
                    const double *d_us, double *d_partial, int32_t n_blocks,
                    double *d_out, const double *d_x, const double *d_fx,
                    int64_t n, int32_t d, void *stream);
/* The d x d part of the update on the device (one block; replaces the host's
 * pints/_optimisers/_xnes.py:170-182, and pints/_optimisers/_snes.py:150-166
 * with separable = 1): d_state = [mu (d), A (d, d) row-major, Philox offset
 * (uint64)] is updated in place from d_sums (what pb200_xnes_sums wrote):
 *   mu += eta_mu A Gd,  A = A expm(eta_A / 2 (Gm - sum_us I))
 * (expm: scaled Taylor series of degree 18 + squaring; separable: the diagonal
 * of A times exp(eta_A / 2 (Gm_aa - sum_us))), offset += offset_step.
 * d_best = [f_best, x_best (d), f_guessed, generations] (d + 3 doubles, start
 * it as [+inf, x0, +inf, 0]) follows :160-165; d_hist (optional, hist_rows
 * rows of d + 2 doubles) keeps [f_best, x_best, f_guessed] per generation,
 * row = generation mod hist_rows, for a host that synchronises only every
 * few generations. */
int pb200_xnes_update(double *d_state, const double *d_sums, double *d_best,
                      double *d_hist, int32_t hist_rows, double eta_mu,
                      double eta_A, double sum_us, int32_t d,
                      uint64_t offset_step, int32_t separable, void *stream);

/* Ranking of the scores for the device-resident optimisers: ascending argsort
 * of n <= PB200_ARGSORT_MAX doubles (`order = np.argsort(fx)`,
 * pints/_optimisers/_xnes.py:159), NaN last, ties by index, deterministic.
 * d_ws: pb200_argsort_workspace_bytes(n) bytes; *d_status becomes 1 if the
 * sample sort met a bucket it cannot hold (sort another way then), else 0. */
#define PB200_ARGSORT_MAX 262144
int64_t pb200_argsort_workspace_bytes(int64_t n);
int pb200_argsort_f64(const double *d_values, int64_t *d_order, int64_t n,
                      void *d_ws, int64_t ws_bytes, int32_t *d_status,
                      void *stream);

/* Counter-based RNG (Philox4x32-10): n doubles, N(0,1) or U(0,1), for
 * (seed, stream_id, offset).  Used only in device-RNG sampler mode. */
int pb200_philox_normal(double *d_out, int64_t n, uint64_t seed,
                        uint64_t offset, void *stream);
int pb200_philox_uniform(double *d_out, int64_t n, uint64_t seed,
                         uint64_t offset, void *stream);
/* two distinct indices != j, uniform over the other n_total - 1 chains */
int pb200_philox_pairs(int32_t *d_r1, int32_t *d_r2, int64_t n,
                       int64_t first, int64_t n_total, uint64_t seed,
                       uint64_t offset, void *stream);

/*
 * FP64 FMA-pipe probe: every thread runs `iters` rounds of 8 independent
 * dependent-FMA chains.  Executes 16 * iters flop per thread.  The caller
 * times it with events; used as the measured FP64 roofline denominator.
 */
int pb200_fp64_probe(double *d_sink, int32_t blocks, int32_t threads,
                     int64_t iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PINTS_B200_H */
