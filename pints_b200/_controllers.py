"""
Callers of the hot path.

1. ``cuda_evaluators()``: a context manager that makes the *unmodified*
   reference controllers build a ``CudaEvaluator``.  The reference has no
   ``set_evaluator()``: every controller constructs
   ``pints.SequentialEvaluator(f)`` / ``pints.ParallelEvaluator(f, ...)`` by name
   inside ``run()`` (pints/_optimisers/__init__.py:679-690,
   pints/_mcmc/__init__.py:572-584), and those names are looked up on the
   ``pints`` module at call time, so rebinding them for the duration of the run
   is enough.

2. ``XNES``: the exponential natural evolution strategy with vectorised
   ``ask``/``tell`` that consume NumPy's global stream exactly like the
   reference's per-point loops (pints/_optimisers/_xnes.py:63-91, :147-182): at
   population 65536 the reference's Python loops cost 0.7 s per generation
   against well under a millisecond of GPU work.

3. ``CudaOptimisationController`` / ``CudaMCMCController``: small stand-alone
   controllers (same constructor arguments and ``run()`` results as the
   reference's) for environments without the reference, with the whole MCMC
   iteration -- propose, evaluate, accept, adapt, store -- enqueued on the GPU.
"""
import contextlib
import os
import time

import numpy as np

from ._compat import pints_or_none
from ._util import vector


@contextlib.contextmanager
def cuda_evaluators(**options):
    """While active, ``pints.SequentialEvaluator``, ``pints.ParallelEvaluator``
    and ``pints.MultiSequentialEvaluator`` construct ``CudaEvaluator`` objects,
    so ``OptimisationController.run()`` / ``MCMCController.run()`` of the
    reference score their batches on the GPU.  ``options`` go to
    ``CudaEvaluator`` (``devices=``, ``rtol=`` ...)."""
    pints = pints_or_none()
    if pints is None:
        raise RuntimeError('the reference package `pints` is not importable')
    from ._evaluation import CudaEvaluator

    # the reference's signatures (pints/_evaluation.py:178-187, :470-471):
    # n_workers / max_tasks_per_worker / n_fn_evals_per_worker have no meaning
    # on the GPU; `args` keeps its meaning, positional or by keyword
    def make(function, args=None):
        return CudaEvaluator(function, args=args, **options)

    def make_parallel(function, n_workers=None, max_tasks_per_worker=500,
                      n_numpy_threads=1, args=None):
        return CudaEvaluator(function, args=args, **options)

    def make_multi(functions, args=None):
        return _MultiCudaEvaluator(functions, args, options)

    saved = (pints.SequentialEvaluator, pints.ParallelEvaluator,
             pints.MultiSequentialEvaluator)
    pints.SequentialEvaluator = make
    pints.ParallelEvaluator = make_parallel
    pints.MultiSequentialEvaluator = make_multi
    try:
        yield
    finally:
        (pints.SequentialEvaluator, pints.ParallelEvaluator,
         pints.MultiSequentialEvaluator) = saved


class _MultiCudaEvaluator(object):
    """Position i goes to function i (pints/_evaluation.py:416-454).  Positions
    that share a function (the same object) are scored in one launch; the
    launches of distinct functions are all enqueued, each on its evaluator's
    own streams, before the first result is awaited."""

    def __init__(self, functions, args=None, options=None):
        from ._evaluation import CudaEvaluator
        for function in functions:
            if not callable(function):
                raise ValueError('The given functions must be callable.')
        self._n = len(functions)
        self._groups = []                 # (evaluator, rows it scores)
        seen = {}
        for i, f in enumerate(functions):
            if id(f) not in seen:
                seen[id(f)] = len(self._groups)
                self._groups.append((CudaEvaluator(f, args=args, as_list=False,
                                                   **(options or {})), []))
            self._groups[seen[id(f)]][1].append(i)

    def evaluate(self, positions):
        try:
            len(positions)
        except TypeError:
            raise ValueError(
                'The argument `positions` must be a sequence of input values'
                ' to the evaluator\'s function.')
        if len(positions) != self._n:
            raise ValueError('Number of positions does not equal number of '
                             'functions.')
        pending = [(ev, rows, ev.begin([positions[i] for i in rows]))
                   for ev, rows in self._groups]
        out = [None] * self._n
        for ev, rows, work in pending:
            scores, _ = ev.finish(work)
            for i, v in zip(rows, scores.tolist()):
                out[i] = v
        return out


# ----------------------------------------------------------------------
# host-side population optimisers with vectorised ask / tell
# ----------------------------------------------------------------------
# Where the reference is importable they are subclasses of its own classes that
# replace only the per-sample Python loops of ask() and tell(); otherwise the
# stand-alone restatements are used.
if pints_or_none() is not None:
    from ._vectorised_optimisers import XNES, SNES, PSO, BareCMAES    # noqa: E402,F401
else:
    from ._standalone_optimisers import XNES, SNES, PSO, BareCMAES    # noqa: E402,F401


# ----------------------------------------------------------------------
# stand-alone controllers
# ----------------------------------------------------------------------

class CudaOptimisationController(object):
    """Minimal counterpart of ``pints.OptimisationController``
    (pints/_optimisers/__init__.py:340-1115) around a ``CudaEvaluator``:
    ``ask -> evaluate on the GPU -> tell`` until the iteration limit or no
    significant change.  Log-pdfs are maximised, error measures minimised, as
    in the reference (:414-417)."""

    def __init__(self, function, x0, sigma0=None, boundaries=None,
                 method=None, sharded=False, group=None, **evaluator_options):
        from ._evaluation import CudaEvaluator
        from ._objectives import LogPDF
        from ._compat import base
        self._function = function
        # log-pdfs are maximised, error measures (including
        # ProbabilityBasedError) minimised
        ref_logpdf = base('LogPDF')
        is_logpdf = isinstance(function, LogPDF) or (
            ref_logpdf is not object and isinstance(function, ref_logpdf))
        self._minimising = not is_logpdf
        method = XNES if method is None else method
        self._optimiser = method(x0, sigma0, boundaries)
        self._evaluator = CudaEvaluator(function, as_list=False,
                                        **evaluator_options)
        # sharded (one process per GPU, torch.distributed initialised, device-
        # resident optimiser only): every rank draws the same population (same
        # Philox key), scores its own block, the scores are exchanged by the
        # kernel itself (ScoreExchange) and every rank applies the same update
        self._sharded = bool(sharded)
        self._group = group
        if self._sharded and not getattr(self._optimiser, 'device_resident',
                                         False):
            raise ValueError('sharded=True needs a device-resident optimiser '
                             '(DeviceXNES); ShardedEvaluator covers the host '
                             'optimisers.')
        self._sign = 1.0 if self._minimising else -1.0
        self._max_iterations = 10000
        self._unchanged_max = 200
        self._unchanged_threshold = 1e-11
        self._iterations = 0
        self._evaluations = 0
        self._time = None
        self._has_run = False
        # device-resident optimisers: generations between two synchronisations
        self._sync_every = 8

    def optimiser(self):
        return self._optimiser

    def evaluator(self):
        return self._evaluator

    def set_max_iterations(self, iterations=10000):
        if iterations is not None:
            iterations = int(iterations)
            if iterations < 0:
                raise ValueError(
                    'Maximum number of iterations cannot be negative.')
        self._max_iterations = iterations

    def set_max_unchanged_iterations(self, iterations=200, threshold=1e-11):
        if iterations is not None:
            iterations = int(iterations)
            if iterations < 0:
                raise ValueError(
                    'Maximum number of iterations cannot be negative.')
        threshold = float(threshold)
        if threshold < 0:
            raise ValueError('Minimum significant change cannot be negative.')
        self._unchanged_max = iterations
        self._unchanged_threshold = threshold

    def iterations(self):
        return self._iterations

    def evaluations(self):
        return self._evaluations

    def time(self):
        return self._time

    def run(self):
        """Runs the optimisation; returns ``(x_best, f_best)`` in the user's
        sign convention."""
        if self._has_run:
            raise RuntimeError("Controller is valid for single use only")
        self._has_run = True
        if self._max_iterations is None and self._unchanged_max is None:
            raise ValueError('At least one stopping criterion must be set.')
        start = time.perf_counter()
        unchanged, f_sig = 0, np.inf
        opt = self._optimiser
        on_device = getattr(opt, 'device_resident', False)
        if on_device:
            # population, scores and ranking stay in HBM (DeviceXNES)
            problem = self._evaluator.device_problem()
            opt._device_arg = problem.device
        exchange = None
        if self._sharded:
            import torch.distributed as dist
            from ._dist import ScoreExchange
            world = dist.get_world_size(self._group)
            rank = dist.get_rank(self._group)
            n = opt.population_size()
            per = (n + world - 1) // world
            lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
            exchange = ScoreExchange(per, problem.device, self._group)
        best = None
        while True:
            # f_best after each generation of this round, oldest first
            if exchange is not None:
                xs = opt.ask()
                fs = exchange.run(problem, xs[lo:hi], n_total=n)
                opt.tell(-fs if self._sign < 0 else fs)
                history = [(opt.f_best(), None, len(fs))]
            elif on_device:
                # a block of generations (ask + score + tell, replayed from a
                # CUDA graph) with ONE synchronisation: nothing in a generation
                # needs the host; the stopping criteria are then applied
                # generation by generation to the history the device kept
                block = self._sync_every
                if self._max_iterations is not None:
                    block = max(1, min(block, self._max_iterations - self._iterations))
                rows = opt.generations(problem, block, negate=self._sign < 0)
                history = [(float(r[0]), r, opt.population_size()) for r in rows]
            else:
                xs = opt.ask()
                fs = self._evaluator.evaluate(xs)
                opt.tell(self._sign * fs if self._sign < 0 else fs)
                history = [(opt.f_best(), None, len(fs))]
            stop = False
            for f_new, row, n_evals in history:
                self._iterations += 1
                self._evaluations += n_evals
                best = row
                if np.abs(f_new - f_sig) >= self._unchanged_threshold:
                    unchanged, f_sig = 0, f_new
                else:
                    unchanged += 1
                if self._max_iterations is not None and \
                        self._iterations >= self._max_iterations:
                    stop = True
                if self._unchanged_max is not None and \
                        unchanged >= self._unchanged_max:
                    stop = True
                if stop:
                    break
            if stop:
                break
        self._time = time.perf_counter() - start
        if best is not None:
            # the generation the criteria stopped at (the device may have run
            # the rest of its block beyond it)
            d = len(best) - 2
            return np.array(best[1:1 + d], copy=True), self._sign * float(best[0])
        self._time = time.perf_counter() - start
        return opt.x_best(), self._sign * opt.f_best()


class CudaMCMCController(object):
    """Counterpart of ``pints.MCMCController`` (pints/_mcmc/__init__.py:242-1119)
    for the two device samplers: the whole iteration -- propose, evaluate,
    accept/reject, adapt, store -- is enqueued on the GPU and only the finished
    chains come back to the host.

    ``sharded=True`` (one process per GPU) runs a contiguous block of the chains
    on every rank with a rank-dependent Philox seed: results are reproducible
    for a given seed AND world size, and differ (statistically equivalent
    chains) between world sizes.

    Parameters follow the reference: ``log_pdf`` (a supported LogPDF, see
    ``CudaEvaluator``), ``chains``, ``x0`` (``chains`` starting points),
    ``sigma0``, ``method`` (``HaarioBardenetACMC`` (default), ``HaarioACMC``,
    ``RaoBlackwellACMC``, ``MetropolisRandomWalkMCMC`` or
    ``DifferentialEvolutionMCMC`` from this package).
    """

    #: replay the iterations from a CUDA graph (see ``run``);
    #: PB200_MCMC_GRAPH=0 in the environment turns it off
    use_graph = os.environ.get('PB200_MCMC_GRAPH', '1') != '0'
    #: run independent chains on a closed-form model as one persistent launch
    #: (``pb200_mcmc_chain_run``); PB200_MCMC_CHAIN_LOOP=0 turns it off
    use_chain_loop = os.environ.get('PB200_MCMC_CHAIN_LOOP', '1') != '0'
    #: ... for at most this many chains per GPU (one warp each, all resident)
    chain_loop_max_chains = 1024
    #: ... for runs with at least this many iterations left: capturing costs
    #: about 2 ms, and only loops whose iteration is shorter on the GPU than
    #: the host's 40-60 us of launch work gain from the replay (7 us per
    #: iteration for a few chains on a closed-form model)
    graph_min_iterations = 1000

    def __init__(self, log_pdf, chains, x0, sigma0=None, method=None,
                 device=None, seed=0, sharded=False, group=None,
                 **evaluator_options):
        from ._evaluation import CudaEvaluator
        from ._samplers import DifferentialEvolutionMCMC, HaarioBardenetACMC
        self._n_chains = int(chains)
        # sharded: one process per GPU (torch.distributed); every rank is given
        # ALL starting points and runs a contiguous block of the chains.  The
        # adaptive-covariance samplers need no exchange at all; differential
        # evolution proposes from the positions of two other chains, so the
        # (n_chains, d) positions are all-gathered once per iteration
        # (SURVEY.md section 8e).
        self._sharded = bool(sharded)
        self._group = group
        self._first, self._n_local = 0, self._n_chains
        if self._sharded:
            import os
            import torch.distributed as dist
            from ._dist import shard_bounds
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            if self._n_chains % world != 0:
                raise ValueError('The number of chains must be a multiple of '
                                 'the number of ranks.')
            lo, hi, _ = shard_bounds(self._n_chains, world, rank)
            self._first, self._n_local = lo, hi - lo
            # distinct device streams per rank.  NB: the Philox counters are
            # indexed by the LOCAL chain number, so a sharded run draws other
            # numbers than the same chains on one GPU (or on another number of
            # GPUs): the chains are statistically, not bitwise, the same.
            seed = int(seed) + 7919 * rank
            if device is None:
                device = 'cuda:%d' % int(os.environ.get('LOCAL_RANK', '0'))
        if self._n_chains < 1:
            raise ValueError('Number of chains must be at least 1.')
        if len(x0) != self._n_chains:
            raise ValueError(
                'Number of initial positions must be equal to number of'
                ' chains.')
        self._evaluator = CudaEvaluator(log_pdf, devices=device, as_list=False,
                                        **evaluator_options)
        self._n_parameters = self._evaluator.n_parameters()
        x0 = np.array([vector(x) for x in x0])
        if x0.shape[1] != self._n_parameters:
            raise ValueError(
                'All initial positions must have the same dimension as the'
                ' given LogPDF.')
        method = HaarioBardenetACMC if method is None else method
        lo, hi = self._first, self._first + self._n_local
        self._gather_positions = False
        if isinstance(method, type) and issubclass(method, HaarioBardenetACMC):
            # the single-chain family: HaarioBardenetACMC, HaarioACMC,
            # RaoBlackwellACMC, MetropolisRandomWalkMCMC (all chains at once)
            s0 = sigma0
            if s0 is not None and np.shape(s0) == (self._n_chains,) + \
                    (self._n_parameters,) * 2:
                s0 = np.asarray(s0)[lo:hi]
            self._sampler = method(x0[lo:hi], s0, device=device, seed=seed)
        elif method is DifferentialEvolutionMCMC:
            self._sampler = method(self._n_local, x0[lo:hi], device=device,
                                   seed=seed, first=lo,
                                   n_total=self._n_chains,
                                   x0_mean=np.mean(x0, axis=0))
            self._gather_positions = self._sharded
        else:
            raise ValueError(
                'method must be one of the device samplers of pints_b200: '
                'HaarioBardenetACMC, HaarioACMC, RaoBlackwellACMC, '
                'MetropolisRandomWalkMCMC or DifferentialEvolutionMCMC')
        self._max_iterations = 10000
        self._initial_phase_iterations = 200 \
            if self._sampler.needs_initial_phase() else None
        self._samples = None
        self._time = None
        self._has_run = False
        self._n_evaluations = 0

    def sampler(self):
        return self._sampler

    def evaluator(self):
        return self._evaluator

    def set_max_iterations(self, iterations=10000):
        if iterations is not None:
            iterations = int(iterations)
            if iterations < 0:
                raise ValueError(
                    'Maximum number of iterations cannot be negative.')
        self._max_iterations = iterations

    def set_initial_phase_iterations(self, iterations=200):
        if self._initial_phase_iterations is None:
            raise NotImplementedError
        iterations = int(iterations)
        if iterations < 0:
            raise ValueError(
                'Number of initial-phase iterations cannot be negative.')
        self._initial_phase_iterations = iterations

    def n_evaluations(self):
        return self._n_evaluations

    def time(self):
        return self._time

    def device_samples(self):
        """The chains as a device tensor ``(n_chains, n_iterations, d)``."""
        return self._samples

    def run(self, to_host=True):
        """Runs the chains; returns them as ``(n_chains, n_iterations, d)``
        (NumPy array, or the device tensor when ``to_host`` is false)."""
        import torch
        if self._has_run:
            raise RuntimeError("Controller is valid for single use only")
        self._has_run = True
        if self._max_iterations is None:
            raise ValueError('At least one stopping criterion must be set.')
        n_iter = self._max_iterations
        sampler, problem = self._sampler, self._evaluator.device_problem()
        device = problem.device
        n_local, d = self._n_local, self._n_parameters
        samples = torch.empty((n_local, n_iter, d), dtype=torch.float64,
                              device=device)
        fx = torch.empty(n_local, dtype=torch.float64, device=device)
        pool = None
        if self._gather_positions:
            import torch.distributed as dist
            pool = torch.empty((self._n_chains, d), dtype=torch.float64,
                               device=device)
        store_in_tell = True        # the samplers write the chain sample in tell()
        switch_at = self._initial_phase_iterations \
            if sampler.needs_initial_phase() else None
        graph_ok = self.use_graph and pool is None and \
            hasattr(sampler, 'plan_iterations')
        # independent chains on a closed-form model, few enough for all their
        # warps to be resident: the whole rest of the run is ONE launch
        chain_ok = self.use_chain_loop and pool is None and \
            hasattr(sampler, 'enqueue_chain_loop') and \
            not problem.spec.is_ode() and \
            n_local <= self.chain_loop_max_chains
        self.graph_replayed = 0
        self.chain_loop_iterations = 0
        start = time.perf_counter()
        with torch.cuda.device(device):
            it = 0
            while it < n_iter:
                if chain_ok and it >= 2 and n_iter - it >= 4 and \
                        sampler.graph_ready():
                    k = n_iter - it
                    rows = torch.from_numpy(sampler.plan_iterations(
                        k, samples, it, switch_at)).to(device)
                    sampler.enqueue_chain_loop(problem, fx, samples, rows, k)
                    self._graph_rows = rows         # alive until the sync below
                    self.chain_loop_iterations = k
                    it = n_iter
                    break
                if graph_ok and it >= 2 and \
                        n_iter - it >= max(4, self.graph_min_iterations) and \
                        sampler.graph_ready():
                    # The rest of the run from a CUDA graph: the scalars that
                    # change from one iteration to the next (Philox offsets,
                    # store offset, gamma, phase) are planned into a device
                    # table, a device counter picks the row, and one captured
                    # iteration -- ask, solve, tell, advance -- is replayed
                    # with no other host work in the loop.
                    k = n_iter - it
                    rows = torch.from_numpy(sampler.plan_iterations(
                        k, samples, it, switch_at)).to(device)
                    counter = torch.zeros(1, dtype=torch.int64, device=device)

                    def one_iteration():
                        xs = sampler.enqueue_ask_it(rows, counter)
                        problem.eval(xs, out=fx)
                        sampler.enqueue_tell_it(fx, samples, rows, counter)
                        sampler._ops.iter_advance(counter)
                    try:
                        from ._engine import capture_graph
                        graph = capture_graph(device, one_iteration)
                        for _ in range(k):
                            graph.replay()
                        self.graph_replayed = k
                    except Exception:
                        # the planned launches work without a graph as well;
                        # start again from the first planned row
                        torch.cuda.synchronize(device)
                        counter.zero_()
                        for _ in range(k):
                            one_iteration()
                    self._graph_rows = rows         # alive until the sync below
                    it = n_iter
                    break
                if self._initial_phase_iterations is not None and \
                        it == self._initial_phase_iterations:
                    sampler.set_initial_phase(False)
                if pool is not None and it > 0:
                    # the one exchange of the sharded DE sampler: everybody's
                    # current positions (n_chains * d * 8 bytes)
                    dist.all_gather_into_tensor(
                        pool, sampler.current().contiguous(),
                        group=self._group)
                    xs = sampler.ask(all_current=pool)
                else:
                    xs = sampler.ask()
                problem.eval(xs, out=fx)
                if store_in_tell:
                    sampler.tell(fx, store=(samples, it))
                else:
                    current, _, _ = sampler.tell(fx)
                    samples[:, it].copy_(current)
                it += 1
            if self._sharded:
                import torch.distributed as dist
                every = torch.empty((self._n_chains, n_iter, d),
                                    dtype=torch.float64, device=device)
                dist.all_gather_into_tensor(every, samples, group=self._group)
                samples = every
            torch.cuda.synchronize(device)
        self._n_evaluations = n_iter * self._n_chains
        self._time = time.perf_counter() - start
        self._samples = samples
        return samples.cpu().numpy() if to_host else samples
