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
import time

import numpy as np
import scipy.linalg

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

    def make(function, *args, **kwargs):
        # n_workers / max_tasks_per_worker / n_fn_evals_per_worker have no
        # meaning on the GPU; `args` keeps its meaning
        return CudaEvaluator(function, args=kwargs.get('args'), **options)

    def make_multi(functions, args=None):
        evs = [CudaEvaluator(f, **options) for f in functions]
        return _MultiCudaEvaluator(evs)

    saved = (pints.SequentialEvaluator, pints.ParallelEvaluator,
             pints.MultiSequentialEvaluator)
    pints.SequentialEvaluator = make
    pints.ParallelEvaluator = make
    pints.MultiSequentialEvaluator = make_multi
    try:
        yield
    finally:
        (pints.SequentialEvaluator, pints.ParallelEvaluator,
         pints.MultiSequentialEvaluator) = saved


class _MultiCudaEvaluator(object):
    """Position i goes to function i (pints/_evaluation.py:416-454)."""

    def __init__(self, evaluators):
        self._evs = evaluators

    def evaluate(self, positions):
        if len(positions) != len(self._evs):
            raise ValueError('Number of positions does not equal number of '
                             'functions.')
        return [ev.evaluate([x])[0] for ev, x in zip(self._evs, positions)]


# ----------------------------------------------------------------------
# vectorised xNES
# ----------------------------------------------------------------------

class XNES(object):
    """Exponential natural evolution strategy, a vectorised restatement of
    ``pints.XNES`` (pints/_optimisers/_xnes.py).

    Same constructor, same ``ask()``/``tell()`` protocol, same hyper-parameters
    (Table 1 of Glasmachers et al. 2010), and the same consumption of
    ``numpy.random``: one ``normal(0, 1, (population, d))`` call draws exactly
    the numbers of the reference's ``population`` calls of ``normal(0, 1, d)``.
    """

    def __init__(self, x0, sigma0=None, boundaries=None):
        self._x0 = vector(x0)
        self._n_parameters = len(self._x0)
        if self._n_parameters < 1:
            raise ValueError('Problem dimension must be greater than zero.')
        self._boundaries = boundaries
        if boundaries is not None:
            if boundaries.n_parameters() != self._n_parameters:
                raise ValueError(
                    'Boundaries must have same dimension as starting point.')
            if not boundaries.check(self._x0):
                raise ValueError(
                    'Initial position must lie within given boundaries.')
        if sigma0 is None:
            # as pints/_optimisers/__init__.py:113-126
            try:
                self._sigma0 = (1 / 6) * boundaries.range()
            except AttributeError:
                self._sigma0 = (1 / 3) * np.abs(self._x0)
                self._sigma0 = self._sigma0 + (self._sigma0 == 0)
        elif np.isscalar(sigma0):
            sigma0 = float(sigma0)
            if sigma0 <= 0:
                raise ValueError(
                    'Initial standard deviation must be greater than zero.')
            self._sigma0 = np.ones(self._n_parameters) * sigma0
        else:
            self._sigma0 = vector(sigma0)
            if len(self._sigma0) != self._n_parameters:
                raise ValueError(
                    'Initial standard deviation must be None, scalar, or have'
                    ' dimension ' + str(self._n_parameters) + '.')
            if np.any(self._sigma0 <= 0):
                raise ValueError(
                    'Initial standard deviations must be greater than zero.')
        self._population_size = self._suggested_population_size()
        self._running = False
        self._ready_for_tell = False
        self._x_best = self._x0
        self._f_best = np.inf
        self._f_guessed = np.inf

    # -- PopulationBasedOptimiser interface ---------------------------
    def _suggested_population_size(self):
        return 4 + int(3 * np.log(self._n_parameters))

    def population_size(self):
        return self._population_size

    def set_population_size(self, population_size=None):
        if self._running:
            raise Exception('Cannot change population size during run.')
        if population_size is None:
            population_size = self._suggested_population_size()
        population_size = int(population_size)
        if population_size < 1:
            raise ValueError('Population size must be at least 1.')
        self._population_size = population_size

    def suggested_population_size(self, round_up_to_multiple_of=None):
        p = self._suggested_population_size()
        if round_up_to_multiple_of is not None:
            n = int(round_up_to_multiple_of)
            if n > 1:
                p = n * (((p - 1) // n) + 1)
        return p

    def running(self):
        return self._running

    def f_best(self):
        return self._f_best

    def f_guessed(self):
        return self._f_guessed

    def x_best(self):
        return self._x_best

    def x_guessed(self):
        return self._mu if self._running else self._x0

    def needs_sensitivities(self):
        return False

    def stop(self):
        return False

    @classmethod
    def name(cls):
        return 'Exponential Natural Evolution Strategy (xNES)'

    def _initialise(self):
        d, n = self._n_parameters, self._population_size
        self._eta_mu = 1
        self._eta_A = 0.6 * (3 + np.log(d)) * d ** -1.5
        us = np.maximum(0, np.log(n / 2 + 1) - np.log(1 + np.arange(n)))
        us /= np.sum(us)
        us -= 1 / n
        self._us = us
        self._mu = np.array(self._x0, copy=True)
        self._A = np.eye(d) * self._sigma0
        self._I = np.eye(d)
        self._running = True

    def ask(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        n, d = self._population_size, self._n_parameters
        self._zs = np.random.normal(0, 1, (n, d))
        self._xs = self._mu + self._zs @ self._A.T
        if self._boundaries is not None:
            inside = self._boundaries.check(self._xs) \
                if self._xs.ndim == 2 and _vector_check(self._boundaries) \
                else np.array([self._boundaries.check(x) for x in self._xs])
            self._bounded_ids = np.nonzero(inside)
            self._bounded_xs = self._xs[self._bounded_ids]
        else:
            self._bounded_xs = self._xs
        self._bounded_xs.setflags(write=False)
        return self._bounded_xs

    def tell(self, fx):
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        n = self._population_size
        fx = np.asarray(fx, dtype=float)
        if self._boundaries is not None and len(fx) < n:
            full = np.ones((n,)) * np.inf
            full[self._bounded_ids] = fx
            fx = full
        order = np.argsort(fx)
        zs = self._zs[order]
        self._zs = zs
        # centre:  mu += eta_mu * A (sum_i u_i z_i)
        Gd = np.dot(self._us, zs)
        self._mu = self._mu + self._eta_mu * np.dot(self._A, Gd)
        # shape:   Gm = sum_i u_i (z_i z_i^T - I)
        Gm = (zs * self._us[:, None]).T @ zs - np.sum(self._us) * self._I
        self._A = np.dot(self._A, scipy.linalg.expm(0.5 * self._eta_A * Gm))
        self._f_guessed = fx[order[0]]
        if self._f_guessed < self._f_best:
            self._x_best = np.array(self._xs[order[0]], copy=True)
            self._f_best = fx[order[0]]


def _vector_check(boundaries):
    """True when ``boundaries.check`` accepts an (n, d) array (the mirror's
    RectangularBoundaries does, the reference's returns a single bool)."""
    return type(boundaries).__module__.startswith('pints_b200')


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
                 method=None, **evaluator_options):
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
        self._sign = 1.0 if self._minimising else -1.0
        self._max_iterations = 10000
        self._unchanged_max = 200
        self._unchanged_threshold = 1e-11
        self._iterations = 0
        self._evaluations = 0
        self._time = None
        self._has_run = False

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
        while True:
            xs = opt.ask()
            fs = self._evaluator.evaluate(xs)
            opt.tell(self._sign * fs if self._sign < 0 else fs)
            self._evaluations += len(fs)
            self._iterations += 1
            f_new = opt.f_best()
            if np.abs(f_new - f_sig) >= self._unchanged_threshold:
                unchanged, f_sig = 0, f_new
            else:
                unchanged += 1
            if self._max_iterations is not None and \
                    self._iterations >= self._max_iterations:
                break
            if self._unchanged_max is not None and \
                    unchanged >= self._unchanged_max:
                break
        self._time = time.perf_counter() - start
        return opt.x_best(), self._sign * opt.f_best()


class CudaMCMCController(object):
    """Counterpart of ``pints.MCMCController`` (pints/_mcmc/__init__.py:242-1119)
    for the two device samplers: the whole iteration -- propose, evaluate,
    accept/reject, adapt, store -- is enqueued on the GPU and only the finished
    chains come back to the host.

    Parameters follow the reference: ``log_pdf`` (a supported LogPDF, see
    ``CudaEvaluator``), ``chains``, ``x0`` (``chains`` starting points),
    ``sigma0``, ``method`` (``HaarioBardenetACMC`` (default), ``HaarioACMC``,
    ``RaoBlackwellACMC``, ``MetropolisRandomWalkMCMC`` or
    ``DifferentialEvolutionMCMC`` from this package).
    """

    def __init__(self, log_pdf, chains, x0, sigma0=None, method=None,
                 device=None, seed=0, **evaluator_options):
        from ._evaluation import CudaEvaluator
        from ._samplers import DifferentialEvolutionMCMC, HaarioBardenetACMC
        self._n_chains = int(chains)
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
        if isinstance(method, type) and issubclass(method, HaarioBardenetACMC):
            # the single-chain family: HaarioBardenetACMC, HaarioACMC,
            # RaoBlackwellACMC, MetropolisRandomWalkMCMC (all chains at once)
            self._sampler = method(x0, sigma0, device=device, seed=seed)
        elif method is DifferentialEvolutionMCMC:
            self._sampler = method(self._n_chains, x0, device=device,
                                   seed=seed)
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
        samples = torch.empty(
            (self._n_chains, n_iter, self._n_parameters), dtype=torch.float64,
            device=device)
        fx = torch.empty(self._n_chains, dtype=torch.float64, device=device)
        start = time.perf_counter()
        with torch.cuda.device(device):
            for it in range(n_iter):
                if self._initial_phase_iterations is not None and \
                        it == self._initial_phase_iterations:
                    sampler.set_initial_phase(False)
                xs = sampler.ask()
                problem.eval(xs, out=fx)
                current, _, _ = sampler.tell(fx)
                samples[:, it].copy_(current)
            torch.cuda.synchronize(device)
        self._n_evaluations = n_iter * self._n_chains
        self._time = time.perf_counter() - start
        self._samples = samples
        return samples.cpu().numpy() if to_host else samples
