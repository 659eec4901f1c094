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
import scipy.linalg

from ._compat import pints_or_none
import functools

from ._util import single_thread_blas, vector


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
# vectorised xNES
# ----------------------------------------------------------------------

def _with_single_thread_blas(method):
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        with single_thread_blas():
            return method(self, *args, **kwargs)
    return wrapped


class _PopulationOptimiser(object):
    """What the vectorised optimisers share with
    ``pints.PopulationBasedOptimiser`` (pints/_optimisers/__init__.py:24-312):
    constructor checks, default ``sigma0``, population-size handling."""

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

    def __init_subclass__(cls, **kwargs):
        # ask() and tell() run their small matrix products on one BLAS thread
        super().__init_subclass__(**kwargs)
        for name in ('ask', 'tell'):
            method = cls.__dict__.get(name)
            if method is not None:
                setattr(cls, name, _with_single_thread_blas(method))

    def _select(self, xs):
        """``(ids, user_xs)``: the points of the population that lie inside
        the boundaries, as the reference filters them before evaluation
        (pints/_optimisers/_xnes.py:79-87); ``ids`` is None when all do (no
        copy is made then)."""
        if self._boundaries is None:
            return None, xs
        if _vector_check(self._boundaries):
            # whole population inside (the common case): two reductions
            # instead of a row-wise test and a gather
            # (column by column: numpy's axis-0 reduction of an (n, 3) array
            # is ten times slower than three strided 1-d ones)
            lo, hi = self._boundaries.lower(), self._boundaries.upper()
            if all(xs[:, k].min() >= lo[k] and xs[:, k].max() < hi[k]
                   for k in range(xs.shape[1])):
                return None, xs
        ids = np.nonzero(self._inside(xs))
        return ids, xs[ids]

    def _inside(self, xs):
        """Row-wise ``boundaries.check`` of an (n, d) array."""
        if _vector_check(self._boundaries):
            return self._boundaries.check(xs)
        return np.array([self._boundaries.check(x) for x in xs])

    def _full_scores(self, fx, ids):
        """Scores of the whole population: out-of-bounds points count as inf
        (pints/_optimisers/_xnes.py:154-157 and the same lines elsewhere)."""
        fx = np.asarray(fx, dtype=float)
        if ids is not None and len(fx) < self._population_size:
            full = np.ones((self._population_size,)) * np.inf
            full[ids] = fx
            fx = full
        return fx


class XNES(_PopulationOptimiser):
    """Exponential natural evolution strategy, a vectorised restatement of
    ``pints.XNES`` (pints/_optimisers/_xnes.py).

    Same constructor, same ``ask()``/``tell()`` protocol, same hyper-parameters
    (Table 1 of Glasmachers et al. 2010), and the same consumption of
    ``numpy.random``: one ``normal(0, 1, (population, d))`` call draws exactly
    the numbers of the reference's ``population`` calls of ``normal(0, 1, d)``.
    """

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
        self._xs = self._zs @ self._A.T
        self._xs += self._mu                # in place: the sum commutes
        self._bounded_ids, self._bounded_xs = self._select(self._xs)
        self._bounded_xs.setflags(write=False)
        return self._bounded_xs

    def tell(self, fx):
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        n = self._population_size
        fx = self._full_scores(fx, self._bounded_ids)
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


class SNES(_PopulationOptimiser):
    """Separable natural evolution strategy, vectorised restatement of
    ``pints.SNES`` (pints/_optimisers/_snes.py:56-166): one
    ``normal(0, 1, (population, d))`` call per generation instead of
    ``population`` calls, same stream."""

    @classmethod
    def name(cls):
        return 'Seperable Natural Evolution Strategy (SNES)'

    def _initialise(self):
        d, n = self._n_parameters, self._population_size
        self._eta_mu = 1
        self._eta_sigmas = 0.2 * (3 + np.log(d)) * d ** -0.5
        us = np.maximum(0, np.log(n / 2 + 1) - np.log(1 + np.arange(n)))
        us /= np.sum(us)
        us -= 1 / n
        self._us = us
        self._mu = np.array(self._x0, copy=True)
        self._sigmas = np.array(self._sigma0, copy=True)
        self._running = True

    def ask(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        n, d = self._population_size, self._n_parameters
        self._ss = np.random.normal(0, 1, (n, d))
        self._xs = self._mu + self._sigmas * self._ss
        self._user_ids, self._user_xs = self._select(self._xs)
        self._user_xs.setflags(write=False)
        return self._user_xs

    def tell(self, fx):
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        fx = self._full_scores(fx, self._user_ids)
        order = np.argsort(fx)
        self._ss = self._ss[order]
        self._mu = self._mu + self._eta_mu * self._sigmas * np.dot(
            self._us, self._ss)
        self._sigmas = self._sigmas * np.exp(
            0.5 * self._eta_sigmas * np.dot(self._us, self._ss**2 - 1))
        self._f_guessed = fx[order[0]]
        if self._f_guessed < self._f_best:
            self._x_best = np.array(self._xs[order[0]], copy=True)
            self._f_best = fx[order[0]]


class PSO(_PopulationOptimiser):
    """Particle swarm optimisation, vectorised restatement of ``pints.PSO``
    (pints/_optimisers/_pso.py:84-271).  The reference loops over particles in
    ``tell()`` and draws ``uniform(0, almax, d)`` then ``uniform(0, agmax, d)``
    for each; here one ``uniform(0, 1, (population, 2, d))`` call draws the same
    numbers in the same order (the legacy ``uniform(0, a)`` is ``a * u``)."""

    def __init__(self, x0, sigma0=None, boundaries=None):
        super().__init__(x0, sigma0, boundaries)
        self.set_local_global_balance()

    @classmethod
    def name(cls):
        return 'Particle Swarm Optimisation (PSO)'

    def set_local_global_balance(self, r=0.5):
        if self._running:
            raise Exception('Cannot change settings during run.')
        r = float(r)
        if r < 0 or r > 1:
            raise ValueError('Parameter r must be in the range 0-1.')
        _amax = 4.1
        self._almax = r * _amax
        self._agmax = _amax - self._almax

    def f_best(self):
        return self._fg if self._running else np.inf

    def f_guessed(self):
        return self.f_best()

    def x_best(self):
        if self._running:
            return np.array(self._pg, copy=True)
        return np.array(self._x0, copy=True)

    def x_guessed(self):
        return self.x_best()

    def _filter(self):
        self._user_ids, self._user_xs = self._select(self._xs)

    def _initialise(self):
        n, d = self._population_size, self._n_parameters
        xs = [np.array(self._x0, copy=True)]
        if self._boundaries is not None:
            try:
                xs.extend(self._boundaries.sample(n - 1))
            except NotImplementedError:
                pass
        if len(xs) < n:
            xs.extend(np.random.normal(self._x0, self._sigma0, (n - 1, d)))
        self._xs = np.array(xs)
        self._vs = 1e-1 * self._sigma0 * np.random.uniform(0, 1, (n, d))
        self._fl = np.full(n, np.inf)
        self._pl = np.array(self._xs, copy=True)
        self._fg = np.inf
        self._pg = np.array(self._xs[0], copy=True)
        # In the reference the initial local bests and the initial global best
        # are VIEWS of the position rows (pints/_optimisers/_pso.py:140-146):
        # until a particle first improves, its local best moves with it, and
        # until the first finite score the global best is "wherever particle 0
        # is now" -- including, inside one tell(), for the particles updated
        # after particle 0.  Reproduced with these two flags.
        self._pl_follows = np.ones(n, dtype=bool)
        self._pg_follows = True
        self._filter()
        self._running = True

    def ask(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        return np.copy(self._user_xs)

    def tell(self, fx):
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        n, d = self._population_size, self._n_parameters
        fx = self._full_scores(fx, self._user_ids)
        better = fx < self._fl
        self._fl = np.where(better, fx, self._fl)
        self._pl_follows &= ~better
        self._pl = np.where((better | self._pl_follows)[:, None], self._xs,
                            self._pl)
        u = np.random.uniform(0, 1, (n, 2, d))
        al = 0.0 + (self._almax - 0.0) * u[:, 0]
        ag = 0.0 + (self._agmax - 0.0) * u[:, 1]

        def move(rows, pg):
            vs = self._vs[rows] + (al[rows] * (self._pl[rows] - self._xs[rows])
                                   + ag[rows] * (pg - self._xs[rows]))
            if self._boundaries is not None:
                # going out of bounds: halve the speed (_pso.py:248-251)
                out = ~np.asarray(self._inside(self._xs[rows] + vs), dtype=bool)
                vs = np.where(out[:, None], vs * 0.5, vs)
            self._vs[rows] = vs
            self._xs[rows] = self._xs[rows] + vs

        if self._pg_follows:
            move(slice(0, 1), self._xs[0])            # pg is particle 0 itself
            move(slice(1, n), np.array(self._xs[0]))  # ... after its move
        else:
            move(slice(0, n), self._pg)
        self._pl = np.where(self._pl_follows[:, None], self._xs, self._pl)
        self._filter()
        i = np.argmin(self._fl)
        if self._fl[i] < self._fg:
            self._fg = self._fl[i]
            self._pg = np.array(self._pl[i], copy=True)
            self._pg_follows = False
        elif self._pg_follows:
            self._pg = np.array(self._xs[0], copy=True)


class BareCMAES(_PopulationOptimiser):
    """Bare-bones CMA-ES, vectorised restatement of ``pints.BareCMAES``
    (pints/_optimisers/_cmaes_bare.py:86-357): the per-sample Python loops
    (sampling, rotation/scaling, negative-weight rescaling, outer products)
    become array expressions; the random stream is consumed identically."""

    def __init__(self, x0, sigma0=None, boundaries=None):
        super().__init__(x0, sigma0, boundaries)
        from scipy.special import gamma
        n = self._n_parameters
        self._eta = np.min(self._sigma0)
        self._C = np.identity(n)
        self._R = np.identity(n)
        self._S = np.identity(n)
        self._e = np.sqrt(2) * gamma((n + 1) / 2) / gamma(n / 2)

    @classmethod
    def name(cls):
        return 'Bare-bones CMA-ES'

    def cov(self, decomposed=False):
        if decomposed:
            return np.array(self._R, copy=True), np.array(self._S, copy=True)
        return np.array(self._C, copy=True)

    def mean(self):
        return np.array(self._mu if self._running else self._x0, copy=True)

    def stop(self):
        cond = np.diagonal(self._S)
        cond = (np.max(cond) / np.min(cond)) ** 2
        if cond > 1e14:
            return 'Ill-conditioned covariance matrix'
        return False

    def _initialise(self):
        # constants of the tutorial (Hansen 2016), as
        # pints/_optimisers/_cmaes_bare.py:134-228
        n, npo = self._n_parameters, self._population_size
        self._iterations = 0
        self._mu = np.array(self._x0, copy=True)
        self._parent_pop_size = npo // 2
        npa = self._parent_pop_size
        W = np.log((npo + 1) / 2.) - np.log(1 + np.arange(npo))
        self._W = W
        self._muEff = np.sum(W[:npa]) ** 2 / np.sum(np.square(W[:npa]))
        self._muEffMinus = np.sum(W[npa:]) ** 2 / np.sum(np.square(W[npa:]))
        self._pc = np.zeros(n)
        self._psig = np.zeros(n)
        self._cm = 1
        self._ccov = (4 + self._muEff / n) / (n + 4 + 2 * self._muEff / n)
        self._csig = (2 + self._muEff) / (n + 5 + self._muEff)
        self._c1 = 2 / ((n + 1.3) ** 2 + self._muEff)
        self._cmu = min(
            2 * (self._muEff - 2 + 1 / self._muEff)
            / ((n + 2) ** 2 + self._muEff),
            1 - self._c1)
        self._dsig = 1 + self._csig + 2 * max(
            0, np.sqrt((self._muEff - 1) / (n + 1)) - 1)
        alpha_mu = 1 + self._c1 / self._cmu
        alpha_mueff = 1 + 2 * self._muEffMinus / (self._muEff + 2)
        alpha_pos_def = (1 - self._c1 - self._cmu) / (n * self._cmu)
        sum_pos = np.sum(W[W > 0])
        sum_neg = np.sum(W[W < 0])
        scale_pos = 1 / sum_pos
        scale_neg = min(alpha_mu, alpha_mueff, alpha_pos_def) / -sum_neg
        pos, neg = W > 0, W < 0
        W[pos] *= scale_pos
        W[neg] *= scale_neg
        self._running = True

    def ask(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        n, d = self._population_size, self._n_parameters
        self._zs = np.random.normal(0, 1, (n, d))
        self._ys = self._zs @ self._R.dot(self._S).T
        self._xs = self._mu + self._eta * self._ys
        self._user_ids, self._user_xs = self._select(self._xs)
        self._user_xs.setflags(write=False)
        return self._user_xs

    def tell(self, fx):
        from numpy.linalg import norm
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        self._iterations += 1
        n, npa = self._n_parameters, self._parent_pop_size
        fx = self._full_scores(fx, self._user_ids)
        order = np.argsort(fx)
        xs, zs, ys = self._xs[order], self._zs[order], self._ys[order]
        W = self._W
        self._mu = self._mu + self._cm * np.sum(
            (xs[:npa] - self._mu) * W[:npa, None], axis=0)
        zmeans = np.sum(zs[:npa] * W[:npa, None], 0)
        ymeans = np.sum(ys[:npa] * W[:npa, None], 0)
        c = np.sqrt(self._csig * (2 - self._csig) * self._muEff)
        self._psig = (1 - self._csig) * self._psig + c * self._R.dot(zmeans)
        cond = (
            norm(self._psig)
            / np.sqrt(1 - (1 - self._csig) ** (2 * (self._iterations + 1)))
            < (1.4 + 2 / (n + 1)) * self._e)
        h_sig = 1 if cond else 0
        delta_sig = (1 - h_sig) * self._ccov * (2 - self._ccov)
        c = np.sqrt(self._ccov * (2 - self._ccov) * self._muEff)
        self._pc = (1 - self._ccov) * self._pc + h_sig * c * ymeans
        # negative weights are rescaled by n / ||R * z_i||_F^2, where R * z_i
        # scales the columns of R: its squared Frobenius norm is
        # sum_k z_ik^2 sum_j R_jk^2
        col2 = np.sum(self._R ** 2, axis=0)
        norms2 = (zs ** 2) @ col2
        temp_weights = np.where(W < 0, W * n / norms2, W)
        rank1 = self._c1 * np.outer(self._pc, self._pc)
        rankmu = self._cmu * (ys * temp_weights[:, None]).T @ ys
        self._C = rank1 + rankmu + self._C * (
            1 + delta_sig * self._c1 - self._c1 - self._cmu * sum(W))
        self._C = np.triu(self._C) + np.triu(self._C, 1).T
        self._eta = self._eta * np.exp(
            self._csig / self._dsig * (norm(self._psig) / self._e - 1))
        eig = np.linalg.eigh(self._C)
        self._S = np.sqrt(np.diag(eig[0]))
        self._R = eig[1]
        self._f_guessed = fx[order[0]]
        if self._f_guessed < self._f_best:
            self._f_best = self._f_guessed
            self._x_best = np.array(xs[0], copy=True)


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
        while True:
            if exchange is not None:
                xs = opt.ask()
                fs = exchange.run(problem, xs[lo:hi], n_total=n)
                opt.tell(-fs if self._sign < 0 else fs)
                self._evaluations += len(fs)
            elif on_device:
                # ask + score + tell with one synchronisation (CUDA graph)
                opt.generation(problem, negate=self._sign < 0)
                self._evaluations += opt.population_size()
            else:
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
            seed = int(seed) + 7919 * rank          # distinct device streams
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
