"""
Host-side population optimisers for installations WITHOUT the reference
package: stand-alone XNES, SNES, PSO and BareCMAES with the vectorised
``ask`` / ``tell`` of ``_vectorised_optimisers.py`` plus the constructor,
hyper-parameter set-up and accessors that module otherwise inherits from
``pints.XNES`` / ``pints.SNES`` / ``pints.PSO`` / ``pints.BareCMAES``
(pints/_optimisers/_xnes.py, _snes.py, _pso.py, _cmaes_bare.py and
pints/_optimisers/__init__.py:24-312).  The update rules and their constants
are those of the reference, in the same order on the same ``numpy.random``
stream, so this module is by necessity a restatement; it is imported only when
``import pints`` fails (``pints_b200/_controllers.py``).
"""
import functools

import numpy as np
import scipy.linalg

from ._util import single_thread_blas, vector


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
