"""
Host-side population optimisers for populations of GPU size: ``pints.XNES``,
``pints.SNES``, ``pints.PSO`` and ``pints.BareCMAES`` with the per-sample Python
loops of ``ask()`` and ``tell()`` replaced by array expressions.

At population 65 536 the reference's loops cost 0.7 s per generation against
well under a millisecond of GPU work; the hot path needs
its callers to keep up.  Everything else -- constructor checks, default
``sigma0``, hyper-parameters, ``_initialise``, accessors, stopping -- IS the
reference's code: each class below subclasses the reference class and overrides
only what loops over the population.  The overrides consume ``numpy.random``
exactly as the loops do (``normal(0, 1, d)`` called ``n`` times draws the
numbers of one ``normal(0, 1, (n, d))`` call, in the same order), so a run is
the reference's run: ``tests/test_controllers_cpu.py`` compares trajectories
and the generator state after every generation.

This module needs the reference (``import pints``); without it
``pints_b200/_standalone_optimisers.py`` provides the same classes.
"""
import functools

import numpy as np
import pints
import scipy.linalg

from ._util import single_thread_blas


def _with_single_thread_blas(method):
    @functools.wraps(method)
    def wrapped(self, *args, **kwargs):
        with single_thread_blas():
            return method(self, *args, **kwargs)
    return wrapped


def _vector_check(boundaries):
    """True when ``boundaries.check`` accepts an (n, d) array (this package's
    RectangularBoundaries does, the reference's returns a single bool)."""
    return type(boundaries).__module__.startswith('pints_b200')


class _Vectorised(object):
    """What the vectorised ask() / tell() share: the boundary filter of a
    whole population and the padding of its scores."""

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
            # whole population inside (the common case): per-column extrema
            # instead of a row-wise test and a gather
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

    def _asked(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        return self._population_size, self._n_parameters

    def _told(self):
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False


class XNES(_Vectorised, pints.XNES):
    """``pints.XNES`` (pints/_optimisers/_xnes.py) with the loops of ``ask``
    (:73-76) and ``tell`` (:170-173) as matrix products."""

    def ask(self):
        n, d = self._asked()
        self._zs = np.random.normal(0, 1, (n, d))
        self._xs = self._zs @ self._A.T
        self._xs += self._mu                # in place: the sum commutes
        self._bounded_ids, self._bounded_xs = self._select(self._xs)
        self._bounded_xs.setflags(write=False)
        return self._bounded_xs

    def tell(self, fx):
        self._told()
        fx = self._full_scores(fx, self._bounded_ids)
        order = np.argsort(fx)
        zs = self._zs = self._zs[order]
        # centre: sum_i u_i z_i; shape: sum_i u_i (z_i z_i^T - I)
        Gd = np.dot(self._us, zs)
        Gm = (zs * self._us[:, None]).T @ zs - np.sum(self._us) * self._I
        self._mu = self._mu + self._eta_mu * np.dot(self._A, Gd)
        self._A = np.dot(self._A, scipy.linalg.expm(0.5 * self._eta_A * Gm))
        self._f_guessed = fx[order[0]]
        if self._f_guessed < self._f_best:
            self._x_best = np.array(self._xs[order[0]], copy=True)
            self._f_best = fx[order[0]]


class SNES(_Vectorised, pints.SNES):
    """``pints.SNES`` (pints/_optimisers/_snes.py) with the sampling loop
    (:66-68) as one array expression."""

    def ask(self):
        n, d = self._asked()
        self._ss = np.random.normal(0, 1, (n, d))
        self._xs = self._mu + self._sigmas * self._ss
        self._user_ids, self._user_xs = self._select(self._xs)
        self._user_xs.setflags(write=False)
        return self._user_xs

    def tell(self, fx):
        self._told()
        fx = self._full_scores(fx, self._user_ids)
        order = np.argsort(fx)
        self._ss = self._ss[order]
        step = np.dot(self._us, self._ss)
        spread = np.dot(self._us, self._ss**2 - 1)
        self._mu = self._mu + self._eta_mu * self._sigmas * step
        self._sigmas = self._sigmas * np.exp(0.5 * self._eta_sigmas * spread)
        self._f_guessed = fx[order[0]]
        if self._f_guessed < self._f_best:
            self._x_best = np.array(self._xs[order[0]], copy=True)
            self._f_best = fx[order[0]]


class PSO(_Vectorised, pints.PSO):
    """``pints.PSO`` (pints/_optimisers/_pso.py) with the particle loop of
    ``tell`` (:234-256) as array expressions.  The reference draws
    ``uniform(0, almax, d)`` then ``uniform(0, agmax, d)`` per particle; one
    ``uniform(0, 1, (population, 2, d))`` call draws the same numbers in the
    same order (the legacy ``uniform(0, a)`` is ``a * u``)."""

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
        # after particle 0.  Arrays cannot alias row by row; two flags say
        # which rows still follow.
        self._pl_follows = np.ones(n, dtype=bool)
        self._pg_follows = True
        self._filter()
        self._running = True

    def ask(self):
        self._asked()
        return np.copy(self._user_xs)

    def tell(self, fx):
        self._told()
        n, d = self._population_size, self._n_parameters
        fx = self._full_scores(fx, self._user_ids)
        better = fx < self._fl
        self._fl = np.where(better, fx, self._fl)
        self._pl_follows &= ~better
        moved_with = (better | self._pl_follows)[:, None]
        self._pl = np.where(moved_with, self._xs, self._pl)
        u = np.random.uniform(0, 1, (n, 2, d))
        al = 0.0 + (self._almax - 0.0) * u[:, 0]
        ag = 0.0 + (self._agmax - 0.0) * u[:, 1]

        def move(rows, pg):
            pull = al[rows] * (self._pl[rows] - self._xs[rows]) \
                + ag[rows] * (pg - self._xs[rows])
            vs = self._vs[rows] + pull
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


class BareCMAES(_Vectorised, pints.BareCMAES):
    """``pints.BareCMAES`` (pints/_optimisers/_cmaes_bare.py) with the
    per-sample loops of ``ask`` (:97-104) and ``tell`` (:316-328: weighted
    sums, negative-weight rescaling, rank-mu outer products) as array
    expressions."""

    def ask(self):
        n, d = self._asked()
        self._zs = np.random.normal(0, 1, (n, d))
        self._ys = self._zs @ self._R.dot(self._S).T
        self._xs = self._mu + self._eta * self._ys
        self._user_ids, self._user_xs = self._select(self._xs)
        self._user_xs.setflags(write=False)
        return self._user_xs

    def tell(self, fx):
        from numpy.linalg import norm
        self._told()
        self._iterations += 1
        n, npa = self._n_parameters, self._parent_pop_size
        fx = self._full_scores(fx, self._user_ids)
        order = np.argsort(fx)
        xs, zs, ys = self._xs[order], self._zs[order], self._ys[order]
        W = self._W
        best = W[:npa, None]
        self._mu = self._mu + self._cm * np.sum((xs[:npa] - self._mu) * best, axis=0)
        zmeans = np.sum(zs[:npa] * best, 0)
        ymeans = np.sum(ys[:npa] * best, 0)
        c = np.sqrt(self._csig * (2 - self._csig) * self._muEff)
        self._psig = (1 - self._csig) * self._psig + c * self._R.dot(zmeans)
        damped = 1 - (1 - self._csig) ** (2 * (self._iterations + 1))
        h_sig = 1 if norm(self._psig) / np.sqrt(damped) \
            < (1.4 + 2 / (n + 1)) * self._e else 0
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
