"""
Device-resident MCMC samplers: the companion kernels of the evaluation path.

  HaarioBardenetACMC          all chains of pints.HaarioBardenetACMC at once
                              (pints/_mcmc/_adaptive_covariance.py:109-131,
                              :236-303; pints/_mcmc/_haario_bardenet_ac.py:59-73)
  DifferentialEvolutionMCMC   pints.DifferentialEvolutionMCMC
                              (pints/_mcmc/_differential_evolution.py:89-124,
                              :146-192, :283-332)

Both keep the reference's ask/tell protocol, but for *all* chains in one call and
with the state living in HBM: ``ask()`` returns the ``(n_chains, d)`` proposals as
a device tensor, ``tell(fx)`` takes the ``(n_chains,)`` log-pdfs as a device
tensor.  The accept/reject rule and the adaptation arithmetic are bit-exact with
the reference given the same proposals, scores and uniforms (``replay`` inputs;
see tests/test_gpu_samplers.py).  With ``rng='philox'`` proposals and uniforms
come from a counter-based generator on the device; the Haario-Bardenet proposal
then uses a Cholesky factor where NumPy uses an SVD, which is a different but
equally distributed draw (statistical parity only).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._engine import _ptr, _stream_ptr, require_cuda


def _f64(x, device):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device)


class _DeviceOps(object):
    """Thin wrappers of the sampler entry points of the C ABI."""

    def __init__(self, device):
        self.device = require_cuda(device)
        self.lib = _lib.load()

    def _call(self, fn, *args):
        with torch.cuda.device(self.device):
            _lib.check(fn(*args, _stream_ptr(self.device)))

    def mh_accept(self, cur, cur_f, prop, fx, log_u, accepted, require_finite):
        n, d = cur.shape
        self._call(self.lib.pb200_mh_accept, _ptr(cur), _ptr(cur_f), _ptr(prop),
                   _ptr(fx), _ptr(log_u), _ptr(accepted), n, d,
                   1 if require_finite else 0)

    def mh_accept_ratio(self, cur, cur_f, prop, fx, log_u, accepted, log_ratio,
                        require_finite):
        n, d = cur.shape
        self._call(self.lib.pb200_mh_accept_ratio, _ptr(cur), _ptr(cur_f),
                   _ptr(prop), _ptr(fx), _ptr(log_u), _ptr(accepted),
                   _ptr(log_ratio), n, d, 1 if require_finite else 0)

    def ac_adapt(self, variant, cur, prev, prop, accepted, log_ratio, mu, sigma,
                 log_lambda, gamma, target):
        n, d = cur.shape
        self._call(self.lib.pb200_ac_adapt, int(variant), _ptr(cur),
                   _ptr(prev), _ptr(prop), _ptr(accepted), _ptr(log_ratio),
                   _ptr(mu), _ptr(sigma), _ptr(log_lambda), n, d, float(gamma),
                   float(target))

    def ac_ask(self, cur, sigma, log_lambda, prop, seed, offset):
        n, d = cur.shape
        self._call(self.lib.pb200_ac_ask, _ptr(cur), _ptr(sigma),
                   _ptr(log_lambda), _ptr(prop), n, d, ctypes.c_uint64(seed),
                   ctypes.c_uint64(offset))

    def de_ask(self, cur_all, first, b_star, prop, gamma, seed, off_eps,
               off_pairs):
        n, d = prop.shape
        self._call(self.lib.pb200_de_ask, _ptr(cur_all), int(first),
                   int(cur_all.shape[0]), _ptr(b_star), _ptr(prop), n, d,
                   float(gamma), ctypes.c_uint64(seed),
                   ctypes.c_uint64(off_eps), ctypes.c_uint64(off_pairs))

    def ac_tell(self, variant, cur, cur_f, prop, fx, accepted, acc_count, mu,
                sigma, log_lambda, samples, stride, sample_offset, gamma,
                target, seed, offset, require_finite=True):
        n, d = cur.shape
        self._call(self.lib.pb200_ac_tell, int(variant),
                   1 if require_finite else 0, _ptr(cur), _ptr(cur_f),
                   _ptr(prop), _ptr(fx), _ptr(accepted), _ptr(acc_count),
                   _ptr(mu), _ptr(sigma), _ptr(log_lambda),
                   _ptr(samples) if samples is not None else None,
                   int(stride), int(sample_offset), n, d, float(gamma),
                   float(target), ctypes.c_uint64(seed),
                   ctypes.c_uint64(offset))

    # the same launches with per-iteration scalars from a device table
    def ac_ask_it(self, cur, sigma, log_lambda, prop, seed, rows, counter):
        n, d = cur.shape
        self._call(self.lib.pb200_ac_ask_it, _ptr(cur), _ptr(sigma),
                   _ptr(log_lambda), _ptr(prop), n, d, ctypes.c_uint64(seed),
                   _ptr(rows), _ptr(counter))

    def de_ask_it(self, cur_all, first, b_star, prop, seed, rows, counter):
        n, d = prop.shape
        self._call(self.lib.pb200_de_ask_it, _ptr(cur_all), int(first),
                   cur_all.shape[0], _ptr(b_star), _ptr(prop), n, d,
                   ctypes.c_uint64(seed), _ptr(rows), _ptr(counter))

    def ac_tell_it(self, variant, cur, cur_f, prop, fx, accepted, acc_count,
                   mu, sigma, log_lambda, samples, stride, target, seed, rows,
                   counter, require_finite=True):
        n, d = cur.shape
        self._call(self.lib.pb200_ac_tell_it, int(variant),
                   1 if require_finite else 0, _ptr(cur), _ptr(cur_f),
                   _ptr(prop), _ptr(fx), _ptr(accepted), _ptr(acc_count),
                   _ptr(mu), _ptr(sigma), _ptr(log_lambda), _ptr(samples),
                   int(stride), n, d, float(target), ctypes.c_uint64(seed),
                   _ptr(rows), _ptr(counter))

    def iter_advance(self, counter):
        self._call(self.lib.pb200_iter_advance, _ptr(counter))

    def chain_run(self, problem, cur, cur_f, prop, fx, accepted, acc_count,
                  mu, sigma, log_lambda, samples, stride, target, seed, rows,
                  first_row, n_iterations, require_finite=True):
        n, d = cur.shape
        self._call(self.lib.pb200_mcmc_chain_run, problem._handle,
                   1 if require_finite else 0, _ptr(cur), _ptr(cur_f),
                   _ptr(prop), _ptr(fx), _ptr(accepted), _ptr(acc_count),
                   _ptr(mu), _ptr(sigma), _ptr(log_lambda), _ptr(samples),
                   int(stride), n, d, float(target), ctypes.c_uint64(seed),
                   _ptr(rows), int(first_row), int(n_iterations))

    def hbac_adapt(self, cur, accepted, mu, sigma, log_lambda, gamma, target):
        n, d = cur.shape
        self._call(self.lib.pb200_hbac_adapt, _ptr(cur), _ptr(accepted),
                   _ptr(mu), _ptr(sigma), _ptr(log_lambda), n, d,
                   float(gamma), float(target))

    def hbac_propose(self, cur, sigma, log_lambda, z, prop):
        n, d = cur.shape
        self._call(self.lib.pb200_hbac_propose, _ptr(cur), _ptr(sigma),
                   _ptr(log_lambda), _ptr(z), _ptr(prop), n, d)

    def de_propose(self, cur_all, first, eps, r1, r2, prop, gamma):
        n, d = prop.shape
        self._call(self.lib.pb200_de_propose, _ptr(cur_all), int(first),
                   _ptr(eps), _ptr(r1), _ptr(r2), _ptr(prop), n, d,
                   float(gamma))

    def normal(self, out, seed, offset):
        self._call(self.lib.pb200_philox_normal, _ptr(out), out.numel(),
                   ctypes.c_uint64(seed), ctypes.c_uint64(offset))

    def uniform(self, out, seed, offset):
        self._call(self.lib.pb200_philox_uniform, _ptr(out), out.numel(),
                   ctypes.c_uint64(seed), ctypes.c_uint64(offset))

    def pairs(self, r1, r2, first, n_total, seed, offset):
        self._call(self.lib.pb200_philox_pairs, _ptr(r1), _ptr(r2),
                   r1.numel(), int(first), int(n_total),
                   ctypes.c_uint64(seed), ctypes.c_uint64(offset))


class _GraphIterationsBase(object):
    """See ``_GraphIterations`` below (kept first so that both sampler
    families can inherit from it)."""

    def graph_ready(self):
        return (getattr(self, '_fused', False) and self._rng_mode == 'philox'
                and self._have_current and not self._asked)


class _Philox(object):
    """Bookkeeping of the counter space: every draw gets a fresh offset."""

    def __init__(self, seed):
        self.seed = int(seed) & (2**64 - 1)
        self.offset = 0

    def take(self, n):
        off = self.offset
        self.offset += (int(n) + 1) // 2 + 1
        return off


class HaarioBardenetACMC(_GraphIterationsBase):
    """All chains of the Haario-Bardenet adaptive covariance sampler
    (pints/_mcmc/_haario_bardenet_ac.py).

    Parameters
    ----------
    x0 : (n_chains, d) starting points
    sigma0 : None, (d,), (d, d) or (n_chains, d, d)
        Initial proposal covariance; ``None`` gives the reference's default
        ``diag(0.01 |x0|)`` per chain (zeros replaced by one,
        pints/_mcmc/__init__.py:84-91).
    rng : 'philox' (device draws) or 'replay' (the caller supplies proposals
        and log-uniforms to ``ask``/``tell``; used for bit-exact parity tests
        and for reproducing a NumPy-seeded reference run).
    """

    _VARIANT = _lib.AC_HAARIO_BARDENET

    def __init__(self, x0, sigma0=None, device=None, rng='philox', seed=0,
                 eta=0.6, target_acceptance=0.234, fused=True):
        # fused: with device draws, one launch per ask() and one per tell()
        # (pb200_ac_ask / pb200_ac_tell) instead of eight small ones; the
        # chains are identical either way
        self._fused = bool(fused)
        x0 = np.array(x0, dtype=np.float64, copy=True)
        if x0.ndim != 2:
            raise ValueError('x0 must have shape (n_chains, n_parameters).')
        self._n, self._d = x0.shape
        if self._d > _lib.MAX_PARAMS:
            raise ValueError('At most %d parameters are supported.'
                             % _lib.MAX_PARAMS)
        if rng not in ('philox', 'replay'):
            raise ValueError("rng must be 'philox' or 'replay'")
        self._ops = _DeviceOps(device)
        dev = self._ops.device
        self._rng_mode = rng
        self._rng = _Philox(seed)
        n, d = self._n, self._d
        if sigma0 is None:
            s = np.abs(x0)
            s[s == 0] = 1
            sigma = np.zeros((n, d, d))
            sigma[:, np.arange(d), np.arange(d)] = 0.01 * s
        else:
            s = np.array(sigma0, dtype=np.float64)
            if s.shape == (n, d, d):
                sigma = s.copy()
            elif s.size == d:
                sigma = np.tile(np.diag(s.reshape(d)), (n, 1, 1))
            else:
                sigma = np.tile(s.reshape((d, d)), (n, 1, 1))
        self._x0 = _f64(x0, dev)
        self._sigma = _f64(sigma, dev)
        self._mu = self._x0.clone()
        self._log_lambda = torch.zeros(n, dtype=torch.float64, device=dev)
        self._current = torch.empty_like(self._x0)
        self._current_f = torch.empty(n, dtype=torch.float64, device=dev)
        self._proposed = torch.empty_like(self._x0)
        self._z = torch.empty_like(self._x0)
        self._log_u = torch.empty(n, dtype=torch.float64, device=dev)
        self._log_ratio = torch.zeros(n, dtype=torch.float64, device=dev)
        self._previous = torch.empty_like(self._x0)
        self._accepted = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._eta = float(eta)
        self._target = float(target_acceptance)
        self._adaptive = False
        self._adaptations = 1
        self._gamma = 1.0
        self._iterations = 0
        self._acceptance_count = torch.zeros(n, dtype=torch.int64, device=dev)
        self._started = False
        self._have_current = False
        self._asked = False

    # -- reference-shaped accessors -----------------------------------
    def n_chains(self):
        return self._n

    def n_parameters(self):
        return self._d

    def needs_initial_phase(self):
        return True

    def in_initial_phase(self):
        return not self._adaptive

    def set_initial_phase(self, initial_phase):
        self._adaptive = not bool(initial_phase)

    def eta(self):
        return self._eta

    def set_eta(self, eta):
        if eta <= 0:
            raise ValueError('eta should be greater than zero')
        self._eta = float(eta)

    def target_acceptance_rate(self):
        return self._target

    def set_target_acceptance_rate(self, rate=0.234):
        rate = float(rate)
        if rate <= 0:
            raise ValueError('Target acceptance rate must be greater than 0.')
        if rate > 1:
            raise ValueError('Target acceptance rate cannot exceed 1.')
        self._target = rate

    def acceptance_rate(self):
        """Per-chain measured acceptance rate (host array)."""
        if self._iterations == 0:
            return np.zeros(self._n)
        return self._acceptance_count.cpu().numpy() / float(self._iterations)

    def current(self):
        return self._current

    def current_log_pdfs(self):
        return self._current_f

    def state(self):
        """Host copies of (mu, sigma, log_lambda), for tests and diagnostics."""
        return (self._mu.cpu().numpy(), self._sigma.cpu().numpy(),
                self._log_lambda.cpu().numpy())

    # -- graph replay (see CudaMCMCController) --------------------------
    def plan_iterations(self, k, samples, first_it, switch_at=None):
        """Rows of the device table for iterations ``first_it ..
        first_it + k - 1`` (``(k, 8)`` int64), advancing the host bookkeeping
        as ``k`` eager ask/tell rounds would; the initial phase ends at
        iteration ``switch_at``."""
        rows = np.zeros((k, 8), dtype=np.int64)
        step = samples.stride(1)
        for i in range(k):
            it = first_it + i
            if switch_at is not None and it == switch_at:
                self.set_initial_phase(False)
            rows[i, 0] = self._rng.take(self._z.numel())
            rows[i, 2] = self._rng.take(self._n)
            variant = -1
            if self._adaptive and self._VARIANT is not None:
                self._gamma = (self._adaptations + 1) ** -self._eta
                self._adaptations += 1
                variant = self._VARIANT
            rows[i, 3] = it * step
            rows[i, 5] = _f64_bits(self._gamma)
            rows[i, 6] = variant
            self._iterations += 1
        return rows

    def enqueue_ask_it(self, rows, counter, all_current=None):
        self._ops.ac_ask_it(self._current, self._sigma, self._log_lambda,
                            self._proposed, self._rng.seed, rows, counter)
        return self._proposed

    def enqueue_chain_loop(self, problem, fx, samples, rows, k):
        """The ``k`` planned iterations of ``rows`` in ONE launch: a persistent
        warp per chain proposes, scores the proposal on ``problem`` (a
        closed-form ``DeviceProblem``) and accepts / adapts / stores
        (``pb200_mcmc_chain_run``).  Same chains as ``k`` rounds of
        ``enqueue_ask_it`` / ``problem.eval`` / ``enqueue_tell_it``."""
        self._ops.chain_run(problem, self._current, self._current_f,
                            self._proposed, fx, self._accepted,
                            self._acceptance_count, self._mu, self._sigma,
                            self._log_lambda, samples, samples.stride(0),
                            self._target, self._rng.seed, rows, 0, k)

    def enqueue_tell_it(self, fx, samples, rows, counter):
        variant = self._VARIANT if self._VARIANT is not None else -1
        self._ops.ac_tell_it(variant, self._current, self._current_f,
                             self._proposed, fx, self._accepted,
                             self._acceptance_count, self._mu, self._sigma,
                             self._log_lambda, samples, samples.stride(0),
                             self._target, self._rng.seed, rows, counter)

    @classmethod
    def name(cls):
        return 'Haario-Bardenet adaptive covariance MCMC'

    # -- ask / tell -----------------------------------------------------
    def ask(self, proposed=None):
        """Proposals for all chains (device tensor ``(n_chains, d)``).  In
        replay mode pass the proposals drawn elsewhere."""
        if not self._started:
            self._started = True
            self._proposed.copy_(self._x0)
        elif not self._asked:
            if proposed is not None:
                self._proposed.copy_(torch.as_tensor(proposed).to(
                    self._proposed.device))
            elif self._rng_mode == 'replay':
                raise RuntimeError('replay mode: ask() needs the proposals')
            elif self._fused:
                off = self._rng.take(self._z.numel())
                self._ops.ac_ask(self._current, self._sigma, self._log_lambda,
                                 self._proposed, self._rng.seed, off)
            else:
                off = self._rng.take(self._z.numel())
                self._ops.normal(self._z, self._rng.seed, off)
                self._ops.hbac_propose(self._current, self._sigma,
                                       self._log_lambda, self._z,
                                       self._proposed)
        self._asked = True
        return self._proposed

    def tell(self, fx, log_u=None, store=None):
        """Accept/reject and adapt.  ``fx`` holds the log-pdfs of the proposals
        (device tensor or array).  Returns ``(current, current_log_pdfs,
        accepted)`` as device tensors.  ``store = (samples, it)`` also writes
        the new current points to ``samples[:, it]`` (a contiguous device
        tensor ``(n_chains, n_iterations, d)``)."""
        if not self._asked:
            raise RuntimeError('Tell called before proposal was set.')
        self._asked = False
        fx = torch.as_tensor(fx)
        if fx.device != self._current.device or fx.dtype != torch.float64:
            fx = fx.to(device=self._current.device, dtype=torch.float64)
        self._iterations += 1
        if not self._have_current:
            if not bool(torch.isfinite(fx).all()):
                raise ValueError(
                    'Initial point for MCMC must have finite logpdf.')
            self._current.copy_(self._proposed)
            self._current_f.copy_(fx)
            self._accepted.fill_(1)
            self._have_current = True
            if store is not None:
                store[0][:, store[1]].copy_(self._current)
            return self._current, self._current_f, self._accepted
        if log_u is None and self._rng_mode != 'replay' and self._fused:
            # one launch: log-uniform, accept, count, adapt, store
            off = self._rng.take(self._n)
            variant = -1
            if self._adaptive and self._VARIANT is not None:
                self._gamma = (self._adaptations + 1) ** -self._eta
                self._adaptations += 1
                variant = self._VARIANT
            samples, stride, soff = None, 0, 0
            if store is not None:
                samples = store[0]
                stride, soff = samples.stride(0), store[1] * samples.stride(1)
            self._ops.ac_tell(variant, self._current, self._current_f,
                              self._proposed, fx, self._accepted,
                              self._acceptance_count, self._mu, self._sigma,
                              self._log_lambda, samples, stride, soff,
                              self._gamma, self._target, self._rng.seed, off)
            return self._current, self._current_f, self._accepted
        if log_u is not None:
            self._log_u.copy_(torch.as_tensor(log_u).to(self._log_u.device))
        elif self._rng_mode == 'replay':
            raise RuntimeError('replay mode: tell() needs log_u')
        else:
            off = self._rng.take(self._n)
            self._ops.uniform(self._log_u, self._rng.seed, off)
            torch.log_(self._log_u)
        v = self._VARIANT
        if v == _lib.AC_HAARIO_BARDENET or v is None:
            self._ops.mh_accept(self._current, self._current_f, self._proposed,
                                fx, self._log_u, self._accepted, True)
        else:
            if v == _lib.AC_RAO_BLACKWELL:
                self._previous.copy_(self._current)
            self._ops.mh_accept_ratio(
                self._current, self._current_f, self._proposed, fx,
                self._log_u, self._accepted, self._log_ratio, True)
        self._acceptance_count += self._accepted
        if self._adaptive and v is not None:
            # gamma is the same for every chain: computed on the host exactly as
            # the reference does (pints/_mcmc/_adaptive_covariance.py:289)
            self._gamma = (self._adaptations + 1) ** -self._eta
            self._adaptations += 1
            if v == _lib.AC_HAARIO_BARDENET:
                self._ops.hbac_adapt(self._current, self._accepted, self._mu,
                                     self._sigma, self._log_lambda,
                                     self._gamma, self._target)
            else:
                self._ops.ac_adapt(
                    v, self._current, self._previous, self._proposed,
                    self._accepted, self._log_ratio, self._mu, self._sigma,
                    self._log_lambda, self._gamma, self._target)
        if store is not None:
            store[0][:, store[1]].copy_(self._current)
        return self._current, self._current_f, self._accepted


def _f64_bits(v):
    return int(np.float64(v).view(np.int64))


class HaarioACMC(HaarioBardenetACMC):
    """All chains of the Haario adaptive covariance sampler
    (pints/_mcmc/_haario_ac.py): as Haario-Bardenet, but log lambda moves with
    the acceptance probability ``min(1, exp(log_ratio))`` instead of the 0/1
    outcome."""
    _VARIANT = _lib.AC_HAARIO

    @classmethod
    def name(cls):
        return 'Haario adaptive covariance MCMC'


class RaoBlackwellACMC(HaarioBardenetACMC):
    """All chains of the Rao-Blackwell adaptive covariance sampler
    (pints/_mcmc/_rao_blackwell_ac.py): proposals from ``lambda sigma`` with the
    fixed ``lambda = 2.38^2 / d``; sigma adapts towards
    ``p * proposed + (1 - p) * previous``."""
    _VARIANT = _lib.AC_RAO_BLACKWELL

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        # device proposals take exp(log lambda) * sigma: a constant here
        self._lambda = (2.38 ** 2) / self._d
        self._log_lambda.fill_(float(np.log(self._lambda)))

    def state(self):
        mu, sigma, _ = super().state()
        return mu, sigma, np.zeros(self._n)

    @classmethod
    def name(cls):
        return 'Rao-Blackwell adaptive covariance MCMC'


class MetropolisRandomWalkMCMC(HaarioBardenetACMC):
    """All chains of the Metropolis random-walk sampler
    (pints/_mcmc/_metropolis.py): proposals from the fixed ``sigma0``, the same
    accept rule, no adaptation and no initial phase."""
    _VARIANT = None

    def needs_initial_phase(self):
        return False

    def in_initial_phase(self):
        return False

    def set_initial_phase(self, initial_phase):
        raise NotImplementedError

    def acceptance_rate(self):
        # the reference counts from the second tell on
        # (pints/_mcmc/_metropolis.py:146-180)
        if self._iterations <= 1:
            return np.zeros(self._n)
        return self._acceptance_count.cpu().numpy() / float(self._iterations - 1)

    @classmethod
    def name(cls):
        return 'Metropolis random walk MCMC'


class DifferentialEvolutionMCMC(_GraphIterationsBase):
    """Differential-evolution MCMC with the reference's default settings
    (Gaussian error, relative scaling, ``gamma = 2.38 / sqrt(2 d)`` and
    ``gamma = 1`` on every tenth proposal round).

    ``first`` / ``n_total`` describe a shard of a larger chain set: this object
    owns chains ``[first, first + n)`` of ``n_total``; ``ask`` then needs the
    positions of all chains (``all_current``), which the caller all-gathers.
    """

    def __init__(self, chains, x0, device=None, rng='philox', seed=0,
                 first=0, n_total=None, x0_mean=None, fused=True):
        self._fused = bool(fused)      # one launch per ask() / tell()
        x0 = np.array(x0, dtype=np.float64, copy=True)
        if x0.ndim != 2 or x0.shape[0] != int(chains):
            raise ValueError(
                'Number of initial positions must be equal to number of'
                ' chains.')
        self._n, self._d = x0.shape
        self._first = int(first)
        self._n_total = int(n_total) if n_total is not None else self._n
        if self._n_total < 3:
            raise ValueError('Need at least 3 chains.')
        if rng not in ('philox', 'replay'):
            raise ValueError("rng must be 'philox' or 'replay'")
        self._ops = _DeviceOps(device)
        dev = self._ops.device
        self._rng_mode = rng
        self._rng = _Philox(seed)
        self._gamma_default = 2.38 / np.sqrt(2 * self._d)
        self._gamma_switch_rate = 10
        self._b = 0.001
        # scale of the error process: |mean(x0) * b| over ALL chains
        mean = np.mean(x0, axis=0) if x0_mean is None else np.asarray(x0_mean)
        self._b_star = np.abs(mean * self._b)
        self._d_b_star = _f64(self._b_star, dev)
        self._x0 = _f64(x0, dev)
        n, d = self._n, self._d
        self._current = torch.empty_like(self._x0)
        self._current_f = torch.empty(n, dtype=torch.float64, device=dev)
        self._proposed = torch.empty_like(self._x0)
        self._eps = torch.empty_like(self._x0)
        self._r1 = torch.empty(n, dtype=torch.int32, device=dev)
        self._r2 = torch.empty(n, dtype=torch.int32, device=dev)
        self._log_u = torch.empty(n, dtype=torch.float64, device=dev)
        self._accepted = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._iter_count = 0
        self._started = False
        self._have_current = False
        self._asked = False

    def n_chains(self):
        return self._n

    def n_parameters(self):
        return self._d

    def needs_initial_phase(self):
        return False

    def gamma(self):
        return self._gamma_default

    def b_star(self):
        return self._b_star

    def current(self):
        return self._current

    def current_log_pdfs(self):
        return self._current_f

    @classmethod
    def name(cls):
        return 'Differential Evolution MCMC'

    # -- graph replay (see CudaMCMCController) --------------------------
    def plan_iterations(self, k, samples, first_it, switch_at=None):
        """As ``HaarioBardenetACMC.plan_iterations``: error-draw, partner and
        uniform offsets, the gamma schedule (1 on every tenth round)."""
        rows = np.zeros((k, 8), dtype=np.int64)
        step = samples.stride(1)
        for i in range(k):
            gamma = self.next_gamma()
            self._iter_count += 1
            rows[i, 0] = self._rng.take(self._eps.numel())
            rows[i, 1] = self._rng.take(2 * self._n)
            rows[i, 2] = self._rng.take(self._n)
            rows[i, 3] = (first_it + i) * step
            rows[i, 4] = _f64_bits(gamma)
            rows[i, 6] = -1
        return rows

    def enqueue_ask_it(self, rows, counter, all_current=None):
        pool = self._current if all_current is None else all_current
        self._ops.de_ask_it(pool, self._first, self._d_b_star, self._proposed,
                            self._rng.seed, rows, counter)
        return self._proposed

    def enqueue_tell_it(self, fx, samples, rows, counter):
        self._ops.ac_tell_it(-1, self._current, self._current_f,
                             self._proposed, fx, self._accepted, None, None,
                             None, None, samples, samples.stride(0), 0.0,
                             self._rng.seed, rows, counter,
                             require_finite=False)

    def next_gamma(self):
        """The gamma the coming proposal round uses
        (pints/_mcmc/_differential_evolution.py:98-101,118)."""
        if self._iter_count % self._gamma_switch_rate == 0:
            return 1.0
        return float(self._gamma_default)

    def ask(self, eps=None, r1=None, r2=None, all_current=None):
        """Proposals ``x_j + gamma (x_r1 - x_r2) + eps_j`` for the owned chains.
        Replay mode: pass ``eps`` (n, d) and ``r1``/``r2`` (n,) drawn elsewhere.
        Sharded: pass ``all_current`` (n_total, d), the gathered positions."""
        if not self._started:
            self._started = True
            self._proposed.copy_(self._x0)
            self._asked = True
            return self._proposed
        if not self._asked:
            gamma = self.next_gamma()
            self._iter_count += 1
            dev = self._current.device
            if eps is not None:
                self._eps.copy_(torch.as_tensor(eps).to(dev))
                self._r1.copy_(torch.as_tensor(np.asarray(r1, np.int32)).to(dev))
                self._r2.copy_(torch.as_tensor(np.asarray(r2, np.int32)).to(dev))
            elif self._rng_mode == 'replay':
                raise RuntimeError('replay mode: ask() needs eps, r1 and r2')
            elif self._fused:
                off_eps = self._rng.take(self._eps.numel())
                off_pairs = self._rng.take(2 * self._n)
                pool = self._current if all_current is None else all_current
                if pool.shape[0] != self._n_total:
                    raise ValueError('all_current must hold all %d chains'
                                     % self._n_total)
                self._ops.de_ask(pool, self._first, self._d_b_star,
                                 self._proposed, gamma, self._rng.seed,
                                 off_eps, off_pairs)
                self._asked = True
                return self._proposed
            else:
                off = self._rng.take(self._eps.numel())
                self._ops.normal(self._eps, self._rng.seed, off)
                self._eps.mul_(self._d_b_star)
                off = self._rng.take(2 * self._n)
                self._ops.pairs(self._r1, self._r2, self._first,
                                self._n_total, self._rng.seed, off)
            pool = self._current if all_current is None else all_current
            if pool.shape[0] != self._n_total:
                raise ValueError('all_current must hold all %d chains'
                                 % self._n_total)
            self._ops.de_propose(pool, self._first, self._eps, self._r1,
                                 self._r2, self._proposed, gamma)
        self._asked = True
        return self._proposed

    def tell(self, fx, log_u=None, store=None):
        if not self._asked:
            raise RuntimeError('Tell called before proposal was set.')
        self._asked = False
        fx = torch.as_tensor(fx)
        if fx.device != self._current.device or fx.dtype != torch.float64:
            fx = fx.to(device=self._current.device, dtype=torch.float64)
        if not self._have_current:
            if not bool(torch.isfinite(fx).all()):
                raise ValueError(
                    'Initial points for MCMC must have finite logpdf.')
            self._current.copy_(self._proposed)
            self._current_f.copy_(fx)
            self._accepted.fill_(1)
            self._have_current = True
            if store is not None:
                store[0][:, store[1]].copy_(self._current)
            return self._current, self._current_f, self._accepted
        if log_u is None and self._rng_mode != 'replay' and self._fused:
            off = self._rng.take(self._n)
            samples, stride, soff = None, 0, 0
            if store is not None:
                samples = store[0]
                stride, soff = samples.stride(0), store[1] * samples.stride(1)
            self._ops.ac_tell(-1, self._current, self._current_f,
                              self._proposed, fx, self._accepted, None, None,
                              None, None, samples, stride, soff, 0.0, 0.0,
                              self._rng.seed, off, require_finite=False)
            return self._current, self._current_f, self._accepted
        if log_u is not None:
            self._log_u.copy_(torch.as_tensor(log_u).to(self._log_u.device))
        elif self._rng_mode == 'replay':
            raise RuntimeError('replay mode: tell() needs log_u')
        else:
            off = self._rng.take(self._n)
            self._ops.uniform(self._log_u, self._rng.seed, off)
            torch.log_(self._log_u)
        self._ops.mh_accept(self._current, self._current_f, self._proposed, fx,
                            self._log_u, self._accepted, False)
        if store is not None:
            store[0][:, store[1]].copy_(self._current)
        return self._current, self._current_f, self._accepted
