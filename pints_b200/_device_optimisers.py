"""
Device-resident optimiser: the caller of the evaluation path at population
sizes where the host update itself becomes the bottleneck.

``DeviceXNES`` is the exponential natural evolution strategy of
``pints.XNES`` (pints/_optimisers/_xnes.py:63-91 ask, :147-182 tell) with the
population kept in HBM: ``ask()`` returns the ``(population, d)`` candidate
points as a device tensor, ``tell(fx)`` takes their scores as a device tensor.
At population 65 536 the vectorised host ``XNES`` of this package needs 12 ms
per generation (6 ms of them NumPy's Mersenne-Twister normals, which identity
with the reference's random stream requires) against 0.4 ms for scoring the
population; here a generation's update costs a sort and four small kernels,
the d x d part (``expm``) included (``pb200_xnes_update``), so the host only
reads results back -- every generation, or once per block of generations
(``generations``).

Same hyper-parameters, utilities and update equations as the reference.  The
samples come from the counter-based Philox generator of the C ABI instead of
``numpy.random`` -- a different stream, so trajectories agree with the reference
statistically, while *given the same normals and scores* the update is the
reference's (``tests/test_gpu_samplers.py::test_device_xnes_replays_host_xnes``
feeds the device's normals to the host ``XNES`` through ``numpy.random``).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._engine import _ptr, _stream_ptr, capture_graph, require_cuda
from ._util import vector

_SUM_BLOCKS = 148          # one block of ranks per SM


class DeviceArgsort(object):
    """Ascending argsort of ``n`` scores on the device (``pb200_argsort_f64``:
    sample sort, NaN last, ties by index) with its workspace.  Measured on
    B200 against the library sort (``tools/argsort_time.py``): 53 / 62 / 82 /
    137 us against 53 / 116 / 120 / 137 us at n = 4096 / 16 384 / 65 536 /
    262 144; beyond ``PB200_ARGSORT_MAX`` the library sort is used."""

    MAX = 262144

    def __init__(self, n, device, force_native=False):
        self.n = int(n)
        self.device = require_cuda(device)
        self._lib = _lib.load()
        self.native = self.n <= self.MAX
        # starts as the identity: a sort that gives up (status = 1) leaves its
        # slice untouched, and what reads `order` next on the stream must find
        # valid indices until the host has seen the status and sorted again
        self.order = torch.arange(self.n, dtype=torch.int64, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        nbytes = int(self._lib.pb200_argsort_workspace_bytes(self.n))
        self._ws = torch.empty(max(nbytes, 16), dtype=torch.uint8,
                               device=self.device)

    def __call__(self, values):
        """Enqueues the sort of ``values`` (device tensor ``(n,)``, float64,
        contiguous) on the current stream; returns the permutation (int64
        device tensor, overwritten by the next call).  ``status`` is non-zero
        afterwards if the native sort gave up (see ``pb200_argsort_f64``)."""
        if not self.native:
            return torch.sort(values)[1]
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pb200_argsort_f64(
                _ptr(values), _ptr(self.order), self.n, _ptr(self._ws),
                self._ws.numel(), _ptr(self.status),
                _stream_ptr(self.device)))
        return self.order


class DeviceXNES(object):
    """xNES with the population on the GPU.

    Parameters as ``pints.XNES`` plus ``seed`` (Philox key) and ``device``.
    ``boundaries`` must be ``RectangularBoundaries`` (of this package or of the
    reference); a sampled point outside them is never scored -- its row of the
    tensor ``ask()`` returns holds the current mean instead and its score is
    replaced by ``+inf`` in ``tell()``, which ranks it exactly as the reference
    ranks the points it withholds from the evaluator
    (pints/_optimisers/_xnes.py:79-87, :154-157)."""

    device_resident = True
    use_graph = True
    #: generations the device keeps [f_best, x_best, f_guessed] of (the most
    #: one call of ``generations`` can run between two synchronisations)
    HISTORY = 64
    #: rank with pb200_argsort_f64 (False: the library sort)
    native_sort = True

    def __init__(self, x0, sigma0=None, boundaries=None, seed=1, device=None):
        self._x0 = vector(x0)
        self._d = len(self._x0)
        if self._d < 1:
            raise ValueError('Problem dimension must be greater than zero.')
        if self._d > _lib.MAX_PARAMS:
            raise ValueError('DeviceXNES supports at most %d parameters.'
                             % _lib.MAX_PARAMS)
        self._boundaries = boundaries
        if boundaries is not None:
            if boundaries.n_parameters() != self._d:
                raise ValueError(
                    'Boundaries must have same dimension as starting point.')
            if not boundaries.check(self._x0):
                raise ValueError(
                    'Initial position must lie within given boundaries.')
            if not (hasattr(boundaries, 'lower') and hasattr(boundaries, 'upper')):
                raise TypeError('DeviceXNES needs RectangularBoundaries.')
        if sigma0 is None:
            # pints/_optimisers/__init__.py:113-126
            try:
                self._sigma0 = (1 / 6) * boundaries.range()
            except AttributeError:
                self._sigma0 = (1 / 3) * np.abs(self._x0)
                self._sigma0 = self._sigma0 + (self._sigma0 == 0)
        elif np.isscalar(sigma0):
            if float(sigma0) <= 0:
                raise ValueError(
                    'Initial standard deviation must be greater than zero.')
            self._sigma0 = np.ones(self._d) * float(sigma0)
        else:
            self._sigma0 = vector(sigma0)
            if len(self._sigma0) != self._d:
                raise ValueError(
                    'Initial standard deviation must be None, scalar, or have'
                    ' dimension ' + str(self._d) + '.')
            if np.any(self._sigma0 <= 0):
                raise ValueError(
                    'Initial standard deviations must be greater than zero.')
        self._population_size = 4 + int(3 * np.log(self._d))
        self._seed = int(seed) & (2 ** 64 - 1)
        self._offset = 0
        self._device_arg = device
        self._running = False
        self._ready_for_tell = False
        self._x_best = np.array(self._x0, copy=True)
        self._f_best = np.inf
        self._f_guessed = np.inf

    # -- pints.PopulationBasedOptimiser interface -----------------------
    @classmethod
    def name(cls):
        return 'Exponential Natural Evolution Strategy (xNES), device resident'

    def population_size(self):
        return self._population_size

    def set_population_size(self, population_size=None):
        if self._running:
            raise Exception('Cannot change population size during run.')
        if population_size is None:
            population_size = 4 + int(3 * np.log(self._d))
        population_size = int(population_size)
        if population_size < 1:
            raise ValueError('Population size must be at least 1.')
        self._population_size = population_size

    def suggested_population_size(self, round_up_to_multiple_of=None):
        p = 4 + int(3 * np.log(self._d))
        if round_up_to_multiple_of is not None:
            n = int(round_up_to_multiple_of)
            if n > 1:
                p = n * (((p - 1) // n) + 1)
        return p

    def running(self):
        return self._running

    def needs_sensitivities(self):
        return False

    def stop(self):
        return False

    def f_best(self):
        return self._f_best

    def f_guessed(self):
        return self._f_guessed

    def x_best(self):
        return self._x_best

    def x_guessed(self):
        return self._mu if self._running else self._x0

    # -- state -----------------------------------------------------------
    def _initialise(self):
        n, d = self._population_size, self._d
        self._device = require_cuda(self._device_arg)
        self._lib = _lib.load()
        dev = self._device
        # Table 1 of Glasmachers et al. 2010 (pints/_optimisers/_xnes.py:93-121)
        self._eta_mu = 1
        self._eta_A = 0.6 * (3 + np.log(d)) * d ** -1.5
        us = np.maximum(0, np.log(n / 2 + 1) - np.log(1 + np.arange(n)))
        us /= np.sum(us)
        us -= 1 / n
        self._sum_us = float(np.sum(us))
        self._mu = np.array(self._x0, copy=True)
        self._A = self._initial_A()
        self._I = np.eye(d)
        f64 = dict(dtype=torch.float64, device=dev)
        self._d_us = torch.from_numpy(us).to(dev)
        # mu, A, Philox offset: lives on the device, updated there
        self._d_state = torch.empty(d + d * d + 1, **f64)
        self._d_z = torch.empty((n, d), **f64)
        self._d_x = torch.empty((n, d), **f64)
        self._d_x_eval = torch.empty((n, d), **f64)
        self._d_inside = torch.empty(n, dtype=torch.uint8, device=dev)
        self._d_inside_bool = self._d_inside.view(torch.bool)
        self._d_inf = torch.full((), float('inf'), **f64)
        self._d_partial = torch.empty(_SUM_BLOCKS * (d + d * d), **f64)
        # sums, then the best point of the generation and its score
        self._d_out = torch.empty(d + d * d + d + 1, **f64)
        # [f_best, x_best, f_guessed, generations]; per-generation history
        self._d_best = torch.empty(d + 3, **f64)
        self._d_hist = torch.empty((self.HISTORY, d + 2), **f64)
        self._h_state = torch.empty(d + d * d + 1, dtype=torch.float64,
                                    pin_memory=True)
        self._h_best = torch.empty(d + 3, dtype=torch.float64, pin_memory=True)
        self._h_hist = torch.empty((self.HISTORY, d + 2), dtype=torch.float64,
                                   pin_memory=True)
        self._offset_step = (n * d + 1) // 2 + 1
        h = self._h_state.numpy()
        h[:d] = self._mu
        h[d:d + d * d] = self._A.reshape(-1)
        h[d + d * d:].view(np.uint64)[0] = self._offset
        hb = self._h_best.numpy()
        hb[0] = np.inf
        hb[1:1 + d] = self._x0
        hb[d + 1] = np.inf
        hb[d + 2] = 0.0
        with torch.cuda.device(dev):
            self._d_state.copy_(self._h_state, non_blocking=True)
            self._d_best.copy_(self._h_best, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        self._generations_done = 0
        self._graphs = {}             # (problem, sign, cost sort of the launch) -> captured generation
        self._graph_failed = None
        self._eager_generations = 0
        self._sorter = DeviceArgsort(n, dev)
        self._sort_gave_up = False
        self._h_status = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        self._d_lower = self._d_upper = None
        if self._boundaries is not None:
            self._d_lower = torch.from_numpy(np.array(
                self._boundaries.lower(), dtype=np.float64)).to(dev)
            self._d_upper = torch.from_numpy(np.array(
                self._boundaries.upper(), dtype=np.float64)).to(dev)
        self._running = True

    def normals(self):
        """The standard normals of the last ``ask()``, ``(population, d)``."""
        return self._d_z

    def points(self):
        """The sampled points of the last ``ask()`` (rows outside the
        boundaries included), ``(population, d)``."""
        return self._d_x

    def inside(self):
        """Which of them lie inside the boundaries (uint8)."""
        return self._d_inside

    def _initial_A(self):
        return np.eye(self._d) * self._sigma0

    # -- ask / tell ------------------------------------------------------
    def _enqueue_ask(self):
        n, d = self._population_size, self._d
        state = self._d_state.data_ptr()
        _lib.check(self._lib.pb200_xnes_ask(
            ctypes.c_void_p(state), ctypes.c_void_p(state + 8 * d),
            _ptr(self._d_lower), _ptr(self._d_upper), _ptr(self._d_z),
            _ptr(self._d_x), _ptr(self._d_x_eval), _ptr(self._d_inside),
            n, d, ctypes.c_uint64(self._seed),
            ctypes.c_void_p(state + 8 * (d + d * d)),
            _stream_ptr(self._device)))

    #: 1 = the diagonal update of SNES
    _separable = 0

    def _eta_matrix(self):
        return self._eta_A

    def _enqueue_tell(self, fx, read_back=True):
        n, d = self._population_size, self._d
        self._last_fx = fx
        if self._d_lower is not None:
            fx = torch.where(self._d_inside_bool, fx, self._d_inf)
        # ranks (NaN scores rank last): the sample sort of the C ABI, or the
        # library sort for populations beyond its reach
        order = self._sorter(fx.contiguous()) if self.native_sort \
            and self._sorter.native and not self._sort_gave_up \
            else torch.sort(fx)[1]
        stream = _stream_ptr(self._device)
        _lib.check(self._lib.pb200_xnes_sums(
            _ptr(self._d_z), _ptr(order), _ptr(self._d_us),
            _ptr(self._d_partial), _SUM_BLOCKS, _ptr(self._d_out),
            _ptr(self._d_x), _ptr(fx), n, d, stream))
        # the d x d update, the best point so far, the next Philox offset
        # (pints/_optimisers/_xnes.py:160-182) -- on the device as well
        _lib.check(self._lib.pb200_xnes_update(
            _ptr(self._d_state), _ptr(self._d_out), _ptr(self._d_best),
            _ptr(self._d_hist), self.HISTORY, float(self._eta_mu),
            float(self._eta_matrix()), self._sum_us, d,
            ctypes.c_uint64(self._offset_step), self._separable, stream))
        if read_back:
            self._enqueue_read_back()

    def _enqueue_read_back(self):
        self._h_state.copy_(self._d_state, non_blocking=True)
        self._h_best.copy_(self._d_best, non_blocking=True)
        self._h_status.copy_(self._sorter.status, non_blocking=True)

    def _update(self):
        """After a synchronisation: the host's copies of what the device has
        computed (mean, A, best point, scores)."""
        d = self._d
        h = self._h_state.numpy()
        self._mu = np.array(h[:d], copy=True)
        self._A = np.array(h[d:d + d * d], copy=True).reshape((d, d))
        hb = self._h_best.numpy()
        self._f_best = float(hb[0])
        if np.isfinite(self._f_best) or self._f_best == -np.inf:
            self._x_best = np.array(hb[1:1 + d], copy=True)
        self._f_guessed = float(hb[d + 1])
        self._generations_done = int(hb[d + 2])

    def _snapshot(self):
        return self._d_state.clone(), self._d_best.clone()

    def _restore(self, snap):
        self._d_state.copy_(snap[0])
        self._d_best.copy_(snap[1])

    def _sort_failed(self):
        """True if the sample sort met a bucket it cannot hold (never seen; see
        ``pb200_argsort_f64``): the library sort takes over for the rest of the
        run, and the caller repeats the generations concerned from its snapshot
        of the state."""
        if int(self._h_status[0]) == 0:
            return False
        self._sort_gave_up = True
        self._graphs = {}
        self._sorter.status.zero_()
        return True

    def ask(self):
        if not self._running:
            self._initialise()
        self._ready_for_tell = True
        with torch.cuda.device(self._device):
            self._enqueue_ask()
        return self._d_x_eval

    def tell(self, fx):
        """``fx``: device tensor ``(population,)`` of the scores of the rows
        ``ask()`` returned (to be minimised)."""
        if not self._ready_for_tell:
            raise Exception('ask() not called before tell()')
        self._ready_for_tell = False
        with torch.cuda.device(self._device):
            snap = self._snapshot()
            self._enqueue_tell(fx)
            torch.cuda.current_stream(self._device).synchronize()
            if self._sort_failed():
                self._restore(snap)
                self._enqueue_tell(fx)
                torch.cuda.current_stream(self._device).synchronize()
        self._update()

    def generation(self, problem, negate=False):
        """One whole generation -- ask, score on ``problem`` (a
        ``DeviceProblem``), tell -- with a single host synchronisation.
        ``negate``: maximise the scores."""
        self.generations(problem, 1, negate)

    def generations(self, problem, count, negate=False):
        """``count`` (at most ``HISTORY``) whole generations with ONE host
        synchronisation at the end: nothing in a generation needs the host
        (sampling, scoring, ranking, the weighted sums and the d x d update
        are all on the device).  After two eager generations the device work
        of a generation (sampling kernel, sort + solve kernels, ranking, sums,
        update: about ten launches) is captured in a CUDA graph and replayed;
        set ``use_graph = False`` to keep it eager.  Returns the history of the
        block, an array ``(count, d + 2)`` of ``[f_best, x_best, f_guessed]``
        after each generation."""
        if not self._running:
            self._initialise()
        count = int(count)
        if not 1 <= count <= self.HISTORY:
            raise ValueError('count must be between 1 and %d' % self.HISTORY)
        key = (id(problem), bool(negate))
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device)
            snap = self._snapshot()
            first = self._generations_done
            for attempt in range(2):
                for _ in range(count):
                    self._one_generation(problem, negate, key)
                self._enqueue_read_back()
                self._h_hist.copy_(self._d_hist, non_blocking=True)
                stream.synchronize()
                if not self._sort_failed():
                    break
                self._restore(snap)       # and again, with the library sort
        self._update()
        rows = [(first + k) % self.HISTORY for k in range(count)]
        return np.array(self._h_hist.numpy()[rows], copy=True)

    @property
    def _graph(self):
        """A captured generation (any of them), or None."""
        return next(iter(self._graphs.values()), None)

    def _one_generation(self, problem, negate, key):
        # which cost sort the evaluation launch takes is decided when it is
        # ENQUEUED (from the spread of the previous population's keys,
        # pb200_problem_sort_hint), i.e. it is baked into a captured graph: one
        # graph per choice, so that a population that has narrowed since the
        # first generations moves to the sort inside the kernel
        hint = getattr(problem, 'sort_hint', None)
        key = key + (hint() == 1 if hint is not None else False,)
        graph = self._graphs.get(key)
        if graph is not None:
            graph.replay()
        elif self.use_graph and self._eager_generations >= 2 \
                and self._graph_failed is None:
            try:
                graph = capture_graph(
                    self._device,
                    lambda: self._enqueue_generation(problem, negate))
                self._graphs[key] = graph
                graph.replay()
            except Exception as e:          # keep going without a graph
                self._graph_failed = e
                self._graphs = {}
                torch.cuda.synchronize(self._device)
                self._enqueue_generation(problem, negate)
        else:
            self._enqueue_generation(problem, negate)
            self._eager_generations += 1

    def _enqueue_generation(self, problem, negate):
        self._enqueue_ask()
        fx = problem.eval(self._d_x_eval)
        self._enqueue_tell(-fx if negate else fx, read_back=False)


class DeviceSNES(DeviceXNES):
    """Separable natural evolution strategy (``pints.SNES``,
    pints/_optimisers/_snes.py:56-166) with the population on the GPU: the
    same kernels as :class:`DeviceXNES` with a diagonal ``A``; the update uses
    ``sum_r u_r s_r`` and the diagonal of ``sum_r u_r s_r s_r^T``."""

    @classmethod
    def name(cls):
        return 'Seperable Natural Evolution Strategy (SNES), device resident'

    _separable = 1

    def _initialise(self):
        d = self._d
        self._eta_sigmas = 0.2 * (3 + np.log(d)) * d ** -0.5
        super()._initialise()

    def _initial_A(self):
        return np.diag(self._sigma0)

    def _eta_matrix(self):
        return self._eta_sigmas

    def _update(self):
        # pints/_optimisers/_snes.py:150-166, computed by pb200_xnes_update
        super()._update()
        self._sigmas = np.array(np.diagonal(self._A), copy=True)
