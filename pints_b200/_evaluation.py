"""
``CudaEvaluator``: the drop-in for ``pints.SequentialEvaluator`` /
``pints.ParallelEvaluator`` (pints/_evaluation.py:63-123, :126-413, :457-482)
that scores the whole batch of positions in one fused CUDA launch.
"""
import collections.abc
import ctypes

import numpy as np
import torch

from . import _lib
from ._compat import base
from ._engine import DeviceProblem, positions_to_array, require_cuda
from ._spec import ProblemSpec, describe


class Evaluator(base('Evaluator')):
    """Mirror of ``pints.Evaluator`` (pints/_evaluation.py:63-123): same
    constructor checks, same ``evaluate(positions)`` contract.  When the
    reference is importable this class also *is* a ``pints.Evaluator``."""

    def __init__(self, function, args=None):
        if not callable(function):
            raise ValueError('The given function must be callable.')
        self._function = function
        if args is None:
            self._args = ()
        else:
            try:
                len(args)
            except TypeError:
                raise ValueError(
                    'The argument `args` must be either None or a sequence.')
            self._args = args

    def evaluate(self, positions):
        """Evaluates the function for every value in ``positions`` and returns
        the evaluations (a list, as the reference does)."""
        try:
            len(positions)
        except TypeError:
            raise ValueError(
                'The argument `positions` must be a sequence of input values'
                ' to the evaluator\'s function.')
        return self._evaluate(positions)

    def _evaluate(self, positions):
        raise NotImplementedError


class ScoreList(collections.abc.Sequence):
    """What ``evaluate`` returns by default: a read-only, list-compatible view
    of the scores as they came back from the GPU (a float64 ndarray), without
    boxing every score into a Python float up front -- ``tolist()`` of 65 536
    scores costs more than solving them.  It supports what the reference's
    consumers do with the list: ``len(fs)``, ``fs[i]``
    (a Python ``float``), slices, ``iter(fs)`` / ``next`` (boxed lazily, on
    first iteration: pints/_mcmc/__init__.py:733-735), ``np.argsort(fs)``,
    ``np.array(fs)``, ``full[ids] = fs`` (pints/_optimisers/_xnes.py:154-160),
    ``fs == [..]``, ``fs + [..]``, ``min`` / ``max`` / ``sorted``;
    ``fs.tolist()`` is the plain list."""

    __slots__ = ('_a', '_boxed')

    def __init__(self, array):
        self._a = array
        self._boxed = None

    def tolist(self):
        if self._boxed is None:
            self._boxed = self._a.tolist()
        return self._boxed

    def __array__(self, dtype=None, copy=None):
        a = self._a
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __len__(self):
        return self._a.shape[0]

    def __getitem__(self, k):
        if isinstance(k, slice):
            return ScoreList(self._a[k])
        try:
            return self._a.item(k)
        except (TypeError, ValueError):
            return self._a[k]                 # index arrays: an ndarray

    def __iter__(self):
        return iter(self.tolist())

    def __reversed__(self):
        return reversed(self.tolist())

    def __contains__(self, v):
        return v in self.tolist()

    def index(self, *args):
        return self.tolist().index(*args)

    def count(self, v):
        return self.tolist().count(v)

    def __eq__(self, other):
        if isinstance(other, ScoreList):
            other = other._a
        if isinstance(other, (list, tuple, np.ndarray)):
            other = np.asarray(other)
            return other.shape == self._a.shape and bool(
                np.all((other == self._a) | ((other != other) & (self._a != self._a))))
        return NotImplemented

    def __ne__(self, other):
        r = self.__eq__(other)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __add__(self, other):
        return self.tolist() + list(other)

    def __radd__(self, other):
        return list(other) + self.tolist()

    def __repr__(self):
        return repr(self.tolist())


class _SpecFunction(object):
    """Stands in for ``function`` when a CudaEvaluator is built directly from
    a ``ProblemSpec``: there is no Python callable to fall back to."""

    def __init__(self, spec):
        self.spec = spec

    def __call__(self, x):
        raise RuntimeError('a ProblemSpec is evaluated through CudaEvaluator, '
                           'not by calling it')


class _Lane(object):
    """Per-device state of a CudaEvaluator: problem handle, streams, a pinned
    host staging buffer and device workspaces (grown geometrically)."""

    N_STREAMS = 4

    def __init__(self, spec, device, opts, sort=None):
        self.device = device
        self.problem = DeviceProblem(spec, device, sort=sort, **opts)
        self.streams = [torch.cuda.Stream(device=device)
                        for _ in range(self.N_STREAMS)]
        self.stream_ptrs = [ctypes.c_void_p(s.cuda_stream)
                            for s in self.streams]
        self.used = 0
        self.capacity = 0
        self.width = 0
        self.d_steps = None

    def reserve(self, n, d, want_steps):
        if n > self.capacity or d != self.width:
            cap = max(n, int(self.capacity * 1.5), 1024)
            self.h_in = torch.empty((cap, d), dtype=torch.float64,
                                    pin_memory=True)
            self.d_in = torch.empty((cap, d), dtype=torch.float64,
                                    device=self.device)
            self.d_out = torch.empty(cap, dtype=torch.float64,
                                     device=self.device)
            # raw pointers, so a call does no tensor slicing
            self.a_h_in = self.h_in.data_ptr()
            self.a_d_in, self.a_d_out = self.d_in.data_ptr(), self.d_out.data_ptr()
            self.d_steps = None
            self.capacity, self.width = cap, d
        if want_steps and self.d_steps is None:
            self.d_steps = torch.zeros(self.capacity, dtype=torch.int32,
                                       device=self.device)
            self.a_d_steps = self.d_steps.data_ptr()

    #: rows staged into pinned memory (and sent) per slice
    STAGE_ROWS = 32768

    def send(self, lib, xs, lo, hi, st, pinned):
        """Rows ``[lo, hi)`` of the host array ``xs`` to the same rows of
        ``d_in`` on stream ``st``.  Page-locked input is sent from where it
        lies; anything else is staged through the pinned buffer in slices of
        STAGE_ROWS rows by the library's copy threads (``pb200_host_copy``),
        each slice on its way while the next one is staged."""
        V = ctypes.c_void_p
        d = xs.shape[1]
        if pinned:
            _lib.check(lib.pb200_copy_async(
                V(self.a_d_in + lo * d * 8), V(xs.ctypes.data + lo * d * 8),
                (hi - lo) * d * 8, 1, st))
            return
        for a in range(lo, hi, self.STAGE_ROWS):
            b = min(hi, a + self.STAGE_ROWS)
            _lib.check(lib.pb200_host_copy(
                V(self.a_h_in + a * d * 8), V(xs.ctypes.data + a * d * 8),
                (b - a) * d * 8, 0))
            _lib.check(lib.pb200_copy_async(
                V(self.a_d_in + a * d * 8), V(self.a_h_in + a * d * 8),
                (b - a) * d * 8, 1, st))

    @staticmethod
    def is_pinned(lib, xs):
        return bool(xs.flags['C_CONTIGUOUS'] and xs.size >= 4096 and
                    lib.pb200_host_is_pinned(ctypes.c_void_p(xs.ctypes.data)))

    def launch(self, lib, xs, a_scores, a_steps, chunk_rows):
        """Enqueue copy-in, kernel and copy-out for the rows of ``xs``; the
        scores (and step counts, if ``a_steps``) go to the page-locked host
        addresses given.  The batch is cut into at most N_STREAMS compute
        chunks of ``chunk_rows`` rows or more, one fused launch each on its
        own stream."""
        m, d = xs.shape
        want_steps = a_steps is not None
        self.reserve(m, d, want_steps)
        V = ctypes.c_void_p
        n_chunks = max(1, min(len(self.streams), m // max(1, chunk_rows)))
        rows = -(-m // n_chunks)
        self.used = 0
        handle, opts = self.problem._handle, ctypes.byref(self.problem.opts)
        pinned = self.is_pinned(lib, xs)
        with torch.cuda.device(self.device):
            for c in range(n_chunks):
                lo, hi = c * rows, min(m, (c + 1) * rows)
                if hi <= lo:
                    break
                st = self.stream_ptrs[c]
                self.send(lib, xs, lo, hi, st, pinned)
                ws, ws_bytes = None, 0
                if self.problem.wants_sort(hi - lo):
                    t = self.problem.sort_workspace(rows, ('chunk', c))
                    ws, ws_bytes = V(t.data_ptr()), t.numel()
                _lib.check(lib.pb200_eval_sorted(
                    handle, V(self.a_d_in + lo * d * 8), hi - lo, d,
                    V(self.a_d_out + lo * 8),
                    V(self.a_d_steps + lo * 4) if want_steps else None,
                    opts, ws, ws_bytes, st))
                _lib.check(lib.pb200_copy_async(
                    V(a_scores + lo * 8), V(self.a_d_out + lo * 8),
                    (hi - lo) * 8, 0, st))
                if want_steps:
                    _lib.check(lib.pb200_copy_async(
                        V(a_steps + lo * 4), V(self.a_d_steps + lo * 4),
                        (hi - lo) * 4, 0, st))
                self.used = c + 1

    def wait(self):
        for s in self.streams[:self.used]:
            s.synchronize()


class _AddTerm(object):
    """Last item of a pending evaluation: once the lanes before it have been
    waited for, adds the host-side term of a ``TransformedLogPDF`` (the log
    Jacobian determinant, pints/_transformation.py:1166-1169) to the scores."""

    def __init__(self, scores, term):
        self.scores, self.term = scores, term

    def wait(self):
        np.add(self.scores, self.term, out=self.scores)


def pinned_result(n, dtype=torch.float64):
    """A fresh 1-d array in page-locked host memory from torch's caching host
    allocator: the device-to-host copy of a result lands in the array that is
    handed to the caller, with no copy-out; the block goes back to the cache
    when the caller drops the array."""
    return torch.empty(n, dtype=dtype, pin_memory=True).numpy()


def pinned_empty(shape):
    """A float64 ndarray in page-locked host memory.  Positions written into
    such an array (e.g. by a vectorised ``ask()``) are sent to the GPU by
    :class:`CudaEvaluator` from where they lie, without the staging copy that
    ordinary (pageable) arrays need."""
    t = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    return t.numpy()


class CudaEvaluator(Evaluator):
    """
    Evaluates a supported objective for a whole list of positions on the GPU.

    Use it wherever the reference uses ``SequentialEvaluator(f)`` or
    ``ParallelEvaluator(f)``: ``CudaEvaluator(f).evaluate(xs)`` returns the same
    list of ``f(x)`` values, computed as one batched forward solve plus
    likelihood reduction (see ``include/pints_b200.h``: ``pb200_eval``).

    Parameters
    ----------
    function
        ``SumOfSquaresError``, ``GaussianKnownSigmaLogLikelihood`` or
        ``GaussianLogLikelihood`` over a Single/MultiOutputProblem on one of the
        supported toy models -- optionally inside ``LogPosterior`` (uniform
        prior) and/or ``ProbabilityBasedError`` -- from the reference or from
        this package; or a ready ``ProblemSpec``.  Anything else raises
        ``TypeError``: there is no CPU fallback.
    args
        Must be empty (the supported objectives take no extra arguments).
    devices
        CUDA device(s) to use.  With several devices the rows are split into
        contiguous shards, one per device, launched concurrently.
    rtol, atol, max_steps, h0, block
        Controls of the adaptive Dormand-Prince 5(4) integrator for the ODE
        models; the tolerances default to scipy odeint's 1.49012e-8 that the
        reference runs with (pints/toy/_toy_classes.py:225-233).
    lanes_per_vector
        GPU lanes that integrate one parameter vector: 1, or 2 (one lane per
        state component, two-state models only); default automatic.
    wide_interpolation
        ODE models: take the observations that pile up inside long steps four
        at a time (pays when steps span many observations); ``None`` =
        automatic per model, ``True`` / ``False`` force it.  Results do not
        depend on it.
    sort
        Cost-coherent scheduling of the ODE models (vectors with similar
        dynamics share a warp; results do not depend on it): ``True``,
        ``False`` or ``None`` = automatic (on for batches of 1024 rows or more).
    as_list
        ``True`` (default): return a :class:`ScoreList`, a list-compatible
        read-only sequence over the scores as they came back from the GPU
        (boxed into Python floats lazily; ``.tolist()`` is the plain list);
        ``'boxed'``: a plain Python ``list`` of floats, exactly the
        reference's return type; ``False``: a float64 ndarray.
    chunk_rows
        Host batches of several times this many rows are cut into up to four
        compute chunks (one fused launch each, on separate streams), so that
        staging of a later chunk overlaps with the solve of an earlier one.
        Within a chunk the rows are staged and sent in slices of 16384.

    Unlike ``ParallelEvaluator`` (pints/_evaluation.py:310) this evaluator never
    draws from ``numpy.random``, so controller trajectories equal those of a
    ``SequentialEvaluator`` run for the same seed.
    """

    def __init__(self, function, args=None, devices=None, rtol=None, atol=None,
                 max_steps=None, h0=None, block=None, as_list=True,
                 chunk_rows=65536, lanes_per_vector=None, sort=None,
                 wide_interpolation=None):
        if isinstance(function, ProblemSpec):
            spec = function
            function = _SpecFunction(spec)
        else:
            spec = None
        super().__init__(function, args)
        if len(self._args) != 0:
            raise ValueError(
                'CudaEvaluator does not support extra function arguments.')
        self._lib = _lib.load()                # fails now if the library is absent
        self._spec = spec if spec is not None else describe(function)
        self._opts = dict(rtol=rtol, atol=atol, max_steps=max_steps, h0=h0,
                          block=block, lanes=lanes_per_vector,
                          wide_interpolation=wide_interpolation)
        self._sort = sort
        self._as_list = as_list if as_list == 'boxed' else bool(as_list)
        self._chunk_rows = max(1, int(chunk_rows))
        if devices is None or isinstance(devices, (int, str, torch.device)):
            devices = [devices]
        self._device_args = list(devices)
        self._lanes = None
        self.last_steps = None

    # ------------------------------------------------------------------
    def spec(self):
        """The ``ProblemSpec`` derived from the function."""
        return self._spec

    def n_parameters(self):
        return self._spec.n_parameters()

    def _setup(self):
        if self._lanes is None:
            devs = [require_cuda(d) for d in self._device_args]
            self._lanes = [_Lane(self._spec, d, self._opts, self._sort)
                           for d in devs]
        return self._lanes

    def device_problem(self, index=0):
        """The ``DeviceProblem`` behind this evaluator, for callers that keep
        their positions on the device.  Its inputs are model-space
        parameters: an objective wrapped in a transformation can only be
        scored through :meth:`evaluate`, which maps the rows first."""
        if self._spec.transform is not None:
            raise TypeError(
                'the objective is wrapped in a transformation: device-resident '
                'callers work in model space; use CudaEvaluator.evaluate')
        return self._setup()[index].problem

    # ------------------------------------------------------------------
    def evaluate_array(self, positions, return_steps=False):
        """Like :meth:`evaluate` but returns a float64 ndarray (and, on
        request, the attempted RK steps per vector)."""
        try:
            len(positions)
        except TypeError:
            raise ValueError(
                'The argument `positions` must be a sequence of input values'
                ' to the evaluator\'s function.')
        scores, steps = self.finish(self.begin(positions, return_steps))
        if return_steps:
            self.last_steps = steps
            return scores, steps
        return scores

    def begin(self, positions, return_steps=False):
        """Enqueues the evaluation of ``positions`` (copy-in, kernels,
        copy-out) without waiting for it; :meth:`finish` waits and returns the
        result.  One evaluation may be in flight per evaluator; evaluators of
        different functions run on streams of their own, so several can be in
        flight at once."""
        d = self._spec.n_parameters()
        xs = positions_to_array(positions, d)
        n = xs.shape[0]
        if n == 0:
            return (np.empty(0, dtype=np.float64),
                    np.zeros(0, dtype=np.int32) if return_steps else None, [])
        lanes = self._setup()          # raises when there is no GPU
        jacobian = None
        if self._spec.transform is not None:
            # positions arrive in the search space of the controller's
            # transformation: map the rows to model space on the host
            from ._transform import rows_log_jacobian_det, rows_to_model
            transformation, add_jacobian, outer_sign = self._spec.transform
            if add_jacobian:
                jacobian = outer_sign * rows_log_jacobian_det(transformation, xs)
            xs = rows_to_model(transformation, xs)
        # results land in page-locked arrays that are handed to the caller
        scores = pinned_result(n)
        steps = pinned_result(n, torch.int32) if return_steps else None
        g = len(lanes)
        per = (n + g - 1) // g
        work = []
        for k, lane in enumerate(lanes):
            lo, hi = k * per, min(n, (k + 1) * per)
            if hi <= lo:
                continue
            lane.launch(self._lib, xs[lo:hi], scores.ctypes.data + lo * 8,
                        steps.ctypes.data + lo * 4 if return_steps else None,
                        self._chunk_rows)
            work.append(lane)
        if jacobian is not None:
            work.append(_AddTerm(scores, jacobian))
        return scores, steps, work

    @staticmethod
    def finish(pending):
        scores, steps, work = pending
        for lane in work:
            lane.wait()
        return scores, steps

    def _evaluate(self, positions):
        return wrap_scores(self.evaluate_array(positions), self._as_list)


def wrap_scores(scores, as_list):
    """The return type of ``evaluate`` for the ``as_list`` option."""
    if as_list == 'boxed':
        return scores.tolist()
    return ScoreList(scores) if as_list else scores
