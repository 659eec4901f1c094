"""
``CudaEvaluator``: the drop-in for ``pints.SequentialEvaluator`` /
``pints.ParallelEvaluator`` (pints/_evaluation.py:63-123, :126-413, :457-482)
that scores the whole batch of positions in one fused CUDA launch.
"""
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


class _SpecFunction(object):
    """Stands in for ``function`` when a CudaEvaluator is built directly from
    a ``ProblemSpec``: there is no Python callable to fall back to."""

    def __init__(self, spec):
        self.spec = spec

    def __call__(self, x):
        raise RuntimeError('a ProblemSpec is evaluated through CudaEvaluator, '
                           'not by calling it')


class _Lane(object):
    """Per-device state of a CudaEvaluator: problem handle + staging."""

    def __init__(self, spec, device, opts):
        self.device = device
        self.problem = DeviceProblem(spec, device, **opts)
        self.stream = torch.cuda.Stream(device=device)
        self.capacity = 0
        self.h_in = self.h_out = self.d_in = self.d_out = self.d_steps = None
        self.h_steps = None

    def reserve(self, n, d, want_steps):
        if n > self.capacity:
            cap = max(n, int(self.capacity * 1.5), 1024)
            self.h_in = torch.empty((cap, d), dtype=torch.float64,
                                    pin_memory=True)
            self.h_out = torch.empty(cap, dtype=torch.float64, pin_memory=True)
            self.d_in = torch.empty((cap, d), dtype=torch.float64,
                                    device=self.device)
            self.d_out = torch.empty(cap, dtype=torch.float64,
                                     device=self.device)
            self.d_steps = self.h_steps = None
            self.capacity = cap
        if want_steps and self.d_steps is None:
            self.d_steps = torch.zeros(self.capacity, dtype=torch.int32,
                                       device=self.device)
            self.h_steps = torch.zeros(self.capacity, dtype=torch.int32,
                                       pin_memory=True)


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
    as_list
        Return a Python list like the reference (default) or, if ``False``, a
        float64 ndarray (every consumer in the reference accepts either).

    Unlike ``ParallelEvaluator`` (pints/_evaluation.py:310) this evaluator never
    draws from ``numpy.random``, so controller trajectories equal those of a
    ``SequentialEvaluator`` run for the same seed.
    """

    def __init__(self, function, args=None, devices=None, rtol=None, atol=None,
                 max_steps=None, h0=None, block=None, as_list=True):
        if isinstance(function, ProblemSpec):
            spec = function
            function = _SpecFunction(spec)
        else:
            spec = None
        super().__init__(function, args)
        if len(self._args) != 0:
            raise ValueError(
                'CudaEvaluator does not support extra function arguments.')
        _lib.load()                            # fail now if the library is absent
        self._spec = spec if spec is not None else describe(function)
        self._opts = dict(rtol=rtol, atol=atol, max_steps=max_steps, h0=h0,
                          block=block)
        self._as_list = bool(as_list)
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
            self._lanes = [_Lane(self._spec, d, self._opts) for d in devs]
        return self._lanes

    def device_problem(self, index=0):
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
        d = self._spec.n_parameters()
        xs = positions_to_array(positions, d)
        n = xs.shape[0]
        scores = np.empty(n, dtype=np.float64)
        steps = np.zeros(n, dtype=np.int32) if return_steps else None
        if n == 0:
            return (scores, steps) if return_steps else scores
        lanes = self._setup()
        g = len(lanes)
        per = (n + g - 1) // g
        work = []
        for k, lane in enumerate(lanes):
            lo, hi = k * per, min(n, (k + 1) * per)
            if hi <= lo:
                continue
            m = hi - lo
            lane.reserve(m, d, return_steps)
            lane.h_in[:m].numpy()[...] = xs[lo:hi]
            with torch.cuda.device(lane.device), torch.cuda.stream(lane.stream):
                lane.d_in[:m].copy_(lane.h_in[:m], non_blocking=True)
                lane.problem.eval(lane.d_in[:m], out=lane.d_out[:m],
                                  steps=lane.d_steps if return_steps else None)
                lane.h_out[:m].copy_(lane.d_out[:m], non_blocking=True)
                if return_steps:
                    lane.h_steps[:m].copy_(lane.d_steps[:m], non_blocking=True)
            work.append((lane, lo, hi, m))
        for lane, lo, hi, m in work:
            lane.stream.synchronize()
            scores[lo:hi] = lane.h_out[:m].numpy()
            if return_steps:
                steps[lo:hi] = lane.h_steps[:m].numpy()
        if return_steps:
            self.last_steps = steps
            return scores, steps
        return scores

    def _evaluate(self, positions):
        scores = self.evaluate_array(positions)
        return scores.tolist() if self._as_list else scores
