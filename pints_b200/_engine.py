"""
Device side of a problem: uploads the packed series, owns the
``pb200_problem`` handle, and launches the fused kernels on torch-owned device
buffers.  PyTorch is used for device memory, streams and events only.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._spec import ProblemSpec

DEFAULT_RTOL = 1.49012e-8       # scipy odeint's default, used by the reference
DEFAULT_ATOL = 1.49012e-8


def require_cuda(device=None):
    """The torch device to run on; raises when there is no GPU (there is no
    CPU implementation behind this package)."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            'pints_b200 needs a CUDA device (built for NVIDIA B200, sm_100a); '
            'no GPU is visible and there is no CPU fallback.')
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    device = torch.device(device)
    if device.type != 'cuda':
        raise RuntimeError('pints_b200 runs on CUDA devices only, got %s'
                           % device)
    if device.index is None:
        device = torch.device('cuda', torch.cuda.current_device())
    return device


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def capture_graph(device, enqueue):
    """Captures the launches ``enqueue()`` makes (on the stream that is
    current while it runs) in a CUDA graph and returns it; replay with
    ``graph.replay()`` on the caller's current stream.

    ``torch.cuda.graph`` is not used on purpose: on entry it runs the garbage
    collector and ``torch.cuda.empty_cache()``, which in a process holding
    gigabytes of cached blocks costs 60 ms and more -- longer than the loops
    captured here run.  ``capture_begin`` alone puts the allocator in
    graph-private-pool mode, which is all the capture needs."""
    current = torch.cuda.current_stream(device)
    side = torch.cuda.Stream(device)
    side.wait_stream(current)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        graph.capture_begin()
        try:
            enqueue()
        except BaseException:
            try:
                graph.capture_end()
            except Exception:
                pass
            raise
        graph.capture_end()
    current.wait_stream(side)
    return graph


def solver_opts(rtol=None, atol=None, h0=None, max_steps=None, block=None,
                lanes=None, wide_interpolation=None):
    # 0 = the model's default (1.49012e-8; Lotka-Volterra 5e-9), chosen by the
    # library (pb200_api.cu: default_tolerance)
    return _lib.SolverOpts(
        rtol if rtol else 0.0, atol if atol else 0.0,
        h0 if h0 else 0.0, int(max_steps) if max_steps else 0,
        int(block) if block else 0, int(lanes) if lanes else 0,
        0 if wide_interpolation is None else (1 if wide_interpolation else -1))


class DeviceProblem(object):
    """A ``ProblemSpec`` resident on one GPU."""

    #: populations at least this large are sorted by cost before the launch
    SORT_MIN_ROWS = 1024

    def __init__(self, spec, device=None, sort=None, **opts):
        if not isinstance(spec, ProblemSpec):
            raise TypeError('spec must be a ProblemSpec')
        self.spec = spec
        self.device = require_cuda(device)
        self._lib = _lib.load()
        self.opts = solver_opts(**opts)
        # cost-coherent scheduling (pb200_eval_sorted): None = automatic
        self.sort = sort
        self._sort_ws = {}
        with torch.cuda.device(self.device):
            self._series = torch.from_numpy(spec.packed_series()).to(
                self.device)
            mc, n_mc = _lib.doubles(spec.model_consts)
            oc, n_oc = _lib.doubles(spec.obj_consts)
            lo = hi = None
            n_prior = 0
            if spec.prior_lower is not None:
                lo, n_prior = _lib.doubles(spec.prior_lower)
                hi, _ = _lib.doubles(spec.prior_upper)
            handle = ctypes.c_void_p()
            _lib.check(self._lib.pb200_problem_create(
                spec.model, spec.objective, spec.n_times, spec.n_outputs,
                _ptr(self._series), mc, n_mc, oc, n_oc, spec.sign, lo, hi,
                n_prior, spec.log_prior, ctypes.byref(handle)))
        self._handle = handle
        if spec.prior_terms:
            n = len(spec.prior_terms)
            kinds = (ctypes.c_int32 * n)(*[int(t[0]) for t in spec.prior_terms])
            index = (ctypes.c_int32 * n)(*[int(t[1]) for t in spec.prior_terms])
            flat = [float(v) for t in spec.prior_terms for v in t[2:5]]
            consts = (ctypes.c_double * (3 * n))(*flat)
            _lib.check(self._lib.pb200_problem_set_prior_terms(
                handle, n, kinds, index, consts))
        self.n_parameters = spec.n_parameters()
        self.n_model_parameters = spec.n_model_parameters()

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h:
            try:
                self._lib.pb200_problem_destroy(h)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------
    def _check_params(self, d_params, width):
        if d_params.dtype != torch.float64 or d_params.dim() != 2:
            raise ValueError('parameters must be a 2-d float64 tensor')
        if d_params.device != self.device:
            raise ValueError('parameters live on %s, problem on %s'
                             % (d_params.device, self.device))
        if d_params.shape[1] < width:
            raise ValueError(
                'Expected %d parameters per vector, got %d.'
                % (width, d_params.shape[1]))
        if d_params.stride(1) != 1:
            d_params = d_params.contiguous()
        return d_params

    def wants_sort(self, n):
        """Whether a batch of ``n`` rows is scheduled through the cost sort."""
        if not self.spec.is_ode():
            return False
        if self.sort is None:
            return n >= self.SORT_MIN_ROWS
        return bool(self.sort)

    def sort_hint(self):
        """What the last sorted launch saw (``pb200_problem_sort_hint``): 0
        nothing yet, 1 costs close together (the next one-block-per-SM launch sorts
        inside the evaluation kernel), 2 wide (sort kernel)."""
        return int(self._lib.pb200_problem_sort_hint(self._handle))

    def sort_workspace(self, n, key=0):
        """Device workspace of ``pb200_sort_workspace_bytes(n)`` bytes, cached
        per ``key`` (one per stream that may run concurrently)."""
        need = int(self._lib.pb200_sort_workspace_bytes(n))
        ws = self._sort_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, int(1.5 * need) if ws is not None else need),
                             dtype=torch.uint8, device=self.device)
            self._sort_ws[key] = ws
        return ws

    def eval(self, d_params, out=None, steps=None):
        """Scores of the rows of ``d_params`` (device tensor ``(n, d)``);
        asynchronous on the current stream.  Returns a device tensor."""
        d_params = self._check_params(d_params, self.n_parameters)
        n = d_params.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            ws, ws_bytes = None, 0
            if self.wants_sort(n):
                t = self.sort_workspace(n, stream)
                ws, ws_bytes = ctypes.c_void_p(t.data_ptr()), t.numel()
            _lib.check(self._lib.pb200_eval_sorted(
                self._handle, _ptr(d_params), n, d_params.stride(0),
                _ptr(out), _ptr(steps), ctypes.byref(self.opts),
                ws, ws_bytes, ctypes.c_void_p(stream)))
        return out

    def eval_exchange(self, d_params, desc, steps=None):
        """Scores of the rows of ``d_params`` (this rank's block of a multi-GPU
        job; may be empty) stored straight into the staging arrays of all
        ranks, followed by this rank's flags: ``pb200_eval_exchange`` with the
        ``_lib.ExchangeDesc`` ``desc``.  Asynchronous on the current stream;
        :func:`exchange_finish` is the reading side."""
        n = d_params.shape[0]
        if n:
            d_params = self._check_params(d_params, self.n_parameters)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            ws, ws_bytes = None, 0
            if n and self.wants_sort(n):
                t = self.sort_workspace(n, stream)
                ws, ws_bytes = ctypes.c_void_p(t.data_ptr()), t.numel()
            _lib.check(self._lib.pb200_eval_exchange(
                self._handle, _ptr(d_params) if n else None, n,
                d_params.stride(0) if n else self.n_parameters,
                _ptr(steps), ctypes.byref(self.opts), ws, ws_bytes,
                ctypes.byref(desc), ctypes.c_void_p(stream)))

    def simulate(self, d_params, out=None, steps=None):
        """Trajectories ``(n, n_times, n_outputs)`` of the rows of
        ``d_params`` (only the model parameters are read)."""
        d_params = self._check_params(d_params, self.n_model_parameters)
        n = d_params.shape[0]
        if out is None:
            out = torch.empty((n, self.spec.n_times, self.spec.n_outputs),
                              dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pb200_simulate(
                self._handle, _ptr(d_params), n, d_params.stride(0),
                _ptr(out), _ptr(steps), ctypes.byref(self.opts),
                _stream_ptr(self.device)))
        return out

    def score_trajectories(self, d_traj, d_params, out=None):
        """Unfused baseline, second half: objective from stored
        trajectories."""
        d_params = self._check_params(d_params, self.n_parameters)
        n = d_params.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.pb200_score_trajectories(
                self._handle, _ptr(d_traj), _ptr(d_params), n,
                d_params.stride(0), _ptr(out), _stream_ptr(self.device)))
        return out

    def new_steps(self, n):
        return torch.zeros(n, dtype=torch.int32, device=self.device)


def exchange_finish(desc, rows_per_rank, n_total, out):
    """Reading side of the fused exchange: waits for every rank's flag of
    generation ``desc.epoch`` and writes the ``n_total`` scores in row order
    into ``out`` (device tensor): ``pb200_exchange_finish``.  Asynchronous on
    the current stream of ``out``'s device."""
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.pb200_exchange_finish(
            ctypes.byref(desc), int(rows_per_rank), int(n_total), _ptr(out),
            _stream_ptr(out.device)))
    return out


def positions_to_array(positions, width=None):
    """Evaluator input -> contiguous float64 ``(n, d)`` host array.  Accepts
    what the reference's controllers pass: a 2-d (read-only) ndarray or a list
    of 1-d arrays (pints/_mcmc/__init__.py:718)."""
    if isinstance(positions, np.ndarray) and positions.ndim == 2:
        xs = np.ascontiguousarray(positions, dtype=np.float64)
    else:
        n = len(positions)
        if n == 0:
            return np.zeros((0, width or 0), dtype=np.float64)
        xs = np.array([np.asarray(x, dtype=np.float64).reshape(-1)
                       for x in positions], dtype=np.float64)
        if xs.ndim != 2:
            raise ValueError('All positions must have the same length.')
    if width is not None and xs.shape[0] > 0 and xs.shape[1] != width:
        raise ValueError('Expected positions with %d parameters, got %d.'
                         % (width, xs.shape[1]))
    return xs


def fp64_probe(device=None, iters=20000, blocks_per_sm=4, threads=256,
               repeats=3):
    """Measured FP64 FMA throughput (TFLOP/s) of the device: the roofline
    denominator for the fused ODE kernel (MEASURED_PEAKS.json has none)."""
    device = require_cuda(device)
    lib = _lib.load()
    props = torch.cuda.get_device_properties(device)
    blocks = props.multi_processor_count * blocks_per_sm
    sink = torch.zeros(8, dtype=torch.float64, device=device)
    best = 0.0
    with torch.cuda.device(device):
        for r in range(repeats + 1):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.pb200_fp64_probe(_ptr(sink), blocks, threads,
                                            iters, _stream_ptr(device)))
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            flops = 16.0 * iters * blocks * threads
            if r > 0:
                best = max(best, flops / (ms * 1e-3) / 1e12)
    return best
