"""Small host helpers shared by the mirror classes."""
import contextlib

import numpy as np


def vector(x):
    """Copy ``x`` into a read-only 1-d float array, like ``pints.vector``
    (pints/_util.py:77-95)."""
    if np.isscalar(x):
        x = np.array([float(x)])
    else:
        x = np.array(x, copy=True, dtype=float)
    if x.ndim != 1:
        n = np.max(x.shape) if x.size else 0
        if np.prod(x.shape) != n:
            raise ValueError(
                'Unable to convert to 1d vector of scalar values.')
        x = x.reshape((n,))
    x.setflags(write=False)
    return x


def matrix2d(x):
    """Copy ``x`` into a read-only 2-d float array, like ``pints.matrix2d``
    (pints/_util.py:98-111)."""
    x = np.array(x, copy=True, dtype=float)
    if x.ndim == 1:
        x = x.reshape((len(x), 1))
    elif x.ndim != 2:
        raise ValueError('Unable to convert to 2d matrix.')
    x.setflags(write=False)
    return x


def is_batch(x):
    """True when ``x`` is a batch of parameter vectors (2-d), False for one
    vector (1-d)."""
    if isinstance(x, np.ndarray):
        return x.ndim == 2
    try:
        first = x[0]
    except (TypeError, IndexError, KeyError):
        return False
    return np.ndim(first) == 1


_blas_controller = None


def single_thread_blas():
    """Context manager: BLAS calls inside run on one thread.  The host-side
    optimiser updates multiply (population, d) by (d, d) matrices with d of a
    few units; multi-threaded OpenBLAS spends 5-10x longer synchronising its
    pool on such shapes than doing the arithmetic (xNES at population 65 536:
    ask 100 -> 14 ms, tell 42 -> 8 ms on 8 cores).  Results are unchanged (each
    output element is one short dot product whichever thread computes it).
    Without ``threadpoolctl`` (a dependency of the reference,
    pints/_evaluation.py:15) this is a no-op."""
    global _blas_controller
    if _blas_controller is None:
        try:
            import threadpoolctl
            _blas_controller = threadpoolctl.ThreadpoolController()
        except Exception:                                   # pragma: no cover
            _blas_controller = False
    if not _blas_controller:
        return contextlib.nullcontext()
    return _blas_controller.limit(limits=1, user_api='blas')
