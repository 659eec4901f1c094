"""
ctypes binding of libpints_b200.so (the C ABI declared in
``include/pints_b200.h``).

This is the whole foreign-function surface of the package: a maintainer of the
reference who wants the CUDA path binds exactly these symbols (INTEGRATION.md
shows the stub).  There is no fallback: if the shared library cannot be loaded
the import of any compute entry point raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int32,
                    c_int64, c_uint64, c_void_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'lib', 'libpints_b200.so')

# ids, mirrored from include/pints_b200.h
MODEL_LOGISTIC, MODEL_HH_IK, MODEL_CONSTANT = 0, 1, 2
MODEL_FHN, MODEL_LOTKA_VOLTERRA, MODEL_SIR = 3, 4, 5
MODEL_GOODWIN, MODEL_HES1, MODEL_REPRESSILATOR = 6, 7, 8
MODEL_BEELER_REUTER = 9
OBJ_SSE, OBJ_GAUSSIAN_KNOWN_SIGMA, OBJ_GAUSSIAN, OBJ_MSE, OBJ_RMSE = 0, 1, 2, 3, 4
MAX_OUTPUTS, MAX_PARAMS, MAX_EVENTS, MAX_PEERS = 8, 16, 64, 16
SENTINEL_ROWS = 8

MODEL_NAMES = {MODEL_LOGISTIC: 'logistic', MODEL_HH_IK: 'hh_ik',
               MODEL_CONSTANT: 'constant', MODEL_FHN: 'fhn',
               MODEL_LOTKA_VOLTERRA: 'lotka_volterra', MODEL_SIR: 'sir',
               MODEL_GOODWIN: 'goodwin', MODEL_HES1: 'hes1',
               MODEL_REPRESSILATOR: 'repressilator',
               MODEL_BEELER_REUTER: 'beeler_reuter'}
AC_HAARIO_BARDENET, AC_HAARIO, AC_RAO_BLACKWELL = 0, 1, 2
OBJECTIVE_NAMES = {OBJ_SSE: 'sse',
                   OBJ_GAUSSIAN_KNOWN_SIGMA: 'gaussian_known_sigma',
                   OBJ_GAUSSIAN: 'gaussian', OBJ_MSE: 'mse', OBJ_RMSE: 'rmse'}


class SolverOpts(Structure):
    """``pb200_solver_opts``"""
    _fields_ = [('rtol', c_double), ('atol', c_double), ('h0', c_double),
                ('max_steps', c_int32), ('block', c_int32),
                ('lanes', c_int32), ('wide_interpolation', c_int32)]


class ExchangeDesc(Structure):
    """``pb200_exchange_desc``"""
    _fields_ = [('n_ranks', c_int32), ('rank', c_int32),
                ('per_rank', c_int64),
                ('stage', c_void_p * MAX_PEERS),
                ('flags', c_void_p * MAX_PEERS),
                ('stage_multicast', c_void_p), ('epoch', c_uint64)]


class LibraryError(RuntimeError):
    """A call into libpints_b200.so returned an error code."""


# every exported symbol: name -> (restype, argtypes)
_P = c_void_p
SIGNATURES = {
    'pb200_version': (c_int32, []),
    'pb200_last_error': (c_char_p, []),
    'pb200_series_len': (c_int64, [c_int32, c_int32]),
    'pb200_problem_create': (c_int32, [
        c_int32, c_int32, c_int32, c_int32, _P,
        POINTER(c_double), c_int32, POINTER(c_double), c_int32, c_double,
        POINTER(c_double), POINTER(c_double), c_int32, c_double,
        POINTER(c_void_p)]),
    'pb200_problem_destroy': (None, [_P]),
    'pb200_problem_set_prior_terms': (c_int32, [_P, c_int32, POINTER(c_int32),
                                                POINTER(c_int32),
                                                POINTER(c_double)]),
    'pb200_problem_n_parameters': (c_int32, [_P]),
    'pb200_eval': (c_int32, [_P, _P, c_int64, c_int64, _P, _P,
                             POINTER(SolverOpts), _P]),
    'pb200_sort_workspace_bytes': (c_int64, [c_int64]),
    'pb200_problem_sort_hint': (c_int32, [c_void_p]),
    'pb200_eval_sorted': (c_int32, [_P, _P, c_int64, c_int64, _P, _P,
                                    POINTER(SolverOpts), _P, c_int64, _P]),
    'pb200_eval_exchange': (c_int32, [_P, _P, c_int64, c_int64, _P,
                                      POINTER(SolverOpts), _P, c_int64,
                                      POINTER(ExchangeDesc), _P]),
    'pb200_exchange_finish': (c_int32, [POINTER(ExchangeDesc), c_int64,
                                        c_int64, _P, _P]),
    'pb200_eval_host': (c_int32, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P,
                                  _P, c_int64, POINTER(SolverOpts), _P]),
    'pb200_copy_async': (c_int32, [_P, _P, c_int64, c_int32, _P]),
    'pb200_host_copy': (c_int32, [_P, _P, c_int64, c_int32]),
    'pb200_host_is_pinned': (c_int32, [_P]),
    'pb200_simulate': (c_int32, [_P, _P, c_int64, c_int64, _P, _P,
                                 POINTER(SolverOpts), _P]),
    'pb200_score_trajectories': (c_int32, [_P, _P, _P, c_int64, c_int64, _P,
                                           _P]),
    'pb200_mh_accept': (c_int32, [_P, _P, _P, _P, _P, _P, c_int64, c_int32,
                                  c_int32, _P]),
    'pb200_mh_accept_ratio': (c_int32, [_P, _P, _P, _P, _P, _P, _P, c_int64,
                                        c_int32, c_int32, _P]),
    'pb200_ac_adapt': (c_int32, [c_int32, _P, _P, _P, _P, _P, _P, _P, _P,
                                 c_int64, c_int32, c_double, c_double, _P]),
    'pb200_ac_ask': (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_uint64,
                               c_uint64, _P]),
    'pb200_de_ask': (c_int32, [_P, c_int64, c_int64, _P, _P, c_int64, c_int32,
                               c_double, c_uint64, c_uint64, c_uint64, _P]),
    'pb200_ac_tell': (c_int32, [c_int32, c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                _P, c_int64, c_int64, c_int64, c_int32,
                                c_double, c_double, c_uint64, c_uint64, _P]),
    'pb200_ac_ask_it': (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_uint64,
                                  _P, _P, _P]),
    'pb200_de_ask_it': (c_int32, [_P, c_int64, c_int64, _P, _P, c_int64,
                                  c_int32, c_uint64, _P, _P, _P]),
    'pb200_ac_tell_it': (c_int32, [c_int32, c_int32, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _P, c_int64, c_int64, c_int32,
                                   c_double, c_uint64, _P, _P, _P]),
    'pb200_iter_advance': (c_int32, [_P, _P]),
    'pb200_mcmc_chain_run': (c_int32, [_P, c_int32, _P, _P, _P, _P, _P, _P, _P,
                                       _P, _P, _P, c_int64, c_int64, c_int32,
                                       c_double, c_uint64, _P, c_int64,
                                       c_int64, _P]),
    'pb200_xnes_ask': (c_int32, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64,
                                 c_int32, c_uint64, _P, _P]),
    'pb200_xnes_sums': (c_int32, [_P, _P, _P, _P, c_int32, _P, _P, _P,
                                  c_int64, c_int32, _P]),
    'pb200_xnes_update': (c_int32, [_P, _P, _P, _P, c_int32, c_double, c_double,
                                    c_double, c_int32, c_uint64, c_int32, _P]),
    'pb200_argsort_workspace_bytes': (c_int64, [c_int64]),
    'pb200_argsort_f64': (c_int32, [_P, _P, c_int64, _P, c_int64, _P, _P]),
    'pb200_hbac_adapt': (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32,
                                   c_double, c_double, _P]),
    'pb200_hbac_propose': (c_int32, [_P, _P, _P, _P, _P, c_int64, c_int32,
                                     _P]),
    'pb200_de_propose': (c_int32, [_P, c_int64, _P, _P, _P, _P, c_int64,
                                   c_int32, c_double, _P]),
    'pb200_philox_normal': (c_int32, [_P, c_int64, c_uint64, c_uint64, _P]),
    'pb200_philox_uniform': (c_int32, [_P, c_int64, c_uint64, c_uint64, _P]),
    'pb200_philox_pairs': (c_int32, [_P, _P, c_int64, c_int64, c_int64,
                                     c_uint64, c_uint64, _P]),
    'pb200_fp64_probe': (c_int32, [_P, c_int32, c_int32, c_int64, _P]),
}

_lib = None


def load():
    """Load (once) and return the shared library.  Raises if it is missing:
    the product has no other implementation to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryError(
            'libpints_b200.so has not been built (%s). Build it with '
            '`python -m pints_b200.build`; there is no CPU fallback.'
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Turn a non-zero return code into an exception carrying the message."""
    if rc != 0:
        msg = load().pb200_last_error()
        raise LibraryError('libpints_b200 error %d: %s'
                           % (rc, msg.decode() if msg else '?'))


def doubles(values):
    """A ctypes double array (and its length) from any float sequence."""
    values = [float(v) for v in values]
    n = len(values)
    if n == 0:
        return None, 0
    return (c_double * n)(*values), n


__all__ = ['load', 'check', 'doubles', 'SolverOpts', 'ExchangeDesc', 'LibraryError', 'byref',
           'SIGNATURES', 'LIB_PATH']
