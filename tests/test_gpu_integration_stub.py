"""
The minimal ctypes binding shown in INTEGRATION.md section 1, executed: a
maintainer's ~25-line stub over the C ABI gives the same scores as the full
CudaEvaluator of this package.
"""
import ctypes
from ctypes import (POINTER, byref, c_double, c_int32, c_int64, c_void_p)

import numpy as np
import pytest

from oracle import pints_oracle as po

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def test_integration_stub_matches_cuda_evaluator():
    import pints_b200 as pb
    from pints_b200 import build
    pb.require_cuda()
    lib = ctypes.CDLL(build.LIB)
    lib.pb200_last_error.restype = ctypes.c_char_p
    lib.pb200_series_len.restype = c_int64
    lib.pb200_series_len.argtypes = [c_int32, c_int32]
    lib.pb200_problem_create.argtypes = [
        c_int32, c_int32, c_int32, c_int32, c_void_p,
        POINTER(c_double), c_int32, POINTER(c_double), c_int32,
        c_double, POINTER(c_double), POINTER(c_double), c_int32, c_double,
        POINTER(c_void_p)]
    lib.pb200_eval.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                               c_void_p, c_void_p, c_void_p]

    def check(rc):
        if rc:
            raise RuntimeError(lib.pb200_last_error().decode())

    spec = po.headline_fhn_spec()
    model = pb.toy.FitzhughNagumoModel()
    problem = pb.MultiOutputProblem(model, spec.times, spec.values)
    log_likelihood = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)

    # --- the stub of INTEGRATION.md, with `p = log_likelihood._problem` ---
    p = log_likelihood._problem
    nt, no = p.n_times(), p.n_outputs()
    rows = np.zeros((nt + 8, 1 + no))
    rows[:nt, 0] = p.times()
    rows[:nt, 1:] = p.values()
    rows[nt:, 0] = np.inf
    packed = np.zeros(lib.pb200_series_len(nt, no))
    packed[:rows.size] = rows.ravel()
    series = torch.from_numpy(packed).cuda()
    y0 = (c_double * 2)(*p._model._y0)
    oc = (c_double * (2 * no))(*(np.ones(no) * log_likelihood._offset),
                               *(np.ones(no) * log_likelihood._multip))
    handle = c_void_p()
    check(lib.pb200_problem_create(3, 1, nt, no, series.data_ptr(), y0, 2, oc,
                                   2 * no, 1.0, None, None, 0, 0.0,
                                   byref(handle)))
    positions = po.headline_fhn_points(500)
    xs = torch.from_numpy(np.ascontiguousarray(positions, dtype=float)).cuda()
    out = torch.empty(len(xs), dtype=torch.float64, device='cuda')
    check(lib.pb200_eval(handle, xs.data_ptr(), len(xs), xs.stride(0),
                         out.data_ptr(), None, None,
                         torch.cuda.current_stream().cuda_stream))
    scores = out.cpu().tolist()
    # ----------------------------------------------------------------------
    want = pb.CudaEvaluator(log_likelihood).evaluate(positions)
    assert scores == want
    ref = po.scores(spec, positions[:8])
    assert np.max(np.abs(np.array(scores[:8]) - ref) / np.abs(ref)) < 1e-6
    lib.pb200_problem_destroy.argtypes = [c_void_p]
    lib.pb200_problem_destroy(handle)
