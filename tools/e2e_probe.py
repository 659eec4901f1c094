"""End-to-end latency of CudaEvaluator.evaluate for several chunk settings (development tool)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pints_b200 as pb
from oracle import pints_oracle as po
s = po.headline_fhn_spec()
ll = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times, s.values), 0.5)
xs = po.headline_fhn_points(65536)
ref = None
for chunk, sort in ((1 << 30, True), (32768, True), (16384, True), (8192, True),
                    (1 << 30, False), (16384, False)):
    ev = pb.CudaEvaluator(ll, as_list=False, chunk_rows=chunk, sort=sort)
    for _ in range(5): out = ev.evaluate(xs)
    if ref is None: ref = out
    assert np.array_equal(out, ref)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): ev.evaluate(xs)
    dt = (time.perf_counter() - t0) / 50
    print('sort=%s ' % sort, end='')
    print('chunk_rows=%d: %.3f ms per evaluate, %.3e evals/s' % (chunk, dt * 1e3, 65536 / dt))
ev = pb.CudaEvaluator(ll)   # list output
for _ in range(5): ev.evaluate(xs)
t0 = time.perf_counter()
for _ in range(50): ev.evaluate(xs)
dt = (time.perf_counter() - t0) / 50
print('as_list=True default: %.3f ms per evaluate, %.3e evals/s' % (dt * 1e3, 65536 / dt))
