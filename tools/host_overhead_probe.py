import sys, time, numpy as np, torch, cProfile, pstats
sys.path.insert(0, ".")
import pints_b200 as pb
from oracle import pints_oracle as po
s = po.headline_fhn_spec()
ll = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times[:8], s.values[:8]), 0.5)
ev = pb.CudaEvaluator(ll, as_list=False)
xs = pb.pinned_empty((65536, 3)); xs[:] = po.headline_fhn_points(65536)
for _ in range(20): ev.evaluate(xs)
t0 = time.perf_counter()
for _ in range(500): ev.evaluate(xs)
print('per call (8-point series, GPU work tiny): %.1f us' % ((time.perf_counter() - t0) / 500 * 1e6))
pr = cProfile.Profile(); pr.enable()
for _ in range(500): ev.evaluate(xs)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
