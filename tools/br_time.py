"""Beeler-Reuter: time of a population at the default tolerance and at a few
others, optionally with a variant build of the library
(PINTS_B200_LIB=build/variants/x.so, see tools/build_variant.sh)."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from pints_b200 import _lib
if os.environ.get('PINTS_B200_LIB'):
    _lib.LIB_PATH = os.path.abspath(os.environ['PINTS_B200_LIB'])
import pints_b200 as pb  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'beeler_reuter.npz'))
times, xs = g['times'], g['xs']
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tols = [float(v) for v in sys.argv[2:]] or [0.0]
rng = np.random.RandomState(5)
big = np.tile(xs[0], (n, 1)) * (1 + 0.02 * rng.randn(n, xs.shape[1]))
problem = pb.MultiOutputProblem(pb.toy.ActionPotentialModel(), times, g['values'])
for tol in tols:
    kw = dict(rtol=tol, atol=tol) if tol else {}
    ev = pb.CudaEvaluator(pb.GaussianLogLikelihood(problem), as_list=False, **kw)
    dp = ev.device_problem()
    d_x = torch.from_numpy(big).cuda()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    steps = dp.new_steps(n)
    dp.eval(d_x, out=out, steps=steps)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dp.eval(d_x, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    small = ev.evaluate(xs)
    rel = np.max(np.abs(np.asarray(small) - g['truth_scores']) / np.abs(g['truth_scores']))
    print('%s n=%d tol=%g: %.3f ms  %.3e evals/s  mean steps %.0f  %.2f us per step  score rel err %.2e  nan %d'
          % (os.environ.get('PINTS_B200_LIB', 'product'), n, tol, ms, n / ms * 1e3, steps.float().mean().item(),
             ms * 1e3 / steps.float().mean().item(), rel, int(torch.isnan(out).sum())))
