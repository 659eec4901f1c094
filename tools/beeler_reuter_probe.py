"""Beeler-Reuter on the device: error against the tight-tolerance fixture at a
few tolerances, attempted steps, and throughput of a population."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import pints_b200 as pb  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden',
                         'beeler_reuter.npz'))
model = pb.toy.ActionPotentialModel()
times, xs = g['times'], g['xs']
rel = lambda a, b: np.abs(a - b) / np.abs(b)
for tol in (None, 1e-5, 1e-6, 3e-7, 1e-7, 1e-8, 1e-10):
    if tol:
        model.set_solver_tolerances(tol, tol)
    traj = model.simulate(xs[:8, :5], times)
    e = np.abs(traj - g['truth'])
    problem = pb.MultiOutputProblem(model, times, g['values'])
    kw = dict(rtol=tol, atol=tol) if tol else {}
    ev = pb.CudaEvaluator(pb.GaussianLogLikelihood(problem), as_list=False, **kw)
    got, steps = ev.evaluate_array(xs, True)
    print('tol', tol, 'V err', e[..., 0].max(), 'V err/(1+|V|)',
          (e[..., 0] / (1 + np.abs(g['truth'][..., 0]))).max(),
          'Cai rel', (e[..., 1] / g['truth'][..., 1]).max(),
          'score rel', rel(got, g['truth_scores']).max(),
          'vs ref', rel(got, g['scores']).max(),
          'steps', steps.min(), steps.max())
print('reference score rel err', rel(g['scores'], g['truth_scores']).max())

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rng = np.random.RandomState(5)
big = np.tile(xs[0], (n, 1)) * (1 + 0.02 * rng.randn(n, xs.shape[1]))
problem = pb.MultiOutputProblem(pb.toy.ActionPotentialModel(), times, g['values'])
ev = pb.CudaEvaluator(pb.GaussianLogLikelihood(problem), as_list=False)
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, steps = ev.evaluate_array(big, True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('n', n, 'time %.4f s' % dt, '%.3e evals/s' % (n / dt), 'steps mean',
          steps.mean(), 'nan', np.isnan(out).sum())
spec = ev.spec()
if spec is not None:
    fl = spec.flops_per_eval(steps).sum()
    print('algorithmic TFLOP/s %.2f' % (fl / dt / 1e12))
