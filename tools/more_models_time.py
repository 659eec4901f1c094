"""Kernel time of the further ODE toy models (Goodwin, Hes1, Repressilator) on
the fixtures of tests/golden, at two population sizes (development tool)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402

MODELS = [('goodwin', 'GoodwinOscillatorModel'), ('hes1', 'Hes1Model'),
          ('repressilator', 'RepressilatorModel')]
for name, cls in MODELS:
    g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    model = getattr(pb.toy, cls)()
    multi = model.n_outputs() > 1
    problem = (pb.MultiOutputProblem if multi else pb.SingleOutputProblem)(
        model, g['times'], g['values'])
    ev = pb.CudaEvaluator(pb.GaussianLogLikelihood(problem), as_list=False)
    dp = ev.device_problem()
    rng = np.random.RandomState(3)
    for n in (4096, 65536):
        xs = g['xs'][rng.randint(0, len(g['xs']), n)] * (1 + 0.01 * rng.randn(n, g['xs'].shape[1]))
        d = torch.from_numpy(xs).cuda()
        out = torch.empty(n, dtype=torch.float64, device='cuda')
        steps = dp.new_steps(n)
        dp.eval(d, out=out, steps=steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            dp.eval(d, out=out)
        e0.record()
        for _ in range(10):
            dp.eval(d, out=out)
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print('%-14s N=%6d  %.4f ms  %.3e evals/s  mean steps %.0f  nan %d' % (
            name, n, ms, n / ms * 1e3, steps.float().mean().item(),
            int(torch.isnan(out).sum())))
