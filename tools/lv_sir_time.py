"""Lotka-Volterra / SIR kernel time at small and large populations
(development tool for the interpolation-loop variants)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import pints_b200 as pb
from oracle import pints_oracle as po
rng = np.random.RandomState(1)
def problems():
    t = np.linspace(0, 20, 1000)
    m = pb.toy.LotkaVolterraModel(); x = np.array([0.5, 0.15, 1.0, 0.3])
    v = po.lv_simulate(x, t) + rng.normal(0, 0.1, (1000, 2))
    yield 'LV', pb.GaussianLogLikelihood(pb.MultiOutputProblem(m, t, v)), np.hstack([x, [0.1, 0.1]])
    m = pb.toy.SIRModel(); x = np.array([0.026, 0.285, 38.0]); t = np.linspace(0, 21, 1000)
    v = po.sir_simulate(x, t) + rng.normal(0, 1, (1000, 2))
    yield 'SIR', pb.GaussianLogLikelihood(pb.MultiOutputProblem(m, t, v)), np.hstack([x, [1.0, 1.0]])
out_line = []
for name, ll, x in problems():
    dp = pb.CudaEvaluator(ll, as_list=False).device_problem()
    for n in (4096, 65536):
        xs = x * (1 + 0.05 * rng.randn(n, len(x)))
        d = torch.from_numpy(xs).cuda(); out = torch.empty(n, dtype=torch.float64, device='cuda')
        for _ in range(5): dp.eval(d, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): dp.eval(d, out=out)
        e1.record(); e1.synchronize()
        out_line.append('%s N=%d %.1f us (sort hint %d)' % (name, n, e0.elapsed_time(e1) / 50 * 1e3, dp.sort_hint()))
print('; '.join(out_line))
