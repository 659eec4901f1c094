import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
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
    h = po.headline_fhn_spec()
    yield 'FHN', pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times, h.values), 0.5), np.array([0.1, 0.5, 3.0])
for name, ll, x in problems():
    for n in (1024, 4096, 16384):
        xs = x * (1 + 0.05 * rng.randn(n, len(x)))
        d = torch.from_numpy(xs).cuda(); out = torch.empty(n, dtype=torch.float64, device='cuda')
        res = []
        for sort in (True, False):
            dp = pb.CudaEvaluator(ll, as_list=False, sort=sort).device_problem()
            for _ in range(5): dp.eval(d, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50): dp.eval(d, out=out)
            e1.record(); e1.synchronize()
            res.append(e0.elapsed_time(e1) / 50 * 1e3)
        print('%s N=%d: sorted %.1f us, unsorted %.1f us' % (name, n, res[0], res[1]))
