"""Device time of one evaluation with and without the cost sort, small batches (development tool)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pints_b200 as pb
from oracle import pints_oracle as po

def t_eval(dp, d, out, reps=30):
    for _ in range(5): dp.eval(d, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): dp.eval(d, out=out)
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps

s = po.headline_fhn_spec()
fhn = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times, s.values), 0.5)
rng = np.random.RandomState(1)
t = np.linspace(0, 20, 1000)
v = po.lv_simulate([0.5, 0.15, 1.0, 0.3], t) + rng.normal(0, 0.1, (1000, 2))
lv = pb.GaussianLogLikelihood(pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), t, v))
for name, f, base in (('fhn', fhn, [0.1, 0.5, 3]), ('lv', lv, [0.5, 0.15, 1.0, 0.3, 0.1, 0.1])):
    for n in (256, 1024, 2048, 4096, 8192, 16384, 32768):
        xs = np.array(base) * (1 + 0.05 * rng.randn(n, len(base)))
        d = torch.from_numpy(xs).cuda(); out = torch.empty(n, dtype=torch.float64, device='cuda')
        res = []
        for sort in (False, True):
            dp = pb.CudaEvaluator(f, as_list=False, sort=sort).device_problem()
            res.append(t_eval(dp, d, out))
        print('%s n=%6d  unsorted %.4f ms  sorted %.4f ms' % (name, n, res[0], res[1]))
