import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import pints_b200 as pb
from oracle import pints_oracle as po
m = pb.toy.HodgkinHuxleyIKModel(); t = m.suggested_times()
rng = np.random.RandomState(1)
v = po.hh_ik_simulate(m.suggested_parameters(), t) + rng.normal(0, 40, len(t))
ll = pb.GaussianLogLikelihood(pb.SingleOutputProblem(m, t, v))
n = 16384
x0 = np.hstack([np.array(m.suggested_parameters()) * rng.uniform(0.95, 1.05, (n, 5)), 40 * np.ones((n, 1))])
dp = pb.CudaEvaluator(ll, as_list=False).device_problem()
for fused in (True, False):
    s = pb.HaarioBardenetACMC(x0, seed=1, fused=fused)
    fx = torch.empty(n, dtype=torch.float64, device='cuda')
    samples = torch.empty((n, 300, 6), dtype=torch.float64, device='cuda')
    dp.eval(s.ask(), out=fx); s.tell(fx, store=(samples, 0)); s.set_initial_phase(False)
    torch.cuda.synchronize()
    ta = tb = tc = 0.0
    t0 = time.perf_counter()
    for it in range(1, 300):
        a = time.perf_counter(); xs = s.ask(); b = time.perf_counter()
        dp.eval(xs, out=fx); c = time.perf_counter()
        s.tell(fx, store=(samples, it)); d = time.perf_counter()
        ta += b - a; tb += c - b; tc += d - c
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print('fused', fused, 'host enqueue per it: ask %.1f us eval %.1f us tell %.1f us; total enqueue %.1f us; wall incl sync %.1f us' % (ta / 299 * 1e6, tb / 299 * 1e6, tc / 299 * 1e6, (t1 - t0) / 299 * 1e6, (t2 - t0) / 299 * 1e6))
