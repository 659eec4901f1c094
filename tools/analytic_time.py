import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import pints_b200 as pb
from oracle import pints_oracle as po
m = pb.toy.HodgkinHuxleyIKModel(); t = m.suggested_times()
rng = np.random.RandomState(1)
v = po.hh_ik_simulate(m.suggested_parameters(), t) + rng.normal(0, 40, len(t))
ll = pb.GaussianLogLikelihood(pb.SingleOutputProblem(m, t, v))
ev = pb.CudaEvaluator(ll, as_list=False); dp = ev.device_problem()
for n in (16384, 65536):
    x = np.hstack([np.array(m.suggested_parameters()) * (1 + 0.05 * rng.randn(n, 5)), 40 * np.ones((n, 1))])
    d = torch.from_numpy(x).cuda(); out = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(5): dp.eval(d, out=out)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dp.eval(d, out=out)
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1) / 20
    print("HH-IK 4800 pts N=%d: %.4f ms  %.3e point-evals/s" % (n, ms, n * 4800 / ms * 1e3))
# logistic
times = np.linspace(0, 100, 1000); values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 1000)
sse = pb.SumOfSquaresError(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values))
dp = pb.CudaEvaluator(sse, as_list=False).device_problem()
for n in (64, 65536, 1000000):
    x = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(n, 2)); d = torch.from_numpy(x).cuda(); out = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(5): dp.eval(d, out=out)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dp.eval(d, out=out)
    e1.record(); e1.synchronize(); ms = e0.elapsed_time(e1) / 20
    print("Logistic 1000 pts N=%d: %.4f ms  %.3e point-evals/s" % (n, ms, n * 1000 / ms * 1e3))
