import sys, numpy as np, torch
sys.path.insert(0, ".")
import pints_b200 as pb
from oracle import pints_oracle as po
m = pb.toy.HodgkinHuxleyIKModel(); t = m.suggested_times()
rng = np.random.RandomState(1)
v = po.hh_ik_simulate(m.suggested_parameters(), t) + rng.normal(0, 40, len(t))
ll = pb.GaussianLogLikelihood(pb.SingleOutputProblem(m, t, v))
n = 16384
x0 = np.hstack([np.array(m.suggested_parameters()) * rng.uniform(0.95, 1.05, (n, 5)), 40 * np.ones((n, 1))])
mc = pb.CudaMCMCController(ll, n, x0, method=pb.HaarioBardenetACMC, seed=1)
mc.set_max_iterations(int(sys.argv[1]) if len(sys.argv) > 1 else 260)
mc.set_initial_phase_iterations(20)
import time
t0 = time.perf_counter(); mc.run(to_host=False); torch.cuda.synchronize(); print('s per iteration', (time.perf_counter() - t0) / mc._max_iterations)
