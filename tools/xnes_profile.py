import os, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo')
import pints_b200 as pb
from oracle import pints_oracle as po
import scipy.linalg
h = po.headline_fhn_spec()
problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times, h.values)
ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
ev = pb.CudaEvaluator(ll, as_list=False); dp = ev.device_problem()
opt = pb.DeviceXNES([0.1, 0.5, 3.0], [0.005, 0.025, 0.15], b); opt.set_population_size(65536)
def T(f, reps=50):
    for _ in range(5): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e6
xs = opt.ask(); fx = dp.eval(xs); opt.tell(-fx)
print('ask (enqueue+exec) %.1f us' % T(lambda: opt.ask()))
xs = opt.ask()
print('eval %.1f us' % T(lambda: dp.eval(xs)))
fx = -dp.eval(xs)
print('masked_fill %.1f us' % T(lambda: fx.masked_fill(opt._d_inside == 0, float('inf'))))
print("sort (torch) %.1f us" % T(lambda: torch.sort(fx)))
print("sort (pb200_argsort_f64) %.1f us" % T(lambda: opt._sorter(fx)))
def tell():
    opt._ready_for_tell = True; opt.tell(fx)
print('tell total %.1f us' % T(tell))
G = np.random.rand(3, 3) * 0.01
t0 = time.perf_counter()
for _ in range(200): scipy.linalg.expm(G)
print('expm host %.1f us' % ((time.perf_counter() - t0) / 200 * 1e6))
def gen():
    x = opt.ask(); f = dp.eval(x); opt.tell(-f)
print("generation eager %.1f us" % T(gen))
print("generation() %.1f us (graph: %s)" % (T(lambda: opt.generation(dp, True)), opt._graph is not None))
