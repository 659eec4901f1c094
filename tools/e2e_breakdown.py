"""Where the end-to-end time of one evaluate() goes (development tool)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pints_b200 as pb
from oracle import pints_oracle as po

def bench(f, n=200):
    for _ in range(10): f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6

xs = po.headline_fhn_points(65536)
pin = torch.empty((65536, 3), dtype=torch.float64, pin_memory=True)
pin_np = pin.numpy()
out_pin = torch.empty(65536, dtype=torch.float64, pin_memory=True)
out_np = out_pin.numpy()
res = np.empty(65536)
print('np.copyto pageable->pinned 1.5 MB: %.1f us' % bench(lambda: np.copyto(pin_np, xs)))
print('np.copyto pinned->pageable 0.5 MB: %.1f us' % bench(lambda: np.copyto(res, out_np)))
print('np.empty + copy 0.5 MB:            %.1f us' % bench(lambda: out_np.copy()))
d = torch.empty((65536, 3), dtype=torch.float64, device='cuda')
do = torch.empty(65536, dtype=torch.float64, device='cuda')
def h2d():
    d.copy_(pin, non_blocking=True); torch.cuda.synchronize()
def d2h():
    out_pin.copy_(do, non_blocking=True); torch.cuda.synchronize()
def nothing():
    torch.cuda.synchronize()
print('sync only:                         %.1f us' % bench(nothing))
print('H2D 1.5 MB + sync:                 %.1f us' % bench(h2d))
print('D2H 0.5 MB + sync:                 %.1f us' % bench(d2h))
s = po.headline_fhn_spec()
ll = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times, s.values), 0.5)
ev = pb.CudaEvaluator(ll, as_list=False, chunk_rows=1 << 30)
dp = ev.device_problem()
def dev():
    dp.eval(d, out=do); torch.cuda.synchronize()
print('device eval + sync:                %.1f us' % bench(dev, 50))
print('evaluate(xs) single chunk:         %.1f us' % bench(lambda: ev.evaluate(xs), 50))
ev4 = pb.CudaEvaluator(ll, as_list=False)
print('evaluate(xs) default chunks:       %.1f us' % bench(lambda: ev4.evaluate(xs), 50))
print('tolist 65536:                      %.1f us' % bench(lambda: res.tolist(), 50))
