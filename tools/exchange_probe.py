"""Where the time of the fused exchange goes, on ONE GPU: the evaluation kernel
with 1..8 staging arrays (all local: the stores and the signalling tail without
the links) against the plain launch, and the reading-side kernel alone."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pints_b200 as pb
from pints_b200 import _lib
from pints_b200._engine import exchange_finish
from oracle import pints_oracle as po

spec = po.headline_fhn_spec()
f = pb.GaussianKnownSigmaLogLikelihood(
    pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times, spec.values), 0.5)
dp = pb.CudaEvaluator(f, as_list=False).device_problem()
n = 65536
d_x = torch.from_numpy(po.headline_fhn_points(n)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        flush.fill_(i)
        e0[i].record(); fn(); e1[i].record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in zip(e0, e1)]))


print('plain eval %.4f ms' % timed(lambda: dp.eval(d_x)))
for world in (1, 2, 4, 8):
    pairs = world * n
    stage = [torch.zeros(2 * pairs, dtype=torch.float64, device='cuda') for _ in range(world)]
    flags = [torch.zeros(16, dtype=torch.int64, device='cuda') for _ in range(world)]
    out = torch.empty(pairs, dtype=torch.float64, device='cuda')
    st = {'e': 0}

    def desc():
        d = _lib.ExchangeDesc()
        d.n_ranks, d.rank, d.per_rank = world, 0, n
        for q in range(world):
            d.stage[q] = stage[q].data_ptr(); d.flags[q] = flags[q].data_ptr()
        d.epoch = st['e']
        return d

    def producer():
        st['e'] += 1
        dp.eval_exchange(d_x, desc())
    t_prod = timed(producer)
    flags[0][:] = 1 << 60                       # the reading side never waits
    t_fin = timed(lambda: exchange_finish(desc(), n, pairs, out))
    print('  world %d: evaluation kernel with exchange %.4f ms, reading side %.4f ms'
          % (world, t_prod, t_fin))
