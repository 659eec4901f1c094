"""Whole xNES generations with the population sharded over the GPUs of a node
(one process per GPU under torchrun): every rank draws the same population,
scores its block, the scores travel through the fused exchange, every rank
applies the same update.  Population = 65 536 per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P tools/xnes_sharded_time.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402
from oracle import pints_oracle as po  # noqa: E402  (inputs)

rank = int(os.environ.get('LOCAL_RANK', '0'))
world = int(os.environ.get('WORLD_SIZE', '1'))
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
h = po.headline_fhn_spec()
problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times, h.values)
ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
n = 65536 * world
ctl = pb.CudaOptimisationController(
    ll, [0.1, 0.5, 3.0], [0.005, 0.025, 0.15], boundaries=b,
    method=pb.DeviceXNES, sharded=True, devices='cuda:%d' % rank)
ctl.optimiser().set_population_size(n)
ctl.set_max_iterations(40)
ctl.set_max_unchanged_iterations(None)
x, f = ctl.run()
if rank == 0:
    print('DeviceXNES sharded over %d GPU(s), population %d, 40 generations: x = %s f = %.6f'
          % (world, n, np.round(x, 5), f))
# a run's set-up (symmetric-memory rendezvous) takes hundreds of milliseconds
# and varies, so the generations are timed phase by phase instead
# where the time goes (rank 0's view, synchronising after every phase)
from pints_b200._dist import ScoreExchange
opt = pb.DeviceXNES([0.1, 0.5, 3.0], [0.005, 0.025, 0.15], b, device='cuda:%d' % rank)
opt.set_population_size(n)
dp = pb.CudaEvaluator(ll, as_list=False, devices='cuda:%d' % rank).device_problem()
per = n // world
ex = ScoreExchange(per, dp.device, None)
acc = [0.0, 0.0, 0.0]
for it in range(30):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    xs = opt.ask(); torch.cuda.synchronize(); t1 = time.perf_counter()
    fs = ex.run(dp, xs[rank * per:(rank + 1) * per])[:n]; torch.cuda.synchronize(); t2 = time.perf_counter()
    opt.tell(-fs); t3 = time.perf_counter()
    if it >= 5:
        acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2
if rank == 0:
    total = sum(acc) / 25
    print('phases (us): ask %.0f, score + exchange %.0f, tell %.0f; fused exchange: %s; '
          'generation %.3f ms = %.3e evaluations/s'
          % (acc[0] / 25 * 1e6, acc[1] / 25 * 1e6, acc[2] / 25 * 1e6, ex.fused,
             total * 1e3, n / total))
dist.destroy_process_group()
