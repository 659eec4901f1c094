"""Device MCMC loop launch by launch, replayed from a CUDA graph, and as one
persistent launch (independent chains on a closed-form model), on a cheap
model with few chains (host-bound when eager) and on BASELINE configs[2]."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402
from oracle import pints_oracle as po  # noqa: E402  (inputs)

rng = np.random.RandomState(1)
times = np.linspace(0, 100, 1000)
values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 2, 1000)
logistic = pb.LogPosterior(
    pb.GaussianLogLikelihood(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)),
    pb.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
m = pb.toy.HodgkinHuxleyIKModel()
t = m.suggested_times()
v = po.hh_ik_simulate(m.suggested_parameters(), t) + rng.normal(0, 40, len(t))
hh = pb.GaussianLogLikelihood(pb.SingleOutputProblem(m, t, v))
cases = [
    ('logistic, 8 chains', logistic, np.array([0.1, 50, 2.0]), 8, 20000, 'HaarioBardenetACMC'),
    ('logistic, 64 chains', logistic, np.array([0.1, 50, 2.0]), 64, 20000, 'HaarioBardenetACMC'),
    ('logistic, 1024 chains', logistic, np.array([0.1, 50, 2.0]), 1024, 5000, 'HaarioBardenetACMC'),
    ('logistic, 1024 chains (DE)', logistic, np.array([0.1, 50, 2.0]), 1024, 5000,
     'DifferentialEvolutionMCMC'),
    ('HH-IK, 64 chains', hh, np.append(m.suggested_parameters(), 40.0), 64, 2000,
     'HaarioBardenetACMC'),
    ('HH-IK, 16384 chains', hh, np.append(m.suggested_parameters(), 40.0), 16384, 500,
     'HaarioBardenetACMC'),
]
for label, f, x, n, iters, method in cases:
    x0 = x * (1 + 0.02 * rng.randn(n, len(x)))
    for mode in ('eager', 'graph', 'chain loop'):
        mc = pb.CudaMCMCController(f, n, x0, method=getattr(pb, method), seed=1)
        mc.use_chain_loop = mode == 'chain loop'
        mc.use_graph = mode != 'eager'
        mc.graph_min_iterations = 4
        mc.set_max_iterations(iters)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mc.run(to_host=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if mode == 'chain loop' and not mc.chain_loop_iterations:
            continue
        print('%-28s %-10s %8.1f us per iteration (%d iterations, %d replayed, %d in the '
              'chain loop)' % (label, mode, dt / iters * 1e6, iters, mc.graph_replayed,
                               mc.chain_loop_iterations))
