"""One persistent MCMC chain-loop launch on the logistic problem (profiling
target: `ncu -k regex:analytic_kernel`), event-timed."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402
from oracle import pints_oracle as po  # noqa: E402  (inputs)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
rng = np.random.RandomState(1)
times = np.linspace(0, 100, 1000)
values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 2, 1000)
f = pb.LogPosterior(
    pb.GaussianLogLikelihood(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)),
    pb.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
x0 = np.array([0.1, 50, 2.0]) * (1 + 0.02 * rng.randn(n, 3))
for rep in range(2):
    mc = pb.CudaMCMCController(f, n, x0, method=pb.HaarioBardenetACMC, seed=1)
    mc.use_chain_loop = True
    mc.use_graph = True
    mc.graph_min_iterations = 4
    mc.set_max_iterations(iters)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mc.run(to_host=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('logistic, %d chains: %.2f us per iteration (%d iterations, %d in the chain loop)'
          % (n, dt / iters * 1e6, iters, mc.chain_loop_iterations))
