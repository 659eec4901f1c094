"""Wall time of one xNES generation (ask + score + tell) at population 65 536
on the headline problem: host-vectorised XNES against DeviceXNES."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402
from oracle import pints_oracle as po  # noqa: E402  (inputs)

h = po.headline_fhn_spec()
problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times, h.values)
ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
GENERATIONS = 300
for method, native in ((pb.XNES, None), (pb.DeviceXNES, False), (pb.DeviceXNES, True),
                       (pb.DeviceXNES, False), (pb.DeviceXNES, True)):
    np.random.seed(1)
    ctl = pb.CudaOptimisationController(ll, [0.1, 0.5, 3.0], [0.005, 0.025, 0.15],
                                        boundaries=b, method=method)
    ctl.optimiser().set_population_size(n)
    ctl.set_max_iterations(5)
    ctl.set_max_unchanged_iterations(None)
    ctl.run()                                   # warm-up: buffers, JIT-free
    np.random.seed(1)
    ctl = pb.CudaOptimisationController(ll, [0.1, 0.5, 3.0], [0.005, 0.025, 0.15],
                                        boundaries=b, method=method)
    ctl.optimiser().set_population_size(n)
    if native is not None:
        ctl.optimiser().native_sort = native
    ctl.set_max_iterations(GENERATIONS)
    ctl.set_max_unchanged_iterations(None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, f = ctl.run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%-11s native_sort=%s population %d: %.3f ms per generation, %.3e evaluations/s, x = %s f = %.6f'
          % (method.__name__, native, n, dt / GENERATIONS * 1e3, GENERATIONS * n / dt, np.round(x, 5), f))
