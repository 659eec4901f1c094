"""Randomised parity soak of the further ODE toy models (development tool):
random parameter vectors around the suggested ones, device trajectories at a
tightened tolerance against the oracle's LSODA run at 1e-12 / 1e-13.  Prints
the worst error per model, scaled as the parity tests scale it:
|device - truth| / (1e-6 (1 + |truth|))  (must stay below 1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402
from oracle import pints_oracle as po  # noqa: E402

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
CASES = [
    ('goodwin', pb.toy.GoodwinOscillatorModel, np.linspace(0, 100, 200),
     lambda x, t: po.goodwin_simulate(x, t, **po.TRUTH_TOL), 0.05),
    ('hes1', pb.toy.Hes1Model, np.linspace(0, 270, 150),
     lambda x, t: po.hes1_simulate(x, t, **po.TRUTH_TOL).reshape(-1, 1), 0.05),
    ('repressilator', pb.toy.RepressilatorModel, np.linspace(0, 40, 200),
     lambda x, t: po.repressilator_simulate(x, t, **po.TRUTH_TOL), 0.05),
    ('beeler_reuter', pb.toy.ActionPotentialModel, np.arange(0, 400, 1.0),
     lambda x, t: po.beeler_reuter_simulate(x, t, rtol=1e-12, atol=1e-20,
                                            mxstep=1000000), 0.03),
]
for name, cls, times, truth_of, spread in CASES:
    model = cls()
    model.set_solver_options(rtol=1e-10, atol=1e-10) if name != 'beeler_reuter' \
        else model.set_solver_tolerances(1e-10, 1e-10)
    x0 = np.asarray(model.suggested_parameters(), dtype=float)
    worst = 0.0
    for seed in range(seeds):
        rng = np.random.RandomState(100 + seed)
        xs = x0 * (1 + spread * rng.randn(16, len(x0)))
        got = model.simulate(xs, times)
        got = got.reshape(len(xs), len(times), -1)
        for x, g in zip(xs, got):
            truth = np.asarray(truth_of(x, times)).reshape(len(times), -1)
            scale = 1e-6 * (1 + np.abs(truth))
            if name == 'beeler_reuter':
                scale = scale * [1, 1e-6]        # Ca_i is 1e-7 .. 1e-5
            worst = max(worst, float(np.max(np.abs(g - truth) / scale)))
    print('%-14s %d vectors: worst |device - truth| / (1e-6 (1 + |truth|)) = %.3g'
          % (name, seeds * 16, worst))
    assert worst < 1.0, name
print('ok')
