"""Extended randomised parity soak (development tool): the live-oracle test of
tests/test_gpu_parity.py over many seeds, plus random SIR / Lotka-Volterra
problems.  Prints the worst relative errors seen."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pints_b200 as pb
from oracle import pints_oracle as po
import test_gpu_parity as T

worst = {}
def note(k, v):
    worst[k] = max(worst.get(k, 0.0), float(v))

for seed in range(1, int(sys.argv[1]) + 1 if len(sys.argv) > 1 else 21):
    T.test_live_oracle_random_problems(pb, seed)
    rng = np.random.RandomState(1000 + seed)
    # SIR on a ragged grid starting late
    times = np.sort(rng.uniform(1.0, 25, 90))
    x_true = [0.026, 0.285, 38]
    values = po.sir_simulate(x_true, times) + rng.normal(0, 1.0, (90, 2))
    xs = np.hstack([np.array(x_true) * (1 + 0.1 * rng.randn(32, 3)), rng.uniform(0.8, 1.5, (32, 2))])
    spec = po.Spec('sir', times, values, 'gaussian')
    truth = po.scores(po.Spec('sir', times, values, 'gaussian', tol=po.TRUTH_TOL), xs)
    got = T.evaluate(pb, pb.GaussianLogLikelihood(pb.MultiOutputProblem(pb.toy.SIRModel(), times, values)), xs)
    note('sir dev-truth', np.max(T.rel_err(got, truth)))
    note('sir ref-truth', np.max(T.rel_err(po.scores(spec, xs), truth)))
    # Lotka-Volterra, sum of squares
    times = np.linspace(0, 15, 151)
    x_true = [0.5, 0.15, 1.0, 0.3]
    values = po.lv_simulate(x_true, times) + rng.normal(0, 0.2, (151, 2))
    xs = np.array(x_true) * (1 + 0.1 * rng.randn(32, 4))
    spec = po.Spec('lotka_volterra', times, values, 'sse')
    truth = po.scores(po.Spec('lotka_volterra', times, values, 'sse', tol=po.TRUTH_TOL), xs)
    got = T.evaluate(pb, pb.SumOfSquaresError(pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), times, values)), xs)
    note('lv dev-truth', np.max(T.rel_err(got, truth)))
    note('lv ref-truth', np.max(T.rel_err(po.scores(spec, xs), truth)))
print('seeds ok; worst relative errors:', {k: '%.2e' % v for k, v in worst.items()})
