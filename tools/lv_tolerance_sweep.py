"""Lotka-Volterra log-likelihood error against a converged solve as a function
of the DP5(4) tolerance (the sweep behind the model's default of 5e-9,
pb200_api.cu: default_tolerance), on the reference-generated fixtures of
tests/golden/lotka_volterra.npz, with FitzHugh-Nagumo and SIR beside it."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
cases = [('lotka_volterra', pb.toy.LotkaVolterraModel, 'GaussianLogLikelihood'),
         ('fhn', pb.toy.FitzhughNagumoModel, 'GaussianKnownSigmaLogLikelihood'),
         ('sir', pb.toy.SIRModel, 'GaussianLogLikelihood')]
print('| model | tolerance | max rel. error vs truth | median | mean attempted steps | reference (LSODA 1.49e-8) max |')
print('|---|---|---|---|---|---|')
for name, cls, obj in cases:
    g = dict(np.load(os.path.join(G, name + '.npz')))
    problem = pb.MultiOutputProblem(cls(), g['times'], g['values'])
    f = pb.GaussianKnownSigmaLogLikelihood(problem, float(g['sigma'])) \
        if obj == 'GaussianKnownSigmaLogLikelihood' else pb.GaussianLogLikelihood(problem)
    truth, ref, xs = g['truth_scores'], g['scores'], g['xs']
    ok = np.isfinite(truth)
    ref_err = np.max(np.abs(ref[ok] - truth[ok]) / np.abs(truth[ok]))
    for tol in (None, 1.49012e-8, 1e-8, 5e-9, 3e-9, 1e-9):
        ev = pb.CudaEvaluator(f, as_list=False, rtol=tol, atol=tol)
        got, steps = ev.evaluate_array(xs, return_steps=True)
        e = np.abs(got[ok] - truth[ok]) / np.abs(truth[ok])
        print('| %s | %s | %.2e | %.2e | %.0f | %.2e |' % (
            name, 'default' if tol is None else '%.3g' % tol, e.max(), np.median(e),
            steps[ok & (steps > 0)].mean(), ref_err))
