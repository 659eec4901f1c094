#!/usr/bin/env python3
"""
Generates the golden fixtures in this directory by running the REAL reference
(pints-team/pints, imported from /root/reference or $PINTS_REFERENCE).

Run once in the build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so its outputs are frozen here as
small ``.npz`` files.  Nothing in this script uses the oracle or the CUDA path:
every number written is the reference's own.
"""
import os
import sys

import numpy as np

REF = os.environ.get('PINTS_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import pints            # noqa: E402
import pints.toy        # noqa: E402
from scipy.integrate import odeint  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TRUTH = dict(rtol=1e-13, atol=1e-13, mxstep=100000)


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-24s %8.1f KB' % (name + '.npz', os.path.getsize(path) / 1024.0))


def gauss_ll_from(sim, values, sigma):
    """GaussianLogLikelihood of a given (tight-tolerance) trajectory, written
    with the reference's own expression (pints/_log_likelihoods.py:1127-1129)
    so that only the trajectory differs."""
    sigma = np.asarray(sigma)
    nt = len(values)
    logn = 0.5 * nt * np.log(2 * np.pi)
    error = values - sim
    return np.sum(- logn - nt * np.log(sigma)
                  - np.sum(error**2, axis=0) / (2 * sigma**2))


# --------------------------------------------------------------------------
# the five benchmark problems of BASELINE.json / SURVEY.md section 8(d)
# --------------------------------------------------------------------------

def fhn_problem():
    model = pints.toy.FitzhughNagumoModel()
    times = np.linspace(0, 20, 1000)
    np.random.seed(1)
    values = model.simulate([0.1, 0.5, 3], times) \
        + pints.noise.independent(0.5, (1000, 2))
    problem = pints.MultiOutputProblem(model, times, values)
    return model, times, values, problem


def golden_fhn():
    model, times, values, problem = fhn_problem()
    ll = pints.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    rng = np.random.RandomState(2)
    xs = np.array([0.1, 0.5, 3]) * (1 + 0.05 * rng.randn(96, 3))
    scores = np.array(pints.SequentialEvaluator(ll).evaluate(xs))
    # ProbabilityBasedError is what OptimisationController hands the evaluator
    neg = np.array(pints.SequentialEvaluator(
        pints.ProbabilityBasedError(ll)).evaluate(xs[:8]))
    traj = np.array([model.simulate(x, times) for x in xs[:12]])
    truth = np.array([odeint(model._rhs, model._y0, times, (x,), **TRUTH)
                      for x in xs[:12]])
    truth_scores = np.array([
        np.sum(ll._offset + ll._multip * np.sum((values - odeint(
            model._rhs, model._y0, times, (x,), **TRUTH))**2, axis=0))
        for x in xs])
    # wide box (divergent cost), still inside the non-stiff region
    rngw = np.random.RandomState(3)
    xw = rngw.uniform([0.0, 0.0, 1.0], [1.0, 1.0, 5.0], size=(48, 3))
    wide = np.array(pints.SequentialEvaluator(ll).evaluate(xw))
    wide_truth = np.array([
        np.sum(ll._offset + ll._multip * np.sum((values - odeint(
            model._rhs, model._y0, times, (x,), **TRUTH))**2, axis=0))
        for x in xw])
    # sum of squares and unknown-sigma likelihood on the same series
    sse = pints.SumOfSquaresError(problem, weights=[1.0, 2.5])
    sse_scores = np.array(pints.SequentialEvaluator(sse).evaluate(xs[:16]))
    gl = pints.GaussianLogLikelihood(problem)
    xg = np.hstack([xs[:16], 0.5 * (1 + 0.1 * rng.randn(16, 2))])
    xg[3, 3] = -0.1            # sigma <= 0 -> -inf without simulating
    xg[5, 4] = 0.0
    gl_scores = np.array(pints.SequentialEvaluator(gl).evaluate(xg))
    save('fhn', times=times, values=values, xs=xs, scores=scores,
         neg_scores=neg, traj=traj, truth=truth, truth_scores=truth_scores,
         xs_wide=xw, scores_wide=wide, truth_scores_wide=wide_truth,
         sse_weights=np.array([1.0, 2.5]), sse_scores=sse_scores,
         xs_gauss=xg, gauss_scores=gl_scores, sigma=np.array(0.5))


def golden_fhn_unit_test():
    """Inputs of pints/tests/test_toy_fitzhugh_nagumo_model.py:51-59."""
    model = pints.toy.FitzhughNagumoModel([-2, 1.5])
    times = np.linspace(0, 20, 201)
    x = [0.2, 0.4, 2.5]
    values = model.simulate(x, times)
    truth = odeint(model._rhs, model._y0, times, (x,), **TRUTH)
    # times not starting at zero: t = 0 is prepended and dropped
    t2 = np.linspace(0.5, 20, 40)
    v2 = model.simulate(x, t2)
    save('fhn_unit', y0=np.array([-2, 1.5]), times=times, x=np.array(x),
         values=values, truth=truth, times_nz=t2, values_nz=v2)


def golden_logistic():
    model = pints.toy.LogisticModel()
    times = np.linspace(0, 100, 1000)
    np.random.seed(1)
    values = model.simulate([0.1, 50], times) + np.random.normal(0, 1, 1000)
    problem = pints.SingleOutputProblem(model, times, values)
    sse = pints.SumOfSquaresError(problem)
    rng = np.random.RandomState(4)
    xs = np.array([0.1, 50]) * (1 + 0.2 * rng.randn(64, 2))
    xs[7] = [0.1, -1.0]          # k < 0 -> all zeros
    xs[9] = [0.0, 50.0]          # r = 0
    xs[11] = [-0.05, 50.0]       # negative rate
    xs[13] = [0.1, 2.0]          # k == p0 -> c == 0
    xs[15] = [0.1, 0.0]          # k == 0
    scores = np.array(pints.SequentialEvaluator(sse).evaluate(xs))
    traj = np.array([model.simulate(x, times) for x in xs[:16]])
    ll = pints.GaussianKnownSigmaLogLikelihood(problem, 1.0)
    ll_scores = np.array(pints.SequentialEvaluator(ll).evaluate(xs))
    gl = pints.GaussianLogLikelihood(problem)
    xg = np.hstack([xs, 1.0 + 0.1 * rng.rand(64, 1)])
    xg[2, 2] = -1.0
    gl_scores = np.array(pints.SequentialEvaluator(gl).evaluate(xg))
    # p0 == 0 edge case and the reference unit test's exact value
    # (pints/tests/test_toy_logistic_model.py:93-105)
    m0 = pints.toy.LogisticModel(0)
    z = m0.simulate([0.1, 50], times[:10])
    m5 = pints.toy.LogisticModel(0.5)
    tu = np.linspace(0, 20, 21)
    vu = m5.simulate([0.1, 12], tu)
    save('logistic', times=times, values=values, xs=xs, scores=scores,
         traj=traj, ll_scores=ll_scores, xs_gauss=xg, gauss_scores=gl_scores,
         p0zero=z, unit_times=tu, unit_values=vu)


def golden_hh():
    model = pints.toy.HodgkinHuxleyIKModel()
    times = model.suggested_times()
    p = np.array(model.suggested_parameters(), dtype=float)
    np.random.seed(1)
    values = model.simulate(p, times) + np.random.normal(0, 40, times.shape)
    problem = pints.SingleOutputProblem(model, times, values)
    gl = pints.GaussianLogLikelihood(problem)
    rng = np.random.RandomState(5)
    xs = np.hstack([p * (1 + 0.05 * rng.randn(48, 5)),
                    40 * (1 + 0.05 * rng.randn(48, 1))])
    xs[4, 5] = -3.0
    scores = np.array(pints.SequentialEvaluator(gl).evaluate(xs))
    traj = np.array([model.simulate(x[:5], times) for x in xs[:6]])
    # samples exactly on event times and past the end of the protocol
    te = np.array([0, 45, 90, 95.5, 100, 190, 200, 1190, 1199.75, 1200,
                   1250.5, 1300], dtype=float)
    ve = np.array([model.simulate(x[:5], te) for x in xs[:6]])
    # 1000-point variant used for like-for-like throughput
    t1k = np.linspace(0, 1200, 1000)
    np.random.seed(2)
    v1k = model.simulate(p, t1k) + np.random.normal(0, 40, t1k.shape)
    gl1k = pints.GaussianLogLikelihood(
        pints.SingleOutputProblem(model, t1k, v1k))
    s1k = np.array(pints.SequentialEvaluator(gl1k).evaluate(xs[:16]))
    save('hh_ik', times=times, values=values, xs=xs, scores=scores, traj=traj,
         event_times=te, event_values=ve, times_1k=t1k, values_1k=v1k,
         scores_1k=s1k, events=model._events.astype(float),
         voltages=model._voltages.astype(float))


def golden_lv():
    model = pints.toy.LotkaVolterraModel()
    times = np.linspace(0, 20, 1000)
    p = model.suggested_parameters()
    np.random.seed(1)
    clean = model.simulate(p, times)
    values = clean + np.random.normal(0, 0.1, clean.shape)
    problem = pints.MultiOutputProblem(model, times, values)
    gl = pints.GaussianLogLikelihood(problem)
    rng = np.random.RandomState(6)
    xs = np.hstack([p * (1 + 0.05 * rng.randn(48, 4)),
                    0.1 * (1 + 0.1 * rng.randn(48, 2))])
    scores = np.array(pints.SequentialEvaluator(gl).evaluate(xs))
    traj = np.array([model.simulate(x[:4], times) for x in xs[:8]])
    truth = np.array([odeint(model._rhs, model._y0, times, (x[:4],), **TRUTH)
                      for x in xs[:8]])
    truth_scores = np.array([gauss_ll_from(
        odeint(model._rhs, model._y0, times, (x[:4],), **TRUTH), values,
        x[4:]) for x in xs])
    # reference unit test (pints/tests/test_toy_lotka_volterra_model.py:42-55)
    mu = pints.toy.LotkaVolterraModel([3, 5])
    tu = np.linspace(0, 5, 101)
    vu = mu.simulate([1, 2, 2, 0.5], tu)
    # the real 21-point hare/lynx series, on log scale as in the example
    ts = model.suggested_times()
    vs = np.log(model.suggested_values())
    gls = pints.GaussianLogLikelihood(pints.MultiOutputProblem(model, ts, vs))
    ss = np.array(pints.SequentialEvaluator(gls).evaluate(xs[:16]))
    save('lotka_volterra', times=times, values=values, xs=xs, scores=scores,
         traj=traj, truth=truth, truth_scores=truth_scores, unit_times=tu, unit_values=vu,
         unit_y0=np.array([3.0, 5.0]), unit_x=np.array([1, 2, 2, 0.5]),
         small_times=ts, small_values=vs, small_scores=ss)


def golden_sir():
    model = pints.toy.SIRModel()
    times = np.linspace(1, 21, 1000)
    p = np.array(model.suggested_parameters(), dtype=float)
    np.random.seed(1)
    clean = model.simulate(p, times)
    values = clean + np.random.normal(0, 1, clean.shape)
    problem = pints.MultiOutputProblem(model, times, values)
    gl = pints.GaussianLogLikelihood(problem)
    rng = np.random.RandomState(7)
    xs = np.hstack([p * (1 + 0.05 * rng.randn(48, 3)),
                    1.0 * (1 + 0.1 * rng.randn(48, 2))])
    xs[0, 2] = 38.9              # truncated to 38 by the integer y0
    xs[1, 2] = 37.2
    scores = np.array(pints.SequentialEvaluator(gl).evaluate(xs))
    traj = np.array([model.simulate(x[:3], times) for x in xs[:8]])

    def truth_one(x):
        y0 = np.array([int(x[2]), 1.0, 0.0])
        return odeint(model._rhs, y0, times, (x[0], x[1]), **TRUTH)[:, 1:]
    truth = np.array([truth_one(x) for x in xs[:8]])
    truth_scores = np.array([gauss_ll_from(truth_one(x), values, x[3:])
                             for x in xs])
    # float y0 (no truncation) and the reference unit test
    # (pints/tests/test_toy_sir_model.py:48-60)
    mf = pints.toy.SIRModel([100, 10, 1])
    tu = np.linspace(0, 10, 101)
    vu = mf.simulate([0.05, 0.4, 100], tu)
    vf = mf.simulate([0.05, 0.4, 77.7], tu)
    ts = np.asarray(model.suggested_times(), dtype=float)
    vs = np.asarray(model.suggested_values(), dtype=float)
    gls = pints.GaussianLogLikelihood(pints.MultiOutputProblem(model, ts, vs))
    ss = np.array(pints.SequentialEvaluator(gls).evaluate(xs[:16]))
    save('sir', times=times, values=values, xs=xs, scores=scores, traj=traj,
         truth=truth, truth_scores=truth_scores, unit_times=tu, unit_values=vu, unit_values_frac=vf,
         unit_y0=np.array([100.0, 10.0, 1.0]), small_times=ts,
         small_values=vs, small_scores=ss)


def golden_constant():
    """The likelihood known answers of the reference's tests are stated on
    ConstantModel (pints/tests/test_log_likelihoods.py:1926-1961, :2100-2171;
    pints/tests/test_error_measures.py:738-827)."""
    out = {}
    m1 = pints.toy.ConstantModel(1)
    t = np.asarray([1, 2, 3], dtype=float)
    p1 = pints.SingleOutputProblem(m1, t, [1, 1, 1])
    out['single_times'] = t
    out['single_values'] = np.array([1.0, 1.0, 1.0])
    out['single_sse_x'] = np.array([[1.0], [2.0], [3.0], [0.5]])
    out['single_sse'] = np.array(
        [pints.SumOfSquaresError(p1)(x) for x in out['single_sse_x']])
    m3 = pints.toy.ConstantModel(3)
    v3 = np.array([[1, 4, 9], [2, 3, 8], [0.5, 5, 10]], dtype=float)
    p3 = pints.MultiOutputProblem(m3, t, v3)
    x3 = np.array([[1.0, 2.0, 3.0], [0.0, 1.5, 2.5], [2.0, 2.0, 2.0]])
    out['multi_values'] = v3
    out['multi_x'] = x3
    out['multi_sse'] = np.array([pints.SumOfSquaresError(p3)(x) for x in x3])
    out['multi_sse_weighted'] = np.array(
        [pints.SumOfSquaresError(p3, weights=[1, 2, 3])(x) for x in x3])
    out['multi_sigma'] = np.array([3.5, 1.0, 12.0])
    gk = pints.GaussianKnownSigmaLogLikelihood(p3, out['multi_sigma'])
    out['multi_known'] = np.array([gk(x) for x in x3])
    gl = pints.GaussianLogLikelihood(p3)
    xg = np.hstack([x3, np.array([[3.5, 1, 12], [1, 2, 3], [1, -1, 2.0]])])
    out['multi_gauss_x'] = xg
    out['multi_gauss'] = np.array([gl(x) for x in xg])
    save('constant', **out)


# --------------------------------------------------------------------------
# sampler replays: every draw and decision of the reference, step by step
# --------------------------------------------------------------------------

def _golden_ode(name, model, times, true_x, sigma, spread, truth_one, seed):
    """Trajectories and GaussianLogLikelihood scores of one more ODE toy
    model, with tight-tolerance counterparts."""
    rng = np.random.RandomState(seed)
    clean = np.asarray(model.simulate(true_x, times))
    noisy = clean + rng.normal(0, sigma, clean.shape)
    multi = clean.ndim == 2
    problem = (pints.MultiOutputProblem if multi
               else pints.SingleOutputProblem)(model, times, noisy)
    ll = pints.GaussianLogLikelihood(problem)
    no = problem.n_outputs()
    n = 48
    xs = np.array(true_x, dtype=float) * (1 + spread * rng.randn(n, len(true_x)))
    xg = np.hstack([xs, sigma * (1 + 0.1 * rng.rand(n, no))])
    scores = np.array(pints.SequentialEvaluator(ll).evaluate(xg))
    traj = np.array([model.simulate(x, times) for x in xs[:8]])
    truth = np.array([truth_one(x) for x in xs])
    truth_scores = np.array([
        gauss_ll_from(t, noisy, xg[k, -no:]) for k, t in enumerate(truth)])
    save(name, times=times, values=noisy, xs=xg, scores=scores, traj=traj,
         truth=truth[:8], truth_scores=truth_scores)


def golden_goodwin():
    model = pints.toy.GoodwinOscillatorModel()
    times = np.linspace(0, 100, 400)

    def truth_one(x):
        return odeint(model._rhs, model._y0, times, (x,), **TRUTH)
    _golden_ode('goodwin', model, times, model.suggested_parameters(), 0.02,
                0.02, truth_one, 31)


def golden_hes1():
    model = pints.toy.Hes1Model()
    times = np.linspace(0, 270, 300)

    def truth_one(x):
        return odeint(model._rhs, model._y0, times, (x,), **TRUTH)[:, 0]
    _golden_ode('hes1', model, times, model.suggested_parameters(), 0.2, 0.03,
                truth_one, 32)


def golden_repressilator():
    model = pints.toy.RepressilatorModel()
    times = np.linspace(0, 40, 400)

    def truth_one(x):
        return odeint(model._rhs, model._y0, times, tuple(x), **TRUTH)[:, :3]
    _golden_ode('repressilator', model, times, [1, 1000, 5, 2], 2.0, 0.02,
                truth_one, 33)


def golden_beeler_reuter():
    """Beeler-Reuter action potential (pints/toy/_beeler_reuter_model.py): the
    reference at its default (1e-4, 1e-6) tolerances and a tight-tolerance
    counterpart; per-output noise levels (mV and mM differ by 1e7)."""
    model = pints.toy.ActionPotentialModel()
    times = model.suggested_times()
    y0 = [model._v0, model._cai0, model._m0, model._h0, model._j0, model._d0,
          model._f0, model._x10]

    def truth_one(x):
        # Ca_i is 1e-7 .. 7e-6: the truth run uses relative control only
        return odeint(model._rhs, y0, times, (x,), hmax=2.0, rtol=1e-12,
                      atol=1e-20, mxstep=1000000)[:, :2]
    _golden_ode('beeler_reuter', model, times, model.suggested_parameters(),
                np.array([1.0, 2e-7]), 0.02, truth_one, 34)


def golden_composed_prior():
    """LogPosterior with a ComposedLogPrior (uniform box on the first two FHN
    parameters, Gaussian on the third and on sigma_1, uniform on sigma_2)."""
    model, times, values, problem = fhn_problem()
    ll = pints.GaussianLogLikelihood(problem)
    prior = pints.ComposedLogPrior(
        pints.UniformLogPrior([0.05, 0.3], [0.15, 0.7]),
        pints.GaussianLogPrior(3.0, 0.2),
        pints.GaussianLogPrior(0.5, 0.05),
        pints.UniformLogPrior([0.4], [0.6]))
    post = pints.LogPosterior(ll, prior)
    rng = np.random.RandomState(41)
    xs = np.array([0.1, 0.5, 3, 0.5, 0.5]) * (1 + 0.05 * rng.randn(40, 5))
    xs[3, 0] = 0.2          # outside the first box
    xs[5, 4] = 0.7          # outside the last box
    xs[7, 3] = -0.1         # sigma <= 0, inside the Gaussian's support
    scores = np.array(pints.SequentialEvaluator(post).evaluate(xs))
    priors = np.array([prior(x) for x in xs])
    # logistic (closed form: 1e-12) with a single Gaussian prior per parameter
    lmodel = pints.toy.LogisticModel()
    ltimes = np.linspace(0, 100, 200)
    lvalues = lmodel.simulate([0.1, 50], ltimes) + rng.normal(0, 1, 200)
    lproblem = pints.SingleOutputProblem(lmodel, ltimes, lvalues)
    lpost = pints.LogPosterior(
        pints.GaussianKnownSigmaLogLikelihood(lproblem, 1.0),
        pints.ComposedLogPrior(pints.GaussianLogPrior(0.1, 0.02),
                               pints.GaussianLogPrior(50, 5)))
    lxs = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(32, 2))
    lscores = np.array(pints.SequentialEvaluator(lpost).evaluate(lxs))
    save('composed_prior', xs=xs, scores=scores, priors=priors,
         ltimes=ltimes, lvalues=lvalues, lxs=lxs, lscores=lscores)


def golden_error_measures():
    """MeanSquaredError / RootMeanSquaredError / NormalisedRootMeanSquaredError
    of the reference on the logistic, FHN and HH-IK problems."""
    model = pints.toy.LogisticModel()
    times = np.linspace(0, 100, 1000)
    np.random.seed(1)
    values = model.simulate([0.1, 50], times) + np.random.normal(0, 1, 1000)
    problem = pints.SingleOutputProblem(model, times, values)
    rng = np.random.RandomState(14)
    xs = np.array([0.1, 50]) * (1 + 0.2 * rng.randn(48, 2))
    ev = pints.SequentialEvaluator
    out = dict(times=times, values=values, xs=xs)
    out['mse'] = np.array(ev(pints.MeanSquaredError(problem)).evaluate(xs))
    out['mse_w'] = np.array(
        ev(pints.MeanSquaredError(problem, [2.5])).evaluate(xs))
    out['rmse'] = np.array(
        ev(pints.RootMeanSquaredError(problem)).evaluate(xs))
    out['nrmse'] = np.array(
        ev(pints.NormalisedRootMeanSquaredError(problem)).evaluate(xs))
    # multi-output, weighted, on the ODE model
    fmodel, ftimes, fvalues, fproblem = fhn_problem()
    fx = np.array([0.1, 0.5, 3]) * (1 + 0.05 * rng.randn(24, 3))
    out['fhn_xs'] = fx
    out['fhn_mse'] = np.array(
        ev(pints.MeanSquaredError(fproblem, [1.0, 0.25])).evaluate(fx))
    truth = np.array([odeint(fmodel._rhs, fmodel._y0, ftimes, (x,), **TRUTH)
                      for x in fx])
    ninv = 1.0 / np.prod(fvalues.shape)
    out['fhn_mse_truth'] = np.array(
        [ninv * np.sum(np.array([1.0, 0.25])
                       * np.sum((t - fvalues)**2, axis=0), axis=0)
         for t in truth])
    save('error_measures', **out)


def _logistic_posterior():
    model = pints.toy.LogisticModel()
    times = np.linspace(0, 100, 50)
    np.random.seed(11)
    values = model.simulate([0.1, 50], times) + np.random.normal(0, 2, 50)
    problem = pints.SingleOutputProblem(model, times, values)
    return pints.GaussianLogLikelihood(problem), times, values


def _golden_adaptive(cls, name, seed, n_iter=160, warm=40):
    """A single-chain sampler class of the reference, several independent
    chains driven exactly as MCMCController does
    (pints/_mcmc/__init__.py:706-788): all asks, then all tells; adaptation
    (where the class has an initial phase) switched on after `warm` iterations.
    """
    ll, times, values = _logistic_posterior()
    n_chains, d = 6, 3
    rng = np.random.RandomState(21)
    x0 = np.array([0.1, 50, 2]) * (1 + 0.02 * rng.randn(n_chains, d))
    np.random.seed(seed)
    samplers = [cls(x) for x in x0]
    adaptive_class = samplers[0].needs_initial_phase()
    rec = dict(
        proposed=np.zeros((n_iter, n_chains, d)),
        fx=np.zeros((n_iter, n_chains)),
        log_u=np.full((n_iter, n_chains), np.nan),
        accepted=np.zeros((n_iter, n_chains), dtype=np.uint8),
        current=np.zeros((n_iter, n_chains, d)),
        current_f=np.zeros((n_iter, n_chains)),
        mu=np.zeros((n_iter, n_chains, d)),
        sigma=np.zeros((n_iter, n_chains, d, d)),
        log_lambda=np.zeros((n_iter, n_chains)),
        gamma=np.zeros(n_iter),
        adaptive=np.zeros(n_iter, dtype=np.uint8),
        z=np.full((n_iter, n_chains, d), np.nan),
    )
    for it in range(n_iter):
        if it == warm and adaptive_class:
            for s in samplers:
                s.set_initial_phase(False)
        rec['adaptive'][it] = it >= warm and adaptive_class
        # ask: record the standard normals each multivariate_normal consumes
        xs = []
        for c, s in enumerate(samplers):
            if it > 0:
                probe = np.random.RandomState()
                probe.set_state(np.random.get_state())
                rec['z'][it, c] = probe.standard_normal(d)
            xs.append(s.ask())
        rec['proposed'][it] = np.array(xs)
        fxs = [ll(x) for x in xs]
        # make one proposal non-finite to exercise the "no uniform drawn" rule
        if it == 7:
            fxs[2] = -np.inf
        if it == 9:
            fxs[4] = np.nan
        rec['fx'][it] = fxs
        for c, s in enumerate(samplers):
            if it > 0 and np.isfinite(fxs[c]):
                probe = np.random.RandomState()
                probe.set_state(np.random.get_state())
                rec['log_u'][it, c] = np.log(probe.uniform(0, 1))
            cur, cur_f, acc = s.tell(fxs[c])
            rec['accepted'][it, c] = acc
            rec['current'][it, c] = cur
            rec['current_f'][it, c] = cur_f
            if adaptive_class:
                rec['mu'][it, c] = s._mu
                rec['sigma'][it, c] = s._sigma
                rec['log_lambda'][it, c] = getattr(s, '_log_lambda', 0.0)
        rec['gamma'][it] = getattr(samplers[0], '_gamma', 0.0)
    save(name, x0=x0, times=times, values=values, warm=np.array(warm),
         sigma0=np.array([s._sigma0 for s in samplers]), **rec)


def golden_hbac():
    """HaarioBardenetACMC (pints/_mcmc/_haario_bardenet_ac.py)."""
    _golden_adaptive(pints.HaarioBardenetACMC, 'hbac_replay', 12)


def golden_haario():
    """HaarioACMC (pints/_mcmc/_haario_ac.py)."""
    _golden_adaptive(pints.HaarioACMC, 'haario_replay', 14)


def golden_rao_blackwell():
    """RaoBlackwellACMC (pints/_mcmc/_rao_blackwell_ac.py)."""
    _golden_adaptive(pints.RaoBlackwellACMC, 'rao_blackwell_replay', 15)


def golden_metropolis():
    """MetropolisRandomWalkMCMC (pints/_mcmc/_metropolis.py)."""
    _golden_adaptive(pints.MetropolisRandomWalkMCMC, 'metropolis_replay', 16,
                     n_iter=120)


def golden_de():
    """DifferentialEvolutionMCMC (default settings), one multi-chain sampler."""
    ll, times, values = _logistic_posterior()
    n, n_iter, d = 8, 45, 3
    rng = np.random.RandomState(22)
    x0 = np.array([0.1, 50, 2]) * (1 + 0.02 * rng.randn(n, d))
    np.random.seed(13)
    s = pints.DifferentialEvolutionMCMC(n, x0)
    rec = dict(
        proposed=np.zeros((n_iter, n, d)), fx=np.zeros((n_iter, n)),
        eps=np.zeros((n_iter, n, d)), r1=np.zeros((n_iter, n), dtype=np.int32),
        r2=np.zeros((n_iter, n), dtype=np.int32), gamma=np.zeros(n_iter),
        log_u=np.full((n_iter, n), np.nan),
        accepted=np.zeros((n_iter, n), dtype=np.uint8),
        current=np.zeros((n_iter, n, d)), current_f=np.zeros((n_iter, n)))
    for it in range(n_iter):
        if it > 0:
            probe = np.random.RandomState()
            probe.set_state(np.random.get_state())
            gamma = 1.0 if (s._iter_count % s._gamma_switch_rate == 0) \
                else s._gamma
            rec['gamma'][it] = gamma
            for j in range(n):
                rec['eps'][it, j] = probe.normal(0, s._b_star, (d,))
                others = list(range(n))
                others.pop(j)
                r1, r2 = probe.choice(others, 2, replace=False)
                rec['r1'][it, j], rec['r2'][it, j] = r1, r2
        xs = s.ask()
        rec['proposed'][it] = xs
        fxs = np.array([ll(x) for x in xs])
        if it == 5:
            fxs[1] = np.nan           # NaN compares false -> rejected
        if it == 6:
            fxs[3] = -np.inf
        rec['fx'][it] = fxs
        if it > 0:
            probe = np.random.RandomState()
            probe.set_state(np.random.get_state())
            rec['log_u'][it] = np.log(probe.uniform(size=n))
        cur, cur_f, acc = s.tell(fxs)
        rec['accepted'][it] = acc
        rec['current'][it] = cur
        rec['current_f'][it] = cur_f
    save('de_replay', x0=x0, times=times, values=values,
         b_star=np.asarray(s._b_star), **rec)


if __name__ == '__main__':
    print('reference:', pints.__file__, pints.__version__)
    only = sys.argv[1:]
    if only:
        for name in only:
            globals()['golden_' + name]()
        sys.exit(0)
    golden_fhn()
    golden_fhn_unit_test()
    golden_logistic()
    golden_hh()
    golden_lv()
    golden_sir()
    golden_constant()
    golden_goodwin()
    golden_hes1()
    golden_repressilator()
    golden_beeler_reuter()
    golden_error_measures()
    golden_composed_prior()
    golden_hbac()
    golden_haario()
    golden_rao_blackwell()
    golden_metropolis()
    golden_de()
