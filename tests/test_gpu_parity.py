"""
Parity of the CUDA path (through the C ABI, ``pints_b200._lib`` ->
``libpints_b200.so``) against

  * the golden fixtures produced by the real reference (tests/golden), and
  * the CPU oracle run live on the same seeded inputs.

Tolerances (BASELINE.json north_star):
  closed-form models   1e-12 relative on scores and trajectories
  ODE models           scores 1e-6 relative; trajectories |d| <= 1e-5 + 1e-6|y|
                       against the reference's LSODA; and the device must be
                       at least as close to a tight-tolerance solve ("truth") as
                       the reference itself is.
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pints_oracle as po

pytestmark = pytest.mark.gpu

torch = pytest.importorskip('torch')

ANALYTIC_RTOL = 1e-12
ODE_SCORE_RTOL = 1e-6
TRAJ_ATOL, TRAJ_RTOL = 1e-5, 1e-6


@pytest.fixture(scope='module')
def pb():
    import pints_b200
    pints_b200.require_cuda()
    return pints_b200


def rel_err(got, want):
    got, want = np.asarray(got, float), np.asarray(want, float)
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    with np.errstate(invalid='ignore', divide='ignore'):
        r = np.abs(got - want) / np.abs(want)
    r[same] = 0.0
    return r


def evaluate(pb, f, xs, **kw):
    return pb.CudaEvaluator(f, as_list=False, **kw).evaluate(xs)


# --------------------------------------------------------------- closed form

def test_logistic_scores_and_trajectories(pb):
    g = load_golden('logistic')
    model = pb.toy.LogisticModel()
    problem = pb.SingleOutputProblem(model, g['times'], g['values'])
    got = evaluate(pb, pb.SumOfSquaresError(problem), g['xs'])
    assert np.nanmax(rel_err(got, g['scores'])) < ANALYTIC_RTOL
    assert np.isnan(got[15]) and np.isnan(g['scores'][15])
    got = evaluate(pb, pb.GaussianKnownSigmaLogLikelihood(problem, 1.0),
                   g['xs'])
    assert np.nanmax(rel_err(got, g['ll_scores'])) < ANALYTIC_RTOL
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs_gauss'])
    assert got[2] == -np.inf
    assert np.nanmax(rel_err(got, g['gauss_scores'])) < ANALYTIC_RTOL
    traj = model.simulate(g['xs'][:16], g['times'])
    assert traj.shape == (16, 1000)
    assert np.nanmax(rel_err(traj, g['traj'])) < ANALYTIC_RTOL
    # single-vector call keeps the reference's shape
    one = model.simulate(g['xs'][0], g['times'])
    assert one.shape == (1000,)
    assert np.array_equal(one, traj[0])
    # reference unit test values (pints/tests/test_toy_logistic_model.py:93-105)
    v = pb.toy.LogisticModel(0.5).simulate([0.1, 12], g['unit_times'])
    assert np.max(rel_err(v, g['unit_values'])) < ANALYTIC_RTOL
    assert abs(v[1] - 0.5501745273651227) < 1e-15
    assert np.all(pb.toy.LogisticModel(0).simulate([0.1, 50], [0, 1, 2]) == 0)
    with pytest.raises(ValueError):
        model.simulate([0.1, 50], [-1, 0, 1])


def test_constant_model_reference_known_answers(pb):
    g = load_golden('constant')
    t = g['single_times']
    p1 = pb.SingleOutputProblem(pb.toy.ConstantModel(1), t, g['single_values'])
    got = evaluate(pb, pb.SumOfSquaresError(p1), g['single_sse_x'])
    assert list(got) == [0.0, 3.0, 12.0, 0.75]
    p3 = pb.MultiOutputProblem(pb.toy.ConstantModel(3), t, g['multi_values'])
    assert np.array_equal(evaluate(pb, pb.SumOfSquaresError(p3), g['multi_x']),
                          g['multi_sse'])
    got = evaluate(pb, pb.SumOfSquaresError(p3, weights=[1, 2, 3]),
                   g['multi_x'])
    assert np.max(rel_err(got, g['multi_sse_weighted'])) < ANALYTIC_RTOL
    got = evaluate(pb, pb.GaussianKnownSigmaLogLikelihood(
        p3, g['multi_sigma']), g['multi_x'])
    assert np.max(rel_err(got, g['multi_known'])) < ANALYTIC_RTOL
    got = evaluate(pb, pb.GaussianLogLikelihood(p3), g['multi_gauss_x'])
    assert got[2] == -np.inf
    assert np.max(rel_err(got, g['multi_gauss'])) < ANALYTIC_RTOL
    # pints/tests/test_log_likelihoods.py:1909-1961
    data_multi = np.array([[10.7, 3.5, 3.8], [1.1, 3.2, -1.4],
                           [9.3, 0.0, 4.5], [1.2, -3, -10]])
    times = [1, 2, 3, 4]
    ll = pb.GaussianKnownSigmaLogLikelihood(pb.SingleOutputProblem(
        pb.toy.ConstantModel(1), times, np.arange(1, 5) / 5.0), 1.5)
    assert abs(ll([-1]) - -7.3420590096957925) < 1e-12
    ll = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(
        pb.toy.ConstantModel(3), times, data_multi), 1)
    assert abs(ll([0, 0, 0]) - -196.9122623984561) < 1e-10


def test_hh_ik(pb):
    g = load_golden('hh_ik')
    model = pb.toy.HodgkinHuxleyIKModel()
    assert np.array_equal(model._events, g['events'])
    assert np.array_equal(model._voltages, g['voltages'])
    traj = model.simulate(g['xs'][:6, :5], g['times'])
    assert traj.shape == (6, 4800)
    assert np.max(rel_err(traj, g['traj'])) < ANALYTIC_RTOL
    # samples on event times / past the end of the protocol
    ev = model.simulate(g['xs'][:6, :5], g['event_times'])
    assert np.max(rel_err(ev, g['event_values'])) < ANALYTIC_RTOL
    problem = pb.SingleOutputProblem(model, g['times'], g['values'])
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'])
    assert got[4] == -np.inf
    assert np.max(rel_err(got, g['scores'])) < ANALYTIC_RTOL
    problem = pb.SingleOutputProblem(model, g['times_1k'], g['values_1k'])
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'][:16])
    assert np.max(rel_err(got, g['scores_1k'])) < ANALYTIC_RTOL
    # the decay recurrence of the kernel needs equally spaced points inside a
    # segment: irregular and repeated times, a grid whose spacing is not exact
    # in binary, and segments with fewer than 32 points take the other paths
    rng = np.random.RandomState(3)
    for t in (np.sort(np.concatenate([rng.uniform(0, 1300, 700),
                                      [0, 0, 90, 90, 1200.0]])),
              np.linspace(0, 1199.7, 3999), np.arange(0, 1200, 20.0),
              np.arange(0.1, 1200, 0.3)):
        want = np.array([po.hh_ik_simulate(x, t) for x in g['xs'][:4, :5]])
        got = model.simulate(g['xs'][:4, :5], t)
        assert got.shape == want.shape
        assert np.max(rel_err(got, want)) < ANALYTIC_RTOL, len(t)
    # reference unit test (pints/tests/test_toy_hh_ik_model.py:44-74)
    v = model.simulate(model.suggested_parameters(), model.suggested_times())
    assert abs(v[0] - 3.790799999999997) < 1e-7
    assert abs(v[390] - 15.9405) < 1e-2
    with pytest.raises(ValueError):
        model.simulate(model.suggested_parameters(), [-1, 0, 1])


# ----------------------------------------------------------------- ODE models

def near(value, known):
    """Known answers of the reference's unit tests are LSODA outputs rounded
    to 5-6 places; compare with the stated trajectory tolerance."""
    return abs(value - known) <= TRAJ_ATOL + TRAJ_RTOL * abs(known)


def check_ode_scores(got, ref, truth, truth_rtol=ODE_SCORE_RTOL):
    """Log-likelihood parity for the ODE models.

    The reference integrates with LSODA at rtol = atol = 1.49e-8, and its own
    log-likelihoods deviate from a tight-tolerance solve ("truth") by up to
    8e-7 (FHN), 1.9e-5 (Lotka-Volterra) and 1.2e-7 (SIR) on these fixtures, so
    "1e-6 against the reference" is only meaningful where the reference itself
    is that accurate.  Asserted here, per vector:
      * device vs truth  <= truth_rtol (1e-6 unless stated), and
      * device vs reference <= truth_rtol + the reference's own error on that
        vector (so wherever the reference is exact to 1e-6 the two agree to
        2e-6, and to 1e-6 on the headline fixtures, asserted separately).
    """
    dev_truth = rel_err(got, truth)
    ref_truth = rel_err(ref, truth)
    dev_ref = rel_err(got, ref)
    assert np.max(dev_truth) < truth_rtol, np.max(dev_truth)
    assert np.all(dev_ref <= truth_rtol + ref_truth), \
        (np.max(dev_ref), np.max(ref_truth))
    return np.max(dev_truth), np.max(ref_truth), np.max(dev_ref)


def check_ode(traj_dev, traj_ref, truth, beats_reference=True):
    """Trajectory parity: within 1e-5 + 1e-6 |y| of the reference, within
    1e-6 (1 + |y|) of the truth, and (where stated) closer to the truth than
    the reference is."""
    d = np.abs(traj_dev - traj_ref)
    assert np.all(d <= TRAJ_ATOL + TRAJ_RTOL * np.abs(traj_ref)), d.max()
    e = np.abs(traj_dev - truth)
    assert np.all(e <= 1e-6 * (1 + np.abs(truth))), e.max()
    err_dev = e.max()
    err_ref = np.abs(traj_ref - truth).max()
    if beats_reference:
        assert err_dev <= err_ref * 1.05 + 1e-9, (err_dev, err_ref)
    return err_dev, err_ref


def test_fhn_headline(pb):
    g = load_golden('fhn')
    model = pb.toy.FitzhughNagumoModel()
    problem = pb.MultiOutputProblem(model, g['times'], g['values'])
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, float(g['sigma']))
    ev = pb.CudaEvaluator(ll, as_list=False)
    got, steps = ev.evaluate_array(g['xs'], return_steps=True)
    assert np.max(rel_err(got, g['scores'])) < ODE_SCORE_RTOL
    # closer to the tight-tolerance truth than the reference is
    check_ode_scores(got, g['scores'], g['truth_scores'])
    assert np.max(rel_err(got, g['truth_scores'])) <= \
        np.max(rel_err(g['scores'], g['truth_scores']))
    assert steps.min() > 200 and steps.max() < 1000
    # ProbabilityBasedError flips the sign, bit for bit
    neg = evaluate(pb, pb.ProbabilityBasedError(ll), g['xs'][:8])
    assert np.array_equal(neg, -got[:8])
    assert np.max(rel_err(neg, g['neg_scores'])) < ODE_SCORE_RTOL
    traj = model.simulate(g['xs'][:12], g['times'])
    assert traj.shape == (12, 1000, 2)
    check_ode(traj, g['traj'], g['truth'])
    # other objectives on the same series
    got = evaluate(pb, pb.SumOfSquaresError(problem, g['sse_weights']),
                   g['xs'][:16])
    assert np.max(rel_err(got, g['sse_scores'])) < ODE_SCORE_RTOL
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs_gauss'])
    assert got[3] == -np.inf and got[5] == -np.inf
    assert np.max(rel_err(got, g['gauss_scores'])) < ODE_SCORE_RTOL


def test_fhn_wide_box(pb):
    g = load_golden('fhn')
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), g['times'],
                                    g['values'])
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, float(g['sigma']))
    got = evaluate(pb, ll, g['xs_wide'])
    assert np.all(np.isfinite(got))
    # against the truth the device is tight everywhere in the box ...
    assert np.max(rel_err(got, g['truth_scores_wide'])) < ODE_SCORE_RTOL
    # ... and at least as good as the reference
    assert np.max(rel_err(got, g['truth_scores_wide'])) <= \
        np.max(rel_err(g['scores_wide'], g['truth_scores_wide'])) + 1e-9


def test_fhn_reference_unit_test(pb):
    # pints/tests/test_toy_fitzhugh_nagumo_model.py:51-59
    g = load_golden('fhn_unit')
    model = pb.toy.FitzhughNagumoModel(g['y0'])
    v = model.simulate(g['x'], g['times'])
    assert v.shape == (201, 2)
    assert v[0, 0] == -2 and v[0, 1] == 1.5
    assert near(v[200, 0], 1.675726)
    assert near(v[200, 1], -0.226142)
    check_ode(v, g['values'], g['truth'])
    # times not starting at zero: solve still starts at t = 0
    v = model.simulate(g['x'], g['times_nz'])
    assert np.all(np.abs(v - g['values_nz'])
                  <= TRAJ_ATOL + TRAJ_RTOL * np.abs(g['values_nz']))
    with pytest.raises(ValueError):
        model.simulate(g['x'], [-1, 0, 1])


def test_lotka_volterra(pb):
    g = load_golden('lotka_volterra')
    model = pb.toy.LotkaVolterraModel()
    traj = model.simulate(g['xs'][:8, :4], g['times'])
    check_ode(traj, g['traj'], g['truth'])
    problem = pb.MultiOutputProblem(model, g['times'], g['values'])
    f = pb.GaussianLogLikelihood(problem)
    got = evaluate(pb, f, g['xs'])
    # sigma = 0.1 amplifies trajectory error (the reference's LSODA is 1.9e-5
    # from the truth here); the default call -- Lotka-Volterra's default
    # tolerance is 5e-9, see default_tolerance() in pb200_api.cu -- is within
    # 1e-6 of the truth, with margin
    check_ode_scores(got, g['scores'], g['truth_scores'])
    assert np.max(rel_err(got, g['truth_scores'])) < 5e-7
    assert np.max(rel_err(got, g['truth_scores'])) < \
        0.1 * np.max(rel_err(g['scores'], g['truth_scores']))
    # the default is what an explicit 5e-9 gives
    tight = evaluate(pb, f, g['xs'], rtol=5e-9, atol=5e-9)
    assert np.array_equal(tight, got)
    # the 21-point hare/lynx series on log scale
    problem = pb.MultiOutputProblem(model, model.suggested_times(),
                                    np.log(model.suggested_values()))
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'][:16])
    truth = po.scores(po.Spec('lotka_volterra', g['small_times'],
                              g['small_values'], 'gaussian',
                              tol=po.TRUTH_TOL), g['xs'][:16])
    check_ode_scores(got, g['small_scores'], truth)
    # pints/tests/test_toy_lotka_volterra_model.py:42-55
    v = pb.toy.LotkaVolterraModel(g['unit_y0']).simulate(g['unit_x'],
                                                         g['unit_times'])
    assert v[0, 0] == 3 and v[0, 1] == 5
    assert near(v[1, 0], 1.929494)
    assert near(v[1, 1], 4.806542)
    assert near(v[100, 0], 1.277762)
    assert near(v[100, 1], 0.000529)


def test_sir(pb):
    g = load_golden('sir')
    model = pb.toy.SIRModel()
    traj = model.simulate(g['xs'][:8, :3], g['times'])
    check_ode(traj, g['traj'], g['truth'], beats_reference=False)
    # S0 = 38.9 is truncated to 38 by the default integer y0
    a = model.simulate([0.026, 0.285, 38.0], g['times'])
    b = model.simulate([0.026, 0.285, 38.9], g['times'])
    assert np.array_equal(a, b)
    problem = pb.MultiOutputProblem(model, g['times'], g['values'])
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'])
    check_ode_scores(got, g['scores'], g['truth_scores'])
    problem = pb.MultiOutputProblem(
        model, np.asarray(model.suggested_times(), float),
        np.asarray(model.suggested_values(), float))
    got = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'][:16])
    truth = po.scores(po.Spec('sir', g['small_times'], g['small_values'],
                              'gaussian', tol=po.TRUTH_TOL), g['xs'][:16])
    check_ode_scores(got, g['small_scores'], truth)
    # pints/tests/test_toy_sir_model.py:48-60, float y0 keeps S0 as given
    mf = pb.toy.SIRModel(g['unit_y0'])
    v = mf.simulate([0.05, 0.4, 100], g['unit_times'])
    assert v[0, 0] == 10 and v[0, 1] == 1
    assert near(v[1, 0], 15.61537)
    assert near(v[100, 1], 108.542466)
    vf = mf.simulate([0.05, 0.4, 77.7], g['unit_times'])
    assert np.all(np.abs(vf - g['unit_values_frac'])
                  <= TRAJ_ATOL + TRAJ_RTOL * np.abs(g['unit_values_frac']))


# ------------------------------------------- live oracle, seeded random inputs

def test_two_lanes_per_vector(pb):
    """The pair kernel (one lane per state component) solves the same problem:
    same parity bars, and agreement with the one-lane kernel far inside them."""
    g = load_golden('fhn')
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), g['times'],
                                    g['values'])
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, float(g['sigma']))
    one = evaluate(pb, ll, g['xs'], lanes_per_vector=1)
    two = evaluate(pb, ll, g['xs'], lanes_per_vector=2)
    assert np.max(rel_err(two, g['scores'])) < ODE_SCORE_RTOL
    check_ode_scores(two, g['scores'], g['truth_scores'])
    assert np.max(rel_err(two, one)) < 1e-7
    wide = evaluate(pb, ll, g['xs_wide'], lanes_per_vector=2)
    assert np.max(rel_err(wide, g['truth_scores_wide'])) < ODE_SCORE_RTOL
    g = load_golden('lotka_volterra')
    problem = pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), g['times'],
                                    g['values'])
    f = pb.GaussianLogLikelihood(problem)
    two = evaluate(pb, f, g['xs'], lanes_per_vector=2, rtol=5e-9, atol=5e-9)
    check_ode_scores(two, g['scores'], g['truth_scores'])
    # three-state model: the option is ignored, not an error
    g = load_golden('sir')
    problem = pb.MultiOutputProblem(pb.toy.SIRModel(), g['times'], g['values'])
    f = pb.GaussianLogLikelihood(problem)
    assert np.array_equal(evaluate(pb, f, g['xs'], lanes_per_vector=2),
                          evaluate(pb, f, g['xs'], lanes_per_vector=1))


def test_cost_sorted_scheduling_changes_nothing(pb):
    """pb200_eval_sorted groups similar vectors into warps; scores and step
    counts are per vector and must not depend on the grouping, bit for bit."""
    spec = po.headline_fhn_spec()
    xs = np.concatenate([po.headline_fhn_points(3000),
                         np.random.RandomState(3).uniform(
                             [0, 0, 1], [1, 1, 5], (1100, 3))])
    xs[17, 2] = -1.0                               # sigma-free, but c < 0: NaN path
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    a, sa = pb.CudaEvaluator(ll, sort=False).evaluate_array(xs, True)
    b, sb = pb.CudaEvaluator(ll, sort=True).evaluate_array(xs, True)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(sa, sb)
    c = pb.CudaEvaluator(ll, sort=True, lanes_per_vector=2,
                         chunk_rows=1 << 30).evaluate_array(xs)
    assert np.nanmax(rel_err(c, a)) < 1e-7
    # device-resident path, odd sizes, prior box deciding some rows early
    import torch
    prior = pb.UniformLogPrior([0, 0, 1.5], [1, 1, 4.5])
    post = pb.LogPosterior(ll, prior)
    for n in (1, 31, 1025, 4100):
        dp_s = pb.CudaEvaluator(post, sort=True).device_problem()
        dp_u = pb.CudaEvaluator(post, sort=False).device_problem()
        d_x = torch.from_numpy(np.ascontiguousarray(xs[:n])).cuda()
        u = dp_u.eval(d_x).cpu().numpy()
        v = dp_s.eval(d_x).cpu().numpy()
        assert np.array_equal(u, v, equal_nan=True), n
        assert np.isneginf(u).any() or n < 3001
    # other models take the generic key
    g = load_golden('sir')
    problem = pb.MultiOutputProblem(pb.toy.SIRModel(), g['times'], g['values'])
    f = pb.GaussianLogLikelihood(problem)
    assert np.array_equal(evaluate(pb, f, g['xs'], sort=True),
                          evaluate(pb, f, g['xs'], sort=False))
    g = load_golden('lotka_volterra')
    problem = pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), g['times'],
                                    g['values'])
    f = pb.GaussianLogLikelihood(problem)
    assert np.array_equal(evaluate(pb, f, g['xs'], sort=True),
                          evaluate(pb, f, g['xs'], sort=False))


def test_block_local_sort_and_its_hint(pb):
    """One-block-per-SM launches sort inside the kernel, block by block, when
    the previous launch found its vectors' costs (attempted steps) close
    together, and with sort_kernel when they were far apart
    (pb200_problem::sort_hint); whichever runs, scores and step
    counts are per vector, bit for bit those of the unsorted launch.  Sizes:
    migrating layout (65 536, 20 000: 14 and 5 warps per SM), dense classes
    (80 000, 100 000), a ragged last run of rows."""
    import torch
    spec = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    rng = np.random.RandomState(11)
    narrow = po.headline_fhn_points(100000)
    wide = rng.uniform([0, 0, 1], [1, 1, 5], (65536, 3))
    for n in (65536, 20000 - 7, 80000, 100000):
        dp_s = pb.CudaEvaluator(ll, sort=True).device_problem()
        dp_u = pb.CudaEvaluator(ll, sort=False).device_problem()
        d_x = torch.from_numpy(np.ascontiguousarray(narrow[:n])).cuda()
        want = dp_u.eval(d_x).cpu().numpy()
        for _ in range(3):          # global sort, then (hint = narrow) the block-local one
            got = dp_s.eval(d_x).cpu().numpy()
            torch.cuda.synchronize()
            assert np.array_equal(got, want, equal_nan=True), n
    # the models with the generic cost key (and the wide interpolation loop)
    for name, model in (('lotka_volterra', pb.toy.LotkaVolterraModel()),
                        ('sir', pb.toy.SIRModel())):
        g = load_golden(name)
        f = pb.GaussianLogLikelihood(pb.MultiOutputProblem(model, g['times'], g['values']))
        centre = g['xs'][0]
        for spread in (0.01, 0.2):              # costs close together, far apart
            xs = centre * (1 + spread * rng.uniform(-1, 1, (30000, len(centre))))
            if name == 'sir':
                xs[:, 2] = np.round(xs[:, 2])
            d_x = torch.from_numpy(np.ascontiguousarray(xs)).cuda()
            dp_s = pb.CudaEvaluator(f, sort=True).device_problem()
            want = pb.CudaEvaluator(f, sort=False).device_problem().eval(d_x).cpu().numpy()
            hints = []
            for _ in range(3):
                got = dp_s.eval(d_x).cpu().numpy()
                torch.cuda.synchronize()
                hints.append(dp_s.sort_hint())
                assert np.array_equal(got, want, equal_nan=True), (name, spread)
            assert hints[-1] in (1, 2), hints
    # a wide population in between flips the hint back and forth
    dp_s = pb.CudaEvaluator(ll, sort=True).device_problem()
    dp_u = pb.CudaEvaluator(ll, sort=False).device_problem()
    seen = []
    for xs in (narrow[:65536], narrow[:65536], wide, wide, narrow[:65536], narrow[:65536]):
        xs = xs.copy()
        xs[123, 2] = -1.0                          # NaN key
        d_x = torch.from_numpy(np.ascontiguousarray(xs)).cuda()
        got = dp_s.eval(d_x).cpu().numpy()
        torch.cuda.synchronize()
        seen.append(dp_s.sort_hint())
        assert np.array_equal(got, dp_u.eval(d_x).cpu().numpy(), equal_nan=True)
    # what each launch measured: 1 = costs close together, 2 = far apart
    assert seen == [1, 1, 2, 2, 1, 1], seen


def test_mean_squared_and_root_mean_squared_errors(pb):
    """(f) row 3: MeanSquaredError / RootMeanSquaredError /
    NormalisedRootMeanSquaredError epilogues (pints/_error_measures.py:74-161,
    :201-229) against the reference's own numbers."""
    g = load_golden('error_measures')
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), g['times'],
                                     g['values'])
    for f, key in ((pb.MeanSquaredError(problem), 'mse'),
                   (pb.MeanSquaredError(problem, [2.5]), 'mse_w'),
                   (pb.RootMeanSquaredError(problem), 'rmse'),
                   (pb.NormalisedRootMeanSquaredError(problem), 'nrmse')):
        got = evaluate(pb, f, g['xs'])
        assert np.max(rel_err(got, g[key])) < ANALYTIC_RTOL, key
    spec = po.headline_fhn_spec()
    fproblem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                     spec.values)
    got = evaluate(pb, pb.MeanSquaredError(fproblem, [1.0, 0.25]), g['fhn_xs'])
    check_ode_scores(got, g['fhn_mse'], g['fhn_mse_truth'])
    with pytest.raises(ValueError):
        pb.RootMeanSquaredError(fproblem)
    with pytest.raises(ValueError):
        pb.MeanSquaredError(fproblem, [1.0])
    # inside ProbabilityBasedError-free minimisation they are plain errors:
    # the sign stays +1 and the evaluator returns them as they are
    ev = pb.CudaEvaluator(pb.RootMeanSquaredError(problem), as_list='boxed')
    out = ev.evaluate(g['xs'][:4])
    assert isinstance(out, list) and np.allclose(out, g['rmse'][:4], rtol=1e-12)


@pytest.mark.parametrize('name,cls', [('goodwin', 'GoodwinOscillatorModel'),
                                      ('hes1', 'Hes1Model'),
                                      ('repressilator', 'RepressilatorModel')])
def test_more_ode_models(pb, name, cls):
    """(f) row 4: further toy ODE models behind the same DP5(4) template
    (pints/toy/_goodwin_oscillator_model.py, _hes1_michaelis_menten.py,
    _repressilator_model.py), same parity bars as FHN / LV / SIR."""
    g = load_golden(name)
    model = getattr(pb.toy, cls)()
    no = model.n_outputs()
    d = model.n_parameters()
    traj = model.simulate(g['xs'][:8, :d], g['times'])
    ref = g['traj'].reshape(traj.shape)
    truth = g['truth'].reshape(traj.shape)
    # within 1e-6 (1 + |y|) of the tight-tolerance truth; against the
    # reference the bar is widened by the reference's own error (its LSODA run
    # drifts by up to 2e-5 over these longer integrations)
    e = np.abs(traj - truth)
    assert np.all(e <= 1e-6 * (1 + np.abs(truth))), e.max()
    ref_err = np.abs(ref - truth)
    assert np.all(np.abs(traj - ref)
                  <= TRAJ_ATOL + TRAJ_RTOL * np.abs(ref) + ref_err)
    problem = (pb.MultiOutputProblem if no > 1 else pb.SingleOutputProblem)(
        model, g['times'], g['values'])
    got, steps = pb.CudaEvaluator(pb.GaussianLogLikelihood(problem),
                                  as_list=False).evaluate_array(g['xs'], True)
    # The noise of these fixtures is small against the signal, which amplifies
    # trajectory error in the log-likelihood (the reference's own LSODA scores
    # are off by 1e-5 .. 1e-3 relative; for Hes1 the log-likelihood itself
    # nearly cancels to zero, |LL| ~ 50).  At the matched tolerance the device
    # error must be of the order of the reference's own (within 5x: DP5(4) and
    # LSODA have different error constants per problem); with the tolerance
    # tightened the device meets the plain 1e-6 bar.
    dev_truth = np.max(rel_err(got, g['truth_scores']))
    ref_truth = np.max(rel_err(g['scores'], g['truth_scores']))
    print(name, 'matched tolerance: device', dev_truth, 'reference', ref_truth)
    assert dev_truth <= 5 * max(ODE_SCORE_RTOL, ref_truth), \
        (dev_truth, ref_truth)
    assert np.all(rel_err(got, g['scores'])
                  <= 6 * max(ODE_SCORE_RTOL, ref_truth))
    tight = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'],
                     rtol=1e-11, atol=1e-11)
    print(name, 'tight:', np.max(rel_err(tight, g['truth_scores'])))
    assert np.max(rel_err(tight, g['truth_scores'])) < ODE_SCORE_RTOL
    assert steps.min() > 20
    # sorted scheduling changes nothing here either
    again = evaluate(pb, pb.GaussianLogLikelihood(problem), g['xs'], sort=True)
    assert np.array_equal(again, got)
    # a single vector, and the reference object if it can be imported
    one = model.simulate(g['xs'][0, :d], g['times'])
    assert np.array_equal(np.asarray(one).reshape(traj[0].shape), traj[0])


def test_beeler_reuter(pb):
    """(f) row 4: the Beeler-Reuter action potential model
    (pints/toy/_beeler_reuter_model.py), a stiff system with a time-dependent
    right-hand side (the stimulus), on the 4-stage Rosenbrock kernel
    (br_rosenbrock_kernel).  The reference solves it with LSODA at rtol 1e-4 /
    atol 1e-6 and is itself 0.03 mV and 1.7e-3 (relative, log-likelihood) away
    from the converged solution; the device is checked against that converged
    solution."""
    g = load_golden('beeler_reuter')
    model = pb.toy.ActionPotentialModel()
    assert model.n_parameters() == 5 and model.n_outputs() == 2
    xs, times, truth = g['xs'], g['times'], g['truth']
    ref_err = np.abs(g['traj'] - truth)
    # default tolerance (1e-7): three orders closer than the reference
    traj = model.simulate(xs[:8, :5], times)
    assert traj.shape == truth.shape
    e = np.abs(traj - truth)
    print('V err', e[..., 0].max(), 'reference', ref_err[..., 0].max())
    assert e[..., 0].max() < 5e-5 and e[..., 0].max() < 0.002 * ref_err[..., 0].max()
    assert np.max(e[..., 1] / truth[..., 1]) < 1e-6
    assert np.all(e <= 1e-6 * (1 + np.abs(truth)) * [1, 1e-6])
    # and as far from the reference as the reference is from the truth
    assert np.abs(traj - g['traj'])[..., 0].max() <= 1.01 * ref_err[..., 0].max()
    problem = pb.MultiOutputProblem(model, times, g['values'])
    ll = pb.GaussianLogLikelihood(problem)
    got, steps = pb.CudaEvaluator(ll, as_list=False).evaluate_array(xs, True)
    dev = np.max(rel_err(got, g['truth_scores']))
    ref = np.max(rel_err(g['scores'], g['truth_scores']))
    print('score error: device', dev, 'reference', ref)
    assert dev < ODE_SCORE_RTOL and dev < 0.001 * ref
    assert np.all(rel_err(got, g['scores']) <= 1.1 * ref)
    # a stiff integrator: the step follows the action potential, not the
    # sodium gate's time constant (the explicit DP5(4) needs 7 700 steps)
    assert 1200 < steps.min() and steps.max() < 2200
    # fourth order: ten times tighter, ten times closer, 2.3 times the steps
    t8, s8 = pb.CudaEvaluator(ll, as_list=False, rtol=1e-8, atol=1e-8).evaluate_array(xs, True)
    assert np.max(rel_err(t8, g['truth_scores'])) < 0.12 * ODE_SCORE_RTOL
    assert s8.max() < 2.6 * steps.max()
    # the reference's own accuracy takes a quarter of the default's steps
    loose, sl = pb.CudaEvaluator(ll, as_list=False, rtol=1e-5, atol=1e-5).evaluate_array(xs, True)
    assert np.max(rel_err(loose, g['truth_scores'])) < ref and sl.max() < 400
    model.set_solver_tolerances(1e-10, 1e-10)
    t2 = model.simulate(xs[:8, :5], times)
    assert np.all(np.abs(t2 - truth) <= 1e-6 * (1 + np.abs(truth)) * [1, 1e-6])
    # one vector keeps the reference's shape; the stimulus needs times[0]
    one = model.simulate(xs[0, :5], times)
    assert one.shape == (len(times), 2) and np.array_equal(one, t2[0])
    late = model.simulate(xs[0, :5], times[10:])        # starts at t = 5: no stimulus
    assert np.abs(late[:, 0] - model.initial_conditions()[0]).max() < 2.0
    # sorted scheduling and an explicit block size change nothing
    again = evaluate(pb, ll, xs, sort=True)
    assert np.array_equal(again, got)
    assert np.array_equal(evaluate(pb, ll, xs, block=32), got)
    with pytest.raises(ValueError):
        pb.toy.ActionPotentialModel([-80.0, -1e-7])


def test_composed_prior(pb):
    """(f) row 2: ComposedLogPrior (uniform + Gaussian sub-priors) fused into
    the epilogue: outside any uniform support, or sigma <= 0, scores -inf
    without a solve; inside, the terms are added in the reference's order."""
    g = load_golden('composed_prior')
    h = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times,
                                    h.values)
    prior = pb.ComposedLogPrior(
        pb.UniformLogPrior([0.05, 0.3], [0.15, 0.7]),
        pb.GaussianLogPrior(3.0, 0.2), pb.GaussianLogPrior(0.5, 0.05),
        pb.UniformLogPrior([0.4], [0.6]))
    post = pb.LogPosterior(pb.GaussianLogLikelihood(problem), prior)
    got, steps = pb.CudaEvaluator(post, as_list=False).evaluate_array(
        g['xs'], True)
    for k in (3, 5, 7):
        assert got[k] == -np.inf and steps[k] == 0
    finite = np.isfinite(g['scores'])
    assert np.array_equal(np.isfinite(got), finite)
    assert np.max(rel_err(got[finite], g['scores'][finite])) < ODE_SCORE_RTOL
    neg = evaluate(pb, pb.ProbabilityBasedError(post), g['xs'])
    assert np.array_equal(neg, -got)
    # closed form: 1e-12, two Gaussian sub-priors
    lproblem = pb.SingleOutputProblem(pb.toy.LogisticModel(), g['ltimes'],
                                      g['lvalues'])
    lpost = pb.LogPosterior(
        pb.GaussianKnownSigmaLogLikelihood(lproblem, 1.0),
        pb.ComposedLogPrior(pb.GaussianLogPrior(0.1, 0.02),
                            pb.GaussianLogPrior(50, 5)))
    got = evaluate(pb, lpost, g['lxs'])
    assert np.max(rel_err(got, g['lscores'])) < ANALYTIC_RTOL
    # the mirror priors evaluate like the reference's on the host too
    for x, want in zip(g['xs'][:6], g['priors'][:6]):
        assert prior(x) == want

    class OddPrior(pb.LogPrior):           # a user's prior: refused, not guessed
        def n_parameters(self):
            return 1

        def __call__(self, x):
            return 0.0

    with pytest.raises(TypeError):
        pb.describe(pb.LogPosterior(
            pb.GaussianKnownSigmaLogLikelihood(lproblem, 1.0),
            pb.ComposedLogPrior(pb.GaussianLogPrior(0.1, 0.02), OddPrior())))


def test_host_buffer_entry_points(pb):
    """pb200_eval_host / pb200_copy_async / pb200_host_is_pinned through ctypes:
    the one-call host-buffer path gives what the device-pointer path gives."""
    import ctypes
    import torch
    from pints_b200 import _lib
    lib = _lib.load()
    spec = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    ev = pb.CudaEvaluator(ll, as_list=False)
    dp = ev.device_problem()
    n = 3000
    xs = pb.pinned_empty((n, 3))
    xs[:] = po.headline_fhn_points(n)
    V = ctypes.c_void_p
    assert lib.pb200_host_is_pinned(V(xs.ctypes.data)) == 1
    assert lib.pb200_host_is_pinned(V(np.zeros(8).ctypes.data)) == 0
    h_out = pb.pinned_empty((n,))
    h_steps = torch.zeros(n, dtype=torch.int32).pin_memory()
    d_in = torch.empty((n, 3), dtype=torch.float64, device='cuda')
    d_out = torch.empty(n, dtype=torch.float64, device='cuda')
    d_steps = torch.zeros(n, dtype=torch.int32, device='cuda')
    ws = torch.empty(int(lib.pb200_sort_workspace_bytes(n)), dtype=torch.uint8,
                     device='cuda')
    stream = V(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pb200_eval_host(
        dp._handle, V(xs.ctypes.data), n, 3, V(h_out.ctypes.data),
        V(h_steps.data_ptr()), V(d_in.data_ptr()), V(d_out.data_ptr()),
        V(d_steps.data_ptr()), V(ws.data_ptr()), ws.numel(),
        ctypes.byref(dp.opts), stream))
    torch.cuda.synchronize()
    want, steps = ev.evaluate_array(xs, True)
    assert np.array_equal(h_out, want)
    assert np.array_equal(h_steps.numpy(), steps)
    # workspace too small / misaligned is refused with a message
    rc = lib.pb200_eval_sorted(dp._handle, V(d_in.data_ptr()), n, 3,
                               V(d_out.data_ptr()), None, ctypes.byref(dp.opts),
                               V(ws.data_ptr()), 64, stream)
    assert rc == -1 and b'workspace' in lib.pb200_last_error()
    back = np.empty(n)
    _lib.check(lib.pb200_copy_async(V(h_out.ctypes.data), V(d_out.data_ptr()),
                                    n * 8, 0, stream))
    torch.cuda.synchronize()
    assert np.array_equal(h_out, want)
    assert lib.pb200_copy_async(None, None, -1, 0, stream) == -1


@pytest.mark.parametrize('n', [18945, 30000, 47000, 61569, 65536, 66304,
                               80000, 85000, 90000, 97000, 100000, 110000])
def test_migrating_layout_is_bit_identical(pb, n):
    """Populations of 5..14 warps per SM run one block per SM with warp
    migration between the schedulers; 17..24 warps per SM
    (the last six sizes) run the dense register classes, with or without
    migration; an explicit block size selects the plain layout.  Same scores
    and step counts, bit for bit, including vectors decided early (prior
    box), vectors that fail (NaN) and the ragged last block."""
    spec = po.headline_fhn_spec()
    rng = np.random.RandomState(n)
    xs = po.headline_fhn_points(n, seed=n % 97)
    wide = rng.uniform([0, 0, 1], [1, 1, 5], (n // 7, 3))
    xs[rng.permutation(n)[:len(wide)]] = wide
    xs[5] = [0.1, 0.5, 9.0]                       # outside the prior box
    xs[n - 3] = [0.1, 0.5, np.nan]                # NaN parameter -> NaN score
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    post = pb.LogPosterior(pb.GaussianKnownSigmaLogLikelihood(problem, 0.5),
                           pb.UniformLogPrior([0, 0, 0.5], [1, 1, 6]))
    a, sa = pb.CudaEvaluator(post, chunk_rows=1 << 30).evaluate_array(xs, True)
    b, sb = pb.CudaEvaluator(post, chunk_rows=1 << 30,
                             block=64).evaluate_array(xs, True)
    assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(sa, sb)
    assert a[5] == -np.inf and sa[5] == 0 and np.isnan(a[n - 3])
    assert np.isfinite(a).sum() >= n - 2


def test_maximum_sizes(pb):
    """The largest problem the ABI admits: 8 outputs (PB200_MAX_OUTPUTS), hence
    16 parameters with the unknown-sigma likelihood (PB200_MAX_PARAMS); one
    output more is refused, not truncated."""
    from pints_b200 import _lib
    no = _lib.MAX_OUTPUTS
    times = np.linspace(0, 10, 257)
    rng = np.random.RandomState(12)
    model = pb.toy.ConstantModel(no)
    truth = 1.0 + np.arange(no)
    clean = po.constant_simulate(truth, times, False)
    values = clean + rng.normal(0, 0.3, clean.shape)
    problem = pb.MultiOutputProblem(model, times, values)
    ll = pb.GaussianLogLikelihood(problem)
    assert ll.n_parameters() == _lib.MAX_PARAMS
    xs = np.hstack([truth * (1 + 0.1 * rng.randn(300, no)),
                    0.3 * (1 + 0.1 * rng.rand(300, no))])
    got = evaluate(pb, ll, xs)
    spec = po.Spec('constant', times, values, 'gaussian')
    want = po.scores(spec, xs)
    assert np.max(rel_err(got, want)) < ANALYTIC_RTOL
    sse = pb.SumOfSquaresError(problem, weights=np.arange(1, no + 1))
    want = po.scores(po.Spec('constant', times, values, 'sse',
                             np.arange(1, no + 1)), xs[:, :no])
    assert np.max(rel_err(evaluate(pb, sse, xs[:, :no]), want)) < ANALYTIC_RTOL
    # a box prior over all 16 parameters
    post = pb.LogPosterior(ll, pb.UniformLogPrior([0] * 16, [20] * 16))
    out = evaluate(pb, post, xs)
    assert np.all(np.isfinite(out))
    big = pb.toy.ConstantModel(no + 1)
    with pytest.raises(TypeError):
        pb.describe(pb.SumOfSquaresError(pb.MultiOutputProblem(
            big, times, np.zeros((len(times), no + 1)))))


@pytest.mark.parametrize('nt', [1, 2, 3, 7])
def test_migrating_layout_with_very_short_series(pb, nt):
    """The hand-over point of the migrating layout is a fraction of the
    observations: series of one to a few points must still work."""
    spec = po.headline_fhn_spec()
    times = spec.times[[0, 400, 650, 700, 900, 950, 999][:nt]] if nt > 1 \
        else spec.times[[500]]
    values = spec.values[[0, 400, 650, 700, 900, 950, 999][:nt]] if nt > 1 \
        else spec.values[[500]]
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), times, values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    xs = po.headline_fhn_points(20000, seed=nt)
    a = pb.CudaEvaluator(ll, chunk_rows=1 << 30, as_list=False).evaluate(xs)
    b = pb.CudaEvaluator(ll, chunk_rows=1 << 30, as_list=False,
                         block=64).evaluate(xs)
    assert np.array_equal(a, b) and np.all(np.isfinite(a))
    want = po.scores(po.Spec('fhn', times, values, 'gaussian_known_sigma', 0.5),
                     xs[:4])
    assert np.max(rel_err(a[:4], want)) < ODE_SCORE_RTOL


def test_wide_interpolation_changes_nothing(pb):
    """The four-at-a-time pending-observation loop (on by default for the
    smooth models, off for FitzHugh-Nagumo) only regroups the same
    interpolations: scores and step counts are bit-identical either way."""
    g = load_golden('lotka_volterra')
    problem = pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), g['times'],
                                    g['values'])
    f = pb.GaussianLogLikelihood(problem)
    a, sa = pb.CudaEvaluator(f, wide_interpolation=True).evaluate_array(
        g['xs'], True)
    b, sb = pb.CudaEvaluator(f, wide_interpolation=False).evaluate_array(
        g['xs'], True)
    c = evaluate(pb, f, g['xs'])                     # automatic (= on)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    assert np.array_equal(a, c)
    spec = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    xs = po.headline_fhn_points(20000)
    a = evaluate(pb, ll, xs, wide_interpolation=True, chunk_rows=1 << 30)
    b = evaluate(pb, ll, xs, wide_interpolation=False, chunk_rows=1 << 30)
    assert np.array_equal(a, b)


@pytest.mark.parametrize('seed', [101, 102])
def test_live_oracle_random_problems(pb, seed):
    rng = np.random.RandomState(seed)
    # FHN, ragged time grid that does not start at zero, unknown sigma
    times = np.sort(rng.uniform(0.3, 18, 137))
    values = po.fhn_simulate([0.15, 0.45, 2.7], times) \
        + rng.normal(0, 0.3, (137, 2))
    xs = np.hstack([np.array([0.15, 0.45, 2.7]) * (1 + 0.1 * rng.randn(24, 3)),
                    rng.uniform(0.2, 0.5, (24, 2))])
    spec = po.Spec('fhn', times, values, 'gaussian')
    f = pb.GaussianLogLikelihood(pb.MultiOutputProblem(
        pb.toy.FitzhughNagumoModel(), times, values))
    truth = po.scores(po.Spec('fhn', times, values, 'gaussian',
                              tol=po.TRUTH_TOL), xs)
    check_ode_scores(evaluate(pb, f, xs), po.scores(spec, xs), truth)
    # logistic with a posterior (uniform prior box) and sign flip
    times = np.linspace(0, 80, 300)
    values = po.logistic_simulate([0.12, 40], times) + rng.normal(0, 2, 300)
    xs = np.array([0.12, 40]) * (1 + 0.5 * rng.randn(64, 2))
    lower, upper = np.array([0.0, 10.0]), np.array([0.3, 60.0])
    spec = po.Spec('logistic', times, values, 'gaussian_known_sigma', 2.0,
                   sign=-1.0, prior_box=(lower, upper))
    post = pb.LogPosterior(
        pb.GaussianKnownSigmaLogLikelihood(pb.SingleOutputProblem(
            pb.toy.LogisticModel(), times, values), 2.0),
        pb.UniformLogPrior(lower, upper))
    got = evaluate(pb, pb.ProbabilityBasedError(post), xs)
    want = po.scores(spec, xs)
    assert np.array_equal(np.isinf(got), np.isinf(want))
    assert np.any(np.isinf(want)) and not np.all(np.isinf(want))
    assert np.nanmax(rel_err(got, want)) < ANALYTIC_RTOL


# ------------------------------------------------------ fused vs unfused path

def test_fused_equals_unfused(pb):
    g = load_golden('fhn')
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), g['times'],
                                    g['values'])
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    ev = pb.CudaEvaluator(ll, as_list=False)
    fused = ev.evaluate(g['xs'])
    dp = ev.device_problem()
    d_x = torch.from_numpy(g['xs']).to(dp.device)
    traj = dp.simulate(d_x)
    unfused = dp.score_trajectories(traj, d_x).cpu().numpy()
    assert np.max(rel_err(unfused, fused)) < 1e-12


# ------------------------------------------------------------------ edge cases

def test_edge_cases(pb):
    g = load_golden('fhn')
    model = pb.toy.FitzhughNagumoModel()
    problem = pb.MultiOutputProblem(model, g['times'], g['values'])
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    ev = pb.CudaEvaluator(ll)
    assert ev.evaluate([]) == []
    one = ev.evaluate([g['xs'][0]])
    assert isinstance(one, pb.ScoreList) and len(one) == 1
    assert isinstance(one[0], float) and one.tolist() == [one[0]]
    boxed = pb.CudaEvaluator(ll, as_list='boxed').evaluate([g['xs'][0]])
    assert type(boxed) is list and boxed == one
    assert abs(one[0] - g['scores'][0]) < 1e-6 * abs(g['scores'][0])
    assert ll(g['xs'][0]) == one[0]                 # scalar __call__
    # list of 1-d arrays, as MCMCController passes (pints/_mcmc/__init__.py:718)
    lst = ev.evaluate([np.array(x) for x in g['xs'][:5]])
    assert lst == ev.evaluate(g['xs'][:5])
    # read-only input, as optimisers pass
    ro = np.array(g['xs'][:5]); ro.setflags(write=False)
    assert ev.evaluate(ro) == lst
    with pytest.raises(ValueError):
        ev.evaluate(1)
    with pytest.raises(ValueError):
        ev.evaluate(np.zeros((3, 2)))
    # duplicate time points (allowed for multi-output problems)
    t = np.array([0.0, 0.5, 0.5, 1.0, 1.0, 1.0, 2.5])
    v = model.simulate(g['xs'][0], t)
    assert np.array_equal(v[1], v[2]) and np.array_equal(v[3], v[5])
    assert np.array_equal(v[0], [-1.0, 1.0])
    # a vector the explicit solver cannot finish within the step cap -> NaN,
    # never a hang; its neighbours are unaffected
    bad = np.array(g['xs'][:4]); bad[1] = [0.1, 0.5, 1e7]
    capped = pb.CudaEvaluator(ll, max_steps=2000, as_list=False).evaluate(bad)
    assert np.isnan(capped[1]) and np.all(np.isfinite(capped[[0, 2, 3]]))
    assert np.array_equal(capped[[0, 2, 3]],
                          np.array(ev.evaluate(g['xs'][:4]))[[0, 2, 3]])
    nan_in = np.array(g['xs'][:3]); nan_in[2, 0] = np.nan
    out = pb.CudaEvaluator(ll, max_steps=5000, as_list=False).evaluate(nan_in)
    assert np.isnan(out[2]) and np.all(np.isfinite(out[:2]))


def test_long_series_uses_global_memory_path(pb):
    # 6000 x 3 doubles = 144 KB > the shared-memory staging limit
    rng = np.random.RandomState(5)
    times = np.linspace(0, 20, 6000)
    values = po.fhn_simulate([0.1, 0.5, 3], times) + rng.normal(0, 0.5, (6000, 2))
    xs = po.headline_fhn_points(8)
    spec = po.Spec('fhn', times, values, 'gaussian_known_sigma', 0.5)
    f = pb.GaussianKnownSigmaLogLikelihood(pb.MultiOutputProblem(
        pb.toy.FitzhughNagumoModel(), times, values), 0.5)
    assert np.max(rel_err(evaluate(pb, f, xs), po.scores(spec, xs))) \
        < ODE_SCORE_RTOL
    # closed form with a long series (30000 points = 480 KB)
    times = np.linspace(0, 100, 30000)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 30000)
    lx = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(8, 2))
    spec = po.Spec('logistic', times, values, 'sse')
    f = pb.SumOfSquaresError(pb.SingleOutputProblem(
        pb.toy.LogisticModel(), times, values))
    assert np.max(rel_err(evaluate(pb, f, lx), po.scores(spec, lx))) \
        < ANALYTIC_RTOL


# ------------------------------------------- full size, size-free properties

def test_full_population_properties(pb):
    spec = po.headline_fhn_spec()
    n = 65536
    xs = po.headline_fhn_points(n)
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    ev = pb.CudaEvaluator(ll, as_list=False)
    scores, steps = ev.evaluate_array(xs, return_steps=True)
    assert scores.shape == (n,) and np.all(np.isfinite(scores))
    # order independence: every vector scores the same wherever it sits
    perm = np.random.RandomState(0).permutation(n)
    assert np.array_equal(ev.evaluate(xs[perm]), scores[perm])
    # the golden points are the first rows of this population
    g = load_golden('fhn')
    assert np.max(rel_err(scores[:96], g['scores'])) < ODE_SCORE_RTOL
    # the likelihood is maximal near the generating parameters
    best = xs[np.argmax(scores)]
    assert np.all(np.abs(best / np.array([0.1, 0.5, 3]) - 1) < 0.2)
    # sum of squares and known-sigma LL are affinely related:
    #   LL = sum_o offset_o + multip * SSE      (same sigma for both outputs)
    sse = evaluate(pb, pb.SumOfSquaresError(problem), xs[:4096])
    off = ll._offset.sum()
    assert np.max(rel_err(off + ll._multip[0] * sse, scores[:4096])) < 1e-12
    assert 300 < steps.mean() < 600
