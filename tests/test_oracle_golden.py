"""
Pins the oracle (oracle/pints_oracle.py) to the reference:

* the known answers held by the reference's own unit tests (file:line cited at
  each check), and
* the golden fixtures produced by running the real reference
  (tests/golden/make_golden.py).

All CPU, all exact unless the reference's own test gives a number of places.
"""
import numpy as np
import pytest

from oracle import pints_oracle as po
from conftest import load_golden


def spec_fhn(g):
    return po.Spec('fhn', g['times'], g['values'], 'gaussian_known_sigma',
                   float(g['sigma']))


# ------------------------------------------------------------------ models

def test_logistic_reference_known_answers():
    # pints/tests/test_toy_logistic_model.py:93-105 (exact equality)
    times = np.linspace(0, 20, 21)
    v = po.logistic_simulate([0.1, 12], times, p0=0.5)
    assert v[0] == 0.5
    assert v[1] == 0.5501745273651227
    closed = 12 / (1 + (12 / 0.5 - 1) * np.exp(-0.1 * times))
    assert np.all(v == closed)
    # :20-66 edge cases
    assert np.all(po.logistic_simulate([0.1, 50], times, p0=0) == 0)
    assert np.all(po.logistic_simulate([0.1, -1], times) == 0)
    with pytest.raises(ValueError):
        po.logistic_simulate([0.1, 50], [-1, 0, 1])


def test_logistic_golden():
    g = load_golden('logistic')
    assert np.all(po.logistic_simulate([0.1, 12], g['unit_times'], 0.5)
                  == g['unit_values'])
    for x, ref in zip(g['xs'], g['traj']):
        assert np.array_equal(po.logistic_simulate(x, g['times']), ref,
                              equal_nan=True)
    s = po.Spec('logistic', g['times'], g['values'], 'sse')
    out = po.scores(s, g['xs'])
    assert np.array_equal(out, g['scores'], equal_nan=True)
    assert np.isnan(out[15])            # k == 0 -> 0/0 at t = 0
    s = po.Spec('logistic', g['times'], g['values'], 'gaussian_known_sigma',
                1.0)
    assert np.array_equal(po.scores(s, g['xs']), g['ll_scores'],
                          equal_nan=True)
    s = po.Spec('logistic', g['times'], g['values'], 'gaussian')
    out = po.scores(s, g['xs_gauss'])
    assert np.array_equal(out, g['gauss_scores'], equal_nan=True)
    assert out[2] == -np.inf


def test_hh_ik_reference_known_answers():
    # pints/tests/test_toy_hh_ik_model.py:44-74
    events, voltages, v_hold = po.hh_ik_protocol()
    assert len(events) == 24 and events[0] == 90 and events[-1] == 1200
    times = np.arange(1200 * 4) / 4
    v = po.hh_ik_simulate((0.01, 10, 10, 0.125, 80), times)
    assert v.shape == (4800,)
    assert abs(v[0] - 3.790799999999997) < 1e-7
    assert abs(v[1] - 3.83029) < 1e-2
    assert abs(v[390] - 15.9405) < 1e-2
    assert round(abs(v[4790] - 3862.8), 0) == 0        # places=0
    with pytest.raises(ValueError):
        po.hh_ik_simulate((0.01, 10, 10, 0.125, 80), [-1, 0])


def test_hh_ik_golden():
    g = load_golden('hh_ik')
    events, voltages, _ = po.hh_ik_protocol()
    assert np.all(events == g['events']) and np.all(voltages == g['voltages'])
    for x, ref in zip(g['xs'], g['traj']):
        assert np.all(po.hh_ik_simulate(x[:5], g['times']) == ref)
    for x, ref in zip(g['xs'], g['event_values']):
        assert np.all(po.hh_ik_simulate(x[:5], g['event_times']) == ref)
    s = po.Spec('hh_ik', g['times'], g['values'], 'gaussian')
    out = po.scores(s, g['xs'])
    assert np.array_equal(out, g['scores'])
    assert out[4] == -np.inf
    s = po.Spec('hh_ik', g['times_1k'], g['values_1k'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['scores_1k'])


def test_fhn_reference_known_answer():
    # pints/tests/test_toy_fitzhugh_nagumo_model.py:51-59
    times = np.linspace(0, 20, 201)
    v = po.fhn_simulate([0.2, 0.4, 2.5], times, y0=[-2, 1.5])
    assert v.shape == (201, 2)
    assert v[0, 0] == -2 and v[0, 1] == 1.5
    assert round(abs(v[200, 0] - 1.675726), 6) == 0
    assert round(abs(v[200, 1] - -0.226142), 6) == 0
    with pytest.raises(ValueError):
        po.fhn_simulate([0.2, 0.4, 2.5], [-1, 0, 1])


def test_fhn_golden():
    g = load_golden('fhn_unit')
    v = po.fhn_simulate(g['x'], g['times'], y0=g['y0'])
    assert np.array_equal(v, g['values'])
    v = po.fhn_simulate(g['x'], g['times_nz'], y0=g['y0'])
    assert np.array_equal(v, g['values_nz'])
    g = load_golden('fhn')
    for x, ref in zip(g['xs'], g['traj']):
        assert np.array_equal(po.fhn_simulate(x, g['times']), ref)
    s = spec_fhn(g)
    assert np.array_equal(po.scores(s, g['xs'][:24]), g['scores'][:24])
    s.sign = -1.0
    assert np.array_equal(po.scores(s, g['xs'][:8]), g['neg_scores'])
    s = po.Spec('fhn', g['times'], g['values'], 'sse', g['sse_weights'])
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['sse_scores'])
    s = po.Spec('fhn', g['times'], g['values'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs_gauss']), g['gauss_scores'])
    # the "truth" (tight LSODA) is reproducible through the oracle too
    t = po.fhn_simulate(g['xs'][0], g['times'], **po.TRUTH_TOL)
    assert np.array_equal(t, g['truth'][0])


def test_headline_problem_is_the_golden_one():
    g = load_golden('fhn')
    s = po.headline_fhn_spec()
    assert np.array_equal(s.times, g['times'])
    assert np.array_equal(s.values, g['values'])
    assert np.array_equal(po.headline_fhn_points(96), g['xs'])


def test_lotka_volterra():
    # pints/tests/test_toy_lotka_volterra_model.py:42-55
    times = np.linspace(0, 5, 101)
    v = po.lv_simulate([1, 2, 2, 0.5], times, y0=[3, 5])
    assert v[0, 0] == 3 and v[0, 1] == 5
    assert round(abs(v[1, 0] - 1.929494), 6) == 0
    assert round(abs(v[1, 1] - 4.806542), 6) == 0
    assert round(abs(v[100, 0] - 1.277762), 6) == 0
    assert round(abs(v[100, 1] - 0.000529), 6) == 0
    g = load_golden('lotka_volterra')
    assert np.array_equal(v, g['unit_values'])
    for x, ref in zip(g['xs'], g['traj']):
        assert np.array_equal(po.lv_simulate(x[:4], g['times']), ref)
    s = po.Spec('lotka_volterra', g['times'], g['values'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['scores'][:16])
    s = po.Spec('lotka_volterra', g['small_times'], g['small_values'],
                'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['small_scores'])


def test_sir():
    # pints/tests/test_toy_sir_model.py:48-60
    times = np.linspace(0, 10, 101)
    v = po.sir_simulate([0.05, 0.4, 100], times, y0=[100, 10, 1])
    assert v[0, 0] == 10 and v[0, 1] == 1
    assert round(abs(v[1, 0] - 15.61537), 5) == 0
    assert round(abs(v[1, 1] - 1.50528), 5) == 0
    assert round(abs(v[100, 0] - 2.45739), 5) == 0
    assert round(abs(v[100, 1] - 108.542466), 5) == 0
    g = load_golden('sir')
    assert np.array_equal(v, g['unit_values'])
    vf = po.sir_simulate([0.05, 0.4, 77.7], times, y0=[100, 10, 1])
    assert np.array_equal(vf, g['unit_values_frac'])
    for x, ref in zip(g['xs'], g['traj']):
        assert np.array_equal(po.sir_simulate(x[:3], g['times']), ref)
    # the integer y0 truncates S0 (pints/toy/_sir_model.py:80,108-109)
    a = po.sir_simulate([0.026, 0.285, 38.0], g['times'])
    b = po.sir_simulate([0.026, 0.285, 38.9], g['times'])
    assert np.array_equal(a, b)
    s = po.Spec('sir', g['times'], g['values'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['scores'][:16])
    s = po.Spec('sir', g['small_times'], g['small_values'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:16]), g['small_scores'])


# -------------------------------------------------------------- objectives

def test_objective_reference_known_answers():
    t = [1, 2, 3]
    # pints/tests/test_error_measures.py:738-765: single output, data 1,1,1
    s = po.Spec('constant', t, [1, 1, 1], 'sse')
    assert s.score([1]) == 0 and s.score([2]) == 3 and s.score([3]) == 12
    # closed form of the pre-computed parts
    # (pints/_log_likelihoods.py:1027-1032)
    off, mul = po.gaussian_known_sigma_consts(3, 1, 1.5)
    assert off[0] == -0.5 * 3 * np.log(2 * np.pi) - 3 * np.log(1.5)
    assert mul[0] == -1 / (2.0 * 1.5**2)
    with pytest.raises(ValueError):
        po.gaussian_known_sigma_consts(3, 1, 0.0)
    with pytest.raises(ValueError):
        po.gaussian_known_sigma_consts(3, 2, [1.0, 2.0, 3.0])


DATA_MULTI = np.array([[10.7, 3.5, 3.8], [1.1, 3.2, -1.4], [9.3, 0.0, 4.5],
                       [1.2, -3, -10]])


def test_known_sigma_reference_goldens():
    # pints/tests/test_log_likelihoods.py:1909-1961
    times = [1, 2, 3, 4]
    single = np.arange(1, 5) / 5.0
    s = po.Spec('constant', times, single, 'gaussian_known_sigma', 1.5)
    assert abs(s.score([-1]) - -7.3420590096957925) < 1e-7
    s = po.Spec('constant', times, DATA_MULTI, 'gaussian_known_sigma', 1)
    assert abs(s.score([0, 0, 0]) - -196.9122623984561) < 1e-7


def test_gaussian_reference_goldens():
    # pints/tests/test_log_likelihoods.py:2078-2171, :2271-2277
    times = [1, 2, 3, 4]
    single = np.array([1, 2, 3, 4]) / 5.0
    multi = np.array(DATA_MULTI)
    np.random.seed(42)
    sigma = 0.1
    single += np.random.normal(0, sigma, single.shape)
    multi += np.random.normal(0, sigma, multi.shape)
    s = po.Spec('constant', times, single, 'gaussian')
    k = po.Spec('constant', times, single, 'gaussian_known_sigma', sigma)
    assert abs(s.score([2, sigma]) - -421.8952711914118) < 1e-7
    assert abs(s.score([2, sigma]) - k.score([2])) < 1e-7
    s = po.Spec('constant', times, multi, 'gaussian')
    k = po.Spec('constant', times, multi, 'gaussian_known_sigma', sigma)
    assert abs(s.score([0, 0, 0, 0.1, 0.1, 0.1]) - k.score([0, 0, 0])) < 1e-7
    assert abs(s.score([0, 0, 0, 3.5, 1, 12]) - -50.75425117450455) < 1e-7
    assert s.score([0, 0, 0, 1, -1, 1]) == -np.inf
    assert s.score([0, 0, 0, 1, 0, 1]) == -np.inf


def test_constant_golden():
    g = load_golden('constant')
    t = g['single_times']
    s = po.Spec('constant', t, g['single_values'], 'sse')
    assert np.array_equal(po.scores(s, g['single_sse_x']), g['single_sse'])
    assert list(g['single_sse']) == [0.0, 3.0, 12.0, 0.75]
    s = po.Spec('constant', t, g['multi_values'], 'sse')
    assert np.array_equal(po.scores(s, g['multi_x']), g['multi_sse'])
    s = po.Spec('constant', t, g['multi_values'], 'sse', [1, 2, 3])
    assert np.array_equal(po.scores(s, g['multi_x']), g['multi_sse_weighted'])
    s = po.Spec('constant', t, g['multi_values'], 'gaussian_known_sigma',
                g['multi_sigma'])
    assert np.array_equal(po.scores(s, g['multi_x']), g['multi_known'])
    s = po.Spec('constant', t, g['multi_values'], 'gaussian')
    out = po.scores(s, g['multi_gauss_x'])
    assert np.array_equal(out, g['multi_gauss'])
    # unknown-sigma LL at the known sigma equals the known-sigma LL
    # (pints/tests/test_log_likelihoods.py:2139-2171)
    assert abs(out[0] - g['multi_known'][0]) < 1e-12 * abs(out[0])
    assert out[2] == -np.inf


def test_prior_box_and_sign():
    g = load_golden('logistic')
    s = po.Spec('logistic', g['times'], g['values'], 'gaussian_known_sigma',
                1.0, prior_box=([0, 0], [1, 100]))
    base = po.Spec('logistic', g['times'], g['values'],
                   'gaussian_known_sigma', 1.0)
    x = [0.1, 50]
    assert s.score(x) == -np.log(100.0) + base.score(x)
    assert s.score([0.1, 100]) == -np.inf          # upper bound is exclusive
    assert s.score([0.0, 50]) != -np.inf           # lower bound is inclusive
    assert s.score([1.0, 50]) == -np.inf


def test_prior_box_matches_reference(pints_ref):
    pints = pints_ref
    g = load_golden('logistic')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    post = pints.LogPosterior(
        pints.GaussianKnownSigmaLogLikelihood(problem, 1.0),
        pints.UniformLogPrior([0, 0], [1, 100]))
    s = po.Spec('logistic', g['times'], g['values'], 'gaussian_known_sigma',
                1.0, prior_box=(np.array([0., 0.]), np.array([1., 100.])))
    for x in ([0.1, 50], [0.0, 50], [1.0, 50], [0.1, 100], [0.5, 99.9],
              [-0.1, 3]):
        assert s.score(x) == post(x)


def test_sequential_evaluate_contract():
    # pints/tests/test_evaluators.py:40-91
    f = lambda x: float(x) ** 2          # noqa: E731
    assert po.sequential_evaluate(f, [1, 2, 3]) == [1, 4, 9]
    assert po.sequential_evaluate(f, []) == []
    with pytest.raises(ValueError):
        po.sequential_evaluate(f, 1)
    g = lambda x, y, z: x + y * z        # noqa: E731
    assert po.sequential_evaluate(g, [1, 2], args=(10, 20)) == [201, 202]


# ---------------------------------------------------------------- samplers

def _ll_logistic(g):
    s = po.Spec('logistic', g['times'], g['values'], 'gaussian')
    return s.score


def test_hbac_replay_matches_reference():
    g = load_golden('hbac_replay')
    ll = _ll_logistic(g)
    n_iter, n_chains, d = g['proposed'].shape
    np.random.seed(12)
    chains = [po.HaarioBardenetChain(x) for x in g['x0']]
    for c, ch in enumerate(chains):
        assert np.array_equal(ch.sigma, g['sigma0'][c])
    for it in range(n_iter):
        if it == int(g['warm']):
            for ch in chains:
                ch.adaptive = True
        xs = [ch.ask() for ch in chains]
        assert np.array_equal(np.array(xs), g['proposed'][it])
        fxs = [ll(x) for x in xs]
        if it == 7:
            fxs[2] = -np.inf
        if it == 9:
            fxs[4] = np.nan
        assert np.array_equal(np.array(fxs), g['fx'][it], equal_nan=True)
        for c, ch in enumerate(chains):
            cur, cur_f, acc = ch.tell(fxs[c])
            assert bool(acc) == bool(g['accepted'][it, c])
            assert np.array_equal(cur, g['current'][it, c])
            assert cur_f == g['current_f'][it, c]
            assert np.array_equal(ch.mu, g['mu'][it, c])
            assert np.array_equal(ch.sigma, g['sigma'][it, c])
            assert ch.log_lambda == g['log_lambda'][it, c]
            if ch.last_log_u is None:
                assert np.isnan(g['log_u'][it, c])
            else:
                assert ch.last_log_u == g['log_u'][it, c]
    assert g['accepted'][1:].mean() > 0.05


def test_hbac_proposal_is_svd_of_scaled_sigma():
    """multivariate_normal(mean, C) on the legacy stream
    is ``mean + z @ (sqrt(s)[:, None] * Vt)`` with ``U, s, Vt = svd(C)``."""
    g = load_golden('hbac_replay')
    it = int(g['warm']) + 20
    for c in range(g['x0'].shape[0]):
        cov = g['sigma'][it - 1, c] * np.exp(g['log_lambda'][it - 1, c])
        _, s, vt = np.linalg.svd(cov)
        x = g['current'][it - 1, c] + g['z'][it, c] @ (np.sqrt(s)[:, None] * vt)
        assert np.allclose(x, g['proposed'][it, c], rtol=1e-13, atol=0)


def test_de_replay_matches_reference():
    g = load_golden('de_replay')
    ll = _ll_logistic(g)
    n_iter, n, d = g['proposed'].shape
    np.random.seed(13)
    s = po.DifferentialEvolutionChains(n, g['x0'])
    for it in range(n_iter):
        xs = s.ask()
        assert np.array_equal(xs, g['proposed'][it])
        if it > 0:
            assert np.array_equal(s.last['eps'], g['eps'][it])
            assert np.array_equal(s.last['r1'], g['r1'][it])
            assert np.array_equal(s.last['r2'], g['r2'][it])
            assert s.last['gamma'] == g['gamma'][it]
        fxs = np.array([ll(x) for x in xs])
        if it == 5:
            fxs[1] = np.nan
        if it == 6:
            fxs[3] = -np.inf
        assert np.array_equal(fxs, g['fx'][it], equal_nan=True)
        cur, cur_f, acc = s.tell(fxs)
        assert np.array_equal(acc, g['accepted'][it].astype(bool))
        assert np.array_equal(cur, g['current'][it])
        assert np.array_equal(cur_f, g['current_f'][it])
        if it > 0:
            assert np.array_equal(s.last['log_u'], g['log_u'][it])
    assert np.array_equal(s.b_star, g['b_star'])
    # gamma is 1 on every tenth proposal round, 2.38/sqrt(2d) otherwise
    assert g['gamma'][1] == 1.0 and g['gamma'][11] == 1.0
    assert g['gamma'][2] == 2.38 / np.sqrt(2 * d)


# ------------------------------------------- live cross-check when mounted

def test_oracle_equals_live_reference(pints_ref):
    pints = pints_ref
    rng = np.random.RandomState(99)
    times = np.linspace(0, 15, 60)
    m = pints.toy.FitzhughNagumoModel()
    for x in np.array([0.1, 0.5, 3]) * (1 + 0.2 * rng.randn(5, 3)):
        assert np.array_equal(m.simulate(x, times), po.fhn_simulate(x, times))
    m = pints.toy.LotkaVolterraModel()
    for x in np.array([0.5, 0.15, 1.0, 0.3]) * (1 + 0.2 * rng.randn(5, 4)):
        assert np.array_equal(m.simulate(x, times), po.lv_simulate(x, times))
    m = pints.toy.SIRModel()
    ts = np.linspace(1, 21, 60)
    for x in np.array([0.026, 0.285, 38]) * (1 + 0.1 * rng.randn(5, 3)):
        assert np.array_equal(m.simulate(x, ts), po.sir_simulate(x, ts))
    m = pints.toy.HodgkinHuxleyIKModel(0.4)
    th = np.sort(rng.uniform(0, 1300, 200))
    for x in np.array([0.01, 10, 10, 0.125, 80]) * (1 + 0.2 * rng.randn(5, 5)):
        assert np.array_equal(m.simulate(x, th),
                              po.hh_ik_simulate(x, th, n0=0.4))


def test_error_measures_golden():
    """MeanSquaredError, RootMeanSquaredError, NormalisedRootMeanSquaredError
    (pints/_error_measures.py:74-161, :201-229): the oracle reproduces the
    reference exactly; known answers from pints/tests/test_error_measures.py."""
    g = load_golden('error_measures')
    for objective, obj, key in (('mse', None, 'mse'), ('mse', [2.5], 'mse_w'),
                                ('rmse', None, 'rmse'),
                                ('nrmse', None, 'nrmse')):
        s = po.Spec('logistic', g['times'], g['values'], objective, obj)
        assert np.array_equal(po.scores(s, g['xs']), g[key]), key
    s = po.Spec('fhn', np.linspace(0, 20, 1000), po.headline_fhn_spec().values,
                'mse', [1.0, 0.25])
    assert np.array_equal(po.scores(s, g['fhn_xs']), g['fhn_mse'])
    # reference unit tests on ConstantModel (multi-output: output o = p_o (o+1))
    # pints/tests/test_error_measures.py MeanSquaredError: values all ones...
    times = np.array([1.0, 2.0, 3.0])
    values = np.array([[1.0, 4.0], [1.0, 4.0], [1.0, 4.0]])
    s = po.Spec('constant', times, values, 'mse')
    assert po.scores(s, [[1.0, 2.0]])[0] == 0.0
    assert po.scores(s, [[2.0, 2.0]])[0] == 0.5          # (1^2 * 3) / 6
    s1 = po.Spec('constant', times, values[:, 0], 'rmse')
    assert po.scores(s1, [[3.0]])[0] == 2.0
    s1 = po.Spec('constant', times, values[:, 0], 'nrmse')
    assert po.scores(s1, [[3.0]])[0] == 2.0
    with pytest.raises(ValueError):
        po.Spec('constant', times, values, 'rmse')


@pytest.mark.parametrize('name', ['goodwin', 'hes1', 'repressilator',
                                  'beeler_reuter'])
def test_more_ode_models_golden(name):
    """(f) row 4: Goodwin, Hes1 and Repressilator restated in the oracle are
    bit-identical to the reference (same odeint call, same right-hand side)."""
    g = load_golden(name)
    s = po.Spec(name, g['times'], g['values'], 'gaussian')
    assert np.array_equal(po.scores(s, g['xs'][:12]), g['scores'][:12])
    no = s.n_outputs
    for x, ref in zip(g['xs'][:3], g['traj']):
        sim = s.trajectory(x[:-no])
        assert np.array_equal(np.asarray(sim).reshape(ref.shape), ref)


def test_more_ode_models_reference_unit_tests():
    # pints/tests/test_toy_goodwin_oscillator_model.py: suggested run, shapes
    t = np.linspace(0, 100, 200)
    v = po.goodwin_simulate([2, 4, 0.12, 0.08, 0.1], t)
    assert v.shape == (200, 3) and np.all(v[0] == po.GOODWIN_Y0)
    # pints/tests/test_toy_hes1_michaelis_menten_model.py:27-32 values at the
    # suggested times start from m0 = 2
    v = po.hes1_simulate([2.4, 0.025, 0.11, 6.9], np.arange(0, 270, 30))
    assert v.shape == (9,) and v[0] == 2
    # pints/tests/test_toy_repressilator_model.py: starts at y0[:3]
    v = po.repressilator_simulate([1, 1000, 5, 2], np.linspace(0, 40, 400))
    assert v.shape == (400, 3) and np.all(v[0] == 0)
    # pints/tests/test_toy_beeler_reuter_model.py:27-55: initial values, shapes
    # and the all-states variant
    p0 = np.log([4.0, 0.003, 0.09, 0.35, 0.8])
    t = np.arange(0, 100, 0.5)
    v = po.beeler_reuter_simulate(p0, t)
    assert v.shape == (200, 2) and v[0, 0] == -84.622 and v[0, 1] == 2e-7
    assert v[:, 0].max() > 20 and np.all(v[:, 1] > 0)     # an action potential
    a = po.beeler_reuter_simulate(p0, t, all_states=True)
    assert a.shape == (200, 8) and np.array_equal(a[:, :2], v)
    v = po.beeler_reuter_simulate(p0, t, y0=[-80, 1e-5])
    assert v[0, 0] == -80 and v[0, 1] == 1e-5


def test_composed_prior_golden():
    """(f) row 2: LogPosterior over a ComposedLogPrior of uniform and Gaussian
    sub-priors (pints/_log_priors.py:170-219): exact."""
    g = load_golden('composed_prior')
    parts = [('uniform', [0.05, 0.3], [0.15, 0.7]), ('gaussian', 3.0, 0.2),
             ('gaussian', 0.5, 0.05), ('uniform', [0.4], [0.6])]
    for x, want in zip(g['xs'], g['priors']):
        assert po.composed_log_prior(parts, x) == want
    h = po.headline_fhn_spec()
    s = po.Spec('fhn', h.times, h.values, 'gaussian', prior_parts=parts)
    assert np.array_equal(po.scores(s, g['xs'][:10]), g['scores'][:10])
    assert g['scores'][3] == -np.inf and g['scores'][5] == -np.inf
    assert g['scores'][7] == -np.inf
    s = po.Spec('logistic', g['ltimes'], g['lvalues'], 'gaussian_known_sigma',
                1.0, prior_parts=[('gaussian', 0.1, 0.02), ('gaussian', 50, 5)])
    assert np.array_equal(po.scores(s, g['lxs']), g['lscores'])
