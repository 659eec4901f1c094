"""
The UNMODIFIED reference on the GPU: pints.OptimisationController and
pints.MCMCController (from the reference install in baseline/_ref), with
`pints_b200.cuda_evaluators()` making them build CudaEvaluator objects
(pints/_optimisers/__init__.py:679-690, pints/_mcmc/__init__.py:572-584), run
on the real device and are compared with their own SequentialEvaluator runs:
same trajectory, same final numpy.random state (CudaEvaluator draws nothing,
unlike ParallelEvaluator, pints/_evaluation.py:310), and reference-class
objectives scored by CudaEvaluator agree with the reference's Python path.
Skipped where the reference install is absent.
"""
import numpy as np
import pytest

from conftest import load_golden, reference_pints

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def pints():
    p = reference_pints()
    if p is None:
        pytest.skip('reference install (baseline/_ref) not available')
    return p


@pytest.fixture(scope='module')
def pb():
    import pints_b200
    pints_b200.require_cuda()
    return pints_b200


def fhn_likelihood(pints):
    g = load_golden('fhn')
    problem = pints.MultiOutputProblem(pints.toy.FitzhughNagumoModel(),
                                       g['times'], g['values'])
    return g, pints.GaussianKnownSigmaLogLikelihood(problem, float(g['sigma']))


def test_cuda_evaluator_on_reference_objects(pints, pb):
    """CudaEvaluator built from reference-class objects against the
    reference's own evaluators on the same points."""
    g, ll = fhn_likelihood(pints)
    xs = g['xs'][:48]
    ev = pb.CudaEvaluator(ll)
    assert isinstance(ev, pints.Evaluator)
    got = np.array(ev.evaluate(xs))
    ref = np.array(pints.SequentialEvaluator(ll).evaluate(xs))
    assert np.array_equal(ref, g['scores'][:48])        # the fixtures are the reference
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-6
    # ParallelEvaluator gives the reference's numbers too (it is what bench.py
    # times as the CPU arm)
    par = np.array(pints.ParallelEvaluator(ll, n_workers=2).evaluate(xs[:8]))
    assert np.array_equal(par, ref[:8])
    # closed form, through ProbabilityBasedError and a LogPosterior: 1e-12
    h = load_golden('logistic')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), h['times'],
                                        h['values'])
    post = pints.LogPosterior(pints.GaussianLogLikelihood(problem),
                              pints.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
    rng = np.random.RandomState(3)
    xs = np.array([0.1, 50, 1.0]) * (1 + 0.2 * rng.randn(64, 3))
    xs[5, 2] = -1.0                                      # outside the prior
    for f in (post, pints.ProbabilityBasedError(post)):
        got = np.array(pb.CudaEvaluator(f).evaluate(xs))
        ref = np.array(pints.SequentialEvaluator(f).evaluate(xs))
        assert np.isinf(ref[5]) and got[5] == ref[5]
        ok = np.isfinite(ref)
        assert np.max(np.abs(got[ok] - ref[ok]) / np.abs(ref[ok])) < 1e-12


def run_optimisation(pints, pb, f, x0, bounds, pop, iters, use_cuda, seed=5):
    np.random.seed(seed)
    opt = pints.OptimisationController(f, x0, boundaries=bounds,
                                       method=pints.XNES)
    opt.optimiser().set_population_size(pop)
    opt.set_max_iterations(iters)
    opt.set_max_unchanged_iterations(None)
    opt.set_log_to_screen(False)
    if use_cuda:
        with pb.cuda_evaluators():
            x, fx = opt.run()
    else:
        x, fx = opt.run()
    return x, fx, opt.evaluations(), np.random.get_state()[1].copy()


def test_reference_xnes_population_512_closed_form(pints, pb):
    """pints.OptimisationController + pints.XNES, population 512, on the GPU:
    the closed-form scores agree to 1e-12, so the ranking -- and with it the
    whole trajectory -- is the SequentialEvaluator run's."""
    g = load_golden('logistic')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    ll = pints.GaussianKnownSigmaLogLikelihood(problem, 1.0)
    bounds = pints.RectangularBoundaries([0, 10], [1, 100])
    before = pints.SequentialEvaluator
    x1, f1, n1, s1 = run_optimisation(pints, pb, ll, [0.2, 30], bounds, 512, 15, False)
    x2, f2, n2, s2 = run_optimisation(pints, pb, ll, [0.2, 30], bounds, 512, 15, True)
    assert pints.SequentialEvaluator is before                  # names restored
    assert n1 == n2 and np.array_equal(s1, s2)                  # no extra RNG draws
    assert np.allclose(x1, x2, rtol=1e-9, atol=0) and abs(f1 - f2) <= 1e-10 * abs(f1)
    # parallel mode is rerouted as well
    np.random.seed(5)
    opt = pints.OptimisationController(ll, [0.2, 30], boundaries=bounds, method=pints.XNES)
    opt.optimiser().set_population_size(512)
    opt.set_max_iterations(15)
    opt.set_max_unchanged_iterations(None)
    opt.set_log_to_screen(False)
    opt.set_parallel(4)
    with pb.cuda_evaluators():
        x3, f3 = opt.run()
    assert np.array_equal(x3, x2) and f3 == f2


def test_reference_xnes_on_fitzhugh_nagumo(pints, pb):
    """The headline objective (FHN, ODE) under the reference's controller: the
    GPU scores are within 1e-6 of LSODA's, the optimiser consumes the same
    random numbers, and both runs arrive at the same optimum to the ODE
    tolerance."""
    g, ll = fhn_likelihood(pints)
    bounds = pints.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
    x1, f1, n1, s1 = run_optimisation(pints, pb, ll, [0.3, 0.7, 2.5], bounds, 48, 12, False)
    x2, f2, n2, s2 = run_optimisation(pints, pb, ll, [0.3, 0.7, 2.5], bounds, 48, 12, True)
    assert n1 == n2 and np.array_equal(s1, s2)
    assert abs(f1 - f2) <= 2e-6 * abs(f1)
    assert np.allclose(x1, x2, rtol=1e-4)


def test_reference_mcmc_controllers(pints, pb):
    """pints.MCMCController with HaarioBardenetACMC (one sampler per chain, a
    list of arrays per iteration) and DifferentialEvolutionMCMC (one ndarray):
    chains through the GPU evaluator equal the SequentialEvaluator chains."""
    g = load_golden('hbac_replay')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    post = pints.LogPosterior(pints.GaussianLogLikelihood(problem),
                              pints.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))

    def run(method, n_chains, use_cuda):
        np.random.seed(6)
        x0 = [g['x0'][k % len(g['x0'])] * (1 + 0.001 * k) for k in range(n_chains)]
        mc = pints.MCMCController(post, n_chains, x0, method=method)
        mc.set_max_iterations(300)
        if method is pints.HaarioBardenetACMC:
            mc.set_initial_phase_iterations(50)
        mc.set_log_to_screen(False)
        if use_cuda:
            with pb.cuda_evaluators():
                chains = mc.run()
        else:
            chains = mc.run()
        return chains, np.random.get_state()[1].copy()

    for method, n_chains in ((pints.HaarioBardenetACMC, 6),
                             (pints.DifferentialEvolutionMCMC, 8)):
        a, sa = run(method, n_chains, False)
        b, sb = run(method, n_chains, True)
        assert a.shape == b.shape == (n_chains, 300, 3)
        assert np.array_equal(sa, sb)
        # closed-form log-likelihoods agree to 1e-12: a decision could only
        # flip on a tie of that size; allow for the chains' own rounding
        assert np.allclose(a, b, rtol=1e-9, atol=0)
        assert (np.diff(a[:, :, 0], axis=1) != 0).mean() > 0.05      # they moved


def test_reference_multi_sequential_evaluator_route(pints, pb):
    """MCMCController with one log-pdf per chain builds a
    MultiSequentialEvaluator (pints/_mcmc/__init__.py:572-584); under
    cuda_evaluators() positions that share a function go out in one launch."""
    g = load_golden('logistic')
    p1 = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'], g['values'])
    p2 = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                   g['values'] * 1.01)
    f1 = pints.GaussianKnownSigmaLogLikelihood(p1, 1.0)
    f2 = pints.GaussianKnownSigmaLogLikelihood(p2, 1.0)
    fs = [f1, f2, f1, f1, f2]
    xs = [np.array([0.1, 50.0]) * (1 + 0.01 * k) for k in range(5)]
    ref = pints.MultiSequentialEvaluator(fs).evaluate(xs)
    with pb.cuda_evaluators():
        ev = pints.MultiSequentialEvaluator(fs)
        got = ev.evaluate(xs)
        assert len(ev._groups) == 2
        with pytest.raises(ValueError):
            ev.evaluate(xs[:3])
    assert isinstance(got, list) and len(got) == 5
    assert np.allclose(got, ref, rtol=1e-12, atol=0)


def test_reference_controllers_with_a_transformation(pints, pb):
    """`transformation=...` makes the reference's controllers wrap the
    objective (pints/_optimisers/__init__.py:392-406, pints/_mcmc/__init__.py:
    349-361); CudaEvaluator unwraps it, maps the rows on the host and adds the
    Jacobian term of a TransformedLogPDF: same runs as with SequentialEvaluator."""
    g = load_golden('logistic')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    ll = pints.GaussianLogLikelihood(problem)
    post = pints.LogPosterior(ll, pints.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
    t = pints.ComposedTransformation(
        pints.LogTransformation(1),
        pints.RectangularBoundariesTransformation([10, 0.1], [100, 10]))
    # the wrapped objectives themselves, on a batch of search-space points
    rng = np.random.RandomState(2)
    qs = np.array([t.to_search(x) for x in
                   np.array([0.1, 50, 1.0]) * (1 + 0.1 * rng.randn(40, 3))])
    for f in (t.convert_log_pdf(post), t.convert_log_pdf(ll),
              pints.ProbabilityBasedError(t.convert_log_pdf(post))):
        got = np.array(pb.CudaEvaluator(f).evaluate(qs))
        ref = np.array(pints.SequentialEvaluator(f).evaluate(qs))
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12

    def optimise(use_cuda):
        np.random.seed(3)
        opt = pints.OptimisationController(post, [0.2, 30, 2.0], transformation=t,
                                           method=pints.XNES)
        opt.optimiser().set_population_size(64)
        opt.set_max_iterations(20)
        opt.set_max_unchanged_iterations(None)
        opt.set_log_to_screen(False)
        if use_cuda:
            with pb.cuda_evaluators():
                return opt.run() + (np.random.get_state()[1].copy(),)
        return opt.run() + (np.random.get_state()[1].copy(),)

    x1, f1, s1 = optimise(False)
    x2, f2, s2 = optimise(True)
    assert np.array_equal(s1, s2)
    assert np.allclose(x1, x2, rtol=1e-9, atol=0) and abs(f1 - f2) <= 1e-10 * abs(f1)

    def sample(use_cuda):
        np.random.seed(4)
        x0 = [np.array([0.1, 50, 1.0]) * (1 + 0.01 * k) for k in range(4)]
        mc = pints.MCMCController(post, 4, x0, transformation=t,
                                  method=pints.HaarioBardenetACMC)
        mc.set_max_iterations(200)
        mc.set_initial_phase_iterations(40)
        mc.set_log_to_screen(False)
        if use_cuda:
            with pb.cuda_evaluators():
                return mc.run()
        return mc.run()

    a, b = sample(False), sample(True)
    assert a.shape == b.shape == (4, 200, 3)
    assert np.allclose(a, b, rtol=1e-9, atol=0)
