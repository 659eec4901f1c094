"""
Host-side callers of the hot path, on the CPU:

* the vectorised XNES consumes numpy.random like the reference's XNES and
  follows the same trajectory;
* `cuda_evaluators()` makes the unmodified reference controllers construct a
  CudaEvaluator and hands it the batch formats they really use;
* with a test-only oracle backend standing in for the GPU launch, a reference
  OptimisationController / MCMCController run through CudaEvaluator is
  identical to its SequentialEvaluator run (so CudaEvaluator neither reorders
  nor consumes random numbers).
"""
import numpy as np
import pytest

import pints_b200 as pb
from conftest import load_golden
from oracle import pints_oracle as po


def oracle_backend(monkeypatch):
    """Test-only: replace the device launch of CudaEvaluator by the CPU oracle
    (the product itself has no such path)."""
    from pints_b200 import _evaluation, _lib
    names = _lib.MODEL_NAMES
    seen = []

    def fake(self, positions, return_steps=False):
        s = self._spec
        xs = _evaluation.positions_to_array(positions, s.n_parameters())
        seen.append(type(positions).__name__)
        values = s.values if s.multi_output else s.values[:, 0]
        consts = {}
        if s.model == _lib.MODEL_LOGISTIC:
            consts = {'p0': s.model_consts[0]}
        elif s.model in (_lib.MODEL_FHN, _lib.MODEL_LOTKA_VOLTERRA):
            consts = {'y0': s.model_consts[:2]}
        objective = _lib.OBJECTIVE_NAMES[s.objective]
        if objective == 'sse':
            obj = s.obj_consts
        elif objective == 'gaussian_known_sigma':
            # recover sigma from multip = -1 / (2 sigma^2)
            no = s.n_outputs
            obj = np.sqrt(-1 / (2 * np.array(s.obj_consts[no:])))
        else:
            obj = None
        box = None
        if s.prior_lower is not None:
            box = (np.array(s.prior_lower), np.array(s.prior_upper))
        spec = po.Spec(names[s.model], s.times, values, objective, obj,
                       consts=consts, sign=s.sign, prior_box=box)
        return po.scores(spec, xs)

    monkeypatch.setattr(_evaluation.CudaEvaluator, 'evaluate_array', fake)
    return seen


def test_vectorised_xnes_matches_reference(pints_ref):
    pints = pints_ref
    f = lambda x: float(np.sum((np.asarray(x) - [1.0, -2.0, 0.5])**2))  # noqa
    for bounded in (False, True):
        b_ref = pints.RectangularBoundaries([-5, -5, -5], [5, 5, 5]) \
            if bounded else None
        b_new = pb.RectangularBoundaries([-5, -5, -5], [5, 5, 5]) \
            if bounded else None
        np.random.seed(3)
        ref = pints.XNES([0.5, 0.5, 0.5], 2.0, b_ref)
        ref.set_population_size(40)
        ref_trace = []
        for _ in range(15):
            xs = ref.ask()
            ref.tell([f(x) for x in xs])
            ref_trace.append((np.array(xs), np.array(ref._mu), np.array(ref._A),
                              ref.f_best()))
        state_ref = np.random.get_state()[1].copy()
        np.random.seed(3)
        new = pb.XNES([0.5, 0.5, 0.5], 2.0, b_new)
        new.set_population_size(40)
        for k in range(15):
            xs = new.ask()
            assert xs.shape == ref_trace[k][0].shape
            assert np.allclose(xs, ref_trace[k][0], rtol=1e-11, atol=1e-12)
            new.tell(np.array([f(x) for x in xs]))
            assert np.allclose(new._mu, ref_trace[k][1], rtol=1e-10, atol=1e-12)
            assert np.allclose(new._A, ref_trace[k][2], rtol=1e-9, atol=1e-12)
            assert np.isclose(new.f_best(), ref_trace[k][3], rtol=1e-10)
        # identical consumption of the global stream
        assert np.array_equal(np.random.get_state()[1], state_ref)
    assert new.name() == ref.name()
    assert new._suggested_population_size() == ref._suggested_population_size()


@pytest.mark.parametrize('name', ['SNES', 'PSO', 'BareCMAES'])
def test_vectorised_optimisers_match_reference(pints_ref, name):
    """(f) row 1: vectorised host ask()/tell() of the other population-based
    optimisers.  Same trajectories as the reference classes (to rounding: BLAS
    products replace per-sample loops) and IDENTICAL consumption of the global
    numpy.random stream, generation by generation."""
    pints = pints_ref
    target = np.array([1.0, -2.0, 0.5])
    f = lambda x: float(np.sum((np.asarray(x) - target)**2))  # noqa
    for bounded in (False, True):
        b_ref = pints.RectangularBoundaries([-5, -5, -5], [5, 5, 5]) \
            if bounded else None
        b_new = pb.RectangularBoundaries([-5, -5, -5], [5, 5, 5]) \
            if bounded else None
        np.random.seed(5)
        ref = getattr(pints, name)([0.5, 0.5, 0.5], 2.0, b_ref)
        ref.set_population_size(24)
        trace = []
        for _ in range(25):
            xs = ref.ask()
            ref.tell([f(x) for x in xs])
            trace.append((np.array(xs), np.array(ref.x_best()), ref.f_best(),
                          np.random.get_state()[1].copy()))
        np.random.seed(5)
        new = getattr(pb, name)([0.5, 0.5, 0.5], 2.0, b_new)
        new.set_population_size(24)
        for k in range(25):
            xs = new.ask()
            assert xs.shape == trace[k][0].shape, (name, bounded, k)
            assert np.allclose(xs, trace[k][0], rtol=1e-8, atol=1e-10), k
            new.tell(np.array([f(x) for x in xs]))
            assert np.allclose(new.x_best(), trace[k][1], rtol=1e-8, atol=1e-10)
            assert np.isclose(new.f_best(), trace[k][2], rtol=1e-8, atol=1e-12)
            assert np.array_equal(np.random.get_state()[1], trace[k][3]), k
        assert new.name() == ref.name()
        assert new.population_size() == ref.population_size()
        assert not new.needs_sensitivities()
    # the optimisers do converge
    assert new.f_best() < (0.5 if name == 'PSO' else 1e-2)
    if name == 'BareCMAES':
        assert np.allclose(new.cov(), ref.cov(), rtol=1e-6, atol=1e-12)
        assert new.stop() is False
    if name == 'PSO':
        with pytest.raises(ValueError):
            new2 = pb.PSO([1, 2])
            new2.set_local_global_balance(1.5)


def test_xnes_argument_checks():
    with pytest.raises(ValueError):
        pb.XNES([])
    with pytest.raises(ValueError):
        pb.XNES([1, 2], -1.0)
    with pytest.raises(ValueError):
        pb.XNES([1, 2], [1, 2, 3])
    with pytest.raises(ValueError):
        pb.XNES([3, 0], None, pb.RectangularBoundaries([0, 0], [1, 1]))
    opt = pb.XNES([1, 2, 3])
    assert opt.population_size() == 4 + int(3 * np.log(3))
    with pytest.raises(Exception):
        opt.tell([1, 2, 3])
    opt.set_population_size(10)
    opt.ask()
    with pytest.raises(Exception):
        opt.set_population_size(20)


def test_reference_optimisation_controller_through_cuda_evaluator(
        pints_ref, monkeypatch):
    pints = pints_ref
    seen = oracle_backend(monkeypatch)
    g = load_golden('logistic')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    ll = pints.GaussianKnownSigmaLogLikelihood(problem, 1.0)
    bounds = pints.RectangularBoundaries([0, 10], [1, 100])

    def run(use_cuda):
        np.random.seed(5)
        opt = pints.OptimisationController(ll, [0.2, 30], boundaries=bounds,
                                           method=pints.XNES)
        opt.set_max_iterations(12)
        opt.set_log_to_screen(False)
        if use_cuda:
            with pb.cuda_evaluators():
                x, f = opt.run()
        else:
            x, f = opt.run()
        return x, f, np.random.get_state()[1].copy()

    before = pints.SequentialEvaluator
    x1, f1, s1 = run(False)
    x2, f2, s2 = run(True)
    assert pints.SequentialEvaluator is before          # restored
    assert np.array_equal(x1, x2) and f1 == f2
    assert np.array_equal(s1, s2)                       # no extra RNG draws
    assert seen and set(seen) == {'ndarray'}


def test_reference_mcmc_controller_through_cuda_evaluator(
        pints_ref, monkeypatch):
    pints = pints_ref
    seen = oracle_backend(monkeypatch)
    g = load_golden('hbac_replay')
    problem = pints.SingleOutputProblem(pints.toy.LogisticModel(), g['times'],
                                        g['values'])
    post = pints.LogPosterior(pints.GaussianLogLikelihood(problem),
                              pints.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))

    def run(method, use_cuda):
        np.random.seed(6)
        mc = pints.MCMCController(post, 4, g['x0'][:4], method=method)
        mc.set_max_iterations(30)
        if method is pints.HaarioBardenetACMC:
            mc.set_initial_phase_iterations(10)
        mc.set_log_to_screen(False)
        if use_cuda:
            with pb.cuda_evaluators():
                return mc.run()
        return mc.run()

    for method in (pints.HaarioBardenetACMC, pints.DifferentialEvolutionMCMC):
        del seen[:]
        a = run(method, False)
        b = run(method, True)
        assert np.array_equal(a, b)
        # single-chain samplers hand over a list of per-chain arrays,
        # multi-chain samplers an ndarray
        want = 'list' if method is pints.HaarioBardenetACMC else 'ndarray'
        assert set(seen) == {want}


def test_cuda_evaluators_needs_the_reference(monkeypatch):
    from pints_b200 import _controllers
    monkeypatch.setattr(_controllers, 'pints_or_none', lambda: None)
    with pytest.raises(RuntimeError):
        with pb.cuda_evaluators():
            pass
