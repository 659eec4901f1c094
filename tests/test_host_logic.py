"""
CPU-only checks: the C-ABI library builds, loads and exports every symbol the
header declares; the host-side mirror validates like the reference; the
introspection turns objective graphs (mirror objects and, when mounted, real
reference objects) into the same problem description; and nothing silently
falls back to the CPU.
"""
import os
import re

import numpy as np
import pytest

import pints_b200 as pb
from pints_b200 import _lib
from conftest import ROOT, load_golden


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'pints_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'\b(pb200_[a-z0-9_]+)\s*\(', text)
    return sorted(set(names))


def test_library_builds_loads_and_exports_the_header():
    from pints_b200 import build
    path = build.build()
    assert os.path.exists(path)
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), name + ' is declared but not exported'
    assert sorted(_lib.SIGNATURES) == declared
    # every entry point is tied to the reference code it replaces
    with open(os.path.join(ROOT, 'INTEGRATION.md')) as f:
        doc = f.read()
    for name in declared:
        assert name in doc, name + ' is not documented in INTEGRATION.md'
    assert lib.pb200_version() == 100
    assert lib.pb200_series_len(1000, 2) == 3 * (1000 + _lib.SENTINEL_ROWS)
    # (3 + 8) rows * 3 doubles = 33, padded to whole 16-byte units
    assert lib.pb200_series_len(3, 2) == 34


def test_library_is_sm100a_only():
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH],
                         stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r'sm_\d+a?', out))
    assert archs == {'sm_100a'}, archs


def test_problem_create_argument_errors():
    import ctypes
    lib = _lib.load()
    h = ctypes.c_void_p()
    oc, n_oc = _lib.doubles([1.0, 1.0])
    mc, n_mc = _lib.doubles([-1.0, 1.0])
    # unknown model
    rc = lib.pb200_problem_create(99, 0, 10, 2, None, mc, n_mc, oc, n_oc, 1.0,
                                  None, None, 0, 0.0, ctypes.byref(h))
    assert rc == -2 and b'unknown model' in lib.pb200_last_error()
    # wrong number of outputs for FHN
    rc = lib.pb200_problem_create(_lib.MODEL_FHN, 0, 0, 3, None, mc, n_mc, oc,
                                  n_oc, 1.0, None, None, 0, 0.0,
                                  ctypes.byref(h))
    assert rc == -1
    # bad sign
    rc = lib.pb200_problem_create(_lib.MODEL_FHN, 0, 0, 2, None, mc, n_mc, oc,
                                  n_oc, 0.5, None, None, 0, 0.0,
                                  ctypes.byref(h))
    assert rc == -1 and b'sign' in lib.pb200_last_error()
    # a valid (empty-series) problem can be created and destroyed without a GPU
    rc = lib.pb200_problem_create(_lib.MODEL_FHN, 0, 0, 2, None, mc, n_mc, oc,
                                  n_oc, 1.0, None, None, 0, 0.0,
                                  ctypes.byref(h))
    assert rc == 0 and h.value
    assert lib.pb200_problem_n_parameters(h) == 3
    # n = 0 is a no-op everywhere, n < 0 an error
    assert lib.pb200_eval(h, None, 0, 3, None, None, None, None) == 0
    assert lib.pb200_eval(h, None, -1, 3, None, None, None, None) == -1
    lib.pb200_problem_destroy(h)
    with pytest.raises(_lib.LibraryError):
        _lib.check(-1)


def fhn_objects(mod, toy):
    g = load_golden('fhn')
    model = toy.FitzhughNagumoModel()
    problem = mod.MultiOutputProblem(model, g['times'], g['values'])
    return g, model, problem


def test_describe_mirror_objects():
    g, model, problem = fhn_objects(pb, pb.toy)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    s = pb.describe(ll)
    assert (s.model, s.objective, s.sign) == (_lib.MODEL_FHN, 1, 1.0)
    assert s.model_consts == [-1.0, 1.0]
    assert s.n_parameters() == 3 and s.n_outputs == 2 and s.n_times == 1000
    off = -0.5 * 1000 * np.log(2 * np.pi) - 1000 * np.log(0.5)
    assert s.obj_consts == [off, off, -2.0, -2.0]
    packed = s.packed_series()
    assert packed.shape == (3 * (1000 + _lib.SENTINEL_ROWS),)
    assert np.array_equal(packed[:3000].reshape(1000, 3)[:, 0], g['times'])
    assert np.array_equal(packed[:3000].reshape(1000, 3)[:, 1:], g['values'])
    assert np.all(packed[3000::3] == np.inf) and not packed[3001::3].any()
    s = pb.describe(pb.ProbabilityBasedError(ll))
    assert s.sign == -1.0
    s = pb.describe(pb.GaussianLogLikelihood(problem))
    assert s.objective == 2 and s.n_parameters() == 5
    assert s.obj_consts == [0.5 * 1000 * np.log(2 * np.pi)]
    s = pb.describe(pb.SumOfSquaresError(problem, [1, 2.5]))
    assert s.objective == 0 and s.obj_consts == [1.0, 2.5]
    post = pb.LogPosterior(ll, pb.UniformLogPrior([0, 0, 0], [1, 2, 10]))
    s = pb.describe(pb.ProbabilityBasedError(post))
    assert s.prior_lower == [0, 0, 0] and s.prior_upper == [1, 2, 10]
    assert s.log_prior == -np.log(20.0) and s.sign == -1.0
    # SIR: the default integer y0 sets the truncation flag
    assert pb.describe_model(pb.toy.SIRModel())[1] == [38.0, 1.0, 0.0, 1.0]
    assert pb.describe_model(pb.toy.SIRModel([38, 1, 0]))[1][3] == 0.0
    mid, consts = pb.describe_model(pb.toy.HodgkinHuxleyIKModel())
    assert mid == _lib.MODEL_HH_IK and consts[:5] == [0.3, 36, -88, -75, 24]
    assert len(consts) == 5 + 48 and consts[5] == 90 and consts[28] == 1200


def test_unsupported_functions_raise_type_error():
    g, model, problem = fhn_objects(pb, pb.toy)

    class MyModel(pb.toy.FitzhughNagumoModel):     # user subclass: not trusted
        pass

    with pytest.raises(TypeError):
        pb.describe(pb.SumOfSquaresError(
            pb.MultiOutputProblem(MyModel(), g['times'], g['values'])))
    with pytest.raises(TypeError):
        pb.describe(lambda x: 0.0)
    with pytest.raises(TypeError):
        pb.CudaEvaluator(lambda x: 0.0)
    with pytest.raises(ValueError):
        pb.CudaEvaluator(3)                        # not callable
    with pytest.raises(ValueError):
        pb.CudaEvaluator(pb.SumOfSquaresError(problem), args=[1])
    with pytest.raises(ValueError):
        pb.CudaEvaluator(pb.SumOfSquaresError(problem), args=1)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('needs a machine without a GPU')
    g, model, problem = fhn_objects(pb, pb.toy)
    ev = pb.CudaEvaluator(pb.SumOfSquaresError(problem))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ev.evaluate(g['xs'][:4])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model.simulate(g['xs'][0], g['times'])
    with pytest.raises(ValueError):
        ev.evaluate(7)             # contract errors come first, as in pints


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'pints_b200')
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, name)).read()
                assert 'oracle' not in text.replace('no oracle', ''), name


def test_mirror_validation_matches_reference_messages():
    t = [0, 1, 2]
    with pytest.raises(ValueError, match='Population size'):
        pb.toy.LogisticModel(-1)
    with pytest.raises(ValueError, match='size 2'):
        pb.toy.FitzhughNagumoModel([1, 2, 3])
    with pytest.raises(ValueError, match='> 0 and < 1'):
        pb.toy.HodgkinHuxleyIKModel(1.0)
    with pytest.raises(ValueError, match='negative'):
        pb.toy.LotkaVolterraModel([-1, 0])
    with pytest.raises(ValueError, match='size 3'):
        pb.toy.SIRModel([1, 1])
    with pytest.raises(ValueError, match='negative'):
        pb.toy.SIRModel([1, 1, -1])
    with pytest.raises(ValueError):
        pb.toy.ConstantModel(0)
    m = pb.toy.LogisticModel()
    with pytest.raises(ValueError, match='increasing'):
        pb.SingleOutputProblem(m, [0, 1, 1], [0, 0, 0])
    with pytest.raises(ValueError, match='negative'):
        pb.SingleOutputProblem(m, [-1, 1, 2], [0, 0, 0])
    with pytest.raises(ValueError, match='same length'):
        pb.SingleOutputProblem(m, t, [0, 0])
    with pytest.raises(ValueError, match='single-output'):
        pb.SingleOutputProblem(pb.toy.FitzhughNagumoModel(), t, [0, 0, 0])
    f = pb.toy.FitzhughNagumoModel()
    pb.MultiOutputProblem(f, [0, 1, 1], np.zeros((3, 2)))   # ties allowed
    with pytest.raises(ValueError, match='non-decreasing'):
        pb.MultiOutputProblem(f, [0, 2, 1], np.zeros((3, 2)))
    with pytest.raises(ValueError, match='shape'):
        pb.MultiOutputProblem(f, t, np.zeros((3, 3)))
    p = pb.MultiOutputProblem(f, t, np.zeros((3, 2)))
    with pytest.raises(ValueError, match='greater than zero'):
        pb.GaussianKnownSigmaLogLikelihood(p, 0)
    with pytest.raises(ValueError, match='length n_outputs'):
        pb.GaussianKnownSigmaLogLikelihood(p, [1, 2, 3])
    with pytest.raises(ValueError, match='weights'):
        pb.SumOfSquaresError(p, [1, 2, 3])
    with pytest.raises(ValueError, match='exceed'):
        pb.RectangularBoundaries([0, 0], [1, 0])
    b = pb.RectangularBoundaries([0, 0], [1, 2])
    assert b.check([0, 0]) and not b.check([1, 0]) and not b.check([0.5, 2])
    assert list(b.check(np.array([[0, 0], [1, 0]]))) == [True, False]
    prior = pb.UniformLogPrior(b)
    assert prior([0.5, 1]) == -np.log(2.0) and prior([2, 1]) == -np.inf
    with pytest.raises(ValueError, match='same dimension'):
        pb.LogPosterior(pb.GaussianLogLikelihood(p), prior)
    assert pb.GaussianLogLikelihood(p).n_parameters() == 5


# ------------------------------------------------- against the real reference

def test_describe_reference_objects_equals_mirror(pints_ref):
    pints = pints_ref
    g = load_golden('fhn')

    def build(mod, toy):
        out = []
        m = toy.FitzhughNagumoModel()
        p = mod.MultiOutputProblem(m, g['times'], g['values'])
        ll = mod.GaussianKnownSigmaLogLikelihood(p, 0.5)
        out += [ll, mod.ProbabilityBasedError(ll), mod.GaussianLogLikelihood(p),
                mod.SumOfSquaresError(p, [1, 2.5]),
                mod.LogPosterior(ll, mod.UniformLogPrior([0, 0, 0],
                                                         [1, 2, 10]))]
        gl = load_golden('logistic')
        pl = mod.SingleOutputProblem(toy.LogisticModel(3.0), gl['times'],
                                     gl['values'])
        out += [mod.SumOfSquaresError(pl), mod.GaussianLogLikelihood(pl)]
        gh = load_golden('hh_ik')
        ph = mod.SingleOutputProblem(toy.HodgkinHuxleyIKModel(0.4),
                                     gh['times'], gh['values'])
        out += [mod.GaussianLogLikelihood(ph)]
        gs = load_golden('sir')
        out += [mod.GaussianLogLikelihood(mod.MultiOutputProblem(
            toy.SIRModel(), gs['times'], gs['values'])),
            mod.GaussianLogLikelihood(mod.MultiOutputProblem(
                toy.SIRModel([30, 2, 1]), gs['times'], gs['values'])),
            mod.GaussianLogLikelihood(mod.MultiOutputProblem(
                toy.LotkaVolterraModel(), gs['times'], gs['values']))]
        gc = load_golden('constant')
        out += [mod.GaussianKnownSigmaLogLikelihood(mod.MultiOutputProblem(
            toy.ConstantModel(3), gc['single_times'], gc['multi_values']),
            gc['multi_sigma'])]
        return out

    for ref_f, mir_f in zip(build(pints, pints.toy), build(pb, pb.toy)):
        a, b = pb.describe(ref_f), pb.describe(mir_f)
        for key in ('model', 'model_consts', 'objective', 'obj_consts', 'sign',
                    'prior_lower', 'prior_upper', 'log_prior', 'n_times',
                    'n_outputs', 'multi_output'):
            assert getattr(a, key) == getattr(b, key), (key, type(ref_f))
        assert np.array_equal(a.packed_series(), b.packed_series())


def test_cuda_evaluator_is_a_pints_evaluator(pints_ref):
    pints = pints_ref
    assert issubclass(pb.CudaEvaluator, pints.Evaluator)
    g = load_golden('fhn')
    m = pints.toy.FitzhughNagumoModel()
    p = pints.MultiOutputProblem(m, g['times'], g['values'])
    ll = pints.GaussianKnownSigmaLogLikelihood(p, 0.5)
    ev = pb.CudaEvaluator(pints.ProbabilityBasedError(ll))
    assert isinstance(ev, pints.Evaluator) and ev.spec().sign == -1.0
    # mirror objectives pass the reference's isinstance checks
    mm = pb.toy.FitzhughNagumoModel()
    mp = pb.MultiOutputProblem(mm, g['times'], g['values'])
    mll = pb.GaussianKnownSigmaLogLikelihood(mp, 0.5)
    assert isinstance(mll, pints.LogPDF) and isinstance(mm, pints.ForwardModel)
    pints.ProbabilityBasedError(mll)
    pints.LogPosterior(mll, pints.UniformLogPrior([0, 0, 0], [1, 1, 10]))
    # unsupported reference objects are refused, not evaluated on the CPU
    with pytest.raises(TypeError):
        pb.CudaEvaluator(pints.StudentTLogLikelihood(p))
    # ... while the mean-squared family of the reference is on the path
    spec = pb.describe(pints.MeanSquaredError(p, [1.0, 0.5]))
    assert spec.objective == _lib.OBJ_MSE
    assert spec.obj_consts == [1.0, 0.5, 1.0 / 2000]
    with pytest.raises(TypeError):
        pb.CudaEvaluator(pints.toy.RosenbrockError())
    with pytest.raises(TypeError):
        pb.CudaEvaluator(pints.GaussianLogLikelihood(pints.SingleOutputProblem(
            pints.toy.SimpleHarmonicOscillatorModel(), g['times'],
            np.zeros(1000))))
    # ... while the further ODE toy models of the reference are on the path
    spec = pb.describe(pints.GaussianLogLikelihood(pints.MultiOutputProblem(
        pints.toy.GoodwinOscillatorModel(), g['times'], np.zeros((1000, 3)))))
    assert spec.model == _lib.MODEL_GOODWIN and spec.n_parameters() == 8
    assert spec.model_consts == [0.0054, 0.053, 1.93]
    spec = pb.describe(pints.SumOfSquaresError(pints.SingleOutputProblem(
        pints.toy.Hes1Model(), g['times'], np.zeros(1000))))
    assert spec.model == _lib.MODEL_HES1
    assert spec.model_consts == [2.0, 5.0, 3.0, 0.03]
    spec = pb.describe(pints.SumOfSquaresError(pints.MultiOutputProblem(
        pints.toy.RepressilatorModel(), g['times'], np.zeros((1000, 3)))))
    assert spec.model == _lib.MODEL_REPRESSILATOR
    assert spec.model_consts == [0, 0, 0, 2, 1, 3]
    br = pints.toy.ActionPotentialModel()
    spec = pb.describe(pints.GaussianLogLikelihood(pints.MultiOutputProblem(
        br, br.suggested_times(), np.zeros((800, 2)))))
    assert spec.model == _lib.MODEL_BEELER_REUTER and spec.n_parameters() == 7
    assert spec.model_consts == [-84.622, 2e-7, 0.01, 0.99, 0.98, 0.003, 0.99,
                                 0.0004, 1.0, 50.0, 25.0, 1000.0, 2.0]
    mirror = pb.toy.ActionPotentialModel()
    assert pb.describe(pb.SumOfSquaresError(pb.MultiOutputProblem(
        mirror, br.suggested_times(), np.zeros((800, 2))))).model_consts \
        == spec.model_consts
    assert np.array_equal(mirror.suggested_parameters(),
                          br.suggested_parameters())
    assert np.array_equal(mirror.suggested_times(), br.suggested_times())


def test_device_xnes_host_side_contract():
    """Constructor checks and the population-size interface of DeviceXNES are
    those of pints.PopulationBasedOptimiser (pints/_optimisers/__init__.py:
    24-312); nothing touches the GPU before the first ask()."""
    b = pb.RectangularBoundaries([0, 0], [1, 2])
    opt = pb.DeviceXNES([0.5, 1.0], boundaries=b)
    assert opt.population_size() == 4 + int(3 * np.log(2))
    assert np.allclose(opt._sigma0, [1 / 6, 2 / 6])
    assert not opt.running() and not opt.needs_sensitivities()
    assert np.array_equal(opt.x_guessed(), [0.5, 1.0])
    assert opt.f_best() == np.inf
    opt.set_population_size(100)
    assert opt.population_size() == 100
    assert opt.suggested_population_size(8) == 8
    with pytest.raises(ValueError):
        opt.set_population_size(0)
    with pytest.raises(ValueError):
        pb.DeviceXNES([0.5, 3.0], boundaries=b)          # outside
    with pytest.raises(ValueError):
        pb.DeviceXNES([0.5], boundaries=b)                # dimension
    with pytest.raises(ValueError):
        pb.DeviceXNES([0.5, 1.0], sigma0=-1.0)
    with pytest.raises(ValueError):
        pb.DeviceXNES([0.5, 1.0], sigma0=[1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        pb.DeviceXNES(np.ones(17))                        # more than 16 parameters
    assert np.allclose(pb.DeviceXNES([0.0, -3.0])._sigma0, [1.0, 1.0])
    with pytest.raises(Exception):
        opt.tell(None)                                    # before ask()
    # sharding needs a device-resident optimiser
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(),
                                     np.linspace(0, 10, 20), np.zeros(20))
    with pytest.raises(ValueError):
        pb.CudaOptimisationController(pb.SumOfSquaresError(problem),
                                      [0.1, 10.0], method=pb.XNES, sharded=True)


def test_score_list_serves_every_consumer_of_the_reference():
    """What the reference does with the list an evaluator returns (SURVEY.md
    section 8b): len, indexing, iteration, np.argsort, np.array, assignment
    into a padded array, comparison -- on the lazy ScoreList."""
    from pints_b200 import ScoreList
    a = np.array([3.5, -1.0, np.inf, 2.25, np.nan])
    fs = ScoreList(a)
    plain = a.tolist()
    assert len(fs) == 5 and fs[1] == -1.0 and type(fs[1]) is float
    assert fs[-1] != fs[-1]                                     # nan
    assert isinstance(fs[1:3], ScoreList) and fs[1:3] == plain[1:3]
    assert list(fs)[:4] == plain[:4] and fs.tolist()[:4] == plain[:4]
    it = iter(fs)                                # pints/_mcmc/__init__.py:733-735
    assert next(it) == 3.5 and next(it) == -1.0
    assert np.array_equal(np.argsort(fs), np.argsort(a))        # _xnes.py:160
    order = np.argsort(fs)
    assert fs[order[0]] == -1.0                                 # _xnes.py:177
    full = np.ones(8) * np.inf                                  # _xnes.py:154-157
    full[[0, 2, 4, 6, 7]] = fs
    assert full[2] == -1.0 and full[1] == np.inf
    arr = np.array(fs)                           # _differential_evolution.py:290
    assert arr.dtype == np.float64 and arr[3] == 2.25
    arr[0] = 0.0
    assert fs[0] == 3.5                          # np.array copies
    assert np.asarray(fs, dtype=np.float32).dtype == np.float32
    assert fs == plain and fs == a and not (fs != plain)
    assert fs == ScoreList(a.copy()) and not (fs == plain[:4])
    assert ScoreList(np.empty(0)) == [] and not (ScoreList(np.empty(0)) == [1.0])
    assert fs[:4] + [1.0] == plain[:4] + [1.0]
    assert [1.0] + fs[:4] == [1.0] + plain[:4]
    assert min(fs[:4]) == -1.0 and max(fs[:2]) == 3.5
    assert sorted(fs[:4]) == sorted(plain[:4])
    assert 2.25 in fs and fs.index(2.25) == 3 and fs.count(3.5) == 1
    assert list(reversed(fs[:4])) == plain[:4][::-1]
    assert repr(fs[:2]) == repr(plain[:2])
    assert float(fs[0]) == 3.5 and -fs[1] == 1.0
    with pytest.raises(IndexError):
        fs[7]
    with pytest.raises(TypeError):
        hash(fs)
    # the real consumers, when the reference is importable
    pints = conftest_reference()
    if pints is not None:
        np.random.seed(1)
        opt = pints.XNES([1.0, 1.0], 0.5, pints.RectangularBoundaries([0, 0], [2, 2]))
        for _ in range(3):
            xs = opt.ask()
            opt.tell(ScoreList(np.sum((np.asarray(xs) - 1.2) ** 2, axis=1)))
        assert np.isfinite(opt.f_best())
        de = pints.DifferentialEvolutionMCMC(4, [np.array([0.1, 0.2])] * 4)
        xs = de.ask()
        de.tell(ScoreList(-np.sum(np.asarray(xs) ** 2, axis=1)))
        xs = de.ask()
        assert de.tell(ScoreList(-np.sum(np.asarray(xs) ** 2, axis=1))) is not None


def conftest_reference():
    from conftest import reference_pints
    return reference_pints()


def test_host_copy_threads_copy_exactly():
    lib = _lib.load()
    import ctypes
    rng = np.random.RandomState(3)
    for n in (0, 1, 1000, 65536 * 3 + 5):
        a = rng.rand(n)
        b = np.zeros(n)
        for threads in (0, 1, 3, 8):
            b[:] = 0
            assert lib.pb200_host_copy(ctypes.c_void_p(b.ctypes.data),
                                       ctypes.c_void_p(a.ctypes.data),
                                       a.nbytes, threads) == 0
            assert np.array_equal(a, b)
    assert lib.pb200_host_copy(None, None, 8, 1) != 0


def test_transformations_are_unwrapped_and_mapped_like_the_reference():
    """Transformed{ErrorMeasure,LogLikelihood,LogPDF} (what the reference's
    controllers build for `transformation=...`, pints/_optimisers/__init__.py:
    392-406): describe() peels them, and the batched row mapping equals the
    transformation's own per-vector arithmetic bit for bit."""
    pints = conftest_reference()
    if pints is None:
        pytest.skip('reference not available')
    from pints_b200 import _spec
    from pints_b200._transform import rows_log_jacobian_det, rows_to_model
    rng = np.random.RandomState(8)
    q = rng.randn(257, 3)
    cases = [
        pints.LogTransformation(3), pints.LogitTransformation(3),
        pints.IdentityTransformation(3),
        pints.ScalingTransformation([2.0, 0.5, 10.0]),
        pints.ScalingTransformation([2.0, 0.5, 10.0], [1.0, -1.0, 0.5]),
        pints.RectangularBoundariesTransformation([0, 1, 2], [1, 3, 10]),
        pints.ComposedTransformation(pints.LogTransformation(1),
                                     pints.RectangularBoundariesTransformation([0, 1], [1, 4])),
        pints.ComposedTransformation(pints.LogitTransformation(2),
                                     pints.IdentityTransformation(1)),
    ]
    for t in cases:
        want_p = np.array([t.to_model(row) for row in q])
        want_j = np.array([t.log_jacobian_det(row) for row in q])
        assert np.array_equal(rows_to_model(t, q), want_p), type(t).__name__
        assert np.array_equal(rows_log_jacobian_det(t, q), want_j), type(t).__name__
    # an unknown transformation object goes row by row through its own methods
    class Mine(pints.LogTransformation):
        pass
    mine = Mine(3)
    assert np.array_equal(rows_to_model(mine, q), np.exp(q))
    assert np.array_equal(rows_log_jacobian_det(mine, q),
                          np.array([np.sum(row) for row in q]))

    g, model, problem = fhn_objects(pints, pints.toy)
    ll = pints.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    t = pints.LogTransformation(3)
    spec = _spec.describe(t.convert_log_pdf(ll))
    assert spec.transform[0] is t and spec.transform[1] is False   # a likelihood: no Jacobian
    post = pints.LogPosterior(ll, pints.UniformLogPrior([0, 0, 0], [1, 1, 10]))
    spec = _spec.describe(pints.ProbabilityBasedError(t.convert_log_pdf(post)))
    assert spec.transform[1] is True and spec.transform[2] == -1.0 and spec.sign == -1.0
    assert spec.prior_lower is not None
    spec = _spec.describe(t.convert_error_measure(pints.SumOfSquaresError(problem)))
    assert spec.transform[1] is False and spec.sign == 1.0
    with pytest.raises(TypeError):          # two layers
        _spec.describe(t.convert_log_pdf(t.convert_log_pdf(ll)))
    with pytest.raises(TypeError, match='model space'):
        pb.CudaEvaluator(t.convert_log_pdf(ll)).device_problem()
