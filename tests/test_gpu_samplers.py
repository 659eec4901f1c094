"""
Sampler kernels on the GPU (through the C ABI):

* bit-exact replay of the reference's HaarioBardenetACMC and
  DifferentialEvolutionMCMC: fed the proposals / draws / scores recorded from
  the real reference (tests/golden/*_replay.npz), the device reproduces every
  accept/reject decision, every next position and the adapted mu / sigma /
  log-lambda exactly (==, not approximately);
* the counter-based device RNG and the Cholesky proposal are statistically
  right;
* whole chains run on the device recover the posterior.
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import pints_oracle as po

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def pb():
    import pints_b200
    pints_b200.require_cuda()
    return pints_b200


def host(t):
    return t.cpu().numpy()


def test_hbac_replay_is_bit_exact(pb):
    g = load_golden('hbac_replay')
    n_iter, n, d = g['proposed'].shape
    s = pb.HaarioBardenetACMC(g['x0'], rng='replay')
    assert np.array_equal(s.state()[1], g['sigma0'])
    decisions = 0
    for it in range(n_iter):
        if it == int(g['warm']):
            s.set_initial_phase(False)
        xs = s.ask(proposed=g['proposed'][it] if it > 0 else None)
        assert np.array_equal(host(xs), g['proposed'][it])
        cur, cur_f, acc = s.tell(g['fx'][it], log_u=g['log_u'][it])
        assert np.array_equal(host(acc), g['accepted'][it]), it
        assert np.array_equal(host(cur), g['current'][it]), it
        assert np.array_equal(host(cur_f), g['current_f'][it]), it
        mu, sigma, log_lambda = s.state()
        assert np.array_equal(mu, g['mu'][it]), it
        assert np.array_equal(sigma, g['sigma'][it]), it
        assert np.array_equal(log_lambda, g['log_lambda'][it]), it
        if it >= int(g['warm']):
            assert s._gamma == g['gamma'][it]
        decisions += n
    # the fixture exercises both outcomes and the non-finite rule
    assert 0.05 < g['accepted'][1:].mean() < 0.95
    assert g['accepted'][7, 2] == 0 and g['accepted'][9, 4] == 0
    assert decisions == n_iter * n
    rate = s.acceptance_rate()
    # the first point is accepted without counting (reference :249-262)
    assert np.allclose(rate, g['accepted'][1:].sum(0) / n_iter)


@pytest.mark.parametrize('name,fixture', [
    ('HaarioACMC', 'haario_replay'),
    ('RaoBlackwellACMC', 'rao_blackwell_replay'),
    ('MetropolisRandomWalkMCMC', 'metropolis_replay')])
def test_adaptive_family_replay(pb, name, fixture):
    """(f) row 4: the other samplers that share the accept rule.  Fed the
    reference's recorded proposals, scores and log-uniforms, every decision and
    every next position is identical; the adapted state matches to ~1e-15
    (HaarioACMC and RaoBlackwellACMC use exp(log_ratio), where CUDA's exp and
    NumPy's may differ in the last place -- the difference is then carried
    along, hence 1e-12 over 120 adaptation steps)."""
    g = load_golden(fixture)
    n_iter, n, d = g['proposed'].shape
    s = getattr(pb, name)(g['x0'], rng='replay')
    adaptive_class = s.needs_initial_phase()
    for it in range(n_iter):
        if adaptive_class and it == int(g['warm']):
            s.set_initial_phase(False)
        xs = s.ask(proposed=g['proposed'][it] if it > 0 else None)
        assert np.array_equal(host(xs), g['proposed'][it])
        cur, cur_f, acc = s.tell(g['fx'][it], log_u=g['log_u'][it])
        assert np.array_equal(host(acc), g['accepted'][it]), it
        assert np.array_equal(host(cur), g['current'][it]), it
        assert np.array_equal(host(cur_f), g['current_f'][it]), it
        if adaptive_class:
            mu, sigma, log_lambda = s.state()
            assert np.array_equal(mu, g['mu'][it]), it        # no exp in mu
            scale = np.abs(g['sigma'][it]).max(axis=(1, 2), keepdims=True)
            assert np.all(np.abs(sigma - g['sigma'][it]) <= 1e-12 * scale), it
            assert np.allclose(log_lambda, g['log_lambda'][it], rtol=1e-12,
                               atol=1e-14), it
    assert g['accepted'][7, 2] == 0 and g['accepted'][9, 4] == 0
    if name == 'MetropolisRandomWalkMCMC':
        with pytest.raises(NotImplementedError):
            s.set_initial_phase(True)
        assert np.allclose(s.acceptance_rate(),
                           g['accepted'][1:].sum(0) / (n_iter - 1.0))
    else:
        assert g['accepted'][1:].mean() > 0.1
        assert np.allclose(s.acceptance_rate(),
                           g['accepted'][1:].sum(0) / n_iter)


def test_hbac_first_point_must_be_finite(pb):
    s = pb.HaarioBardenetACMC(np.ones((4, 2)), rng='replay')
    s.ask()
    with pytest.raises(ValueError):
        s.tell(np.array([0.0, -np.inf, 0.0, 0.0]))
    s2 = pb.HaarioBardenetACMC(np.ones((4, 2)), rng='replay')
    with pytest.raises(RuntimeError):
        s2.tell(np.zeros(4))


def test_de_replay_is_bit_exact(pb):
    g = load_golden('de_replay')
    n_iter, n, d = g['proposed'].shape
    s = pb.DifferentialEvolutionMCMC(n, g['x0'], rng='replay')
    assert np.array_equal(s.b_star(), g['b_star'])
    for it in range(n_iter):
        if it == 0:
            xs = s.ask()
        else:
            assert s.next_gamma() == g['gamma'][it]
            xs = s.ask(eps=g['eps'][it], r1=g['r1'][it], r2=g['r2'][it])
        assert np.array_equal(host(xs), g['proposed'][it]), it
        cur, cur_f, acc = s.tell(g['fx'][it],
                                 log_u=g['log_u'][it] if it > 0 else None)
        assert np.array_equal(host(acc), g['accepted'][it]), it
        assert np.array_equal(host(cur), g['current'][it]), it
        assert np.array_equal(host(cur_f), g['current_f'][it]), it
    assert g['accepted'][5, 1] == 0 and g['accepted'][6, 3] == 0   # NaN, -inf


def test_de_sharded_equals_whole(pb):
    """Two shards that see the all-gathered positions propose exactly what one
    sampler owning all chains proposes."""
    g = load_golden('de_replay')
    n_iter, n, d = g['proposed'].shape
    h = n // 2
    mean = g['x0'].mean(axis=0)
    a = pb.DifferentialEvolutionMCMC(h, g['x0'][:h], rng='replay', first=0,
                                     n_total=n, x0_mean=mean)
    b = pb.DifferentialEvolutionMCMC(n - h, g['x0'][h:], rng='replay', first=h,
                                     n_total=n, x0_mean=mean)
    for it in range(12):
        if it == 0:
            xa, xb = a.ask(), b.ask()
        else:
            pool = torch.cat([a.current(), b.current()])
            xa = a.ask(eps=g['eps'][it][:h], r1=g['r1'][it][:h],
                       r2=g['r2'][it][:h], all_current=pool)
            xb = b.ask(eps=g['eps'][it][h:], r1=g['r1'][it][h:],
                       r2=g['r2'][it][h:], all_current=pool)
        both = np.vstack([host(xa), host(xb)])
        assert np.array_equal(both, g['proposed'][it]), it
        lu = g['log_u'][it] if it > 0 else np.zeros(n)
        a.tell(g['fx'][it][:h], log_u=lu[:h] if it > 0 else None)
        b.tell(g['fx'][it][h:], log_u=lu[h:] if it > 0 else None)
        assert np.array_equal(
            np.vstack([host(a.current()), host(b.current())]),
            g['current'][it])


def test_philox_streams(pb):
    from pints_b200._samplers import _DeviceOps
    ops = _DeviceOps(None)
    n = 1 << 20
    z = torch.empty(n, dtype=torch.float64, device=ops.device)
    u = torch.empty(n, dtype=torch.float64, device=ops.device)
    ops.normal(z, 1234, 0)
    ops.uniform(u, 1234, 0)
    z, u = host(z), host(u)
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1) < 4e-3
    assert abs((z**3).mean()) < 2e-2 and abs((z**4).mean() - 3) < 5e-2
    assert 0 < u.min() and u.max() < 1
    assert abs(u.mean() - 0.5) < 2e-3 and abs(u.var() - 1 / 12) < 1e-3
    assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 5e-3
    # reproducible, and different for another seed / offset
    z2 = torch.empty(n, dtype=torch.float64, device=ops.device)
    ops.normal(z2, 1234, 0)
    assert np.array_equal(host(z2), z)
    ops.normal(z2, 1235, 0)
    assert not np.array_equal(host(z2), z)
    ops.normal(z2, 1234, 7)
    assert np.array_equal(host(z2)[:100], z[14:114])
    # pairs: distinct, never the chain itself, uniform over the others
    m, total, first = 30000, 10, 3
    r1 = torch.empty(m, dtype=torch.int32, device=ops.device)
    r2 = torch.empty(m, dtype=torch.int32, device=ops.device)
    total_big = m + first
    ops.pairs(r1, r2, first, total_big, 99, 0)
    a, b = host(r1), host(r2)
    me = first + np.arange(m)
    assert np.all(a != b) and np.all(a != me) and np.all(b != me)
    assert a.min() >= 0 and a.max() < total_big and b.max() < total_big
    r1s = torch.empty(7, dtype=torch.int32, device=ops.device)
    r2s = torch.empty(7, dtype=torch.int32, device=ops.device)
    counts = np.zeros((total, total))
    for rep in range(300):
        ops.pairs(r1s, r2s, first, total, 5, rep * 16)
        for j, (p, q) in enumerate(zip(host(r1s), host(r2s))):
            assert p != q and p != first + j and q != first + j
            assert 0 <= p < total and 0 <= q < total
            counts[first + j, p] += 1
    row = counts[first]
    assert row[first] == 0 and row.sum() == 300
    assert np.all(row[np.arange(total) != first] > 10)


def test_hbac_device_proposals_have_the_right_covariance(pb):
    rng = np.random.RandomState(0)
    d, n = 4, 200000
    a = rng.randn(d, d)
    cov = a @ a.T + 0.1 * np.eye(d)
    x0 = np.tile(rng.randn(d), (n, 1))
    s = pb.HaarioBardenetACMC(x0, sigma0=cov, rng='philox', seed=11)
    s.ask()
    s.tell(np.zeros(n))
    s._log_lambda.fill_(0.3)
    xs = host(s.ask())
    emp = np.cov((xs - x0).T)
    want = cov * np.exp(0.3)
    assert np.allclose(emp, want, rtol=0.03, atol=0.03 * np.abs(want).max())
    assert np.allclose((xs - x0).mean(0), 0, atol=0.02 * np.sqrt(want.max()))


def logistic_posterior(pb, n_times=200):
    rng = np.random.RandomState(3)
    times = np.linspace(0, 100, n_times)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 2, n_times)
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)
    post = pb.LogPosterior(pb.GaussianLogLikelihood(problem),
                           pb.UniformLogPrior([0, 10, 0.1], [1, 100, 20]))
    return post


@pytest.mark.parametrize('method', ['hbac', 'de', 'haario', 'rao_blackwell'])
def test_device_mcmc_recovers_the_posterior(pb, method):
    post = logistic_posterior(pb)
    n_chains, n_iter = 256, 1500
    rng = np.random.RandomState(4)
    x0 = np.array([0.1, 50, 2]) * (1 + 0.05 * rng.randn(n_chains, 3))
    cls = {'hbac': pb.HaarioBardenetACMC, 'de': pb.DifferentialEvolutionMCMC,
           'haario': pb.HaarioACMC, 'rao_blackwell': pb.RaoBlackwellACMC}[method]
    mc = pb.CudaMCMCController(post, n_chains, x0, method=cls, seed=7)
    mc.set_max_iterations(n_iter)
    chains = mc.run()
    assert chains.shape == (n_chains, n_iter, 3)
    assert np.all(np.isfinite(chains))
    assert np.array_equal(chains[:, 0], x0)            # first sample is x0
    tail = chains[:, n_iter // 2:].reshape(-1, 3)
    mean, sd = tail.mean(0), tail.std(0)
    assert abs(mean[0] - 0.1) < 0.01 and abs(mean[1] - 50) < 1.5
    assert abs(mean[2] - 2) < 0.4
    assert np.all(sd > 0)
    # chains really move, and stay inside the prior box
    moved = (np.diff(chains[:, :, 0], axis=1) != 0).mean()
    assert 0.05 < moved < 0.9
    assert tail[:, 0].min() >= 0 and tail[:, 1].max() < 100
    if method != 'de':
        rate = mc.sampler().acceptance_rate()
        assert 0.1 < rate.mean() < 0.6
    assert mc.n_evaluations() == n_chains * n_iter and mc.time() > 0
    with pytest.raises(RuntimeError):
        mc.run()


def test_cuda_optimisation_controller_with_vectorised_xnes(pb):
    rng = np.random.RandomState(8)
    times = np.linspace(0, 100, 300)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 300)
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)
    np.random.seed(1)
    opt = pb.CudaOptimisationController(
        pb.SumOfSquaresError(problem), [0.3, 20],
        boundaries=pb.RectangularBoundaries([0, 1], [1, 200]), method=pb.XNES)
    opt.optimiser().set_population_size(512)
    opt.set_max_iterations(60)
    x, f = opt.run()
    assert abs(x[0] - 0.1) < 5e-3 and abs(x[1] - 50) < 0.5
    assert f < 1.2 * 300
    # maximising a log-likelihood finds the same point
    np.random.seed(1)
    opt = pb.CudaOptimisationController(
        pb.GaussianKnownSigmaLogLikelihood(problem, 1.0), [0.3, 20],
        boundaries=pb.RectangularBoundaries([0, 1], [1, 200]))
    opt.optimiser().set_population_size(512)
    opt.set_max_iterations(60)
    x2, f2 = opt.run()
    assert np.allclose(x2, x, rtol=1e-3) and f2 < 0


@pytest.mark.parametrize('name', ['SNES', 'PSO', 'BareCMAES'])
def test_cuda_optimisation_controller_with_other_vectorised_optimisers(pb, name):
    rng = np.random.RandomState(8)
    times = np.linspace(0, 100, 300)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 300)
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)
    np.random.seed(2)
    opt = pb.CudaOptimisationController(
        pb.MeanSquaredError(problem), [0.3, 20],
        boundaries=pb.RectangularBoundaries([0, 1], [1, 200]),
        method=getattr(pb, name))
    opt.optimiser().set_population_size(2048)
    opt.set_max_iterations(80)
    x, f = opt.run()
    assert abs(x[0] - 0.1) < 1e-2 and abs(x[1] - 50) < 1.0, (x, f)
    assert f < 1.3


@pytest.mark.parametrize('name', ['HaarioBardenetACMC', 'HaarioACMC',
                                  'RaoBlackwellACMC',
                                  'MetropolisRandomWalkMCMC'])
def test_fused_ask_tell_gives_identical_chains(pb, name):
    """pb200_ac_ask / pb200_ac_tell (one launch each) against the separate
    launches (normals, proposal, uniforms, log, accept, adapt): same draws,
    same arithmetic, so the chains, the adapted state and the acceptance
    counts are identical bit for bit."""
    post = logistic_posterior(pb)
    dp = pb.CudaEvaluator(post, as_list=False).device_problem()
    rng = np.random.RandomState(6)
    n, n_iter = 96, 120
    x0 = np.array([0.1, 50, 2]) * (1 + 0.05 * rng.randn(n, 3))
    out = []
    for fused in (True, False):
        s = getattr(pb, name)(x0, seed=5, fused=fused)
        samples = torch.zeros((n, n_iter, 3), dtype=torch.float64, device='cuda')
        fx = torch.empty(n, dtype=torch.float64, device='cuda')
        for it in range(n_iter):
            if it == 30 and s.needs_initial_phase():
                s.set_initial_phase(False)
            dp.eval(s.ask(), out=fx)
            s.tell(fx, store=(samples, it))
        out.append((host(samples), s.state(), s.acceptance_rate()))
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    assert np.array_equal(out[0][2], out[1][2])
    assert (np.diff(out[0][0][:, :, 0], axis=1) != 0).mean() > 0.02


def test_fused_de_ask_tell_gives_identical_chains(pb):
    """pb200_de_ask / pb200_ac_tell against the separate launches of the
    differential-evolution sampler: identical chains."""
    post = logistic_posterior(pb)
    dp = pb.CudaEvaluator(post, as_list=False).device_problem()
    rng = np.random.RandomState(7)
    n, n_iter = 64, 150
    x0 = np.array([0.1, 50, 2]) * (1 + 0.05 * rng.randn(n, 3))
    out = []
    for fused in (True, False):
        s = pb.DifferentialEvolutionMCMC(n, x0, seed=5, fused=fused)
        samples = torch.zeros((n, n_iter, 3), dtype=torch.float64, device='cuda')
        fx = torch.empty(n, dtype=torch.float64, device='cuda')
        for it in range(n_iter):
            dp.eval(s.ask(), out=fx)
            s.tell(fx, store=(samples, it))
        out.append(host(samples))
    assert np.array_equal(out[0], out[1])
    assert (np.diff(out[0][:, :, 0], axis=1) != 0).mean() > 0.02


def test_device_xnes_replays_host_xnes(pb, monkeypatch):
    """The device-resident xNES draws from Philox instead of numpy.random, so
    its trajectory differs from the reference's; but fed the SAME normals and
    scores its update is the reference's: the host XNES of this package (itself
    checked against pints.XNES, tests/test_controllers_cpu.py) reproduces it
    generation by generation when numpy.random.normal hands out the device's
    normals."""
    x0 = np.array([0.3, 20.0, 1.5])
    lower, upper = [0.0, 1.0, 0.5], [1.0, 200.0, 1.8]
    target = np.array([0.1, 50.0, 1.0])
    n = 2000

    def score(x):                       # a smooth bowl, minimised at `target`
        return np.sum(((x - target) / [0.1, 10.0, 1.0]) ** 2, axis=1)

    dev = pb.DeviceXNES(x0, [0.05, 5.0, 0.3],
                        pb.RectangularBoundaries(lower, upper), seed=11)
    ref = pb.XNES(x0, [0.05, 5.0, 0.3], pb.RectangularBoundaries(lower, upper))
    dev.set_population_size(n)
    ref.set_population_size(n)
    outside = 0
    for generation in range(25):
        d_x = dev.ask()
        z = host(dev.normals())
        # the draws are the documented Philox stream of the C ABI
        if generation == 0:
            from pints_b200._samplers import _DeviceOps
            buf = torch.empty(n * 3, dtype=torch.float64, device=d_x.device)
            _DeviceOps(None).normal(buf, 11, 0)      # seed 11, offset 0
            assert np.array_equal(host(buf).reshape(n, 3), z)
        monkeypatch.setattr(np.random, 'normal', lambda *a, **k: z.copy())
        xs = ref.ask()
        monkeypatch.undo()
        inside = host(dev.inside()).astype(bool)
        outside += int((~inside).sum())
        assert len(xs) == inside.sum()
        # same points where inside; the mean where the reference withholds them
        assert np.allclose(host(dev.points())[inside], xs, rtol=1e-12, atol=1e-12)
        assert np.array_equal(host(d_x)[~inside],
                              np.tile(dev.x_guessed(), ((~inside).sum(), 1)))
        fx = score(host(d_x))           # the same scores to both: same ranks
        ref.tell(fx[inside])
        dev.tell(torch.from_numpy(fx).to(d_x.device))
        assert np.allclose(dev.x_guessed(), ref.x_guessed(), rtol=1e-8)
        assert dev.f_best() == ref.f_best()
        assert np.allclose(dev.x_best(), ref.x_best(), rtol=1e-12)
    assert outside > 0                  # the boundary path was exercised
    assert np.allclose(dev.x_guessed(), target, rtol=0.05)
    with pytest.raises(Exception):
        dev.tell(torch.zeros(n, dtype=torch.float64, device=d_x.device))
    with pytest.raises(ValueError):
        pb.DeviceXNES([0.5, 0.5], boundaries=pb.RectangularBoundaries([0], [1]))


def test_cuda_optimisation_controller_with_device_xnes(pb):
    """BASELINE configs[1] end to end: FitzHugh-Nagumo, known-sigma Gaussian
    log-likelihood, xNES with the population resident on the GPU."""
    h = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), h.times,
                                    h.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
    opt = pb.CudaOptimisationController(ll, [0.3, 0.7, 2.5], boundaries=b,
                                        method=pb.DeviceXNES)
    opt.optimiser().set_population_size(4096)
    opt.set_max_iterations(40)
    x, f = opt.run()
    # the maximum-likelihood point of the noisy series, near the true values
    assert np.allclose(x, [0.1, 0.5, 3.0], rtol=0.15)
    # the maximum is at least as good as the likelihood at the true parameters
    at_truth = pb.CudaEvaluator(ll, as_list=False).evaluate(
        np.array([[0.1, 0.5, 3.0]]))[0]
    assert f >= at_truth - 1e-6 * abs(at_truth)
    assert opt.evaluations() == 40 * 4096
    # the generations after the second were replayed from a CUDA graph, and
    # an eager run (no graph) follows exactly the same trajectory
    assert opt.optimiser()._graph is not None, opt.optimiser()._graph_failed
    eager = pb.CudaOptimisationController(ll, [0.3, 0.7, 2.5], boundaries=b,
                                          method=pb.DeviceXNES)
    eager.optimiser().use_graph = False
    eager.optimiser().set_population_size(4096)
    eager.set_max_iterations(40)
    xe, fe = eager.run()
    assert eager.optimiser()._graph is None
    assert np.array_equal(xe, x) and fe == f
    # ranking with pb200_argsort_f64 instead of the library sort: the same
    # optimum (tied +inf scores of out-of-bounds points may be ordered
    # differently, which only changes the order of a sum)
    for n in (4096, 6000):
        runs = []
        for native in (False, True):
            ctl = pb.CudaOptimisationController(
                ll, [0.3, 0.7, 2.5], boundaries=b, method=pb.DeviceXNES)
            ctl.optimiser().native_sort = native
            ctl.optimiser().set_population_size(n)
            ctl.set_max_iterations(30)
            runs.append(ctl.run())
        assert np.allclose(runs[0][0], runs[1][0], rtol=1e-9)
        assert np.isclose(runs[0][1], runs[1][1], rtol=1e-12)
    # had the sample sort given up (status != 0), the state is put back, the
    # generation is ranked again with the library sort and the run carries on
    # with it: same result as a run that never saw the flag
    problem_dev = pb.CudaEvaluator(ll, as_list=False).device_problem()
    twins = []
    for pretend in (False, True):
        opt = pb.DeviceXNES([0.3, 0.7, 2.5], boundaries=b)
        opt.native_sort = True
        opt.set_population_size(6000)
        opt.generation(problem_dev, negate=True)
        before = opt.x_guessed().copy()
        if pretend:
            real, seen = opt._sort_failed, []

            def flagged(real=real, seen=seen, opt=opt):
                if not seen:
                    seen.append(1)
                    opt._h_status[0] = 1
                return real()
            opt._sort_failed = flagged
        hist = opt.generations(problem_dev, 3, negate=True)
        assert hist.shape == (3, 5) and hist[-1, 0] == opt.f_best()
        assert opt._sort_gave_up == pretend
        assert not np.array_equal(opt.x_guessed(), before)
        assert opt._generations_done == 4
        twins.append((opt.x_guessed().copy(), opt.f_best(), opt.x_best().copy()))
    assert np.allclose(twins[0][0], twins[1][0], rtol=1e-9)
    assert np.isclose(twins[0][1], twins[1][1], rtol=1e-12)
    # minimising the error measure of the same problem finds the same point
    opt = pb.CudaOptimisationController(
        pb.SumOfSquaresError(problem), [0.3, 0.7, 2.5], boundaries=b,
        method=pb.DeviceXNES)
    opt.optimiser().set_population_size(4096)
    opt.set_max_iterations(40)
    x2, f2 = opt.run()
    assert np.allclose(x2, x, rtol=1e-3) and f2 > 0


@pytest.mark.parametrize('pair', [('DeviceXNES', 'XNES'), ('DeviceSNES', 'SNES')])
def test_device_xnes_twelve_parameters(pb, monkeypatch, pair):
    """The sums kernel splits its block differently for large d (d + d^2
    output quantities): twelve parameters, no boundaries, against the host
    XNES / SNES fed the same normals and scores."""
    d, n = 12, 600
    rng = np.random.RandomState(4)
    target = rng.uniform(-1, 1, d)
    x0 = target + 0.3
    dev = getattr(pb, pair[0])(x0, 0.2, seed=5)
    ref = getattr(pb, pair[1])(x0, 0.2)
    dev.set_population_size(n)
    ref.set_population_size(n)
    for generation in range(12):
        d_x = dev.ask()
        z = host(dev.normals())
        monkeypatch.setattr(np.random, 'normal', lambda *a, **k: z.copy())
        xs = ref.ask()
        monkeypatch.undo()
        assert np.allclose(host(d_x), xs, rtol=1e-12, atol=1e-12)
        fx = np.sum((host(d_x) - target) ** 2, axis=1)
        ref.tell(fx)
        dev.tell(torch.from_numpy(fx).to(d_x.device))
        assert np.allclose(dev.x_guessed(), ref.x_guessed(), rtol=1e-8, atol=1e-10)
        assert dev.f_best() == ref.f_best()
    assert dev.f_best() < np.sum(0.3 ** 2 * np.ones(d))


@pytest.mark.parametrize('n', [1, 5, 1000, 4096, 4097, 65536, 200001])
def test_device_argsort(pb, n):
    """pb200_argsort_f64 (sample sort) against numpy's stable argsort: NaN
    last, ties by index, infinities, signed zeros, heavy tails and constant
    arrays (every bucket boundary is a tie then)."""
    rng = np.random.RandomState(n)
    cases = [rng.randn(n), np.exp(8 * rng.randn(n)), np.zeros(n),
             np.round(rng.randn(n), 1), -np.arange(n, dtype=float)]
    special = rng.randn(n)
    special[rng.rand(n) < 0.05] = np.inf
    special[rng.rand(n) < 0.05] = -np.inf
    special[rng.rand(n) < 0.05] = np.nan
    special[rng.rand(n) < 0.05] = 0.0
    special[rng.rand(n) < 0.05] = -0.0
    cases.append(special)
    sorter = pb.DeviceArgsort(n, None, force_native=True)
    assert sorter.native
    for values in cases:
        d = torch.from_numpy(values).to(sorter.device)
        got = host(sorter(d))
        assert int(sorter.status.item()) == 0
        want = np.argsort(values, kind='stable')
        assert np.array_equal(got, want)


@pytest.mark.parametrize('d,n', [(1, 7), (2, 1), (5, 33), (16, 100)])
def test_device_xnes_small_and_odd_shapes(pb, monkeypatch, d, n):
    """One parameter, one sample, sixteen parameters: every block split of the
    sums kernel and the single-block ranking path against the host XNES."""
    target = np.linspace(-1, 1, d)
    x0 = target + 0.5
    dev = pb.DeviceXNES(x0, 0.3, seed=3)
    ref = pb.XNES(x0, 0.3)
    dev.set_population_size(n)
    ref.set_population_size(n)
    for generation in range(15):
        d_x = dev.ask()
        z = host(dev.normals())
        monkeypatch.setattr(np.random, 'normal', lambda *a, **k: z.copy())
        ref.ask()
        monkeypatch.undo()
        fx = np.sum((host(d_x) - target) ** 2, axis=1)
        ref.tell(fx)
        dev.tell(torch.from_numpy(fx).to(d_x.device))
        assert np.allclose(dev.x_guessed(), ref.x_guessed(), rtol=1e-8, atol=1e-10)
    assert dev.f_best() == ref.f_best()


@pytest.mark.parametrize('name', ['HaarioBardenetACMC', 'HaarioACMC',
                                  'RaoBlackwellACMC',
                                  'MetropolisRandomWalkMCMC',
                                  'DifferentialEvolutionMCMC'])
def test_mcmc_graph_replay_gives_identical_chains(pb, name):
    """CudaMCMCController replays the iterations after the second from a CUDA
    graph (per-iteration scalars from a device table): every sample of every
    chain equals the eager run's, across the end of the initial phase."""
    rng = np.random.RandomState(12)
    times = np.linspace(0, 100, 120)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 2, 120)
    problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)
    post = pb.LogPosterior(pb.GaussianLogLikelihood(problem),
                           pb.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
    n_chains = 24
    x0 = np.array([0.1, 50, 2.0]) * (1 + 0.02 * rng.randn(n_chains, 3))
    chains, rates = [], []
    # the same run three ways: one persistent launch for all iterations after
    # the second (independent chains on a closed-form model), replay of a
    # captured iteration from a CUDA graph, and launch by launch
    independent = name != 'DifferentialEvolutionMCMC'
    for mode in ('chain loop', 'graph', 'eager'):
        mc = pb.CudaMCMCController(post, n_chains, x0,
                                   method=getattr(pb, name), seed=3)
        mc.use_chain_loop = mode == 'chain loop'
        mc.use_graph = mode != 'eager'
        mc.graph_min_iterations = 4
        mc.set_max_iterations(90)
        if mc.sampler().needs_initial_phase():
            mc.set_initial_phase_iterations(30)
        chains.append(mc.run())
        in_loop = mode == 'chain loop' and independent
        assert mc.chain_loop_iterations == (88 if in_loop else 0)
        assert mc.graph_replayed == (88 if mode != 'eager' and not in_loop else 0)
        if hasattr(mc.sampler(), 'acceptance_rate'):
            rates.append(mc.sampler().acceptance_rate())
            assert np.all(rates[-1] >= 0) and np.all(rates[-1] <= 1)
    if rates:                       # same host bookkeeping after the run
        assert np.array_equal(rates[0], rates[1])
        assert np.array_equal(rates[0], rates[2])
    assert chains[0].shape == (n_chains, 90, 3)
    assert np.array_equal(chains[0], chains[1])
    assert np.array_equal(chains[0], chains[2])
    assert np.isfinite(chains[0]).all()
    assert len(np.unique(chains[0][:, :, 0])) > n_chains        # they moved
