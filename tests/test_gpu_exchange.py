"""
The fused score exchange of the multi-GPU path (pb200_eval_exchange /
pb200_exchange_finish), exercised on ONE GPU: several "ranks" share the device
and a stream, each scores its block into the staging arrays of all of them
and raises its flags, then every reading side must reproduce the scores of the
whole population in row order -- bit for bit what one pb200_eval of all rows
gives.  The same protocol across real GPUs (NVLink peer mappings, NVSwitch
multicast) is covered by tests/test_gpu_multi.py where two GPUs exist.
"""
import ctypes

import numpy as np
import pytest

from oracle import pints_oracle as po

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.fixture(scope='module')
def pb():
    import pints_b200
    pints_b200.require_cuda()
    return pints_b200


class Job(object):
    """`world` ranks on one device: staging arrays and flag arrays."""

    def __init__(self, world, per, device):
        from pints_b200 import _lib
        self.lib = _lib
        self.world, self.per = world, per
        pairs = world * per
        self.stage = [[torch.zeros(2 * pairs, dtype=torch.float64, device=device)
                       for _ in range(world)] for _ in range(2)]
        self.flags = [torch.zeros(16, dtype=torch.int64, device=device)
                      for _ in range(world)]
        self.epoch = 0

    def desc(self, rank):
        d = self.lib.ExchangeDesc()
        d.n_ranks, d.rank, d.per_rank = self.world, rank, self.per
        k = self.epoch & 1
        for q in range(self.world):
            d.stage[q] = self.stage[k][q].data_ptr()
            d.flags[q] = self.flags[q].data_ptr()
        d.stage_multicast = None
        d.epoch = self.epoch
        return d

    def generation(self, problem, d_x, rows):
        """All ranks score their blocks of d_x (rows per rank), then all read."""
        from pints_b200._engine import exchange_finish
        n = d_x.shape[0]
        self.epoch += 1
        for r in range(self.world):
            lo, hi = min(n, r * rows), min(n, (r + 1) * rows)
            problem.eval_exchange(d_x[lo:hi], self.desc(r))
        outs = []
        for r in range(self.world):
            out = torch.full((n,), -7.0, dtype=torch.float64, device=d_x.device)
            exchange_finish(self.desc(r), rows, n, out)
            outs.append(out)
        return outs


def fhn_problem(pb):
    spec = po.headline_fhn_spec()
    f = pb.GaussianKnownSigmaLogLikelihood(
        pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                              spec.values), 0.5)
    return pb.CudaEvaluator(f, as_list=False).device_problem()


def same_bits(a, b):
    return torch.equal(a.view(torch.int64), b.view(torch.int64))


@pytest.mark.parametrize('world', [1, 2, 3, 8])
def test_exchange_reproduces_one_launch(pb, world):
    dp = fhn_problem(pb)
    job = Job(world, 2048, dp.device)
    # sorted launches (>= 1024 rows per rank), unsorted ones, ragged last
    # blocks, ranks without rows, and a -inf / NaN row; the two staging arrays
    # alternate from generation to generation
    for n, seed in ((2048 * world, 2), (1500 * world - 7, 3), (40, 4),
                    (1, 5), (2048 * world - 1, 6)):
        xs = po.headline_fhn_points(n, seed=seed)
        if n > 8:
            xs[3, 2] = np.nan
        d_x = torch.from_numpy(xs).to(dp.device)
        want = dp.eval(d_x)
        rows = -(-n // world)
        for out in job.generation(dp, d_x, rows):
            assert same_bits(out, want)
    # every rank's counter stands at epoch * 2^20 on every rank
    assert all(f[:world].tolist() == [job.epoch << 20] * world for f in job.flags)


def test_exchange_with_a_closed_form_model(pb):
    times = np.linspace(0, 100, 200)
    rng = np.random.RandomState(1)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 200)
    f = pb.SumOfSquaresError(pb.SingleOutputProblem(pb.toy.LogisticModel(),
                                                    times, values))
    dp = pb.CudaEvaluator(f, as_list=False).device_problem()
    job = Job(4, 512, dp.device)
    for n in (2048, 1999, 5):
        xs = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(n, 2))
        d_x = torch.from_numpy(xs).to(dp.device)
        want = dp.eval(d_x)
        for out in job.generation(dp, d_x, -(-n // 4)):
            assert same_bits(out, want)


def test_exchange_argument_errors(pb):
    from pints_b200 import _lib
    from pints_b200._engine import exchange_finish
    dp = fhn_problem(pb)
    job = Job(2, 64, dp.device)
    job.epoch = 1
    d_x = torch.from_numpy(po.headline_fhn_points(65)).to(dp.device)
    with pytest.raises(_lib.LibraryError, match='do not fit'):
        dp.eval_exchange(d_x, job.desc(0))              # block too long
    bad = job.desc(0)
    bad.epoch = 0
    with pytest.raises(_lib.LibraryError):
        dp.eval_exchange(d_x[:8], bad)
    bad = job.desc(0)
    bad.stage[1] = None
    with pytest.raises(_lib.LibraryError, match='NULL or misaligned'):
        dp.eval_exchange(d_x[:8], bad)
    out = torch.empty(200, dtype=torch.float64, device=dp.device)
    with pytest.raises(_lib.LibraryError, match='do not fit'):
        exchange_finish(job.desc(0), 64, 200, out)      # more than 2 blocks
    with pytest.raises(_lib.LibraryError):
        exchange_finish(job.desc(0), 65, 100, out)      # rows > per_rank
    torch.cuda.synchronize()


def test_vector_that_blows_up_after_the_last_observation(pb):
    """A Lotka-Volterra vector with b < 0 blows up in finite time (T* = 2.049);
    the data end before that, so its score is finite -- the reference's odeint
    never looks past times[-1].  It shares a warp with vectors that take many
    more steps: once it has consumed its observations it must be frozen, not
    integrate into the singularity and be marked as failed."""
    times = np.linspace(0, 1.85, 60)
    rng = np.random.RandomState(4)
    model = pb.toy.LotkaVolterraModel()
    values = po.lv_simulate([0.5, 0.15, 1.0, 0.3], times) + rng.normal(0, 0.1, (60, 2))
    f = pb.GaussianLogLikelihood(pb.MultiOutputProblem(model, times, values))
    xs = np.tile([40.0, 1.5, 40.0, 3.0, 0.5, 0.5], (32, 1))    # fast oscillations
    xs[:, :4] *= 1 + 0.02 * rng.randn(32, 4)
    xs[5] = [0.5, -0.15, 1.0, 0.3, 0.5, 0.5]                    # blows up at 2.049
    ev = pb.CudaEvaluator(f, as_list=False, sort=False)
    got, steps = ev.evaluate_array(xs, return_steps=True)
    alone = pb.CudaEvaluator(f, as_list=False).evaluate(xs[5:6])
    assert np.all(np.isfinite(got))
    assert steps[5] * 3 < np.median(steps)          # it did finish long before the others
    assert got[5] == alone[0]                       # and does not depend on its neighbours
    spec = po.Spec('lotka_volterra', times, values, 'gaussian', tol=po.TRUTH_TOL)
    truth = po.scores(spec, xs[5:6])
    assert abs(got[5] - truth[0]) <= 1e-6 * abs(truth[0])
    # same for the two-lanes-per-vector kernel
    two = pb.CudaEvaluator(f, as_list=False, sort=False,
                           lanes_per_vector=2).evaluate(xs)
    assert np.all(np.isfinite(two)) and abs(two[5] - truth[0]) <= 1e-6 * abs(truth[0])
