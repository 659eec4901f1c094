"""
Multi-GPU checks (need >= 2 GPUs; skipped otherwise): a CudaEvaluator over
several devices and the one-process-per-GPU ShardedEvaluator over NCCL return
exactly what one device returns.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import pints_oracle as po

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def build_function(pb):
    spec = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times,
                                    spec.values)
    return pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)


@pytest.mark.skipif(n_gpus() < 2, reason='needs 2 GPUs')
def test_multi_device_evaluator_equals_single():
    import pints_b200 as pb
    f = build_function(pb)
    xs = po.headline_fhn_points(10001)
    one = pb.CudaEvaluator(f, devices='cuda:0', as_list=False).evaluate(xs)
    g = n_gpus()
    many = pb.CudaEvaluator(f, devices=list(range(g)),
                            as_list=False).evaluate(xs)
    assert np.array_equal(one, many)
    few = pb.CudaEvaluator(f, devices=[0, 1], as_list=False).evaluate(xs[:3])
    assert np.array_equal(few, one[:3])


def _nccl_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        import pints_b200 as pb
        from pints_b200._dist import ShardedEvaluator
        f = build_function(pb)
        single = pb.CudaEvaluator(f, devices='cuda:%d' % rank, as_list=False)
        sharded = ShardedEvaluator(f, as_list=False)
        ok = True
        # several generations of different sizes: the two gathered arrays of
        # the fused exchange alternate, the last rank's block is ragged
        for n, seed in ((4099, 2), (4099, 3), (1031, 4), (7, 5), (4099, 6)):
            xs = po.headline_fhn_points(n, seed=seed)
            ok = ok and np.array_equal(sharded.evaluate(xs),
                                       single.evaluate(xs))
        # on NVLink-connected GPUs the scores travel by peer stores from
        # inside the kernel (pb200_eval_scatter), not by a separate all-gather
        fused = sharded._exchange is not None and sharded._exchange.fused
        out[rank] = (1 if ok else 0) + (2 if fused else 0)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(n_gpus() < 2, reason='needs 2 GPUs')
def test_sharded_evaluator_nccl():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * world)
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert [v & 1 for v in out] == [1] * world
    print('fused exchange used:', [bool(v & 2) for v in out])


def _mcmc_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        import pints_b200 as pb
        times = np.linspace(0, 100, 50)
        rng = np.random.RandomState(11)
        values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 2, 50)
        problem = pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values)
        post = pb.LogPosterior(
            pb.GaussianLogLikelihood(problem),
            pb.UniformLogPrior([0, 10, 0.1], [1, 100, 10]))
        n_chains, n_iter = 256, 1200
        x0 = np.array([0.1, 50, 2]) * (1 + 0.05 * rng.randn(n_chains, 3))
        ok = True
        for method in (pb.DifferentialEvolutionMCMC, pb.HaarioBardenetACMC):
            mc = pb.CudaMCMCController(post, n_chains, x0, method=method,
                                       seed=3, sharded=True)
            mc.set_max_iterations(n_iter)
            chains = mc.run()
            ok = ok and chains.shape == (n_chains, n_iter, 3)
            ok = ok and np.array_equal(chains[:, 0], x0)   # every rank has all
            tail = chains[:, n_iter // 2:].reshape(-1, 3)
            mean = tail.mean(0)
            ok = ok and abs(mean[0] - 0.1) < 0.01 and abs(mean[1] - 50) < 1.5
            # the two halves were produced by different ranks and both moved
            half = n_chains // 2
            for part in (chains[:half], chains[half:]):
                moved = (np.diff(part[:, :, 0], axis=1) != 0).mean()
                ok = ok and 0.02 < moved < 0.95
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(n_gpus() < 2, reason='needs 2 GPUs')
def test_sharded_device_mcmc():
    """Config 4 of BASELINE.json in small: chains sharded over the GPUs; the
    differential-evolution sampler all-gathers the positions every iteration."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * world)
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_mcmc_worker, args=(r, world, port, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def _xnes_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        import pints_b200 as pb
        f = build_function(pb)
        b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
        results = []
        for sharded in (False, True):
            ctl = pb.CudaOptimisationController(
                f, [0.3, 0.7, 2.5], boundaries=b, method=pb.DeviceXNES,
                sharded=sharded, devices='cuda:%d' % rank)
            ctl.optimiser().set_population_size(4099)       # ragged blocks
            ctl.set_max_iterations(25)
            ctl.set_max_unchanged_iterations(None)
            results.append(ctl.run())
        (x1, f1), (x2, f2) = results
        # the sharded run scores the same population block by block: identical
        ok = np.array_equal(x1, x2) and f1 == f2
        ok = ok and np.allclose(x1, [0.1, 0.5, 3.0], rtol=0.15)
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(n_gpus() < 2, reason='needs 2 GPUs')
def test_sharded_device_xnes():
    """DeviceXNES over two GPUs: every rank draws the same population, scores
    its block, the kernel exchanges the scores; same result as one GPU."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * world)
    port = 29800 + os.getpid() % 1000
    procs = [ctx.Process(target=_xnes_worker, args=(r, world, port, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    for p in procs:
        if p.is_alive():
            p.terminate()
    assert list(out) == [1] * world
