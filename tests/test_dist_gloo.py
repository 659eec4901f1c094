"""
world_size-2 checks of the multi-process path on the CPU (gloo): the block
partition, the padded all-gather of scores, and that every rank ends up with
the full, correctly ordered result.  The device launch is replaced by the CPU
oracle here (test-only); the NCCL version of the same code runs in bench.py
and tests/test_gpu_multi.py.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds():
    from pints_b200._dist import shard_bounds
    for n in (0, 1, 2, 7, 8, 9, 65536, 65537):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi, per = shard_bounds(n, world, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                seen += list(range(lo, hi))
            assert seen == list(range(n))


def _worker(rank, world, port, n, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import pints_b200 as pb
        from pints_b200._dist import ShardedEvaluator
        from oracle import pints_oracle as po

        rng = np.random.RandomState(1)
        times = np.linspace(0, 100, 40)
        values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 40)
        spec = po.Spec('logistic', times, values, 'sse')
        f = pb.SumOfSquaresError(pb.SingleOutputProblem(
            pb.toy.LogisticModel(), times, values))

        class CpuSharded(ShardedEvaluator):
            # test-only: oracle instead of the CUDA launch, CPU tensors
            def _width(self):
                return 2

            def _collective_device(self):
                return torch.device('cpu')

            def _local_scores(self, xs, out):
                out[:xs.shape[0]] = torch.from_numpy(po.scores(spec, xs))

        ev = CpuSharded(f)
        xs = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(n, 2))
        got = ev.evaluate(xs)
        want = po.scores(spec, xs)
        ok = isinstance(got, pb.ScoreList) and len(got) == n and \
            np.array_equal(np.array(got), want) and got.tolist() == list(want)
        # list-of-arrays input and an empty batch
        ok = ok and ev.evaluate([x for x in xs]) == got
        ok = ok and ev.evaluate([]) == []
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [1, 7, 64])
def test_sharded_evaluator_world_size_2(n):
    port = 29500 + (os.getpid() % 2000) + n
    out = mp.get_context('spawn').Array('i', [0, 0])
    procs = [mp.get_context('spawn').Process(
        target=_worker, args=(r, 2, port, n, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def _exchange_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pints_b200._dist import ScoreExchange

        class FakeProblem(object):
            """Stands in for DeviceProblem: score of a row = its first entry."""
            def eval(self, d_params, out=None, steps=None):
                out.copy_(d_params[:, 0])

        per = 5
        ex = ScoreExchange(per, 'cpu')
        ok = not ex.fused                      # no peer memory under gloo / CPU
        for gen in range(3):                   # buffers are reused per generation
            n_mine = per if rank == 0 else 3   # ragged last block
            x = torch.arange(n_mine, dtype=torch.float64).reshape(-1, 1) \
                + 100.0 * rank + 1000.0 * gen
            every = ex.run(FakeProblem(), x.repeat(1, 2), n_total=per + 3)
            ok = ok and every.shape[0] == per + 3
            want0 = torch.arange(per, dtype=torch.float64) + 1000.0 * gen
            want1 = torch.arange(3, dtype=torch.float64) + 100.0 + 1000.0 * gen
            ok = ok and torch.equal(every[:per], want0)
            ok = ok and torch.equal(every[per:per + 3], want1)
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_score_exchange_falls_back_to_all_gather_on_cpu():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Array('i', [0] * world)
    port = 29900 + os.getpid() % 1000
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, out))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1] * world


def _host_gather_worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import time
        from pints_b200._dist import HostGather, shard_bounds

        n = 11
        hg = HostGather(16, None, None, timeout=2.0)
        ok = hg.ok and not hg.pinned and hg.capacity == 16
        for gen in range(7):                   # more calls than result areas
            lo, hi, per = shard_bounds(n, world, rank)
            epoch = hg.next_epoch()
            slot = epoch % HostGather.SLOTS
            if rank == 1:
                time.sleep(0.02)               # the other rank has to wait
            hg.data[slot, lo:hi] = np.arange(lo, hi) + 100.0 * gen
            ok = ok and hg.area(epoch) == hg.data[slot].ctypes.data
            hg.publish(epoch)
            hg.wait(epoch)
            ok = ok and np.array_equal(hg.data[slot, :n], np.arange(n) + 100.0 * gen)
        # a rank that never delivers: the other one gives up with an error
        # naming it instead of hanging
        dist.barrier()
        epoch = hg.next_epoch()
        if rank == 0:
            hg.publish(epoch)
            try:
                hg.wait(epoch)
                ok = False
            except RuntimeError as e:
                ok = ok and 'ranks [1]' in str(e)
        dist.barrier()
        hg.close()
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


def test_host_gather_world_size_2():
    """The host-side all-gather of ShardedEvaluator (ranks of one node): one
    shared segment, result areas used in turn, epoch flags, time-out."""
    world = 2
    port = 29500 + (os.getpid() % 2000) + 977
    out = mp.get_context('spawn').Array('i', [0, 0])
    procs = [mp.get_context('spawn').Process(
        target=_host_gather_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]
