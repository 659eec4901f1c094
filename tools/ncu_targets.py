#!/usr/bin/env python3
"""Small workloads that launch each kernel north_star wants ncu evidence for,
a few times each, so that `ncu -k regex:<kernel>` can capture one launch
(profiles/run_ncu.sh drives it):

    python tools/ncu_targets.py unfused    ode_kernel<.., TRAJ> + score_traj_kernel
                                           (trajectories through HBM: the baseline
                                           the fused kernel is set against)
    python tools/ncu_targets.py logistic   analytic_kernel<LogisticEval> (closed form)
    python tools/ncu_targets.py samplers   ac_ask / ac_tell (HaarioBardenetACMC on the
                                           logistic model, 16384 chains), de_ask
                                           (DifferentialEvolutionMCMC, 4096 chains, LV)
    python tools/ncu_targets.py argsort    the kernels of pb200_argsort_f64 at 65536
    python tools/ncu_targets.py exchange   pb200_eval_exchange + pb200_exchange_finish
                                           with 8 ranks' worth of staging on ONE GPU
                                           (the NVLink leg needs several GPUs; ncu
                                           must not wrap a multi-rank job)
Prints event-timed figures of the same launches (not under ncu: run it plain
for those).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pints_b200 as pb                                   # noqa: E402
from oracle import pints_oracle as po                     # noqa: E402  (inputs only)


def timed(f, reps=5, warm=2):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def fhn():
    spec = po.headline_fhn_spec()
    f = pb.GaussianKnownSigmaLogLikelihood(
        pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times, spec.values), 0.5)
    return pb.CudaEvaluator(f, as_list=False).device_problem()


def unfused(n=65536):
    dp = fhn()
    d_x = torch.from_numpy(po.headline_fhn_points(n)).cuda()
    traj = torch.empty((n, 1000, 2), dtype=torch.float64, device='cuda')
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    fused = dp.eval(d_x)
    t_sim = timed(lambda: dp.simulate(d_x, out=traj))
    t_red = timed(lambda: dp.score_trajectories(traj, d_x, out=out))
    t_fused = timed(lambda: dp.eval(d_x, out=out))
    nbytes = traj.numel() * 8
    rel = float(((out - fused).abs() / fused.abs()).max())
    print('unfused: simulate %.3f ms (%.0f GB/s written), reduce %.3f ms (%.0f GB/s '
          'read), fused %.3f ms; unfused/fused = %.2f; max rel diff %.1e'
          % (t_sim, nbytes / t_sim / 1e6, t_red, nbytes / t_red / 1e6, t_fused,
             (t_sim + t_red) / t_fused, rel))


def logistic(n=65536):
    times = np.linspace(0, 100, 1000)
    rng = np.random.RandomState(1)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 1000)
    f = pb.SumOfSquaresError(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values))
    dp = pb.CudaEvaluator(f, as_list=False).device_problem()
    d_x = torch.from_numpy(np.array([0.1, 50]) * (1 + 0.1 * rng.randn(n, 2))).cuda()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    t = timed(lambda: dp.eval(d_x, out=out), reps=20)
    print('logistic: %d vectors x 1000 points %.4f ms = %.3g point-evaluations/s'
          % (n, t, n * 1000 / t * 1e3))


def samplers():
    times = np.linspace(0, 100, 1000)
    rng = np.random.RandomState(1)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 1000)
    ll = pb.GaussianLogLikelihood(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values))
    dp = pb.CudaEvaluator(ll, as_list=False).device_problem()
    n = 16384
    x0 = np.array([0.1, 50, 1.0]) * rng.uniform(0.95, 1.05, (n, 3))
    s = pb.HaarioBardenetACMC(x0, seed=1)
    fx = torch.empty(n, dtype=torch.float64, device='cuda')
    dp.eval(s.ask(), out=fx)
    s.tell(fx)
    s.set_initial_phase(False)

    def it():
        dp.eval(s.ask(), out=fx)
        s.tell(fx)
    t = timed(it, reps=20, warm=5)
    print('HaarioBardenetACMC, %d chains, logistic: %.1f us per iteration' % (n, t * 1e3))
    m = pb.toy.LotkaVolterraModel()
    lt = np.linspace(0, 20, 1000)
    lv = po.lv_simulate([0.5, 0.15, 1.0, 0.3], lt) + rng.normal(0, 0.1, (1000, 2))
    lll = pb.GaussianLogLikelihood(pb.MultiOutputProblem(m, lt, lv))
    ldp = pb.CudaEvaluator(lll, as_list=False).device_problem()
    n = 4096
    x0 = np.array([0.5, 0.15, 1.0, 0.3, 0.1, 0.1]) * rng.uniform(0.98, 1.02, (n, 6))
    de = pb.DifferentialEvolutionMCMC(n, x0, seed=2)
    fx = torch.empty(n, dtype=torch.float64, device='cuda')
    ldp.eval(de.ask(), out=fx)
    de.tell(fx)

    def it2():
        ldp.eval(de.ask(), out=fx)
        de.tell(fx)
    t = timed(it2, reps=20, warm=5)
    print('DifferentialEvolutionMCMC, %d chains, Lotka-Volterra: %.1f us per iteration'
          % (n, t * 1e3))


def argsort(n=65536):
    s = pb.DeviceArgsort(n, None)
    v = torch.randn(n, dtype=torch.float64, device=s.device)
    t = timed(lambda: s(v), reps=50, warm=5)
    t2 = timed(lambda: torch.sort(v), reps=50, warm=5)
    print('argsort of %d doubles: %.1f us (library sort %.1f us)' % (n, t * 1e3, t2 * 1e3))


def exchange(n=65536, world=8):
    """One GPU plays rank 0 of 8: its kernel stores into 8 staging arrays (all
    local here), which is the store traffic of the real job without the links."""
    from pints_b200 import _lib
    from pints_b200._engine import exchange_finish
    dp = fhn()
    d_x = torch.from_numpy(po.headline_fhn_points(n)).cuda()
    pairs = world * n
    stage = [torch.zeros(2 * pairs, dtype=torch.float64, device='cuda') for _ in range(world)]
    flags = [torch.zeros(16, dtype=torch.int64, device='cuda') for _ in range(world)]
    out = torch.empty(pairs, dtype=torch.float64, device='cuda')
    state = {'epoch': 0}

    def desc():
        d = _lib.ExchangeDesc()
        d.n_ranks, d.rank, d.per_rank = world, 0, n
        for q in range(world):
            d.stage[q] = stage[q].data_ptr()
            d.flags[q] = flags[q].data_ptr()
        d.epoch = state['epoch']
        return d

    def it():
        state['epoch'] += 1
        d = desc()
        dp.eval_exchange(d_x, d)
        # the other ranks' flags: raised by hand (they have no kernel here)
        flags[0][1:world] = state['epoch'] << 20
        exchange_finish(d, n, pairs, out)
    t = timed(it, reps=10, warm=3)
    t_local = timed(lambda: dp.eval(d_x), reps=10, warm=3)
    print('exchange on one GPU (8 staging arrays): %.3f ms per generation, plain eval %.3f ms'
          % (t, t_local))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    table = {'unfused': unfused, 'logistic': logistic, 'samplers': samplers,
             'argsort': argsort, 'exchange': exchange}
    for name, fn in table.items():
        if which in ('all', name):
            fn()
