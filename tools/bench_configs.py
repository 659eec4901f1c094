#!/usr/bin/env python3
"""
Measures the other BASELINE.json configurations and the design-evidence numbers
(fused vs unfused, population sweep, wide parameter box, sampler loops) on one
GPU.  Reporting tool: writes gpurun_out/configs.json and prints a table.

    python tools/bench_configs.py [--quick]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pints_b200 as pb                                    # noqa: E402
from pints_b200._engine import fp64_probe                  # noqa: E402
from oracle import pints_oracle as po                      # noqa: E402  (inputs)


def time_device(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def fhn_problem():
    s = po.headline_fhn_spec()
    problem = pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times, s.values)
    return pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)


def main():
    quick = '--quick' in sys.argv
    out = {}
    dev = torch.device('cuda', 0)
    peak = fp64_probe(dev)
    out['fp64_peak_tflops_measured'] = peak
    rows = []

    # ---- config 5: population sweep (FHN) ----
    ll = fhn_problem()
    ev = pb.CudaEvaluator(ll, as_list=False)
    dp = ev.device_problem()
    spec = ev.spec()
    sweep = []
    for n in [1000, 10000, 65536, 100000, 1000000] + ([] if quick else [10000000]):
        xs = po.headline_fhn_points(n)
        d_x = torch.from_numpy(xs).to(dev)
        d_s = torch.empty(n, dtype=torch.float64, device=dev)
        steps = dp.new_steps(n)
        dp.eval(d_x, out=d_s, steps=steps)
        torch.cuda.synchronize()
        flops = float(spec.flops_per_eval(steps.cpu().numpy()).sum())
        ms = time_device(lambda: dp.eval(d_x, out=d_s), reps=5 if n > 1e6 else 10)
        ev.evaluate(xs)                 # first call: buffers, copy thread pool
        t0 = time.perf_counter()
        for _ in range(3):
            ev.evaluate(xs)
        e2e = (time.perf_counter() - t0) / 3
        sweep.append({'n': n, 'kernel_ms': ms, 'evals_per_s': n / (ms * 1e-3),
                      'tflops': flops / (ms * 1e-3) / 1e12,
                      'frac_of_peak': flops / (ms * 1e-3) / 1e12 / peak,
                      'e2e_evals_per_s': n / e2e})
        rows.append(('FHN sweep N=%d' % n, ms, n / (ms * 1e-3),
                     '%.1f%% of FP64 peak, e2e %.3g/s' % (100 * sweep[-1]['frac_of_peak'], n / e2e)))
        del d_x, d_s, steps
    out['fhn_sweep'] = sweep

    # ---- wide box (divergent cost) ----
    n = 65536
    rng = np.random.RandomState(3)
    xw = rng.uniform([0.0, 0.0, 1.0], [1.0, 1.0, 5.0], size=(n, 3))
    d_x = torch.from_numpy(xw).to(dev)
    d_s = torch.empty(n, dtype=torch.float64, device=dev)
    steps = dp.new_steps(n)
    dp.eval(d_x, out=d_s, steps=steps)
    st = steps.cpu().numpy()
    flops = float(spec.flops_per_eval(st).sum())
    ms = time_device(lambda: dp.eval(d_x, out=d_s))
    # the same vectors sorted by the time-scale parameter c (cost proxy)
    order = np.argsort(xw[:, 2], kind='stable')
    d_xs = torch.from_numpy(np.ascontiguousarray(xw[order])).to(dev)
    ms_sorted = time_device(lambda: dp.eval(d_xs, out=d_s))
    out['fhn_wide'] = {'n': n, 'kernel_ms': ms, 'kernel_ms_sorted_by_c': ms_sorted,
                       'mean_steps': float(st.mean()), 'max_steps': int(st.max()),
                       'tflops': flops / (ms * 1e-3) / 1e12,
                       'nonfinite': int((~torch.isfinite(d_s)).sum())}
    rows.append(('FHN wide box U([0,0,1],[1,1,5]) N=65536', ms, n / (ms * 1e-3),
                 'mean steps %.0f max %d; sorted by c: %.3f ms' % (st.mean(), st.max(), ms_sorted)))

    # ---- fused vs unfused (K6) ----
    xs = po.headline_fhn_points(n)
    d_x = torch.from_numpy(xs).to(dev)
    traj = torch.empty((n, 1000, 2), dtype=torch.float64, device=dev)
    ms_fused = time_device(lambda: dp.eval(d_x, out=d_s))
    ms_sim = time_device(lambda: dp.simulate(d_x, out=traj))
    ms_red = time_device(lambda: dp.score_trajectories(traj, d_x, out=d_s))
    bytes_traj = n * 1000 * 2 * 8
    out['fused_vs_unfused'] = {
        'fused_ms': ms_fused, 'simulate_ms': ms_sim, 'reduce_ms': ms_red,
        'unfused_total_ms': ms_sim + ms_red, 'trajectory_bytes': bytes_traj,
        'simulate_write_gbs': bytes_traj / (ms_sim * 1e-3) / 1e9,
        'reduce_read_gbs': bytes_traj / (ms_red * 1e-3) / 1e9}
    rows.append(('FHN fused eval N=65536', ms_fused, n / (ms_fused * 1e-3), '32 B HBM traffic per eval'))
    rows.append(('FHN unfused: simulate to HBM', ms_sim, n / (ms_sim * 1e-3),
                 'writes %.0f GB/s' % out['fused_vs_unfused']['simulate_write_gbs']))
    rows.append(('FHN unfused: reduce from HBM', ms_red, n / (ms_red * 1e-3),
                 'reads %.0f GB/s' % out['fused_vs_unfused']['reduce_read_gbs']))
    del traj

    # ---- config 1: logistic + SSE, population 64 ----
    times = np.linspace(0, 100, 1000)
    rng = np.random.RandomState(1)
    values = po.logistic_simulate([0.1, 50], times) + rng.normal(0, 1, 1000)
    sse = pb.SumOfSquaresError(pb.SingleOutputProblem(pb.toy.LogisticModel(), times, values))
    ev1 = pb.CudaEvaluator(sse)
    x64 = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(64, 2))
    ev1.evaluate(x64)
    t0 = time.perf_counter()
    for _ in range(200):
        ev1.evaluate(x64)
    gen = (time.perf_counter() - t0) / 200
    dp1 = ev1.device_problem()
    for nn in (64, 65536, 1000000):
        xl = np.array([0.1, 50]) * (1 + 0.1 * rng.randn(nn, 2))
        d_xl = torch.from_numpy(xl).to(dev)
        d_sl = torch.empty(nn, dtype=torch.float64, device=dev)
        msl = time_device(lambda: dp1.eval(d_xl, out=d_sl))
        rows.append(('Logistic+SSE 1000 pts N=%d' % nn, msl, nn / (msl * 1e-3),
                     '%.3g point-evals/s' % (nn * 1000 / (msl * 1e-3))))
        out['logistic_n%d' % nn] = {'kernel_ms': msl, 'evals_per_s': nn / (msl * 1e-3)}
    out['logistic_pop64_e2e_generation_ms'] = gen * 1e3
    rows.append(('Logistic+SSE pop 64 end-to-end generation', gen * 1e3, 64 / gen,
                 'reference: 1.8 ms / generation (SURVEY 6)'))

    # ---- config 3: HH-IK + Gaussian LL, 16384 chains ----
    hh = pb.toy.HodgkinHuxleyIKModel()
    th = hh.suggested_times()
    p = np.array(hh.suggested_parameters(), dtype=float)
    rng = np.random.RandomState(1)
    vh = po.hh_ik_simulate(p, th) + rng.normal(0, 40, th.shape)
    glh = pb.GaussianLogLikelihood(pb.SingleOutputProblem(hh, th, vh))
    n_ch = 16384
    x0 = np.hstack([p * rng.uniform(0.95, 1.05, (n_ch, 5)), 40 * rng.uniform(0.95, 1.05, (n_ch, 1))])
    evh = pb.CudaEvaluator(glh, as_list=False)
    dph = evh.device_problem()
    d_xh = torch.from_numpy(x0).to(dev)
    d_sh = torch.empty(n_ch, dtype=torch.float64, device=dev)
    msh = time_device(lambda: dph.eval(d_xh, out=d_sh))
    out['hh_ik_eval'] = {'n': n_ch, 'n_times': len(th), 'kernel_ms': msh,
                         'evals_per_s': n_ch / (msh * 1e-3),
                         'point_evals_per_s': n_ch * len(th) / (msh * 1e-3)}
    rows.append(('HH-IK 4800 pts + Gaussian LL N=16384', msh, n_ch / (msh * 1e-3),
                 '%.3g point-evals/s; reference 0.62 ms per eval' % (n_ch * len(th) / (msh * 1e-3))))
    iters = 100 if quick else 400
    mc = pb.CudaMCMCController(glh, n_ch, x0, method=pb.HaarioBardenetACMC, seed=1)
    mc.set_max_iterations(iters)
    mc.set_initial_phase_iterations(50)
    mc.run(to_host=False)
    per_it = mc.time() / iters
    out['hh_ik_hbac'] = {'chains': n_ch, 'iterations': iters, 'seconds': mc.time(),
                         'ms_per_iteration': per_it * 1e3,
                         'projected_s_for_1e4_iterations': per_it * 1e4,
                         'chain_iterations_per_s': n_ch / per_it}
    rows.append(('HaarioBardenetACMC 16384 chains on HH-IK (device loop)', per_it * 1e3, n_ch / per_it,
                 '1e4 iterations in %.1f s; reference host 112 us/chain-iteration = 18000 s' % (per_it * 1e4)))

    # ---- config 4: LV / SIR + Gaussian LL, DE-MCMC 4096 chains ----
    for name, model, tt, pp, sig in (
            ('LV', pb.toy.LotkaVolterraModel(), np.linspace(0, 20, 1000), [0.5, 0.15, 1.0, 0.3], 0.1),
            ('SIR', pb.toy.SIRModel(), np.linspace(1, 21, 1000), [0.026, 0.285, 38], 1.0)):
        rng = np.random.RandomState(1)
        oname = 'lotka_volterra' if name == 'LV' else 'sir'
        clean = po.simulate(oname, None, pp, tt)
        vv = clean + rng.normal(0, sig, clean.shape)
        gl = pb.GaussianLogLikelihood(pb.MultiOutputProblem(model, tt, vv))
        n_ch = 4096
        x0 = np.hstack([np.array(pp) * rng.uniform(0.98, 1.02, (n_ch, len(pp))),
                        sig * rng.uniform(0.95, 1.05, (n_ch, 2))])
        evm = pb.CudaEvaluator(gl, as_list=False)
        dpm = evm.device_problem()
        d_xm = torch.from_numpy(x0).to(dev)
        d_sm = torch.empty(n_ch, dtype=torch.float64, device=dev)
        stp = dpm.new_steps(n_ch)
        dpm.eval(d_xm, out=d_sm, steps=stp)
        msm = time_device(lambda: dpm.eval(d_xm, out=d_sm))
        big = np.tile(x0, (16, 1))
        d_xb = torch.from_numpy(big).to(dev)
        d_sb = torch.empty(len(big), dtype=torch.float64, device=dev)
        msb = time_device(lambda: dpm.eval(d_xb, out=d_sb))
        iters = 100 if quick else 300
        mc = pb.CudaMCMCController(gl, n_ch, x0, method=pb.DifferentialEvolutionMCMC, seed=2)
        mc.set_max_iterations(iters)
        mc.run(to_host=False)
        per_it = mc.time() / iters
        out[name.lower() + '_de'] = {
            'chains': n_ch, 'eval_kernel_ms': msm, 'mean_steps': float(stp.double().mean()),
            'evals_per_s_n4096': n_ch / (msm * 1e-3), 'evals_per_s_n65536': len(big) / (msb * 1e-3),
            'ms_per_iteration': per_it * 1e3, 'chain_iterations_per_s': n_ch / per_it}
        rows.append(('%s 1000 pts + Gaussian LL N=4096 / 65536' % name, msm, n_ch / (msm * 1e-3),
                     'N=65536: %.3g evals/s; mean steps %.0f' % (len(big) / (msb * 1e-3), stp.double().mean())))
        rows.append(('DifferentialEvolutionMCMC 4096 chains on %s (device loop)' % name, per_it * 1e3,
                     n_ch / per_it, 'reference host ask() alone 1630 ms/iteration'))

    # ---- config 2 as a whole optimisation loop: xNES generations ----
    b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
    out['xnes_generation_ms'] = {}
    for method, gens in ((pb.XNES, 40), (pb.DeviceXNES, 200)):
        per_gen = None
        for attempt in range(2):                  # the first run warms up
            np.random.seed(1)
            ctl = pb.CudaOptimisationController(
                ll, [0.1, 0.5, 3.0], [0.005, 0.025, 0.15], boundaries=b,
                method=method)
            ctl.optimiser().set_population_size(65536)
            ctl.set_max_iterations(gens)
            ctl.set_max_unchanged_iterations(None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctl.run()
            torch.cuda.synchronize()
            per_gen = (time.perf_counter() - t0) / gens
        out['xnes_generation_ms'][method.__name__] = per_gen * 1e3
        rows.append(('xNES generation (ask + score + tell), population 65536, %s'
                     % method.__name__, per_gen * 1e3, 65536 / per_gen,
                     'reference: 0.7 s of host loops per generation + scoring'))

    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'configs.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('| case | ms | units/s | note |')
    print('|---|---|---|---|')
    for name, ms, rate, note in rows:
        print('| %s | %.4f | %.4g | %s |' % (name, ms, rate, note))
    print('FP64 peak measured: %.2f TFLOP/s' % peak)


if __name__ == '__main__':
    main()
