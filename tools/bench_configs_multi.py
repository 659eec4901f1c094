#!/usr/bin/env python3
"""
The BASELINE.json configurations as written, on N GPUs (one process per GPU,
torchrun; N = 1 runs without it):

  configs[4]  FitzHugh-Nagumo population sweep, 10^3 .. 10^7 GLOBAL vectors
              split over the N GPUs (strong scaling: at 65 536 in total and
              N = 8 a GPU gets 8192 vectors and sits on the latency floor),
              scores exchanged so that every rank holds all of them
  configs[1]  one xNES generation (DeviceXNES: ask + score + exchange + tell),
              population 65 536 in total
  configs[2]  HaarioBardenetACMC, 16 384 chains on HH-IK (4800 points),
              chains sharded: no exchange at all
  configs[3]  DifferentialEvolutionMCMC, 4096 chains on Lotka-Volterra / SIR,
              chains sharded: an all-gather of the positions per iteration

    python tools/bench_configs_multi.py [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P tools/bench_configs_multi.py

Device-timed (CUDA events per repetition, median; maximum over ranks).  Rank 0
writes gpurun_out/configs_multi_<N>gpu.json; tools/configs_table.py merges the
files of several N into the table committed under profiles/.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import pints_b200 as pb                                    # noqa: E402
from pints_b200._dist import ScoreExchange, shard_bounds   # noqa: E402
from pints_b200._engine import fp64_probe                  # noqa: E402
from oracle import pints_oracle as po                      # noqa: E402  (inputs)

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
quick = '--quick' in sys.argv


def max_over_ranks(value):
    if world == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def time_device(fn, reps=10, warm=3):
    """Median of per-repetition CUDA-event times, maximum over ranks."""
    for _ in range(warm):
        fn()
    barrier()
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for i in range(reps):
        e0[i].record()
        fn()
        e1[i].record()
    barrier()
    return max_over_ranks(float(np.median([a.elapsed_time(b) for a, b in zip(e0, e1)])))


def main():
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    out = {'n_gpus': world, 'fp64_peak_tflops_measured': fp64_probe(device)}

    # ---- configs[4]: population sweep, global N split over the ranks --------
    s = po.headline_fhn_spec()
    ll = pb.GaussianKnownSigmaLogLikelihood(
        pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), s.times, s.values), 0.5)
    ev = pb.CudaEvaluator(ll, devices=device, as_list=False)
    dp, spec = ev.device_problem(), ev.spec()
    sweep = []
    sizes = [1000, 10000, 65536, 100000, 1000000] + ([] if quick else [10000000])
    for n in sizes:
        lo, hi, per = shard_bounds(n, world, rank)
        xs = po.headline_fhn_points(n)[lo:hi]
        d_x = torch.from_numpy(np.ascontiguousarray(xs)).to(device)
        steps = dp.new_steps(max(hi - lo, 1))
        d_s = torch.empty(max(hi - lo, 1), dtype=torch.float64, device=device)
        dp.eval(d_x, out=d_s[:hi - lo], steps=steps[:hi - lo])
        flops = torch.tensor([float(spec.flops_per_eval(
            steps[:hi - lo].cpu().numpy()).sum())], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(flops)
            ex = ScoreExchange(max(per, 1), device)
            fn = lambda: ex.run(dp, d_x, rows_per_rank=per, n_total=n)      # noqa: E731
            how = 'fused exchange' if ex.fused else 'NCCL all-gather'
        else:
            fn = lambda: dp.eval(d_x, out=d_s[:hi - lo])                    # noqa: E731
            how = 'one GPU'
        ms = time_device(fn, reps=5 if n > 1e6 else 20)
        sweep.append({'n_global': n, 'n_per_gpu': per, 'ms': ms,
                      'evals_per_s': n / (ms * 1e-3),
                      'tflops_total': float(flops.item()) / (ms * 1e-3) / 1e12,
                      'frac_of_peak_per_gpu': float(flops.item()) / (ms * 1e-3) / 1e12
                      / (world * out['fp64_peak_tflops_measured']),
                      'exchange': how})
        del d_x, d_s, steps
        if world > 1:
            del ex
    out['fhn_sweep'] = sweep

    # ---- configs[1]: xNES generations, population 65 536 in total -----------
    b = pb.RectangularBoundaries([0.01, 0.1, 1.0], [1.0, 1.0, 5.0])
    gens = 60 if quick else 200
    per_gen = None
    for attempt in range(2):                      # the first run warms up
        ctl = pb.CudaOptimisationController(
            ll, [0.1, 0.5, 3.0], [0.005, 0.025, 0.15], boundaries=b,
            method=pb.DeviceXNES, sharded=world > 1, devices=device)
        ctl.optimiser().set_population_size(65536)
        ctl.set_max_iterations(gens)
        ctl.set_max_unchanged_iterations(None)
        barrier()
        t0 = time.perf_counter()
        x_best, f_best = ctl.run()
        barrier()
        per_gen = (time.perf_counter() - t0) / gens
    out['xnes'] = {'population': 65536, 'generations': gens,
                   'ms_per_generation': max_over_ranks(per_gen * 1e3),
                   'x_best': [float(v) for v in x_best], 'f_best': float(f_best)}

    # ---- configs[2]: HaarioBardenetACMC 16 384 chains on HH-IK --------------
    hh = pb.toy.HodgkinHuxleyIKModel()
    th = hh.suggested_times()
    p = np.array(hh.suggested_parameters(), dtype=float)
    rng = np.random.RandomState(1)
    vh = po.hh_ik_simulate(p, th) + rng.normal(0, 40, th.shape)
    glh = pb.GaussianLogLikelihood(pb.SingleOutputProblem(hh, th, vh))
    n_ch = 16384
    x0 = np.hstack([p * rng.uniform(0.95, 1.05, (n_ch, 5)),
                    40 * rng.uniform(0.95, 1.05, (n_ch, 1))])
    iters = 100 if quick else 400
    for n_it in (60, iters):                      # the first run warms up
        mc = pb.CudaMCMCController(glh, n_ch, x0, method=pb.HaarioBardenetACMC, seed=1,
                                   sharded=world > 1, device=device)
        mc.set_max_iterations(n_it)
        mc.set_initial_phase_iterations(50)
        barrier()
        mc.run(to_host=False)
        per_it = max_over_ranks(mc.time() / n_it)
        del mc
    out['hh_ik_hbac'] = {'chains': n_ch, 'iterations': iters,
                         'ms_per_iteration': per_it * 1e3,
                         'projected_s_for_1e4_iterations': per_it * 1e4,
                         'chain_iterations_per_s': n_ch / per_it}

    # ---- configs[3]: DE-MCMC 4096 chains on LV / SIR, sharded ---------------
    for name, model, tt, pp, sig in (
            ('lv', pb.toy.LotkaVolterraModel(), np.linspace(0, 20, 1000),
             [0.5, 0.15, 1.0, 0.3], 0.1),
            ('sir', pb.toy.SIRModel(), np.linspace(1, 21, 1000), [0.026, 0.285, 38], 1.0)):
        rng = np.random.RandomState(1)
        clean = po.simulate('lotka_volterra' if name == 'lv' else 'sir', None, pp, tt)
        vv = clean + rng.normal(0, sig, clean.shape)
        gl = pb.GaussianLogLikelihood(pb.MultiOutputProblem(model, tt, vv))
        n_ch = 4096
        x0 = np.hstack([np.array(pp) * rng.uniform(0.98, 1.02, (n_ch, len(pp))),
                        sig * rng.uniform(0.95, 1.05, (n_ch, 2))])
        iters = 100 if quick else 300
        for n_it in (30, iters):                  # the first run warms up
            mc = pb.CudaMCMCController(gl, n_ch, x0, method=pb.DifferentialEvolutionMCMC,
                                       seed=2, sharded=world > 1, device=device)
            mc.set_max_iterations(n_it)
            barrier()
            mc.run(to_host=False)
            per_it = max_over_ranks(mc.time() / n_it)
            del mc
        out[name + '_de'] = {'chains': n_ch, 'iterations': iters,
                             'ms_per_iteration': per_it * 1e3,
                             'chain_iterations_per_s': n_ch / per_it}

    if rank == 0:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        path = os.path.join(ROOT, 'gpurun_out', 'configs_multi_%dgpu.json' % world)
        with open(path, 'w') as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
