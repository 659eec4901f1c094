#!/usr/bin/env python3
"""
BASELINE.json configs[3]: LotkaVolterra / SIR + GaussianLogLikelihood,
DifferentialEvolutionMCMC with 4096 chains sharded over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P tools/bench_de_sharded.py [--iters K]

Each rank runs 4096 / N chains; the positions (4096 x d doubles) are
all-gathered once per iteration.  Reporting tool (not the headline bench).
"""
import argparse
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
from oracle import pints_oracle as po                      # noqa: E402  (synthetic series)


def problems():
    rng = np.random.RandomState(1)
    t = np.linspace(0, 20, 1000)
    x_lv = [0.5, 0.15, 1.0, 0.3]
    v = po.lv_simulate(x_lv, t) + rng.normal(0, 0.1, (1000, 2))
    lv = pb.GaussianLogLikelihood(pb.MultiOutputProblem(pb.toy.LotkaVolterraModel(), t, v))
    t2 = np.linspace(1, 21, 1000)
    x_sir = [0.026, 0.285, 38]
    v2 = po.sir_simulate(x_sir, t2) + rng.normal(0, 1, (1000, 2))
    sir = pb.GaussianLogLikelihood(pb.MultiOutputProblem(pb.toy.SIRModel(), t2, v2))
    return {'lotka_volterra': (lv, x_lv + [0.1, 0.1]), 'sir': (sir, x_sir + [1.0, 1.0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=2000)
    ap.add_argument('--chains', type=int, default=4096)
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    out = {}
    for name, (ll, truth) in problems().items():
        rng = np.random.RandomState(5)
        x0 = np.array(truth) * (1 + 0.02 * rng.randn(args.chains, len(truth)))
        for warm in (True, False):
            mc = pb.CudaMCMCController(ll, args.chains, x0,
                                       method=pb.DifferentialEvolutionMCMC, seed=9,
                                       sharded=world > 1, device='cuda:%d' % local)
            mc.set_max_iterations(50 if warm else args.iters)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            chains = mc.run(to_host=False)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        tail = chains[:, args.iters // 2:].reshape(-1, len(truth)).mean(0).cpu().numpy()
        out[name] = {'n_gpus': world, 'chains': args.chains, 'iterations': args.iters,
                     'seconds': dt, 'ms_per_iteration': 1e3 * dt / args.iters,
                     'chain_iterations_per_s': args.chains * args.iters / dt,
                     'posterior_mean': [float(v) for v in tail]}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
