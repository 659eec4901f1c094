"""FHN kernel time at populations between one and two waves of the 124-register
layout, under the layout switches (PB200_DENSE, PB200_MIGRATE)."""
import os, sys, subprocess
import numpy as np
if len(sys.argv) == 1:
    for env in ({}, {'PB200_DENSE': '0'}, {'PB200_MIGRATE': '0'}):
        print('--- env', env, flush=True)
        subprocess.run([sys.executable, __file__, 'run'], env=dict(os.environ, **env))
    sys.exit(0)
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pints_b200 as pb
from oracle import pints_oracle as po
spec = po.headline_fhn_spec()
f = pb.GaussianKnownSigmaLogLikelihood(
    pb.MultiOutputProblem(pb.toy.FitzhughNagumoModel(), spec.times, spec.values), 0.5)
ev = pb.CudaEvaluator(f, as_list=False)
dp, sp = ev.device_problem(), ev.spec()
for n in (65536, 75776, 80000, 85000, 90000, 94720, 97000, 100000, 104192, 110000, 113664, 120000, 131072, 151552):
    d_x = torch.from_numpy(po.headline_fhn_points(n)).cuda()
    out = torch.empty(n, dtype=torch.float64, device='cuda')
    steps = dp.new_steps(n)
    dp.eval(d_x, out=out, steps=steps)
    flops = float(sp.flops_per_eval(steps.cpu().numpy()).sum())
    for _ in range(3):
        dp.eval(d_x, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dp.eval(d_x, out=out); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print('N=%7d  warps/SM %.2f  %.4f ms  %.3g evals/s  %.2f TFLOP/s' % (n, n / 32 / 148, ms, n / ms * 1e3, flops / ms / 1e9))
