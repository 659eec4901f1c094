#!/usr/bin/env python3
"""
Headline benchmark: FitzHugh-Nagumo log-likelihood evaluations per second
(1000 time points, GaussianKnownSigmaLogLikelihood, population 65536 per GPU --
BASELINE.json configs[1]) on N B200s, with the FP64 roofline of the fused
kernel and the reference's CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N \
        --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one population: score `--pop`
parameter vectors per GPU (inputs resident in HBM), then all-gather the scores
across ranks (the only exchange step of the path).  Prints ONE JSON line.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'fhn_loglikelihood_evals_per_sec'
UNIT = 'evals/s'
WORKLOAD = ('FitzhughNagumoModel (3 params, 1000 times on [0,20]) + '
            'GaussianKnownSigmaLogLikelihood(sigma=0.5), population 65536 '
            'per GPU (XNES-sized), synthetic noisy series seed 1, points '
            '[0.1,0.5,3]*(1+0.05*randn) seed 2')


# dram__bytes_read.sum + dram__bytes_write.sum of one ode_kernel launch at this
# workload, from the `ncu --set full` capture summarised in
# profiles/r01h_ode_kernel_summary.md (algorithmic HBM traffic: 32 B per
# evaluation = 2.1 MB per launch; the path is FP64-bound, not HBM-bound)
TRAFFIC_BYTES_PER_LAUNCH = 1918208


def hbm_side(kernel_ms):
    """The same launch against the HBM roofline (why the bound is not 'hbm'):
    measured DRAM bytes of the launch / its duration, against the driver's
    measured copy bandwidth (MEASURED_PEAKS.json) or the recipe's fallback."""
    peak, source = 6650.0, 'fallback of B200_PROFILING.md'
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peak, source = float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:
        pass
    achieved = TRAFFIC_BYTES_PER_LAUNCH / (kernel_ms * 1e-3) / 1e9
    return {'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'peak_source': source}


# ------------------------------------------------------------ CPU baseline

_worker_spec = None


def _worker_init():
    global _worker_spec
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    from oracle import pints_oracle as po
    _worker_spec = po.headline_fhn_spec()


def _worker_score(chunk):
    return [_worker_spec.score(x) for x in chunk]


class CpuBaseline(object):
    """The reference's CPU path for this workload: the oracle port (scipy LSODA
    with the reference's right-hand side, bit-identical to
    SequentialEvaluator(f).evaluate) fanned out over all host cores with worker
    processes, as ParallelEvaluator does (pints/_evaluation.py:126-413)."""

    def __init__(self, cores=None):
        self.cores = cores or len(os.sched_getaffinity(0))
        ctx = mp.get_context('spawn')
        self.pool = ctx.Pool(self.cores, initializer=_worker_init)
        from oracle import pints_oracle as po
        self.po = po
        # warm-up: spawn + import cost is excluded, as the reference keeps its
        # workers alive between calls
        self.run(po.headline_fhn_points(self.cores * 4, seed=7))

    def run(self, xs):
        per = max(1, min(32, len(xs) // (self.cores * 4) or 1))
        chunks = [xs[i:i + per] for i in range(0, len(xs), per)]
        t0 = time.perf_counter()
        out = self.pool.map(_worker_score, chunks)
        dt = time.perf_counter() - t0
        return dt, np.array([v for part in out for v in part])

    def sample_size(self, seconds):
        dt, _ = self.run(self.po.headline_fhn_points(self.cores * 16, seed=8))
        rate = self.cores * 16 / dt
        return max(self.cores * 8, int(rate * seconds))

    def close(self):
        self.pool.terminate()
        self.pool.join()


def run_reference_arm(args):
    """`--impl reference`: the CPU path alone, same metric / unit / config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base = CpuBaseline()
    m = base.sample_size(6.0)
    xs = base.po.headline_fhn_points(m)
    for _ in range(args.warmup):
        base.run(xs[:base.cores * 8])
    total = 0.0
    for _ in range(args.steps):
        dt, _ = base.run(xs)
        total += dt
    base.close()
    value = m * args.steps / total
    sample = '%d evaluations per step (~6 s of %d-core CPU work)' % (m, base.cores)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'population_per_step': m,
                   'solver': 'scipy odeint (LSODA) rtol=atol=1.49012e-8'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': base.cores,
                         'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ clocks

class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region, through
    NVML in-process (a `nvidia-smi` call takes longer than the whole region);
    falls back to nvidia-smi if NVML cannot be loaded."""
    REASONS = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4,
               'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
               'hw_power_brake_slowdown': 0x80}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index = index
        self.period = period
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.source = 'nvml'
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(
                self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            self.source = 'nvidia-smi'
            self.period = 0.1

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h,
                                                           n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        out = subprocess.run(
            ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
             '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
            stderr=subprocess.DEVNULL, text=True,
            timeout=5).stdout.strip().split(',')
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for name, flag in zip(names, out[2:]):
            if flag.strip().lower().startswith('active'):
                self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=6)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples),
                'source': self.source}


# ----------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun for --gpus %d' % args.gpus)

    # the CPU baseline runs first (rank 0, N=1 only), before CUDA is touched
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        base = CpuBaseline()
        m = base.sample_size(args.cpu_seconds)
        dt, ref_scores = base.run(base.po.headline_fhn_points(m))
        base.close()
        # the same on ONE core, in this process (= SequentialEvaluator)
        spec1 = base.po.headline_fhn_spec()
        xs1 = base.po.headline_fhn_points(256, seed=9)
        t0 = time.perf_counter()
        k1 = 0
        while time.perf_counter() - t0 < 2.0 and k1 < len(xs1):
            spec1.score(xs1[k1])
            k1 += 1
        one_core = k1 / (time.perf_counter() - t0)
        cpu = {'value': m / dt, 'unit': UNIT, 'cores': base.cores,
               'value_1core': one_core,
               'kind': 'port',
               'sample': '%d evaluations of the same workload (%.1f s), oracle '
                         'port = scipy LSODA + reference RHS, %d worker '
                         'processes' % (m, dt, base.cores)}
        cpu_check = (m, ref_scores)

    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    # stdout carries exactly one JSON line: keep NCCL's version banner off it
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    import pints_b200 as pb
    from pints_b200._engine import fp64_probe
    from oracle import pints_oracle as po      # inputs only (synthetic series)

    spec_o = po.headline_fhn_spec()
    model = pb.toy.FitzhughNagumoModel()
    problem = pb.MultiOutputProblem(model, spec_o.times, spec_o.values)
    ll = pb.GaussianKnownSigmaLogLikelihood(problem, 0.5)
    ev = pb.CudaEvaluator(ll, devices=device, as_list=False)
    dp = ev.device_problem()
    spec = ev.spec()

    pop = args.pop
    # every rank scores its own shard of the global population (weak scaling)
    xs_all = po.headline_fhn_points(pop * world)
    xs = np.ascontiguousarray(xs_all[rank * pop:(rank + 1) * pop])
    d_x = torch.from_numpy(xs).to(device)
    d_scores = torch.empty(pop, dtype=torch.float64, device=device)
    exchange = None
    if world > 1:
        from pints_b200._dist import ScoreExchange
        exchange = ScoreExchange(pop, device, fused=not args.nccl_allgather)
    d_steps = dp.new_steps(pop)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def step(steps_buf=None):
        # N = 1: sort + fused solve.  N > 1: the same, and the exchange step
        # (every rank ends up with all scores): fused into the kernel over
        # NVLink peer memory + a device barrier, or solve + NCCL all-gather
        if world > 1:
            exchange.run(dp, d_x, steps=steps_buf)
        else:
            dp.eval(d_x, out=d_scores, steps=steps_buf)

    # parity spot-check against the CPU oracle numbers produced above
    dp.eval(d_x, out=d_scores, steps=d_steps)
    torch.cuda.synchronize(device)
    flops_per_launch = float(spec.flops_per_eval(d_steps.cpu().numpy()).sum())
    mean_steps = float(d_steps.double().mean().item())
    if cpu is not None:
        m, ref_scores = cpu_check
        k = min(m, pop, 512)
        got = d_scores[:k].cpu().numpy()
        rel = np.max(np.abs(got - ref_scores[:k]) / np.abs(ref_scores[:k]))
        cpu['parity_max_rel'] = float(rel)
        # LSODA's own error on this distribution reaches 9.6e-7;
        # the strict 1e-6 bound lives in tests/, this is a sanity stop
        assert rel < 5e-6, 'GPU/CPU parity %g' % rel

    peak_tflops = fp64_probe(device)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()

    # ---- device-timed: K steps, L2 flushed between steps, events per step --
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    k0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xff)
        e0[i].record()
        step()
        k0[i].record()
        e1[i].record()
    barrier()
    wall = time.perf_counter() - wall0
    step_ms = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(e0, k0))

    # ---- end to end through the evaluator: host arrays in, host scores out -
    # (a) positions in page-locked host memory (what the contract times):
    #     H2D from where they lie + sort + fused solve + D2H + copy-out
    xs_pinned = pb.pinned_empty(xs.shape)
    np.copyto(xs_pinned, xs)
    for _ in range(3):
        ev.evaluate(xs_pinned)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = ev.evaluate(xs_pinned)
    barrier()
    e2e_s = time.perf_counter() - t0
    # (b) an ordinary (pageable) ndarray, as numpy's ask() returns it: staged
    #     through pinned memory in slices
    for _ in range(2):
        ev.evaluate(xs)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = ev.evaluate(xs)
    barrier()
    e2e_pageable_s = time.perf_counter() - t0
    clocks = sampler.finish()
    # (c) the same with the reference's return type (a Python list of floats):
    # the boxing of 65 536 floats alone costs about as much as the solve
    ev_list = pb.CudaEvaluator(ll, devices=device)
    ev_list.evaluate(xs)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = ev_list.evaluate(xs)
    barrier()
    e2e_list_s = time.perf_counter() - t0

    if world > 1:
        t = torch.tensor([step_ms, kern_ms, e2e_s, e2e_list_s, e2e_pageable_s],
                         dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, kern_ms, e2e_s, e2e_list_s, e2e_pageable_s = t.tolist()

    if rank == 0:
        total = pop * world * args.steps
        value = total / (step_ms * 1e-3)
        achieved = flops_per_launch / (kern_ms / args.steps * 1e-3) / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': step_ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'population_per_gpu': pop,
                       'global_population': pop * world,
                       'solver': 'DP5(4) rtol=atol=1.49012e-8',
                       'mean_rk_steps': mean_steps,
                       'l2': 'flushed between steps (256 MiB fill)',
                       'parallelism': 'dp%d' % world,
                       'exchange': 'none (one GPU)' if world == 1 else (
                           'fused: kernel stores scores into every rank\'s '
                           'array over NVLink (symmetric memory) + barrier'
                           if exchange.fused else 'NCCL all-gather')},
            'roofline': {
                'bound': 'fp64', 'achieved': achieved, 'peak': peak_tflops,
                'unit': 'TFLOP/s', 'frac': achieved / peak_tflops,
                'traffic': TRAFFIC_BYTES_PER_LAUNCH,
                'peak_source': 'measured live by pb200_fp64_probe (FMA chains);'
                               ' MEASURED_PEAKS.json has no FP64 entry; '
                               'nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2',
                'flop_per_launch': flops_per_launch,
                'kernel_ms': kern_ms / args.steps,
                'hbm': hbm_side(kern_ms / args.steps)},
            'e2e': {'value': total / e2e_s, 'unit': UNIT,
                    'h2d_bytes_per_step': int(xs.nbytes) * world,
                    'd2h_bytes_per_step': int(pop * 8) * world,
                    'api': 'CudaEvaluator(f, as_list=False).evaluate(ndarray) '
                           '-> ndarray; positions in page-locked host memory '
                           '(pints_b200.pinned_empty), scores back in host memory',
                    'value_pageable_input': total / e2e_pageable_s,
                    'api_pageable_input': 'same call with an ordinary ndarray: '
                                          'staged through pinned memory in '
                                          'slices of 32768 rows (multi-threaded copy)',
                    'value_as_list': total / e2e_list_s,
                    'api_as_list': 'CudaEvaluator(f).evaluate(ndarray) -> list '
                                   '(the reference\'s return type)'},
            # per step and rank: the counting-sort kernel + the fused ODE kernel
            'gpu_launches': args.steps * world * 2,
            'clocks': clocks,
            'wall_s_timed_region': wall,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pop', type=int, default=65536,
                    help='parameter vectors per GPU per step')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--nccl-allgather', action='store_true',
                    help='N > 1: exchange the scores with an NCCL all-gather '
                         'after the kernel instead of the fused peer stores')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
