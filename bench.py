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
import gc
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
# profiles/r02_ode_kernel_summary.md (algorithmic HBM traffic: 32 B per
# evaluation = 2.1 MB per launch; the path is FP64-bound, not HBM-bound)
TRAFFIC_BYTES_PER_LAUNCH = 1923328


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


def _worker_truth(chunk):
    # the same solver driven to convergence: the "truth" both arms are set against
    return [_worker_spec.score(x, rtol=1e-13, atol=1e-13) for x in chunk]


def load_reference():
    """The unmodified reference package, installed from /root/reference into
    the git-ignored baseline/_ref (see __graft_entry__.build), or None."""
    path = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(path, 'pints')):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        import pints
        import pints.toy  # noqa: F401
    except Exception:
        return None
    return pints


class PortBaseline(object):
    """The oracle port of the reference's CPU path (scipy LSODA with the
    reference's right-hand side, bit-identical to
    SequentialEvaluator(f).evaluate) fanned out over all host cores with worker
    processes.  Used for the truth sample, as the second CPU figure, and as the
    reference arm where baseline/_ref is absent."""
    kind = 'port'
    what = 'oracle port = scipy LSODA + reference RHS, multiprocessing.Pool'

    def __init__(self, cores=None):
        self.cores = cores or len(os.sched_getaffinity(0))
        ctx = mp.get_context('spawn')
        self.pool = ctx.Pool(self.cores, initializer=_worker_init)
        from oracle import pints_oracle as po
        self.po = po
        # warm-up: spawn + import cost is excluded, as the reference keeps its
        # workers alive between calls
        self.run(po.headline_fhn_points(self.cores * 4, seed=7))

    def _map(self, fn, xs):
        per = max(1, min(32, len(xs) // (self.cores * 4) or 1))
        chunks = [xs[i:i + per] for i in range(0, len(xs), per)]
        t0 = time.perf_counter()
        out = self.pool.map(fn, chunks)
        dt = time.perf_counter() - t0
        return dt, np.array([v for part in out for v in part])

    def run(self, xs):
        return self._map(_worker_score, xs)

    def truth(self, xs):
        return self._map(_worker_truth, xs)[1]

    def sample_size(self, seconds):
        dt, _ = self.run(self.po.headline_fhn_points(self.cores * 16, seed=8))
        rate = self.cores * 16 / dt
        return max(self.cores * 8, int(rate * seconds))

    def close(self):
        self.pool.terminate()
        self.pool.join()


class ReferenceBaseline(object):
    """The reference's own CPU implementation of the path, unmodified:
    pints.ParallelEvaluator over the reference's FitzhughNagumoModel /
    MultiOutputProblem / GaussianKnownSigmaLogLikelihood objects
    (pints/_evaluation.py:126-413), one worker per host core."""
    kind = 'reference'
    what = 'pints.ParallelEvaluator (unmodified reference from baseline/_ref)'

    def __init__(self, pints, cores=None):
        from oracle import pints_oracle as po
        self.po = po
        self.cores = cores or len(os.sched_getaffinity(0))
        spec = po.headline_fhn_spec()          # the synthetic series only
        model = pints.toy.FitzhughNagumoModel()
        problem = pints.MultiOutputProblem(model, spec.times, spec.values)
        self.f = pints.GaussianKnownSigmaLogLikelihood(problem, 0.5)
        self.pints = pints
        self.ev = pints.ParallelEvaluator(self.f, n_workers=self.cores)
        # warm-up: worker start-up is excluded, the reference keeps its
        # workers alive between calls (pints/_evaluation.py:153-154)
        self.run(po.headline_fhn_points(self.cores * 8, seed=7))

    def run(self, xs):
        t0 = time.perf_counter()
        out = self.ev.evaluate(xs)
        dt = time.perf_counter() - t0
        return dt, np.array(out)

    def one_core(self, seconds=2.0):
        """SequentialEvaluator on one core, in this process."""
        ev = self.pints.SequentialEvaluator(self.f)
        xs = self.po.headline_fhn_points(64, seed=9)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < seconds:
            ev.evaluate(xs[:16])
            n += 16
        return n / (time.perf_counter() - t0)

    def sample_size(self, seconds):
        dt, _ = self.run(self.po.headline_fhn_points(self.cores * 16, seed=8))
        rate = self.cores * 16 / dt
        return max(self.cores * 8, int(rate * seconds))

    def close(self):
        try:
            self.ev._stop()
        except Exception:
            pass


def cpu_arm():
    pints = load_reference()
    return ReferenceBaseline(pints) if pints is not None else PortBaseline()


def run_reference_arm(args):
    """`--impl reference`: the CPU path alone, same metric / unit / config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    base = cpu_arm()
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
    sample = '%d evaluations per step (~6 s of %d-core CPU work); %s' % (
        m, base.cores, base.what)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'population_per_step': m,
                   'solver': 'scipy odeint (LSODA) rtol=atol=1.49012e-8'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': base.cores,
                         'kind': base.kind, 'sample': sample},
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

TRUTH_SAMPLE = 4096        # vectors solved to convergence on the CPU (N = 1)
TRUTH_SAMPLE_MULTI = 1024  # the same at N > 1 (rank 0 only, the others wait)
HEAD_START = 8             # untimed steps queued right before the timed ones


def cpu_side(args, world):
    """Everything the host cores do, before CUDA is touched (rank 0 only):
    the reference arm's throughput on a bounded sample, the oracle port beside
    it, and a truth sample (the same LSODA driven to rtol = atol = 1e-13) that
    the GPU scores are checked against afterwards."""
    from oracle import pints_oracle as po
    port = PortBaseline()
    n_truth = TRUTH_SAMPLE if world == 1 else TRUTH_SAMPLE_MULTI
    pick = np.linspace(0, args.pop - 1, n_truth).astype(np.int64)
    xs_all = po.headline_fhn_points(args.pop * world)
    truth = port.truth(xs_all[pick])
    out = {'truth_rows': pick, 'truth': truth, 'cpu': None}
    if world == 1 and not args.no_cpu:
        m = port.sample_size(min(4.0, args.cpu_seconds))
        dt_port, port_scores = port.run(xs_all[:m])
        port_value = m / dt_port
        port.close()
        pints = load_reference()
        if pints is not None:
            base = ReferenceBaseline(pints)
            m = base.sample_size(args.cpu_seconds)
            dt, ref_scores = base.run(xs_all[:m])
            one_core = base.one_core()
            base.close()
            kind, what, value = base.kind, base.what, m / dt
            assert np.array_equal(ref_scores[:len(port_scores)],
                                  port_scores[:len(ref_scores)]), \
                'oracle port and reference disagree'
        else:
            ref_scores, dt = port_scores, dt_port
            kind, what, value, one_core = port.kind, port.what, port_value, None
        out['cpu'] = {
            'value': value, 'unit': UNIT, 'cores': port.cores, 'kind': kind,
            'value_1core': one_core, 'value_port': port_value,
            'sample': '%d evaluations of the same workload (%.1f s); %s; %d '
                      'workers' % (m, dt, what, port.cores)}
        out['ref_rows'] = np.arange(len(ref_scores))
        out['ref'] = ref_scores
    else:
        port.close()
    return out


def rel_err(a, b):
    return np.abs(a - b) / np.abs(b)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torchrun for --gpus %d' % args.gpus)

    host = cpu_side(args, world) if rank == 0 else None

    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    # stdout carries exactly one JSON line: keep NCCL's version banner off it
    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    load_reference()           # mirror classes then derive from the reference's
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

    pop, K = args.pop, args.steps
    # every rank scores its own shard of the global population (weak scaling)
    xs_all = po.headline_fhn_points(pop * world)
    xs = np.ascontiguousarray(xs_all[rank * pop:(rank + 1) * pop])
    d_x = torch.from_numpy(xs).to(device)
    d_scores = torch.empty(pop, dtype=torch.float64, device=device)
    d_steps = dp.new_steps(pop)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    exchanges = {}
    mode = 'none'
    if world > 1:
        from pints_b200._dist import ScoreExchange
        mode = args.exchange
        exchanges['nccl'] = ScoreExchange(pop, device, fused=False)
        exchanges['fused'] = ScoreExchange(pop, device, multicast=False)
        mc = ScoreExchange(pop, device, multicast=True)
        if mc.fused and mc.multicast:
            exchanges['multicast'] = mc
        if not exchanges['fused'].fused:
            del exchanges['fused']
        if mode not in exchanges:
            mode = 'fused' if 'fused' in exchanges else 'nccl'

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)


    def step(which, mid=None):
        # N = 1: sort + fused solve.  N > 1: the same, and the exchange step
        # (every rank ends up with all scores): stored by the kernel into every
        # rank's staging array over NVLink + reading-side kernel, or solve +
        # NCCL all-gather
        if which == 'local':
            dp.eval(d_x, out=d_scores)
            return d_scores
        return exchanges[which].run(dp, d_x, mid_event=mid)

    def timed(which):
        """K steps, L2 flushed between steps, events per step; returns per-step
        ms of the whole step, of its first part (sort + evaluation kernel) and
        of the rest (reading side / all-gather)."""
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        k0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
        for _ in range(max(args.warmup, 3)):
            step(which)
        barrier()
        # the host enqueues ahead of the device, as a controller loop that
        # does not synchronise every generation does: HEAD_START untimed steps
        # run straight into the timed ones (no synchronisation in between) and
        # give it a head start of about 3 ms, so that a host hiccup on one rank
        # (GC, scheduler) is not charged to all ranks' exchange; garbage
        # collection is off inside the region
        gc.collect()
        gc.disable()
        for i in range(HEAD_START):
            flush.fill_(i & 0xff)
            step(which)
        for i in range(K):
            flush.fill_(i & 0xff)
            e0[i].record()
            if which == 'local':
                step(which)
                k0[i].record()
            else:
                step(which, mid=k0[i])
            e1[i].record()
        barrier()
        gc.enable()
        whole = np.array([a.elapsed_time(b) for a, b in zip(e0, e1)])
        first = np.array([a.elapsed_time(b) for a, b in zip(e0, k0)])
        return whole, first, whole - first

    # ---- parity: whole population against the CPU numbers produced above ---
    dp.eval(d_x, out=d_scores, steps=d_steps)
    torch.cuda.synchronize(device)
    flops_per_launch = float(spec.flops_per_eval(d_steps.cpu().numpy()).sum())
    mean_steps = float(d_steps.double().mean().item())
    parity = {}
    if rank == 0:
        got = d_scores.cpu().numpy()
        assert np.all(np.isfinite(got)), 'non-finite score in the population'
        e_truth = rel_err(got[host['truth_rows']], host['truth'])
        parity['truth_sample'] = int(len(e_truth))
        parity['max_rel_vs_truth'] = float(e_truth.max())
        parity['truth'] = 'the same LSODA at rtol = atol = 1e-13'
        if host.get('ref') is not None:
            k = min(len(host['ref']), pop)
            e_ref = rel_err(got[:k], host['ref'][:k])
            ref_truth = None
            parity['reference_sample'] = int(k)
            parity['max_rel_vs_reference'] = float(e_ref.max())
            parity['frac_within_1e-6_of_reference'] = float(np.mean(e_ref <= 1e-6))
            both = np.intersect1d(host['truth_rows'], np.arange(k))
            if len(both):
                at = np.searchsorted(host['truth_rows'], both)
                ref_truth = rel_err(host['ref'][both], host['truth'][at])
                parity['reference_max_rel_vs_truth'] = float(ref_truth.max())
        # north_star: ODE log-likelihoods within 1e-6 relative
        assert parity['max_rel_vs_truth'] < 1e-6, \
            'GPU vs truth parity %g' % parity['max_rel_vs_truth']
    if world > 1:
        # the gathered array of every exchange equals the single-rank scores
        want = torch.empty(pop * world, dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(want, d_scores)
        for name, ex in exchanges.items():
            got_all = ex.run(dp, d_x)
            torch.cuda.synchronize(device)
            same = bool(torch.equal(got_all.view(torch.int64),
                                    want.view(torch.int64)))
            flag = torch.tensor([1 if same else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            parity['gathered_equals_local_bitwise_' + name] = bool(flag.item())
            assert flag.item() == 1, 'exchange %s changed the scores' % name

    peak_tflops = fp64_probe(device)

    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    wall0 = time.perf_counter()
    main_which = 'local' if world == 1 else mode
    whole, first, rest = timed(main_which)
    wall = time.perf_counter() - wall0
    compare = {}
    if world > 1:
        for name in ['local'] + [m for m in exchanges if m != mode]:
            compare[name] = timed(name)

    # ---- end to end through the evaluator: host arrays in, host scores out -
    def e2e_loop(evaluator, arg, warm=3):
        for _ in range(warm):
            evaluator.evaluate(arg)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            evaluator.evaluate(arg)
        barrier()
        return time.perf_counter() - t0

    e2e = {}
    if world == 1:
        # (a) positions in page-locked host memory: H2D from where they lie +
        #     sort + fused solve + D2H into the (page-locked) result array
        xs_pinned = pb.pinned_empty(xs.shape)
        np.copyto(xs_pinned, xs)
        e2e_s = e2e_loop(ev, xs_pinned)
        # (b) an ordinary (pageable) ndarray, as numpy's ask() returns it:
        #     staged through pinned memory in slices by the library's threads
        e2e_pageable_s = e2e_loop(ev, xs)
        # (c) the drop-in's default call: pageable ndarray in, ScoreList out
        ev_default = pb.CudaEvaluator(ll, devices=device)
        e2e_default_s = e2e_loop(ev_default, xs)
        # (d) the reference's exact return type, a list of Python floats
        ev_boxed = pb.CudaEvaluator(ll, devices=device, as_list='boxed')
        e2e_boxed_s = e2e_loop(ev_boxed, xs)
        total = pop * K
        e2e = {'value': total / e2e_s, 'unit': UNIT,
               'h2d_bytes_per_step': int(xs.nbytes),
               'd2h_bytes_per_step': int(pop * 8),
               'api': 'CudaEvaluator(f, as_list=False).evaluate(ndarray) -> '
                      'ndarray; positions in page-locked host memory '
                      '(pints_b200.pinned_empty), scores back in host memory',
               'value_pageable_input': total / e2e_pageable_s,
               'api_pageable_input': 'same call with an ordinary ndarray: '
                                     'staged through pinned memory in slices '
                                     'of 32768 rows (pb200_host_copy threads)',
               'value_as_list': total / e2e_default_s,
               'api_as_list': 'CudaEvaluator(f).evaluate(ndarray) -> ScoreList'
                              ' (the default: list-compatible view, boxed '
                              'lazily); pageable input',
               'value_boxed_list': total / e2e_boxed_s,
               'api_boxed_list': "CudaEvaluator(f, as_list='boxed') -> list of "
                                 'Python floats (the reference\'s exact type)'}
    else:
        # the multi-GPU API: every rank passes the WHOLE population (as an SPMD
        # controller loop does), scores its block, and gets all scores back in
        # host memory -- H2D of the block, kernel + exchange, D2H of all scores
        sharded = pb.ShardedEvaluator(ll, device=device, as_list=False)
        xs_all_pinned = pb.pinned_empty(xs_all.shape)
        np.copyto(xs_all_pinned, xs_all)
        e2e_s = e2e_loop(sharded, xs_all_pinned)
        got_all = sharded.evaluate(xs_all_pinned)
        if rank == 0:
            e_truth = rel_err(got_all[host['truth_rows']], host['truth'])
            parity['e2e_max_rel_vs_truth'] = float(e_truth.max())
        xs_pinned = pb.pinned_empty(xs.shape)
        np.copyto(xs_pinned, xs)
        e2e_private_s = e2e_loop(ev, xs_pinned)
        total = pop * world * K
        hg = getattr(sharded, '_host_gather', None)
        on_host = hg is not None and hg.ok
        e2e = {'h2d_bytes_per_step': int(xs.nbytes) * world,
               # gathered on the host: every score crosses PCIe once; gathered
               # by the GPUs: every rank reads all scores back
               'd2h_bytes_per_step': int(pop * 8) * world * (1 if on_host else world),
               'exchange': ('host: every rank copies its block of scores into a shared, '
                            'page-locked host segment, epoch flags, copy-out (HostGather%s)'
                            % ('' if hg.pinned else ', segment NOT registered with CUDA')
                            if on_host else
                            'device: scores stored into every rank\'s staging array by the '
                            'kernel over NVLink, every rank reads all of them back'),
               'api': 'ShardedEvaluator(f, as_list=False).evaluate(all '
                      'positions, page-locked) on every rank -> all scores in '
                      'every rank\'s host memory (exchange included)',
               'api_no_exchange': 'each rank its own CudaEvaluator on its own '
                                  'block, nothing exchanged (N independent '
                                  'single-GPU calls)'}
    clocks = sampler.finish()

    # ---- gather the per-rank numbers on rank 0 ------------------------------
    def stats(x):
        return {'median': float(np.median(x)), 'p95': float(np.percentile(x, 95)),
                'mean': float(np.mean(x)), 'max': float(np.max(x))}

    # which cost sort the timed steps ran with (the hint is set by the warm-up
    # steps and the population does not change)
    sort_in_kernel = dp.sort_hint() == 1 and os.environ.get('PB200_LOCAL_SORT', '') != '0'
    mine = {'rank': rank, 'step_ms_each': [round(float(v), 4) for v in whole],
            'step_ms': stats(whole), 'kernel_ms': stats(first),
            'exchange_wait_ms': stats(rest), 'sum_step_ms': float(whole.sum()),
            'sum_kernel_ms': float(first.sum())}
    for name, (w, f, r) in compare.items():
        mine['compare_' + name] = {'sum_step_ms': float(w.sum()),
                                   'step_ms_median': float(np.median(w)),
                                   'kernel_ms_median': float(np.median(f))}
    mine['e2e_s'] = e2e_s
    if world > 1:
        mine['e2e_private_s'] = e2e_private_s
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    else:
        everyone = [mine]

    if rank == 0:
        step_ms = max(r['sum_step_ms'] for r in everyone)       # max over ranks
        kern_ms = max(r['sum_kernel_ms'] for r in everyone)
        total = pop * world * K
        value = total / (step_ms * 1e-3)
        achieved = flops_per_launch / (kern_ms / K * 1e-3) / 1e12
        if world > 1:
            e2e['value'] = total / max(r['e2e_s'] for r in everyone)
            e2e['unit'] = UNIT
            e2e['value_no_exchange'] = total / max(r['e2e_private_s'] for r in everyone)
        exchange_text = {
            'none': 'none (one GPU)',
            'fused': 'fused: the kernel stores {score, row} pairs into every '
                     'rank\'s staging array over NVLink (512-byte runs per '
                     'warp), its last block raises flags, a reading-side '
                     'kernel waits for them and files the scores; no barrier '
                     'launch, no NCCL call',
            'multicast': 'fused, one multimem.st per vector through the '
                         'NVSwitch multicast mapping instead of one store per '
                         'rank',
            'nccl': 'solve + NCCL all-gather'}[mode]
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
            'steps': K, 'warmup': max(args.warmup, 3),
            'ms_per_step': step_ms / K, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'population_per_gpu': pop,
                       'global_population': pop * world,
                       'solver': 'DP5(4) rtol=atol=1.49012e-8',
                       'mean_rk_steps': mean_steps,
                       'l2': 'flushed between steps (256 MiB fill)',
                       'queueing': 'host enqueues ahead: %d untimed steps run '
                                   'straight into the timed ones, GC off in '
                                   'the region; per-step CUDA events' % HEAD_START,
                       'parallelism': 'dp%d' % world,
                       'exchange': exchange_text},
            'roofline': {
                'bound': 'fp64', 'achieved': achieved, 'peak': peak_tflops,
                'unit': 'TFLOP/s', 'frac': achieved / peak_tflops,
                'traffic': TRAFFIC_BYTES_PER_LAUNCH,
                'peak_source': 'measured live by pb200_fp64_probe (FMA chains);'
                               ' MEASURED_PEAKS.json has no FP64 entry; '
                               'nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2',
                'flop_per_launch': flops_per_launch,
                'kernel_ms': kern_ms / K,
                'hbm': hbm_side(kern_ms / K)},
            'e2e': e2e,
            'parity': parity,
            # per step and rank: the fused ODE kernel (+ the reading side of
            # the exchange), and the counting-sort kernel in front of it unless
            # the previous launch found the population's costs close together (sort inside the kernel)
            'gpu_launches': K * world * ((1 if world == 1 else 2) + (0 if sort_in_kernel else 1)),
            'cost_sort': ('inside the evaluation kernel, block by block (costs of the population close together)'
                          if sort_in_kernel else 'sort kernel in front of the evaluation kernel'),
            'clocks': clocks,
            'wall_s_timed_region': wall,
            # where the time of a step goes, per rank: kernel_ms = sort +
            # evaluation kernel (with its peer stores and flags when N > 1),
            # exchange_wait_ms = reading side (waits for the slowest rank)
            'per_rank': [{k: v for k, v in r.items()
                          if not k.startswith('compare_') and not k.startswith('e2e')}
                         for r in everyone],
        }
        if world > 1:
            cmp_out = {}
            for name in compare:
                sums = [r['compare_' + name]['sum_step_ms'] for r in everyone]
                cmp_out[name] = {
                    'ms_per_step': max(sums) / K,
                    'step_ms_median_per_rank': [r['compare_' + name]['step_ms_median'] for r in everyone],
                    'kernel_ms_median_per_rank': [r['compare_' + name]['kernel_ms_median'] for r in everyone]}
            cmp_out['_what'] = ('same invocation, same inputs: "local" = sort + '
                                'evaluation kernel alone (no exchange); the '
                                'others = the whole step with that exchange')
            line['exchange_compare'] = cmp_out
        if host['cpu'] is not None:
            line['cpu_baseline'] = host['cpu']
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
    ap.add_argument('--exchange', default=os.environ.get('PB200_BENCH_EXCHANGE', 'fused'),
                    choices=['fused', 'multicast', 'nccl'],
                    help='N > 1: how the scores are exchanged in the headline '
                         'number (the other ways are timed beside it)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
