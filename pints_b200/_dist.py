"""
Multi-GPU evaluation, one process per GPU (``torch.distributed``).

The path shards by parameter vector: evaluations are independent, so the rows
of ``positions`` are split into contiguous blocks, one per rank, with no
data-path collective.  The only exchange step is a small all-gather of the
scores (N float64 in total) so that every rank -- and the host-side optimiser
update, which is O(population) and stays on the CPU -- sees all of them.  This
replaces the task/result queues of ``ParallelEvaluator``
(pints/_evaluation.py:284-372, :489-565).
"""
import contextlib
import ctypes
import os
import socket
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._evaluation import Evaluator
from ._engine import exchange_finish, positions_to_array


def shard_bounds(n, world, rank):
    """Contiguous block partition: rank r owns rows [lo, hi)."""
    per = (n + world - 1) // world if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


class ScoreExchange(object):
    """The exchange step of the sharded path: after it, every rank holds the
    scores of all ranks, in row order.

    With NCCL on NVLink-connected GPUs the exchange is fused into the
    evaluation kernel (``pb200_eval_exchange`` / ``pb200_exchange_finish``,
    see ``include/pints_b200.h``): every rank owns a staging array of
    ``{score, row}`` pairs and a flag array in torch symmetric memory, mapped
    into every process; the kernel stores each score into all staging arrays
    as the vector finishes (contiguous 512-byte runs per warp and destination,
    or one multicast store through NVSwitch), its last block raises the
    rank's flags, and a small kernel on the reading side waits for the flags
    and files the scores under their rows.  No barrier launch, no NCCL call.
    Two staging arrays alternate, so a rank may start the next generation
    while a slower peer still reads the previous one.  Anywhere else (gloo on
    CPU, no peer access) it is a plain ``all_gather_into_tensor``; the choice
    is made collectively, so all ranks take the same path.
    """

    def __init__(self, per_rank, device, group=None, fused=None,
                 multicast=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.per = int(per_rank)
        self.device = torch.device(device)
        self.fused = False
        self.fused_error = None          # why the fused path was not taken
        self.multicast = False
        self._epoch = 0
        if fused is None:
            fused = self.device.type == 'cuda' and self.world > 1 and \
                self.world <= _lib.MAX_PEERS and \
                dist.get_backend(group) == 'nccl'
        if fused:
            self._setup_fused(multicast)
        if not self.fused:
            self._mine = torch.zeros(self.per, dtype=torch.float64,
                                     device=self.device)
            self._every = torch.empty(self.per * self.world,
                                      dtype=torch.float64, device=self.device)

    def _setup_fused(self, multicast):
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm_mem
            g = self.group if self.group is not None else dist.group.WORLD
            pairs = self.per * self.world
            flag_words = -(-self.world // 16) * 16          # whole 128-byte lines
            # one allocation, in doubles: [stage 0][stage 1][flags]
            self._symm = symm_mem.empty(2 * 2 * pairs + flag_words,
                                        dtype=torch.float64,
                                        device=self.device)
            self._symm.zero_()
            self._hdl = symm_mem.rendezvous(self._symm, g)
        except Exception as e:             # no peer mapping on this rank
            ok = 0
            self.fused_error = e
        # every rank must take the same path: one that fell back to the
        # all-gather would wait for peers that spin on its flags
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:
            if self.fused_error is None:
                self.fused_error = RuntimeError(
                    'symmetric memory is not available on a peer rank')
            self._symm = self._hdl = None
            return
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)     # flags are zero everywhere
        pairs = self.per * self.world
        ptrs = [int(q) for q in self._hdl.buffer_ptrs]
        self._stage_ptrs = [[q + k * pairs * 16 for q in ptrs] for k in (0, 1)]
        self._flag_ptrs = [q + 2 * pairs * 16 for q in ptrs]
        mc = 0
        if multicast is None:
            multicast = os.environ.get('PB200_EXCHANGE_MULTICAST', '0') == '1'
        if multicast:
            try:
                mc = int(self._hdl.multicast_ptr or 0)
            except Exception:
                mc = 0
        self._mc_ptrs = [mc + k * pairs * 16 if mc else 0 for k in (0, 1)]
        self.multicast = bool(mc)
        self._every = torch.empty(pairs, dtype=torch.float64,
                                  device=self.device)
        self.fused = True

    def _desc(self, k):
        d = _lib.ExchangeDesc()
        d.n_ranks, d.rank, d.per_rank = self.world, self.rank, self.per
        for q in range(self.world):
            d.stage[q] = self._stage_ptrs[k][q]
            d.flags[q] = self._flag_ptrs[q]
        d.stage_multicast = self._mc_ptrs[k] or None
        d.epoch = self._epoch
        return d

    def run(self, problem, d_params, steps=None, rows_per_rank=None,
            n_total=None, mid_event=None):
        """Scores ``d_params`` (this rank's block, at most ``per_rank`` rows) on
        ``problem`` and returns the scores of all ranks in row order (device
        tensor).  ``rows_per_rank`` (default ``per_rank``) is the block length
        of this generation and ``n_total`` (default ``rows_per_rank * world``)
        the number of rows of the whole population: rank r holds the rows
        ``[r * rows_per_rank, min(n_total, (r + 1) * rows_per_rank))`` --
        ``n_total`` must be given whenever a block is shorter than the others.
        ``mid_event`` (a CUDA event, for timing) is recorded between the
        evaluation kernel and the reading side."""
        n = d_params.shape[0]
        rows = self.per if rows_per_rank is None else int(rows_per_rank)
        total = rows * self.world if n_total is None else int(n_total)
        if rows > self.per or n > rows:
            raise ValueError('block of %d rows does not fit the exchange (%d)'
                             % (max(n, rows), self.per))
        if self.fused:
            self._epoch += 1
            desc = self._desc(self._epoch & 1)
            problem.eval_exchange(d_params, desc, steps=steps)
            if mid_event is not None:
                mid_event.record()
            exchange_finish(desc, rows, total, self._every)
            return self._every[:total]
        if n:
            problem.eval(d_params, out=self._mine[:n], steps=steps)
        if mid_event is not None:
            mid_event.record()
        if rows == self.per:
            dist.all_gather_into_tensor(self._every, self._mine,
                                        group=self.group)
            return self._every[:total]
        every = self._every[:rows * self.world]
        dist.all_gather_into_tensor(every, self._mine[:rows], group=self.group)
        return every[:total]


def _boot_id():
    try:
        with open('/proc/sys/kernel/random/boot_id') as f:
            return f.read().strip()
    except OSError:
        return ''


class HostGather(object):
    """All-gather of the ranks' score blocks on the HOST, for ranks of one
    node: one POSIX shared-memory segment mapped by every process (and
    registered with CUDA as page-locked, so a rank's device-to-host copy lands
    in it directly), ``SLOTS`` result areas used in turn and one epoch flag per
    rank.  A call: every rank copies ITS block of scores into the current
    area, raises its flag, waits for the flags of the others and copies the
    whole area out.

    Why: ``ShardedEvaluator.evaluate`` hands ALL scores to EVERY rank's host.
    Reading them back from each GPU moves world x n x 8 bytes over PCIe in
    lock step (8 ranks x 4 MB: +0.33 ms per call on an 8-GPU box); here every
    score crosses PCIe once and the rest is a host memory copy.  As measured
    so far that copy and the flag wait cost more than they save (see
    ``ShardedEvaluator._host_gather_possible``), so this path is opt-in.

    The segment is unlinked as soon as all ranks have mapped it (nothing is
    left behind if a process dies).  Flags and data are ordinary stores and
    loads: the copy into the area is complete (stream synchronised) before a
    rank's flag is raised, and x86 keeps stores, and loads, in order.  A rank
    can be at most one call ahead of the slowest one (it needs everybody's
    flag to return), so two areas would do.  ``ok`` is False (and ``error``
    says why) when the ranks are not on one node or the segment cannot be
    set up; the decision is collective.
    """

    SLOTS = 3
    HEADER = 4096

    def __init__(self, capacity, group=None, device=None, timeout=120.0):
        from multiprocessing import resource_tracker, shared_memory
        self.group = group
        self.capacity = int(capacity)           # doubles per result area
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.timeout = float(timeout)
        self.epoch = 0
        self.ok = False
        self.error = None
        self.pinned = False
        self._shm = None
        if self.world * 64 > self.HEADER:
            self.error = 'too many ranks'
            return
        cuda = device is not None and torch.device(device).type == 'cuda'
        ctx = torch.cuda.device(device) if cuda else contextlib.nullcontext()
        size = self.HEADER + self.SLOTS * self.capacity * 8
        shm = None
        with ctx:
            where = [None] * self.world
            dist.all_gather_object(where, (socket.gethostname(), _boot_id()), group=group)
            name = None
            if len(set(where)) == 1 and self.rank == 0:
                try:
                    shm = shared_memory.SharedMemory(create=True, size=size)
                    name = shm.name
                except Exception as e:                  # no /dev/shm, too small, ...
                    self.error = e
            box = [name]
            src = dist.get_global_rank(group, 0) if group is not None else 0
            dist.broadcast_object_list(box, src=src, group=group)
            name = box[0]
            mapped = False
            if name is not None:
                try:
                    if self.rank != 0:
                        # this process did not create the segment and must not
                        # unlink it at exit: attach without telling the resource
                        # tracker (Python < 3.13 has no track=False)
                        register = resource_tracker.register
                        resource_tracker.register = lambda *a, **k: None
                        try:
                            shm = shared_memory.SharedMemory(name=name)
                        finally:
                            resource_tracker.register = register
                    mapped = shm.size >= size
                except Exception as e:
                    self.error = e
            everyone = [None] * self.world
            dist.all_gather_object(everyone, bool(mapped), group=group)
            if self.rank == 0 and shm is not None:
                try:
                    shm.unlink()
                except Exception:
                    pass
        if not all(everyone):
            if shm is not None:
                shm.close()
            if self.error is None:
                self.error = 'the ranks are not on one node' if name is None \
                    else 'a rank could not map the segment'
            return
        self._shm = shm
        self.flags = np.ndarray((self.world, 8), dtype=np.int64, buffer=shm.buf)
        self.data = np.ndarray((self.SLOTS, self.capacity), dtype=np.float64,
                               buffer=shm.buf, offset=self.HEADER)
        self.base = self.data.ctypes.data
        if cuda:
            try:
                with torch.cuda.device(device):
                    rc = torch.cuda.cudart().cudaHostRegister(
                        self.base - self.HEADER, size, 0)
                self.pinned = int(rc) == 0
            except Exception as e:      # still correct: the copies are staged by the driver
                self.error = e
        self.ok = True

    def area(self, epoch):
        """Address of the result area of call ``epoch``."""
        return self.base + (epoch % self.SLOTS) * self.capacity * 8

    def next_epoch(self):
        self.epoch += 1
        return self.epoch

    def publish(self, epoch):
        self.flags[self.rank, 0] = epoch

    def wait(self, epoch):
        """Until every rank has published ``epoch``; raises after ``timeout``
        seconds instead of hanging."""
        col = self.flags[:, 0]
        spins, t0 = 0, None
        while not (col >= epoch).all():
            spins += 1
            if (spins & 1023) == 0:
                now = time.monotonic()
                if t0 is None:
                    t0 = now
                elif now - t0 > self.timeout:
                    raise RuntimeError(
                        'HostGather: ranks %s have not delivered call %d after %g s'
                        % (np.nonzero(col < epoch)[0].tolist(), epoch, self.timeout))

    def close(self):
        shm, self._shm = self._shm, None
        if shm is None:
            return
        if self.pinned:
            try:
                torch.cuda.cudart().cudaHostUnregister(self.base - self.HEADER)
            except Exception:
                pass
        self.flags = self.data = None
        try:
            shm.close()
        except Exception:
            pass

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedEvaluator(Evaluator):
    """Every rank calls ``evaluate(positions)`` with the same positions (as an
    SPMD controller loop does); each scores its own block on its own GPU and
    all ranks return the full list of scores.

    Parameters
    ----------
    function
        as for ``CudaEvaluator``
    group
        process group (default: the world group); ``nccl`` on GPUs
    device
        this rank's CUDA device (default: ``cuda:LOCAL_RANK``)
    """

    def __init__(self, function, args=None, group=None, device=None,
                 as_list=True, **options):
        super().__init__(function, args)
        if len(self._args) != 0:
            raise ValueError(
                'ShardedEvaluator does not support extra function arguments.')
        self._group = group
        self._as_list = as_list if as_list == 'boxed' else bool(as_list)
        self._device = device
        self._options = options
        self._local = None
        self._buffers = {}
        self._exchange = None
        self._host_gather = None

    # -- overridable pieces (the gloo tests replace the launch) ----------
    def _make_local(self):
        from ._evaluation import CudaEvaluator
        import os
        device = self._device
        if device is None:
            device = 'cuda:%d' % int(os.environ.get('LOCAL_RANK', '0'))
        return CudaEvaluator(self._function, devices=device, as_list=False,
                             **self._options)

    def _width(self):
        if self._local is None:
            self._local = self._make_local()
        return self._local.n_parameters()

    def _local_scores(self, xs, out):
        """Scores of the local rows into ``out`` (a tensor on the collective's
        device)."""
        problem = self._local.device_problem()
        d_x = torch.from_numpy(xs).to(problem.device, non_blocking=True)
        problem.eval(d_x, out=out[:xs.shape[0]])

    def _buffer(self, name, size, like=None):
        buf = self._buffers.get(name)
        if buf is None or buf.numel() < size:
            device = like if like is not None else self._collective_device()
            buf = torch.empty(max(size, 1), dtype=torch.float64, device=device)
            self._buffers[name] = buf
        return buf

    def _fused_possible(self, device):
        return torch.device(device).type == 'cuda' and \
            dist.get_backend(self._group) == 'nccl' and \
            type(self)._local_scores is ShardedEvaluator._local_scores

    def _collective_device(self):
        if self._local is None:
            self._local = self._make_local()
        return self._local.device_problem().device

    # --------------------------------------------------------------------
    def evaluate_array(self, positions):
        world = dist.get_world_size(self._group)
        rank = dist.get_rank(self._group)
        xs = positions_to_array(positions, self._width())
        n = xs.shape[0]
        if n == 0:
            return np.empty(0)
        lo, hi, per = shard_bounds(n, world, rank)
        device = self._collective_device()
        if self._host_gather_possible(device):
            # ranks of one node: every score crosses PCIe once, into a host
            # segment all ranks map (HostGather)
            hg = self._host_gather
            if hg is None or (hg.ok and hg.capacity < per * world):
                if hg is not None:
                    hg.close()
                hg = self._host_gather = HostGather(
                    max(per * world, 4096), self._group, device)
            if hg.ok:
                return self._evaluate_host_gather(hg, xs, lo, hi, per, n)
        if self._fused_possible(device):
            # fused exchange: the kernel stores the scores into every rank's
            # staging array (symmetric memory over NVLink)
            ex = self._exchange
            if ex is None or ex.per < per:
                self._exchange = None                 # frees the old buffers
                ex = self._exchange = ScoreExchange(
                    max(per, 1024), device, self._group)
            if ex.fused:
                return self._evaluate_fused(ex, xs, lo, hi, per, n)
        mine = self._buffer('mine', per, device)[:per]
        every = self._buffer('all', per * world, device)[:per * world]
        if hi - lo < per:
            mine.zero_()                       # padding rows of the last ranks
        if hi > lo:
            self._local_scores(np.ascontiguousarray(xs[lo:hi]), mine)
        dist.all_gather_into_tensor(every, mine, group=self._group)
        return every[:n].cpu().numpy() if per * world >= n else None

    def _host_gather_possible(self, device):
        # Opt-in (PB200_HOST_GATHER=1).  Measured with bench.py's end-to-end
        # loop (profiles/r02d_*): 2 GPUs 2.70e8 against 2.64e8 evaluations/s
        # for the device-side exchange, 4 GPUs 3.76e8 against 4.38e8, 8 GPUs
        # 4.45e8 against 6.58e8 -- the copy-out of the whole area by every
        # rank and the flag wait in Python cost more than the PCIe traffic
        # they save on those boxes (32 host cores for 8 ranks).
        return torch.device(device).type == 'cuda' and \
            os.environ.get('PB200_HOST_GATHER', '0') == '1' and \
            type(self)._local_scores is ShardedEvaluator._local_scores

    def _evaluate_host_gather(self, hg, xs, lo, hi, per, n):
        """H2D of this rank's rows, ONE fused launch, D2H of its block of
        scores into the shared result area; then the flags, and the whole area
        is copied out into a fresh array (the library's copy threads)."""
        lane = self._local._setup()[0]
        lib = self._local._lib
        epoch = hg.next_epoch()
        area = hg.area(epoch)
        if hi > lo:
            lane.launch(lib, xs[lo:hi], area + lo * 8, None, 1 << 62)
            lane.wait()
        hg.publish(epoch)
        hg.wait(epoch)
        # (a block of torch's caching host allocator: mapped and touched already,
        # where a fresh 4 MB ndarray would fault its pages in during the copy)
        from ._evaluation import pinned_result
        scores = pinned_result(n)
        _lib.check(lib.pb200_host_copy(ctypes.c_void_p(scores.ctypes.data),
                                       ctypes.c_void_p(area), n * 8, 0))
        return scores

    def _evaluate_fused(self, ex, xs, lo, hi, per, n):
        """H2D of this rank's rows, evaluation kernel with the exchange fused
        in, reading side, D2H of all ``n`` scores into a page-locked array that
        is handed to the caller -- all on one stream of the local evaluator."""
        from ._evaluation import _Lane, pinned_result
        lane = self._local._setup()[0]
        lib = self._local._lib
        stream = lane.streams[0]
        m, d = hi - lo, xs.shape[1]
        lane.reserve(max(m, 1), d, False)
        scores = pinned_result(n)
        with torch.cuda.device(lane.device), torch.cuda.stream(stream):
            if m:
                mine = xs[lo:hi]
                lane.send(lib, mine, 0, m, lane.stream_ptrs[0],
                          _Lane.is_pinned(lib, mine))
            every = ex.run(lane.problem, lane.d_in[:m], rows_per_rank=per,
                           n_total=n)
            _lib.check(lib.pb200_copy_async(
                ctypes.c_void_p(scores.ctypes.data),
                ctypes.c_void_p(every.data_ptr()), n * 8, 0,
                lane.stream_ptrs[0]))
        stream.synchronize()
        return scores

    def _evaluate(self, positions):
        from ._evaluation import wrap_scores
        return wrap_scores(self.evaluate_array(positions), self._as_list)
