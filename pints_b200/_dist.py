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
import numpy as np
import torch
import torch.distributed as dist

from ._evaluation import Evaluator
from ._engine import positions_to_array


def shard_bounds(n, world, rank):
    """Contiguous block partition: rank r owns rows [lo, hi)."""
    per = (n + world - 1) // world if n > 0 else 0
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi, per


class ScoreExchange(object):
    """The exchange step of the sharded path: after it, every rank holds the
    scores of all ranks.

    With NCCL on NVLink-connected GPUs the gathered array lives in torch
    symmetric memory, every rank's copy is mapped into every process, and the
    evaluation kernel itself stores each score into all copies
    (``DeviceProblem.eval_scatter`` -> ``pb200_eval_scatter``): the all-gather
    is fused into the solve and only a device-side barrier remains.  Two
    arrays alternate, so a rank may start the next generation while a slower
    peer still reads the previous one.  Anywhere else (gloo on CPU, no peer
    access) it is a plain ``all_gather_into_tensor``.
    """

    def __init__(self, per_rank, device, group=None, fused=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.per = int(per_rank)
        self.device = torch.device(device)
        self.fused = False
        self._turn = 0
        if fused is None:
            fused = self.device.type == 'cuda' and self.world > 1 and \
                self.world <= 16 and dist.get_backend(group) == 'nccl'
        if fused:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                g = group if group is not None else dist.group.WORLD
                self._arrays, self._handles = [], []
                for _ in range(2):
                    t = symm_mem.empty(self.per * self.world,
                                       dtype=torch.float64, device=self.device)
                    self._handles.append(symm_mem.rendezvous(t, g))
                    self._arrays.append(t)
                self.fused = True
            except Exception:                              # no peer mapping
                self.fused = False
        if not self.fused:
            self._mine = torch.zeros(self.per, dtype=torch.float64,
                                     device=self.device)
            self._every = torch.empty(self.per * self.world,
                                      dtype=torch.float64, device=self.device)

    def run(self, problem, d_params, steps=None):
        """Scores ``d_params`` (this rank's block, at most ``per_rank`` rows) on
        ``problem`` and returns the gathered array (device tensor,
        ``per_rank * world`` entries, rank r's block at ``r * per_rank``)."""
        n = d_params.shape[0]
        if self.fused:
            k = self._turn
            self._turn ^= 1
            hdl = self._handles[k]
            if n:
                problem.eval_scatter(d_params, list(hdl.buffer_ptrs),
                                     self.rank * self.per, steps=steps)
            hdl.barrier(channel=0)
            return self._arrays[k]
        if n:
            problem.eval(d_params, out=self._mine[:n], steps=steps)
        dist.all_gather_into_tensor(self._every, self._mine, group=self.group)
        return self._every


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
        self._as_list = bool(as_list)
        self._device = device
        self._options = options
        self._local = None
        self._buffers = {}
        self._exchange = None

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
        if self._fused_possible(device):
            # fused exchange: the kernel stores the scores into every rank's
            # gathered array (symmetric memory over NVLink)
            ex = self._exchange
            if ex is None or ex.per < per:
                ex = self._exchange = ScoreExchange(
                    max(per, 1024), device, self._group)
            if ex.fused:
                problem = self._local.device_problem()
                d_x = torch.from_numpy(
                    np.ascontiguousarray(xs[lo:hi])).to(device,
                                                        non_blocking=True)
                every = ex.run(problem, d_x)
                parts = [every[r * ex.per: r * ex.per + min(per, max(0, n - r * per))]
                         for r in range(world)]
                return torch.cat(parts).cpu().numpy()
        mine = self._buffer('mine', per, device)[:per]
        every = self._buffer('all', per * world, device)[:per * world]
        if hi - lo < per:
            mine.zero_()                       # padding rows of the last ranks
        if hi > lo:
            self._local_scores(np.ascontiguousarray(xs[lo:hi]), mine)
        dist.all_gather_into_tensor(every, mine, group=self._group)
        return every[:n].cpu().numpy() if per * world >= n else None

    def _evaluate(self, positions):
        scores = self.evaluate_array(positions)
        return scores.tolist() if self._as_list else scores
