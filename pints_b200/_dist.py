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
