"""Multi-GPU partitioning helpers (one process per GPU, ``torch.distributed``; NCCL on the B200 box, gloo in the
CPU tests).

Two partitionings exist on this path:

* independent problems (Lasso columns, batches of LPs): strided split ``j = rank (mod world)`` -- the
  reference's own ``num_chunks`` semantics (LassoSolver.py:349-351) -- with NO data-path collective;
* one large problem: contiguous blocks of constraint rows stay resident on their GPU; each rank forms the
  partial Hessian / gradient / barrier sums of its rows and a SUM all-reduce (MAX for the step index) makes
  them global before the replicated factorisation.
"""

import numpy as np
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def strided_columns(K, rank, world_size):
    """Columns of a batch owned by ``rank``: identical to chunk ``rank`` of ``num_chunks = world_size``."""
    return np.arange(K)[rank::world_size]


def row_range(m, rank, world_size):
    """Contiguous block of constraint rows owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(m, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allreduce_max_(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def allreduce_min_(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def gather_columns(local, cols, K, group=None):
    """Assemble the full ``[n, K]`` result on every rank from per-rank column blocks (host tensors)."""
    rank, ws = world()
    if ws == 1:
        return local
    parts = [None] * ws
    dist.all_gather_object(parts, (np.asarray(cols), np.asarray(local)), group=group)
    out = np.zeros((local.shape[0], K))
    for c, blk in parts:
        out[:, c] = blk
    return out
