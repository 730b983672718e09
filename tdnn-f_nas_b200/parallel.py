"""Data parallelism over the sequences of a minibatch.

Every op of the hot path is independent across sequences (the `n` index) except the sums that form
the parameter / architecture-weight deltas, so: rank g takes sequences [g*S/G, (g+1)*S/G) on the same
t grid, runs its slice, and the per-rank deltas are combined with ONE all-reduce(sum) -- the
synchronous replacement of Kaldi's multi-job `nnet3-average` (common.py:144-164; LR x num_jobs followed
by averaging == summing the per-job deltas).  Noise (Gumbel, one-hot) is drawn per minibatch, not per
row, so all ranks share one RNG seed and counter.  Works with any torch.distributed backend (NCCL over
NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_sequences(num_seqs: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the sequences owned by `rank`; sizes differ by at most one."""
    assert 0 <= rank < world
    base, rem = divmod(num_seqs, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rows(num_frames: int, num_seqs: int, rank: int, world: int) -> List[int]:
    """Row numbers (t-major, n fastest: row = t*num_seqs + n) of a rank's sequences in the global matrix."""
    b, e = shard_sequences(num_seqs, rank, world)
    return [t * num_seqs + n for t in range(num_frames) for n in range(b, e)]


def allreduce_deltas(views: Sequence, group=None, async_op: bool = False):
    """Sum the delta buffers (theta of every updatable component, bias tails, alpha vectors) over ranks."""
    import torch.distributed as dist

    works = [dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group, async_op=True) for v in views]
    if async_op:
        return works
    for w in works:
        w.wait()
    return None


def flops_penalty_normaliser(local_rows: int, world: int) -> int:
    """The FLOPs penalty is scale/(R*C) with R the LOCAL row count in the reference (simple.cc:10154);
    summed over `world` ranks it must be divided by the GLOBAL row count instead."""
    return local_rows * world
