"""Data-parallel protocol of the step: the batch is cut into contiguous slices,
one per rank; two small allreduces per step.  Device-agnostic on purpose: `StepEngine` plugs in the
CUDA phases (NCCL), the CPU tests plug in the oracle (gloo)."""
from __future__ import annotations


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced slice [lo, hi) of range(n) for `rank` (first n % world ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def two_phase_step(partial_fn, finish_fn, sums, grad, group=None):
    """partial_fn() fills `sums` (this rank's fp64 partial sums); after the sum-allreduce,
    finish_fn() fills `grad` (this rank's additive gradient share); after the second allreduce every
    rank holds the full gradient.  With group=None and no initialised process group: single rank."""
    import torch.distributed as dist
    partial_fn()
    multi = dist.is_available() and dist.is_initialized() and (group is not None or dist.get_world_size() > 1)
    if multi:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    finish_fn()
    if multi:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
