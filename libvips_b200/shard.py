"""Batch sharding for independent frames across GPUs.

Each image is an independent unit: rank r takes a contiguous block of the batch,
there is no data-path collective.  torch.distributed is only used for the
barrier, the max-over-ranks clock and a one-time agreement check on a shared
frame (every rank rebuilds the same tables from the same host code, so nothing
needs broadcasting; the check proves it).
"""


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of n_items for this rank; blocks differ by at most one item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_agree(dist, tensor):
    """True on every rank iff `tensor` (integer) is identical on all ranks."""
    lo, hi = tensor.clone(), tensor.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return bool((lo == hi).all())


def max_over_ranks(dist, tensor):
    """The slowest rank's time: every multi-GPU number is reported as the max over ranks."""
    out = tensor.clone()
    dist.all_reduce(out, op=dist.ReduceOp.MAX)
    return out


def aggregate_rate(units_per_rank, world, seconds_max):
    """Whole-job throughput: units all ranks processed / slowest rank's time."""
    return units_per_rank * world / seconds_max
