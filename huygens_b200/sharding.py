"""Filter partitioning across the GPUs of one box (SURVEY.md §8 e).

rank k owns the contiguous filters [lo, hi) of the bank; every rank renders the same input to a
partial bank sum and one reduce(SUM) of the output block completes it.  Streams shard without any
communication.  The reduce itself is torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""


def filter_range(n_total, rank, world):
    """Contiguous, balanced split: the first (n_total % world) ranks get one extra filter."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stream_range(n_streams, rank, world):
    return filter_range(n_streams, rank, world)


def reduce_partial_sum(out, dst=0):
    """In-place SUM-reduce of a rank's partial bank sum onto rank `dst` (no-op without a group)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(out, dst=dst, op=dist.ReduceOp.SUM)
    return out
