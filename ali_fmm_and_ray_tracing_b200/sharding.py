"""Source sharding across ranks / devices.

Every source's travel-time field (and the rays traced through it) is independent -- this is
how the reference parallelises too (queue of source indices, ATR:3979-3984, 4641-4643) -- so
ranks take contiguous blocks of sources and there is no collective on the compute path."""


def shard_bounds(n_items, world_size, rank):
    """Half-open range [lo, hi) of the items rank ``rank`` owns: contiguous, balanced to within one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    k, r = divmod(int(n_items), int(world_size))
    lo = rank * k + min(rank, r)
    return lo, lo + k + (1 if rank < r else 0)


def shard_indices(n_items, world_size, rank):
    lo, hi = shard_bounds(n_items, world_size, rank)
    return list(range(lo, hi))
