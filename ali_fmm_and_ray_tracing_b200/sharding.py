"""Source sharding across ranks / devices.

Every source's travel-time field (and the rays traced through it) is independent -- this is
how the reference parallelises too (queue of source indices, ATR:3979-3984, 4641-4643) -- so
ranks and devices take contiguous blocks of sources and there is no collective on the compute
path.  Used by the ``*_parallel`` methods of ``ALI_FMM`` (devices of one process) and by
``bench.py`` (one process per GPU)."""
import numpy as np


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


def split_list(items, parts):
    """``items`` cut into ``parts`` contiguous balanced lists (shard_bounds for every part)."""
    items = list(items)
    return [items[slice(*shard_bounds(len(items), parts, p))] for p in range(parts)]


def receivers_of(trans_pairs):
    """Transducers that need a travel-time field: column j of ``trans_pairs`` has a pair
    (reference: ATR:4334, 4641-4643)."""
    trans_pairs = np.asarray(trans_pairs)
    return [j for j in range(trans_pairs.shape[1]) if np.sum(trans_pairs[:, j]) > 0]


def rank_pairs(trans_pairs, world_size, rank):
    """The pair matrix restricted to the receivers (columns) rank ``rank`` owns: the rank computes
    those receivers' fields and every ray into them, nothing else.  Summing the matrices of all
    ranks gives ``trans_pairs`` back."""
    trans_pairs = np.asarray(trans_pairs)
    rec = receivers_of(trans_pairs)
    mine = [rec[k] for k in shard_indices(len(rec), world_size, rank)]
    out = np.zeros_like(trans_pairs)
    out[:, mine] = trans_pairs[:, mine]
    return out
