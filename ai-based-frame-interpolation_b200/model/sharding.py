"""Frame-pair sharding of a clip across GPUs (one process per GPU, no collective on the inference path).

Pair i is (frame i, frame i+1); every pair is independent (reference model/inference.py:101-122 keeps no state), so a
clip of F frames is cut into contiguous ranges of pairs, one per rank; a rank needs frames [first, last] inclusive,
i.e. neighbouring ranks share exactly one boundary frame.
"""
from __future__ import annotations


def shard_pairs(n_frames: int, world_size: int, rank: int):
    """Returns (first_pair, n_pairs) of `rank`. Ranges are contiguous, disjoint, cover [0, n_frames-1) and differ in
    size by at most one pair."""
    if n_frames < 2 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("need at least two frames and 0 <= rank < world_size")
    total = n_frames - 1
    base, extra = divmod(total, world_size)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def frames_needed(n_frames: int, world_size: int, rank: int):
    """Inclusive frame range [lo, hi] that `rank` must read (empty shard -> None)."""
    first, n = shard_pairs(n_frames, world_size, rank)
    return (first, first + n) if n > 0 else None
