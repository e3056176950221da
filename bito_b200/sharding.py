"""Site-pattern sharding: rank g of G owns the contiguous pattern range
[g*P/G, (g+1)*P/G) of every PLV, log-likelihood row and weight; per-edge and per-PLV scalars are
replicated and made global by all-reduces inside the engine."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(pattern_count: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of rank's shard; shards differ by at most one pattern and cover [0, P) exactly."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside [0, world_size)")
    base, extra = divmod(int(pattern_count), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
