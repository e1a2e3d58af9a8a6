"""Voice sharding for multi-GPU renders (one process per GPU).

Voices are independent until the bus fan-in, so a render shards by voice with no data-path exchange except ONE
float32 sum-reduce of the per-rank partial buses to the root, after which the root applies the bus ops (the bus
GainNode) and writes the result.  The reference sums the fan-in sequentially in connection order
(AudioNodeInput.cs:121-132); per-rank partial sums associate differently, a float32 discrepancy of order
sqrt(V) * 2^-24 of the bus level — inside the 1e-5 gate (measured in tests/test_multirank_cpu.py and on 2 GPUs).
"""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) range of `n_items` owned by `rank`; the first (n_items % world) ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_list(items: List, rank: int, world: int) -> List:
    lo, hi = shard_range(len(items), rank, world)
    return items[lo:hi]
