"""Batch sharding for the multi-GPU path.

Every polynomial / commitment of a batch is independent, so rank r of G owns
the contiguous slice [start, stop) and there is no data-path collective; the
only exchange is the final gather of the outputs.  Commitment randomness is
keyed by the GLOBAL index (seed = base + global index), so the result is the
same bits for every G.
"""
from __future__ import annotations

import numpy as np


def shard_range(count: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice of `count` units for `rank` of `world`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, extra = divmod(count, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def global_seeds(base: int, start: int, stop: int) -> np.ndarray:
    """Seeds of the units [start, stop): base + global index (mod 2^64), never 0."""
    idx = np.arange(start, stop, dtype=np.uint64)
    s = idx + np.uint64(base % (1 << 64))
    s[s == 0] = np.uint64(1 << 63)
    return s


def gather_slices(parts: list[np.ndarray]) -> np.ndarray:
    """Concatenate per-rank outputs in rank order (what the final all-gather yields)."""
    return np.concatenate(parts, axis=0)
