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


def message_planes(p: int, modulus: int) -> int:
    """Digit planes needed to bind a whole field element: smallest L with p^L >= modulus (0 if more than 4).
    Mirrors lsr_lwe_message_planes / lsr_prover_quotient_planes."""
    d = 1
    for planes in range(1, 5):
        d *= p
        if d >= modulus:
            return planes
    return 0


def message_digits(msgs: np.ndarray, p: int, planes: int) -> np.ndarray:
    """Base-p digits of message rows: [rows][len] -> [rows * planes][len], row-major in (row, plane) -- the unit
    order of lsr_prover_commit_quotient and lsr_lwe_commit_digits_batch_device (unit = row * planes + plane,
    digit = (word // p^plane) % p).  A commitment binds its words modulo p; the planes together bind the word."""
    m = np.ascontiguousarray(msgs, dtype=np.uint64)
    m = m.reshape(-1, m.shape[-1])
    out = np.empty((m.shape[0], planes, m.shape[1]), dtype=np.uint64)
    cur = m.copy()
    for plane in range(planes):
        out[:, plane, :] = cur % np.uint64(p)
        cur //= np.uint64(p)
    return out.reshape(m.shape[0] * planes, m.shape[1])
