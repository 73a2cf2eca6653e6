"""Synthetic corpus definition shared by the CUDA generator (csrc/synth.cuh) and the tests.

Counter-based: byte b of 8-byte chunk c of global row r is byte b (little endian) of
splitmix64(splitmix64(seed ^ r*0xD1342543DE82EF95) + c).  Uniform bytes, reproducible per row,
so a shard can generate its rows on the device and a checker can regenerate any row on the host.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_rows(seed: int, first_row: int, nrows: int, dim: int) -> np.ndarray:
    """Rows [first_row, first_row+nrows) of the synthetic corpus as uint8 [nrows, dim]."""
    chunks = (dim + 7) // 8
    rows = np.arange(first_row, first_row + nrows, dtype=np.uint64)
    with np.errstate(over="ignore"):
        rk = _splitmix64(np.uint64(seed) ^ (rows * np.uint64(0xD1342543DE82EF95)))
        x = _splitmix64(rk[:, None] + np.arange(chunks, dtype=np.uint64)[None, :])
    b = x.astype("<u8").view(np.uint8).reshape(nrows, chunks * 8)
    return np.ascontiguousarray(b[:, :dim])


def synth_queries(seed: int, nq: int, dim: int, corpus_rows: int, corpus_seed: int) -> np.ndarray:
    """Bench/test queries: the first half are perturbed corpus rows (so the head of each result
    list is non-trivial), the second half are independent uniform vectors."""
    rng = np.random.default_rng(seed)
    out = np.zeros((nq, dim), np.uint8)
    half = nq // 2
    for i in range(half):
        r = int(rng.integers(0, corpus_rows))
        row = synth_rows(corpus_seed, r, 1, dim)[0].astype(np.int16)
        noise = rng.integers(-24, 25, size=dim)
        out[i] = np.clip(row + noise, 0, 255).astype(np.uint8)
    out[half:] = rng.integers(0, 256, size=(nq - half, dim), dtype=np.uint8)
    return out
