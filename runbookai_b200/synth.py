"""Deterministic synthetic corpora / query batches.

Values are rounded to bf16 so the CPU oracle (as float64) and the GPU index (as bf16) see
identical inputs.  numpy only; no dependency on the oracle or on CUDA.
"""
from __future__ import annotations

import numpy as np


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even float32 -> bf16, returned as uint16 bit patterns."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    rounded = u + (np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1)))
    out = (rounded >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """float array -> same-shape float32 array holding bf16-representable values."""
    return bf16_bits_to_f32(f32_to_bf16_bits(np.asarray(x, dtype=np.float32)))


def seed_for(config: int) -> int:
    return 0x5EED0000 + config


def random_corpus(n: int, d: int, seed: int) -> np.ndarray:
    """i.i.d. N(0,1) rounded to bf16, as uint16 bits [n, d]."""
    rng = np.random.Generator(np.random.Philox(seed))
    out = np.empty((n, d), dtype=np.uint16)
    step = max(1, (1 << 22) // max(d, 1))
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        out[r0:r1] = f32_to_bf16_bits(rng.standard_normal((r1 - r0, d), dtype=np.float32))
    return out


def random_queries(b: int, d: int, seed: int) -> np.ndarray:
    """bf16-exact float32 queries [b, d]."""
    rng = np.random.Generator(np.random.Philox(seed ^ 0xABCDEF))
    return bf16_round(rng.standard_normal((b, d), dtype=np.float32))


def plant_neighbours(corpus_bits: np.ndarray, queries: np.ndarray, per_query: int, seed: int,
                     cos_lo: float = 0.5, cos_hi: float = 0.99) -> np.ndarray:
    """Overwrite `per_query` random rows per query with noisy copies of the query so that
    their cosine lands in [cos_lo, cos_hi] (exercises the `>= minScore` path, S5).
    Returns the planted slots [B, per_query]."""
    rng = np.random.Generator(np.random.Philox(seed ^ 0x51A27))
    n, d = corpus_bits.shape
    b = queries.shape[0]
    slots = rng.choice(n, size=b * per_query, replace=False).reshape(b, per_query)
    for i in range(b):
        q = queries[i].astype(np.float32)
        target = rng.uniform(cos_lo, cos_hi, size=per_query).astype(np.float32)
        # c = q + sigma * noise with ||noise|| ~ ||q||  ->  cos ~ 1/sqrt(1+sigma^2)
        sigma = np.sqrt(1.0 / (target * target) - 1.0)
        noise = rng.standard_normal((per_query, d), dtype=np.float32)
        noise *= np.linalg.norm(q) / np.linalg.norm(noise, axis=1, keepdims=True)
        corpus_bits[slots[i]] = f32_to_bf16_bits(q[None, :] + sigma[:, None] * noise)
    return slots
