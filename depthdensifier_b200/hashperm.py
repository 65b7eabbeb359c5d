"""The deterministic subsample permutation that replaces ``torch.randperm`` (depth_refiner.py:304).

The reference draws a generator- and device-specific random 500-subset.  The CUDA path instead ranks
the indices 0..n-1 by ``(mix32(i ^ seed), i)`` (murmur3 finaliser) and takes the first ``max_pairs``;
this module is the host statement of that rule so tests can feed the SAME permutation to the
reference (``torch.randperm`` monkey-patched) and compare bit for bit."""

from __future__ import annotations

import numpy as np


def mix32(h: np.ndarray) -> np.ndarray:
    h = h.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return h


def hash_perm(n: int, seed: int = 0) -> np.ndarray:
    """Permutation of 0..n-1 used by align_stats_kernel (csrc/align.cu)."""
    i = np.arange(n, dtype=np.uint32)
    with np.errstate(over="ignore"):
        key = mix32(i ^ np.uint32(seed & 0xFFFFFFFF))
    return np.lexsort((i, key)).astype(np.int64)
