"""Spectrum sharding across the GPUs of one box: spectra are independent (runner.rs:311-325 is a pure
par_iter().flat_map), so rank g of G scores the contiguous block [g*n/G, (g+1)*n/G) against its own replica of the index and
rank 0 gathers the Feature rows. There is NO collective on the data path; the gather moves results only."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    return (rank * n) // world, ((rank + 1) * n) // world


def score_sharded(score_fn, batch, report_psms: int, rank: int, world: int, dist=None, dst: int = 0):
    """score_fn(SpectraBatch) -> (features[n_local*report_psms], counts[n_local]). Returns (features, counts) for the WHOLE batch on
    rank `dst` (spectrum indices rebased to the full batch, order identical to a single-process run), None elsewhere."""
    a, b = shard_range(len(batch), world, rank)
    feats, counts = score_fn(batch.slice(a, b))
    feats = feats.copy()
    feats["spectrum"] += np.uint32(a)
    if world == 1 or dist is None:
        return feats, counts
    gathered = [None] * world if rank == dst else None
    dist.gather_object((a, feats, counts), gathered, dst=dst)
    if rank != dst:
        return None
    gathered.sort(key=lambda t: t[0])
    return np.concatenate([g[1] for g in gathered]), np.concatenate([g[2] for g in gathered])
