"""Catalog-partition sharding across the GPUs of one box (one process per GPU).

The hot path shards by independent units: loci.  Each rank takes one contiguous block of the catalog
(balanced by estimated DP area, not locus count), runs it on its own GPU, and the per-read results are
gathered on rank 0 in catalog order -- exactly where the reference heap-merges the results of its worker
processes (strkit/call/call_sample.py:413-420; block building strkit/call/loci.py:193-207).  There is NO
collective on the data path; torch.distributed only carries the final (n_reads x 4 int32) results.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .batcher import ReadBatch

__all__ = ["partition_catalog", "estimated_cost", "count_reads_sharded"]


def estimated_cost(batch: ReadBatch) -> np.ndarray:
    """Per-locus DP area estimate: sum over reads of len(db) * (len(fl) + len(tr) + len(fr))."""
    n1 = batch.lens.sum(axis=1, dtype=np.int64)
    per_read = n1 * n1
    csum = np.concatenate([[0], np.cumsum(per_read)])
    return csum[batch.read_begin[1:]] - csum[batch.read_begin[:-1]]


def partition_catalog(cost: np.ndarray, n_shards: int) -> np.ndarray:
    """Contiguous partition of the catalog into n_shards blocks of roughly equal cost.
    Returns n_shards + 1 locus boundaries (shard r = loci [b[r], b[r+1]))."""
    n = int(cost.shape[0])
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    csum = np.concatenate([[0], np.cumsum(cost.astype(np.float64))])
    targets = csum[-1] * np.arange(1, n_shards) / n_shards
    cuts = np.searchsorted(csum, targets, side="left")
    bounds = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def count_reads_sharded(batch: ReadBatch, compute: Callable[[ReadBatch], np.ndarray], rank: int = 0,
                        world_size: int = 1, group=None) -> np.ndarray | None:
    """Run `compute` (e.g. `lambda b: engine.count_reads(b, rc_params)`) on this rank's catalog partition and
    gather the per-read results on rank 0 (returns None on the other ranks).  Every rank holds the same
    `batch` description (the catalog); only its own partition is touched: the slice handed to `compute` holds the
    partition's bytes only (slice_loci(compact=True)), so a rank uploads 1/world_size of the arena, not all of it."""
    bounds = partition_catalog(estimated_cost(batch), world_size)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = compute(batch.slice_loci(lo, hi, compact=True)) if hi > lo else np.zeros((0, 4), dtype=np.int32)
    if world_size == 1:
        return mine
    import torch
    import torch.distributed as dist

    # results only: 16 bytes per read, gathered on the host side of the pipeline
    counts = [int(batch.read_begin[bounds[r + 1]] - batch.read_begin[bounds[r]]) for r in range(world_size)]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    local = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
    pad = max(counts)
    buf = torch.zeros((pad, 4), dtype=torch.int32, device=dev)
    buf[: local.shape[0]] = local
    gathered = [torch.zeros_like(buf) for _ in range(world_size)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return np.concatenate([g[:c].cpu().numpy() for g, c in zip(gathered, counts)], axis=0)
