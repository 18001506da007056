"""Frame-level data parallelism across the GPUs of one box.

Scans are independent (the reference class holds no per-scan state,
RP/include/recursive_patchwork.hpp:70), so scan f of a stream goes to rank f mod world or to a
contiguous block; there is NO collective on the data path.  The only exchange is the host-side
gather of label buffers, done here with torch.distributed (gloo on CPU, nccl on GPUs) purely as
plumbing."""
from __future__ import annotations

import numpy as np


def shard_range(n_scans: int, rank: int, world: int) -> range:
    """Contiguous block of scans owned by `rank` (blocks differ by at most one scan)."""
    base, rem = divmod(n_scans, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_round_robin(n_scans: int, rank: int, world: int) -> range:
    return range(rank, n_scans, world)


def gather_labels(local_labels, owner_ranges, dist=None, dst: int = 0):
    """Host-side gather of per-scan label arrays to rank `dst`.

    local_labels: list of uint8 arrays for this rank's scans (in shard order).
    owner_ranges: list over ranks of the scan indices each rank owns.
    Returns (on dst) a list of label arrays in global scan order; None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local_labels)
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = [np.asarray(l, np.uint8) for l in local_labels]
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    n_total = sum(len(r) for r in owner_ranges)
    out = [None] * n_total
    for r, idxs in enumerate(owner_ranges):
        for k, f in enumerate(idxs):
            out[f] = gathered[r][k]
    return out
