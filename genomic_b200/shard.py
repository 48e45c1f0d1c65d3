"""Multi-GPU sharding of a cohort: one process per GPU, every rank segments its own
samples -- the arithmetic of a (sample, chromosome) unit never depends on another unit -- and the only
cross-rank step is the final gather of the segment tables (torch.distributed, NCCL on GPUs; gloo in
the CPU tests).  Unit ids are GLOBAL (sample*24 + chromosome-1) so Philox keys, and therefore
results, do not depend on how the cohort is split."""
from __future__ import annotations

import numpy as np


# the fixed 16-sample parity subset of the 1000-sample cohort (BASELINE configs[3]): two samples of each 125-sample shard
PARITY_SUBSET_1000 = [0, 63, 125, 188, 250, 313, 375, 438, 500, 563, 625, 688, 750, 813, 875, 938]


def parity_subset(n_samples: int):
    if n_samples == 1000:
        return list(PARITY_SUBSET_1000)
    if n_samples <= 16:
        return list(range(n_samples))
    return sorted(set(int(k * (n_samples - 1) // 15) for k in range(16)))


def samples_of_rank(n_samples: int, rank: int, world: int):
    """contiguous blocks, remainder to the first ranks"""
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return list(range(lo, lo + base + (1 if rank < rem else 0)))


def pack_table(seg_count, lengths, means, unit_ids) -> np.ndarray:
    """(n_segments, 3) float64 rows: global unit id, length, mean"""
    uid = np.repeat(np.asarray(unit_ids, dtype=np.float64), np.asarray(seg_count, dtype=np.int64))
    return np.stack([uid, np.asarray(lengths, dtype=np.float64), np.asarray(means, dtype=np.float64)], axis=1) \
        if len(lengths) else np.zeros((0, 3))


def gather_tables(table: np.ndarray, dist=None, device=None, capacity: int | None = None) -> np.ndarray:
    """Gather of the ragged (n,3) tables of all ranks; returns their concatenation in rank order.  ONE collective and ONE
    device->host copy: every rank contributes a fixed-capacity block whose first row carries its row count.  `capacity`
    (rows per rank) must be the same on all ranks; only if some rank has more rows than that -- every rank sees it in the
    gathered counts -- is the gather repeated with the largest count."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return table
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    cap = int(capacity) if capacity else 4096
    while True:
        block = np.zeros((cap + 1, 3))
        block[0, 0] = table.shape[0]
        k = min(table.shape[0], cap)
        block[1:1 + k] = table[:k]
        out = torch.empty((world * (cap + 1), 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, torch.from_numpy(block).to(dev))
        host = out.cpu().numpy().reshape(world, cap + 1, 3)
        counts = host[:, 0, 0].astype(np.int64)
        if int(counts.max()) <= cap:
            return np.concatenate([host[r, 1:1 + counts[r]] for r in range(world)], axis=0)
        cap = int(counts.max())
