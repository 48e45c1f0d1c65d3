"""Multi-GPU sharding of a cohort: one process per GPU, every rank segments its own
samples -- the arithmetic of a (sample, chromosome) unit never depends on another unit -- and the only
cross-rank step is the final gather of the segment tables (torch.distributed, NCCL on GPUs; gloo in
the CPU tests).  Unit ids are GLOBAL (sample*24 + chromosome-1) so Philox keys, and therefore
results, do not depend on how the cohort is split."""
from __future__ import annotations

import numpy as np


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


def gather_tables(table: np.ndarray, dist=None, device=None) -> np.ndarray:
    """all_gather of ragged (n,3) tables; returns the concatenation in rank order"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return table
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([table.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    pad = torch.zeros((mx, 3), dtype=torch.float64, device=dev)
    if table.shape[0]:
        pad[: table.shape[0]] = torch.from_numpy(np.ascontiguousarray(table)).to(dev)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)
