"""Multi-GPU plumbing of the identify path: reads sharded across ranks, index replicated (mode 1).

One process per GPU.  The data path has no collective: every rank runs whole batches of its own reads against its
own copy of the index.  The only exchange is the reduction of the per-taxon counts ([numK][taxa] f64 / u64, a few
kB) and of the k-mer total at the end of a job — the reference's per-thread reduction (Compare.hpp:3445-3454)
lifted to ranks.  torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_first_read(step: int, world: int, rank: int, reads_per_step: int) -> int:
    """Global index of the first read of `rank`'s batch in step `step`: ranks take disjoint, contiguous slices of
    each global step (round-robin over steps), so N ranks process N x reads_per_step reads per step."""
    return (step * world + rank) * reads_per_step


def shard_bounds(n_reads: int, world: int, rank: int) -> tuple[int, int]:
    """[lo, hi) slice of a fixed set of n_reads for `rank` (strong-scaling split of one input file)."""
    base, rem = divmod(n_reads, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counts(count_all: np.ndarray, count_unique: np.ndarray, n_kmers: int, dist=None, device=None):
    """Sum the per-rank profile counts over all ranks.  `dist` = torch.distributed (initialised) or None."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return count_all, count_unique, n_kmers
    import torch
    ca = torch.from_numpy(np.ascontiguousarray(count_all, dtype=np.float64))
    cu = torch.from_numpy(np.ascontiguousarray(count_unique).astype(np.int64))
    nk = torch.tensor([n_kmers], dtype=torch.int64)
    if device is not None:
        ca, cu, nk = ca.to(device), cu.to(device), nk.to(device)
    dist.all_reduce(ca)
    dist.all_reduce(cu)
    dist.all_reduce(nk)
    return ca.cpu().numpy(), cu.cpu().numpy().astype(np.uint64), int(nk.item())
