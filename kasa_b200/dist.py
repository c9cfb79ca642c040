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


# ---------------------------------------------------------------------------------------------------------------
# SURVEY.md §8e mode 2: index sharded by k-mer prefix range, reads sharded too.  There is no send/receive staging:
# the owner of a prefix range runs its lookup kernel directly on the peers' sorted runs (streaming NVLink loads),
# and each origin's scatter kernel pulls its match records back the same way (include/kasa_b200.h, kasa_shard_*).
# The host exchanges one small struct per rank and batch; torch.distributed is that plumbing.

def _gather_struct(dist, obj, device=None):
    """all-gather one fixed-size ctypes struct per rank -> list of structs (rank order)."""
    import ctypes as C
    import torch
    raw = bytes(obj)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    on_gpu = dist.get_backend() == "nccl"
    if on_gpu:
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    out = torch.empty(dist.get_world_size() * len(raw), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    buf = out.cpu().numpy().tobytes()
    cls = type(obj)
    return [cls.from_buffer_copy(buf[i * len(raw):(i + 1) * len(raw)]) for i in range(dist.get_world_size())]


def sharded_build_finish(ctx, dist, k_low: int, device=None, fetch: bool = True):
    """Finish a build whose reference sequences were pushed rank by rank (BASELINE config C4): route every record to the
    owner of its prefix range (owners pull contiguous runs from the peers' sorted bricks over NVLink), then sort /
    dedup / run-length encode locally.  Returns (prefix_begin, what Context.build_finish returns for this rank's
    shard); the shards written in rank order are the reference's index file."""
    import time
    import torch
    from . import capi
    t = [time.perf_counter()]
    hist, _ = ctx.build_shard_local()
    t.append(time.perf_counter())
    h = torch.from_numpy(hist.astype(np.int64))
    if dist.get_backend() == "nccl":
        h = h.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.all_reduce(h)
    prefix_begin = capi.shard_plan_hist(h.cpu().numpy().astype(np.uint64), k_low, dist.get_world_size())
    mem = ctx.shard_init(dist.get_rank(), dist.get_world_size(), prefix_begin)
    ctx.shard_connect(_gather_struct(dist, mem, device))
    t.append(time.perf_counter())
    runs = _gather_struct(dist, ctx.build_shard_runs(), device)
    ctx.build_shard_pull(runs)
    dist.barrier()                                            # nobody reads a peer's brick any more
    t.append(time.perf_counter())
    out = ctx.build_finish(fetch=fetch)
    t.append(time.perf_counter())
    sharded_build_finish.last_phase_s = dict(zip(("local_sort_dedup", "plan_and_connect", "pull", "sort_dedup_trie"), np.diff(t).tolist()))
    return prefix_begin, out


class ShardGroup:
    """One rank's handle on a prefix-sharded identify job: `ctx` already holds shard `rank` of the index."""

    def __init__(self, ctx, dist, prefix_begin, device=None):
        self.ctx, self.dist, self.device = ctx, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        mem = ctx.shard_init(self.rank, self.world, prefix_begin)
        ctx.shard_connect(_gather_struct(dist, mem, device))
        dist.barrier()

    def identify_batch_raw(self, rb, res):
        """One batch of this rank's reads against the whole (sharded) index; fills `res` like kasa_identify_batch."""
        runs = self.ctx.shard_sort_raw(rb)                    # extract + sort + where the sorted run changes owner
        all_runs = _gather_struct(self.dist, runs, self.device)
        self.ctx.shard_lookup(all_runs)                       # serves every rank's queries of this shard's range
        self.dist.barrier()                                   # all records have reached their origin
        self.ctx.shard_score_raw(all_runs, res)                # pulls the peers' records, scores this rank's reads


class LocalShardGroup:
    """The same protocol for several contexts in ONE process (tests: n shards on one GPU; or one process driving
    several GPUs): peers are reached through their raw device pointers."""

    def __init__(self, ctxs, prefix_begin):
        self.ctxs = list(ctxs)
        n = len(self.ctxs)
        mems = [c.shard_init(r, n, prefix_begin) for r, c in enumerate(self.ctxs)]
        for c in self.ctxs:
            c.shard_connect(mems)

    def identify_batches(self, batches, read_first_seqs=None):
        """batches[r] = the reads of rank r (as for Context.identify_batch) -> [(n_kmers, scores, gpu_ms)] per rank."""
        runs, n_reads = [], []
        for r, c in enumerate(self.ctxs):
            ru, nr = c.shard_sort(batches[r], read_first_seqs[r] if read_first_seqs is not None else None)
            runs.append(ru)
            n_reads.append(nr)
        for c in self.ctxs:
            c.shard_lookup(runs)
        return [c.shard_score(runs, n_reads[r]) for r, c in enumerate(self.ctxs)]


def local_sharded_build_finish(ctxs, k_low: int):
    """sharded_build_finish for several contexts in one process -> (prefix_begin, [(records, trie)] per shard)."""
    from . import capi
    hist = sum(c.build_shard_local()[0] for c in ctxs)
    prefix_begin = capi.shard_plan_hist(hist, k_low, len(ctxs))
    mems = [c.shard_init(r, len(ctxs), prefix_begin) for r, c in enumerate(ctxs)]
    for c in ctxs:
        c.shard_connect(mems)
    runs = [c.build_shard_runs() for c in ctxs]
    for c in ctxs:
        c.build_shard_pull(runs)
    return prefix_begin, [c.build_finish() for c in ctxs]
