"""World-size-2 CPU test (gloo) of the N>1 host path: reads sharded across ranks, index replicated, per-taxon counts
all-reduced at the end (kasa_b200/dist.py).  The per-rank compute is the CPU oracle here — the point is the sharding
and the reduction, which are the same code bench.py runs over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import golden_util as G
    from kasa_b200.dist import reduce_counts, shard_bounds
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, taxids, _ = G.content("k12")
    idx = G.index("k12")
    reads = [[s] for _, s in G.fastq("k12")]
    lo, hi = shard_bounds(len(reads), world, rank)
    res = O.identify_reads(reads[lo:hi], idx, 12, 12, 7, True, taxids)
    ca, cu, nk = reduce_counts(res.count_all, res.count_unique, res.n_kmers, dist)
    if rank == 0:
        np.savez(out, ca=ca, cu=cu, nk=nk, lo=lo, hi=hi)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reduce_to_the_single_rank_result(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as G
    from oracle import oracle as O
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    _, taxids, _ = G.content("k12")
    reads = [[s] for _, s in G.fastq("k12")]
    want = O.identify_reads(reads, G.index("k12"), 12, 12, 7, True, taxids)
    assert int(got["nk"]) == want.n_kmers == G.n_kmers("k12", "six_12_7")
    assert np.array_equal(got["cu"], want.count_unique)
    np.testing.assert_allclose(got["ca"], want.count_all, rtol=1e-12)


def _shard_worker(rank, world, port, out):
    """prefix-sharded mode (SURVEY.md §8e mode 2) on CPU: rank r owns the index entries of prefix range r (kasa_shard_plan)
    and looks ALL reads up against them (the oracle stands in for the kernels); counts and scores are additive over
    shards because no level-k group is split; the small structs travel through the same gather the GPU path uses."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import golden_util as G
    from kasa_b200 import capi
    from kasa_b200.dist import _gather_struct
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, taxids, _ = G.content("k12")
    idx = G.index("k12")
    eb, pb = capi.shard_plan(idx, 7, world)
    runs = capi.CShardRuns()
    runs.n_kmers = 1000 + rank
    runs.cur = rank & 1
    for i in range(world + 1):
        runs.begin[i] = 10 * rank + i
    got = _gather_struct(dist, runs)
    assert [int(g.n_kmers) for g in got] == [1000 + r for r in range(world)]
    assert [int(g.begin[world]) for g in got] == [10 * r + world for r in range(world)]
    reads = [[s] for _, s in G.fastq("k12")]
    res = O.identify_reads(reads, idx[int(eb[rank]):int(eb[rank + 1])], 12, 12, 7, True, taxids)
    ca = torch.from_numpy(res.count_all.copy())
    cu = torch.from_numpy(res.count_unique.astype(np.int64))
    sc = torch.from_numpy(res.scores.astype(np.float64))
    for t in (ca, cu, sc):
        dist.all_reduce(t)
    if rank == 0:
        np.savez(out, ca=ca.numpy(), cu=cu.numpy(), sc=sc.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_prefix_shards_add_up_to_the_unsharded_result(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as G
    from oracle import oracle as O
    out = str(tmp_path / "s0.npz")
    mp.spawn(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    _, taxids, _ = G.content("k12")
    reads = [[s] for _, s in G.fastq("k12")]
    want = O.identify_reads(reads, G.index("k12"), 12, 12, 7, True, taxids)
    assert np.array_equal(got["cu"].astype(np.uint64), want.count_unique)
    np.testing.assert_allclose(got["ca"], want.count_all, rtol=1e-12)
    np.testing.assert_allclose(got["sc"], want.scores, rtol=1e-5)
