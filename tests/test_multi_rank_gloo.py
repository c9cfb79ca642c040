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
