#!/usr/bin/env python
"""Sharded build (BASELINE config C4 shape, scaled to the GPUs at hand): every rank extracts the k-mers of its own taxa,
owners pull their prefix range over NVLink, sort, dedup, trie.  Run under torchrun; prints one JSON line (rank 0).
  python -m torch.distributed.run --nproc-per-node N tools/bench_build_sharded.py --taxa 50 --genome-len 33400000
Not a bench.py line (build throughput is reported next to the identify metric)."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--taxa", type=int, default=50)
    ap.add_argument("--genome-len", type=int, default=33_400_000)
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import kasa_b200
    from kasa_b200.dist import sharded_build_finish
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29542")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    per = (a.taxa + world - 1) // world
    first, n_mine = rank * per, max(0, min(per, a.taxa - rank * per))
    windows = (a.genome_len + 33)
    # a shard holds ~ total / world records, a rank's brick its own taxa: size for the larger with slack for imbalance
    cap = int(max(n_mine * windows, 1.3 * a.taxa * windows / world)) + 4096
    ctx = kasa_b200.Context(kasa_b200.Params(device=local, max_batch_kmers=cap, max_batch_seqs=1024, max_batch_bases=a.genome_len + 64))
    times = []
    for it in range(a.repeat):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.synth_build_reference(3, n_mine, a.genome_len, first_taxon=first)
        t1 = time.perf_counter()
        pb, (n_entries, n_trie) = sharded_build_finish(ctx, dist, 7, dev, fetch=False)
        torch.cuda.synchronize(); dist.barrier()
        t2 = time.perf_counter()
        tot = torch.tensor([n_entries, n_trie], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        times.append((t2 - t0, t1 - t0, int(tot[0]), int(tot[1]), n_entries, dict(sharded_build_finish.last_phase_s)))
    if rank == 0:
        best = min(times)
        print(json.dumps({"what": "sharded build (extract + route by prefix + sort + dedup + trie)", "n_gpus": world, "taxa": a.taxa,
                          "genome_len": a.genome_len, "bases": a.taxa * a.genome_len, "index_entries": best[2], "trie_records": best[3],
                          "entries_rank0": best[4], "seconds": best[0], "seconds_extract": best[1],
                          "phases_s": best[5], "bases_per_s": a.taxa * a.genome_len / best[0], "all_runs_s": [t[0] for t in times]}), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
