#!/usr/bin/env python
"""Throughput of the BASELINE.json configs that are not bench.py's line (the other configs are
parity-test cases; this tool only reports how fast they run at scale, for DESIGN.md).
  --config c3 : identify, k = 25..7 (128-bit k-mers, 19 levels), --six, 2 x 150 bp pairs, 50 taxa x 4 Mbp index
One JSON line; device-resident inputs, CUDA-event stage times from the library."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3", choices=["c3"])
    ap.add_argument("--pairs-per-step", type=int, default=1_000_000)
    ap.add_argument("--taxa", type=int, default=50)
    ap.add_argument("--genome-len", type=int, default=4_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    import torch
    import kasa_b200
    from kasa_b200 import capi
    torch.cuda.set_device(0)
    K, k_low, read_len = 25, 7, 150
    P = a.pairs_per_step
    kmers_per_pair = 520                                   # 130 per strand and mate at k_low = 7
    n_windows = a.taxa * (a.genome_len + 3 * (K - 1))
    cap = max(P * kmers_per_pair, n_windows) + 4096
    ctx = kasa_b200.Context(kasa_b200.Params(device=0, highest_k=K, k_high=K, k_low=k_low, frames=6, score_mode=capi.SCORE_FAST,
                                             max_batch_kmers=cap, max_batch_seqs=2 * P, max_batch_bases=max(2 * P * read_len, a.genome_len + 128)))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.synth_build_reference(5, a.taxa, a.genome_len)
    n_entries, n_trie = ctx.build_finish(fetch=False)
    taxids = np.array([0] + [1000 + t for t in range(a.taxa)], dtype=np.uint32)
    ctx.index_from_build(taxids)
    build_s = time.perf_counter() - t0
    bufs = [torch.empty(2 * P * read_len, dtype=torch.uint8, device="cuda") for _ in range(3)]
    for b, buf in enumerate(bufs):
        ctx.synth_reads(5, 6, b * 2 * P, 2 * P, a.taxa, a.genome_len, read_len, out=buf.data_ptr(), out_location=capi.MEM_DEVICE)
    seq_start = torch.arange(2 * P, dtype=torch.int64, device="cuda") * read_len
    seq_len = torch.full((2 * P,), read_len, dtype=torch.int32, device="cuda")
    first = torch.arange(0, 2 * P + 1, 2, dtype=torch.int32, device="cuda")        # mates are consecutive and share a read id
    torch.cuda.synchronize()

    def step(i):
        rb = capi.CReadBatch(P, 2 * P, capi.SEQ_ASCII, capi.MEM_DEVICE, bufs[i % 3].data_ptr(), None, seq_start.data_ptr(), seq_len.data_ptr(),
                             first.data_ptr())
        res = capi.CBatchResult(0, None, capi.MEM_DEVICE, 0.0)
        ctx.identify_batch_raw(rb, res)
        return res

    for i in range(a.warmup):
        step(i)
    ctx.stage_times(reset=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    n_kmers = 0
    for i in range(a.steps):
        n_kmers += step(a.warmup + i).n_kmers
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    stages, nb, n_passes, _ = ctx.stage_times(reset=True)
    ca, cu, total = ctx.identify_finish()
    print(json.dumps({"config": "C3: identify k=25..7 --six, 2x150 bp pairs, %d taxa x %d bp (128-bit index)" % (a.taxa, a.genome_len),
                      "pairs_per_s": a.steps * P / (ms * 1e-3), "lookups_per_s": n_kmers / (ms * 1e-3), "ms_per_step": ms / a.steps,
                      "pairs_per_step": P, "kmers_per_step": n_kmers // a.steps, "index_entries": int(n_entries), "index_build_s": build_s,
                      "stage_ms_per_step": {k: v / max(1, nb) for k, v in stages.items()}, "radix_passes_per_step": n_passes // max(1, nb),
                      "identified_frac_k25": float(ca[0].sum() / max(1, total))}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
