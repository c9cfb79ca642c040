#!/usr/bin/env python
"""Probe for DESIGN.md §7 item 1: do the stages of consecutive batches overlap when they run on two streams?

Two contexts on ONE device (each with its own copy of the C2 index and buffers for a quarter-size batch) are fed alternately
through kasa_identify_submit / _wait with page-locked host reads; the same number of quarter batches through one context is the
baseline.  Prints one JSON line.  GPU tool, not part of bench.py's line.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1_675_000)
    ap.add_argument("--taxa", type=int, default=50)
    ap.add_argument("--genome-len", type=int, default=33_400_000)
    ap.add_argument("--batches", type=int, default=24)
    ap.add_argument("--prio", type=int, default=0, help="1: the second context's stream gets the high priority")
    args = ap.parse_args()
    import torch
    import kasa_b200
    from kasa_b200 import capi

    R, RL = args.reads, 150
    KPR = 6 * (RL // 3 - 12 + 1) - 0
    taxids = np.array([0] + [1000 + t for t in range(args.taxa)], dtype=np.uint32)

    # the index is built once by a full-size context (the build brick needs one record per window), fetched, and loaded into two
    # contexts sized for quarter batches
    def params(kmers):
        return kasa_b200.Params(device=0, highest_k=12, k_high=12, k_low=7, frames=6, score_mode=capi.SCORE_ORDERED,
                                max_batch_kmers=kmers, max_batch_seqs=R, max_batch_bases=max(R * RL, args.genome_len + 64))

    b = kasa_b200.Context(params(args.taxa * (args.genome_len + 33) + 4096))
    b.synth_build_reference(21, args.taxa, args.genome_len)
    idx, _ = b.build_finish()
    b.close()

    def make():
        c = kasa_b200.Context(params(R * 264 + 4096))
        c.load_index(idx, taxids)
        return c

    ctxs = [make(), make()]
    if args.prio:
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        s = torch.cuda.Stream(priority=-1)
        ctxs[1].set_stream(s.cuda_stream)
    # page-locked host batches: kasa_identify_submit then returns without waiting for the device (a device-resident batch asks the
    # device for its k-mer total)
    bufs = [torch.empty(R * RL, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
    h_start = np.arange(R, dtype=np.uint64) * RL
    h_len = np.full(R, RL, dtype=np.uint32)
    for b, buf in enumerate(bufs):
        ctxs[0].synth_reads(21, 22, b * R, R, args.taxa, args.genome_len, RL, out=buf.data_ptr(), out_location=capi.MEM_HOST)
    torch.cuda.synchronize()
    keep = []

    def submit(c, step):
        rb = capi.CReadBatch(R, R, capi.SEQ_ASCII, capi.MEM_HOST, bufs[step % 4].data_ptr(), None, h_start.ctypes.data, h_len.ctypes.data, None)
        res = capi.CBatchResult(0, None, capi.MEM_DEVICE, 0.0)
        keep.append((rb, res))
        c.identify_submit_raw(rb, res)

    def run(n_ctx, n):
        inflight = [0] * n_ctx
        for step in range(n):
            i = step % n_ctx
            if inflight[i] >= 2:
                ctxs[i].identify_wait()
                inflight[i] -= 1
            submit(ctxs[i], step)
            inflight[i] += 1
        for i in range(n_ctx):
            while inflight[i]:
                ctxs[i].identify_wait()
                inflight[i] -= 1
        torch.cuda.synchronize()

    out = {"reads_per_batch": R, "batches": args.batches, "prio": args.prio}
    for n_ctx in (1, 2, 1, 2):
        run(n_ctx, 4)
        t0 = time.perf_counter()
        run(n_ctx, args.batches)
        dt = time.perf_counter() - t0
        out.setdefault("ms_per_batch_%dctx" % n_ctx, []).append(dt * 1e3 / args.batches)
    out["reads_per_s_1ctx"] = R / (min(out["ms_per_batch_1ctx"]) * 1e-3)
    out["reads_per_s_2ctx"] = R / (min(out["ms_per_batch_2ctx"]) * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
