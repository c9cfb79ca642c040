#!/usr/bin/env python
"""bench.py — kASA `identify` (k=12..7, --six) throughput on B200: reads/s and k-mer lookups/s, % of HBM roofline.

Workload (BASELINE.json configs[1], "C2"): synthetic 150 bp reads (1 % substitutions, 0.1 % N, 50 % reverse
strand) against a ~20 GB k=12 index (50 taxa x 33.4 Mbp, built on device by this repo's own build path from
seeded synthetic genomes), index replicated per GPU, reads sharded across ranks (no data-path collective; the
per-taxon counts are all-reduced once at the end).  A step = one batch of reads through the whole hot path
(extract -> radix sort -> lookup -> per-read scores).

  python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
  python bench.py --impl reference ...                          the reference's own CPU identify (oracle/_ref/kASA)

One JSON line on stdout (rank 0).  `value` = reads/s with the batch resident in HBM; `e2e` = the same through the
C ABI with host buffers (H2D of the reads and D2H of the score matrix inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN = 150
KMERS_PER_READ = 260          # 130 per strand at K=12, kLow=7
GENOME_SEED, READ_SEED = 3, 4

# Algorithmic bytes per k-mer lookup of the prescribed algorithm, 64-bit path
BYTES_SORT_PASS = 24.0        # one radix pass: read + write of a 12-byte record
BYTES_PER_LOOKUP_MODEL = 233.2


def kmers_per_read(read_len: int) -> int:
    return 2 * (read_len + 3 * (12 - 7) - 3 * 12 + 1)      # both strands, marker of 3 (K - k_low) 'X' (Read.hpp:36-57, 1068-1078)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--taxa", type=int, default=50)
    ap.add_argument("--genome-len", type=int, default=33_400_000)
    ap.add_argument("--reads-per-step", type=int, default=0, help="default: 6.7 M reads of 150 bp (1.742 G k-mers) or as many longer reads")
    ap.add_argument("--read-len", type=int, default=150, help="150 = C2 (the bench line); 10000 = the long reads of C5")
    ap.add_argument("--sharding", default="replicate", choices=["replicate", "prefix"],
                    help="replicate: index copied to every GPU, reads sharded (default, SURVEY.md §8e mode 1); prefix: index "
                         "sharded by k-mer prefix range as well, match records exchanged over NVLink peer memory (mode 2)")
    ap.add_argument("--score-mode", default="ordered", choices=["ordered", "fast", "exact", "none"],
                    help="per-read scoring: ordered = float32 sums in ascending k-mer order per read (the bench line; <= 1e-6 of the reference, "
                         "tests/test_gpu_reference_scale.py; 'fast' is its old name), exact = the reference's flush order (bit-identical to "
                         "kASA -n 1), none = profile counts only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-reads", type=int, default=200_000,
                    help="reads per step of the CPU reference (a bounded sample of the workload; the reference arm shrinks it further "
                         "when --steps is large)")
    ap.add_argument("--cpu-genome-len", type=int, default=4_000_000,
                    help="genome length per taxon of the CPU reference's index (default: BASELINE's C1, 50 x 4 Mbp = 2.4 GB index)")
    a = ap.parse_args()
    if a.reads_per_step <= 0:
        a.reads_per_step = 6_700_000 if a.read_len == READ_LEN else max(1, 6_700_000 * KMERS_PER_READ // kmers_per_read(a.read_len))
    return a


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out["sm_mhz"] = float(np.median(sm)); out["sm_max_mhz"] = float(max(mx)); out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(n_reads: int, n_taxa: int, genome_len: int, steps: int, warmup: int):
    """Times the reference's own identify on the host cores.  Uses oracle/_ref/kASA (unmodified reference sources
    compiled by oracle/Makefile) when present, else the oracle port.  Returns dict(value reads/s, lookups/s, ...)."""
    from kasa_b200 import synth
    cores = os.cpu_count() or 1
    exe = os.path.join(ROOT, "oracle", "_ref", "kASA")
    work = tempfile.mkdtemp(prefix="kasa_cpu_")
    try:
        os.makedirs(os.path.join(work, "tmp"))
        synth.write_reference(os.path.join(work, "ref.fa"), os.path.join(work, "content.txt"), GENOME_SEED, n_taxa, genome_len)
        times, kmers, hot_ns = [], 0, []
        if os.path.exists(exe):
            kind = "reference"
            mem = "16"
            run = lambda cmd: subprocess.run(cmd, cwd=work, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True).stdout
            run([exe, "build", "-c", "content.txt", "-d", "idx", "-i", "ref.fa", "-m", mem, "-n", str(cores), "-t", "tmp/"])
            for s in range(warmup + steps):
                reads = synth.reads_ascii(GENOME_SEED, READ_SEED, s * n_reads, n_reads, n_taxa, genome_len, READ_LEN)
                synth.write_fastq(os.path.join(work, "reads.fq"), reads, first_id=s * n_reads)
                t0 = time.perf_counter()
                out = run([exe, "identify", "-c", "content.txt", "-d", "idx", "-i", "reads.fq", "-p", "prof.csv", "-q", "rtt.jsonl",
                           "--jsonl", "--six", "-r", "-v", "-m", mem, "-n", str(cores), "-t", "tmp/"])
                dt = time.perf_counter() - t0
                if s >= warmup:
                    times.append(dt)
                ns = 0
                for l in out.splitlines():
                    if "Number of k-mers in input" in l:
                        kmers = int(l.split(":")[2].split()[0])
                    if "Time fastq:" in l or "Time compare:" in l:          # -v: read conversion / sort + compare (Compare.hpp:3689-3690)
                        ns += int(l.split(":")[2].split()[0])
                if s >= warmup and ns:
                    hot_ns.append(ns)
            sample = ("oracle/_ref/kASA identify --six -r -n %d: %d reads x %d bp per step vs an index of %d taxa x %d bp (the reference's own "
                      "CPU-runnable configuration C1 for the index; value = reads / wall time incl. process start, index load and JSONL "
                      "output; hot_path_lookups_per_s = k-mers / (Time fastq + Time compare) as the binary reports them with -v)"
                      % (cores, n_reads, READ_LEN, n_taxa, genome_len))
        else:
            kind = "port"
            cores = 1
            from oracle import oracle as O
            fasta = open(os.path.join(work, "ref.fa"), "rb").read()
            acc2tax = {synth.accession(t): 1000 + t for t in range(n_taxa)}
            index, _ = O.build_index(fasta, acc2tax, 12)
            taxids = np.array([0] + [1000 + t for t in range(n_taxa)], dtype=np.uint32)
            n_reads = min(n_reads, 20000)
            for s in range(warmup + steps):
                reads = synth.reads_ascii(GENOME_SEED, READ_SEED, s * n_reads, n_reads, n_taxa, genome_len, READ_LEN)
                t0 = time.perf_counter()
                res = O.identify_reads([[bytes(r)] for r in reads], index, 12, 12, 7, True, taxids)
                dt = time.perf_counter() - t0
                kmers = res.n_kmers
                if s >= warmup:
                    times.append(dt)
            sample = "oracle port (kasa_oracle.c, 1 thread): %d reads per step vs %d taxa x %d bp" % (n_reads, n_taxa, genome_len)
        t = float(np.mean(times))
        hot = kmers / (float(np.mean(hot_ns)) * 1e-9) if hot_ns else None
        return {"value": n_reads / t, "lookups_per_s": kmers / t, "hot_path_lookups_per_s": hot, "unit": "reads/s", "cores": cores, "kind": kind,
                "sample": sample, "ms_per_step": t * 1e3, "n_reads": n_reads}
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a step = a bounded sample: the whole run (index build + W + K identify processes) stays within a few minutes
    n_reads = max(20_000, min(args.cpu_reads, 2_000_000 // max(1, args.steps + args.warmup)))
    r = cpu_reference_run(n_reads, args.taxa, args.cpu_genome_len, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "identify reads/s (k=12..7, six frames)", "value": r["value"], "unit": "reads/s",
        "lookups_per_s": r["lookups_per_s"], "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(args, r["n_reads"]),
        "cpu_baseline": {"value": r["value"], "unit": "reads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                         "hot_path_lookups_per_s": r["hot_path_lookups_per_s"]},
        "e2e": {"value": r["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, reads_per_step):
    how = ("index sharded by k-mer prefix range, reads sharded, match records over NVLink peer memory" if args.sharding == "prefix"
           else "index replicated, reads sharded")
    name = "C2" if args.read_len == READ_LEN else "C5 shape (long reads) on the C2 index"
    return {"workload": "%s: identify k=12..7 --six, %d bp reads vs k=12 index of %d taxa x %d bp (%s)"
                        % (name, args.read_len, args.taxa, args.genome_len, how), "sharding": args.sharding, "score_mode": args.score_mode,
            "reads_per_step_per_gpu": reads_per_step, "kmers_per_read": kmers_per_read(args.read_len),
            "l2_policy": "inputs larger than L2 (each step streams > 10 GB; a fresh read batch per step)"}


# DRAM bytes per k-mer (dram__bytes_read.sum + dram__bytes_write.sum of one launch / k-mers of the launch) from the committed ncu
# captures at this workload's full size (profiles/r02_ncu_*.txt; the radix pass from profiles/r01_ncu_onesweep_kernel.txt)
NCU_DRAM_BYTES = {"extract": (1.142314e9 + 20.847225e9) / 1.742e9, "sort_pass": (3.195906e9 + 3.357949e9) / 2.6e8,
                  "lookup": (51.864335e9 + 76.196481e9) / 1.742e9, "score": (55.928244e9 + 1.429569e9) / 1.742e9}


# ------------------------------------------------------------------------------------------------ our arm
def main_ours(args):
    import torch
    import kasa_b200
    from kasa_b200 import capi
    from kasa_b200.dist import reduce_counts, shard_first_read

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    prefix = args.sharding == "prefix"
    if world > 1 or prefix:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    R = args.reads_per_step
    RL, KPR = args.read_len, kmers_per_read(args.read_len)
    n_windows = args.taxa * (args.genome_len + 33)
    cap = max(R * KPR, n_windows) + 4096
    if args.score_mode == "fast":
        args.score_mode = "ordered"
    score_mode = {"ordered": capi.SCORE_ORDERED, "exact": capi.SCORE_EXACT, "none": capi.SCORE_NONE}[args.score_mode]
    params = kasa_b200.Params(device=local, highest_k=12, k_high=12, k_low=7, frames=6, score_mode=score_mode,
                              max_batch_kmers=cap, max_batch_seqs=R, max_batch_bases=max(R * RL, args.genome_len + 64))
    ctx = kasa_b200.Context(params)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # ---- setup (untimed): build the index on device, generate read batches in HBM
    t0 = time.perf_counter()
    ctx.synth_build_reference(GENOME_SEED, args.taxa, args.genome_len)
    n_entries, n_trie = ctx.build_finish(fetch=False)
    taxids = np.array([0] + [1000 + t for t in range(args.taxa)], dtype=np.uint32)
    group = None
    if prefix:
        from kasa_b200.dist import ShardGroup
        entry_begin, prefix_begin = ctx.index_from_build_shard(taxids, rank, world)     # keep 1/world of the index
        group = ShardGroup(ctx, dist, prefix_begin, torch.device("cuda", local))
    else:
        ctx.index_from_build(taxids)
    build_s = time.perf_counter() - t0
    T = len(taxids)

    n_dev_batches = 4
    dev_reads = [torch.empty(R * RL, dtype=torch.uint8, device="cuda") for _ in range(n_dev_batches)]
    seq_start = torch.arange(R, dtype=torch.int64, device="cuda") * RL
    seq_len = torch.full((R,), RL, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    def gen(step, buf, location):
        first = shard_first_read(step, world, rank, R)        # ranks take disjoint read ranges
        ctx.synth_reads(GENOME_SEED, READ_SEED, first, R, args.taxa, args.genome_len, RL, out=buf, out_location=location)

    for b in range(n_dev_batches):
        gen(b, dev_reads[b].data_ptr(), capi.MEM_DEVICE)

    def run_device_step(step):
        buf = dev_reads[step % n_dev_batches]
        rb = capi.CReadBatch(R, R, capi.SEQ_ASCII, capi.MEM_DEVICE, buf.data_ptr(), None, seq_start.data_ptr(), seq_len.data_ptr(), None)
        res = capi.CBatchResult(0, None, capi.MEM_DEVICE, 0.0)
        if group is not None:
            group.identify_batch_raw(rb, res)
        else:
            ctx.identify_batch_raw(rb, res)
        return res

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident measurement
    for s in range(args.warmup):
        run_device_step(s)
    ctx.stage_times(reset=True)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    gpu_ms_sum, n_kmers = 0.0, 0
    for s in range(args.steps):
        res = run_device_step(args.warmup + s)
        gpu_ms_sum += res.gpu_ms
        n_kmers += res.n_kmers
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else {}
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = ctx.launches - l0
    stages, n_batches, n_passes, stage_kmers = ctx.stage_times(reset=True)
    reads_total = world * args.steps * R
    value = reads_total / (ms_total * 1e-3)
    lookups = world * n_kmers / (ms_total * 1e-3)

    # ---- roofline: every stage's kernel against its algorithmic bytes, the dominant one (longest launch) as the
    # line's `roofline`; durations are CUDA events on the launching stream (kasa_ctx_stage_times)
    peak, peak_src = measured_peak()
    kmers_per_batch = stage_kmers / max(1, n_batches)
    passes_per_batch = n_passes / max(1, n_batches)
    index_bytes_per_lookup = 12.0 * n_entries / max(1.0, kmers_per_batch)
    per = {k: v / max(1, n_batches) for k, v in stages.items()}
    kernels = {
        "extract": ("extract_identify_kernel (translation, k-mer packing, per-read presort)", 12.4, per["extract"], 1.0, NCU_DRAM_BYTES["extract"]),
        "sort_pass": ("onesweep_kernel<256,18> (one 8-bit LSD radix pass over 12-byte records)", BYTES_SORT_PASS,
                      per["sort_passes"] / max(1.0, passes_per_batch), passes_per_batch, NCU_DRAM_BYTES["sort_pass"]),
        "lookup": ("lookup_kernel (sorted k-mers against the index span of their tile, one 32-byte match record each)",
                   20.0 + index_bytes_per_lookup, per["lookup"], 1.0, NCU_DRAM_BYTES["lookup"]),
        "score": ("score_dense_kernel (ordered float32 per-read scores and counts from the match records)" if args.score_mode == "ordered"
                  else "scoring stage (%s)" % args.score_mode, 0.8, per["per_read"], 1.0, NCU_DRAM_BYTES["score"]),
    }
    stage_roofline = {}
    for key, (name, bytes_unit, ms, launches_per_step, ncu_b) in kernels.items():
        ach = bytes_unit * kmers_per_batch / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stage_roofline[key] = {"kernel": name, "bytes_per_unit": bytes_unit, "avg_launch_ms": ms, "launches_per_step": launches_per_step,
                               "achieved": ach, "frac": ach / peak,
                               "dram_bytes_per_unit_ncu": ncu_b, "dram_GBps": (ncu_b * kmers_per_batch / (ms * 1e-3) / 1e9) if (ncu_b and ms > 0) else None}
    dom = max(stage_roofline, key=lambda k: stage_roofline[k]["avg_launch_ms"])
    d = stage_roofline[dom]
    roofline = {"bound": "hbm", "kernel": d["kernel"], "achieved": d["achieved"], "peak": peak, "unit": "GB/s", "frac": d["frac"],
                "traffic": (d["dram_bytes_per_unit_ncu"] * kmers_per_batch) if d["dram_bytes_per_unit_ncu"] else None,
                "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this workload's "
                                  "size (profiles/r02_ncu_*.txt), per k-mer, scaled to this launch",
                "peak_source": peak_src, "bytes_per_unit": d["bytes_per_unit"], "units_per_launch": kmers_per_batch,
                "avg_launch_ms": d["avg_launch_ms"], "share_of_step": d["avg_launch_ms"] * d["launches_per_step"] / max(1e-9, sum(per.values())),
                "note": "dominant kernel = the longest launch of the step; its algorithmic bytes are SURVEY.md §8d's figure for the stage (the "
                        "accumulate stage is counted there as the dense score write-back only: the kernel is issue-bound, not HBM-bound); "
                        "whole_step_model is the figure comparable across stages and rounds",
                "stages": stage_roofline,
                "whole_step_model": {"bytes_per_lookup": BYTES_PER_LOOKUP_MODEL + index_bytes_per_lookup, "achieved_GBps": None}}
    model_bytes = roofline["whole_step_model"]["bytes_per_lookup"]
    roofline["whole_step_model"]["achieved_GBps"] = (n_kmers / args.steps) * model_bytes / (ms_total / args.steps * 1e-3) / 1e9
    roofline["whole_step_model"]["frac"] = roofline["whole_step_model"]["achieved_GBps"] / peak
    stage_ms = {k: v / max(1, n_batches) for k, v in stages.items()}

    # ---- end-to-end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        n_host = 3
        host_reads = [torch.empty(R * RL, dtype=torch.uint8, pin_memory=True) for _ in range(n_host)]
        host_scores = [torch.empty(R * T, dtype=torch.float32, pin_memory=True) for _ in range(3)]
        h_start = (np.arange(R, dtype=np.uint64) * RL)
        h_len = np.full(R, RL, dtype=np.uint32)
        for b in range(n_host):
            gen(1000 + b, host_reads[b].data_ptr(), capi.MEM_HOST)

        def run_host_steps(n_steps):
            """the production loop of the C ABI: kasa_identify_submit / kasa_identify_wait, three batches in flight — the
            next batch's host->device copy and the previous batch's score read-back overlap the kernels and each other"""
            keep = []
            for step in range(n_steps if group is not None else 0):        # prefix-sharded: one batch at a time
                buf = host_reads[step % n_host]
                rb = capi.CReadBatch(R, R, capi.SEQ_ASCII, capi.MEM_HOST, buf.data_ptr(), None, h_start.ctypes.data, h_len.ctypes.data, None)
                res = capi.CBatchResult(0, host_scores[step % 3].data_ptr(), capi.MEM_HOST, 0.0)
                group.identify_batch_raw(rb, res)
            if group is not None:
                return
            for step in range(n_steps):
                buf = host_reads[step % n_host]
                rb = capi.CReadBatch(R, R, capi.SEQ_ASCII, capi.MEM_HOST, buf.data_ptr(), None, h_start.ctypes.data, h_len.ctypes.data, None)
                res = capi.CBatchResult(0, host_scores[step % 3].data_ptr(), capi.MEM_HOST, 0.0)
                ctx.identify_submit_raw(rb, res)
                keep.append((rb, res))
                if step >= 2:
                    ctx.identify_wait()
            for _ in range(min(2, n_steps)):
                ctx.identify_wait()

        run_host_steps(min(2, args.warmup))
        ctx.stage_times(reset=True)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run_host_steps(args.steps)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
        e2e_stages, e2e_batches, _, _ = ctx.stage_times(reset=True)
        e2e = {"value": reads_total / (e2e_ms * 1e-3), "unit": "reads/s",
               "h2d_bytes_per_step": int(R * RL + R * 12), "d2h_bytes_per_step": int(R * T * 4),
               "ms_per_step": e2e_ms / args.steps, "wall_ms_per_step": wall * 1e3 / args.steps,
               "device_ms_per_step": e0.elapsed_time(e1) / args.steps,
               "stage_ms_per_step": {k: v / max(1, e2e_batches) for k, v in e2e_stages.items()}, "checksum": float(host_scores[(args.steps - 1) % 3][: T * 1000].sum()),
               "api": ("kasa_shard_sort/_lookup/_score per batch (pinned host buffers)" if group is not None else
                       "kasa_identify_submit/_wait (3 batches in flight, pinned host buffers)")}

    # ---- job epilogue: reduce the per-taxon counts across ranks (the only collective of this path)
    ca, cu, total_kmers = ctx.identify_finish()
    ca, cu, total_kmers = reduce_counts(ca, cu, total_kmers, dist, torch.device("cuda", local))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.cpu_reads, args.taxa, args.cpu_genome_len, steps=1, warmup=0)
            cpu = {"value": r["value"], "unit": "reads/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "lookups_per_s": r["lookups_per_s"], "hot_path_lookups_per_s": r["hot_path_lookups_per_s"]}
        except Exception as e:            # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "reads/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    if rank == 0:
        line = {
            "metric": "identify reads/s (k=12..7, six frames)", "value": value, "unit": "reads/s", "lookups_per_s": lookups,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": workload_config(args, R), "index_entries": int(n_entries), "index_gb": n_entries * 12 / 1e9,
            "index_entries_this_gpu": int(ctx.index_size),
            "index_build_s": build_s, "stage_ms_per_step": stage_ms, "gpu_ms_per_step_lib": gpu_ms_sum / args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "identified_frac_k12": float(ca[0].sum() / max(1, total_kmers)),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
