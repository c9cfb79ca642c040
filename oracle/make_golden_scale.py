#!/usr/bin/env python
"""Golden outputs of the reference at a scale where its k_low space is SATURATED (tests/golden/scale/*.npz).

TEST INFRASTRUCTURE.  Run in the authoring container after `make -C oracle ref`:
    python oracle/make_golden_scale.py
The small goldens of make_golden.py never reach what BASELINE's C1/C2 are made of: k_low groups naming dozens of taxa,
index spans larger than a lookup tile stages, taxa above k_low by chance.  Here the reference binary (`kASA_scalar -n 1`,
SURVEY.md §8c) builds a 50-taxon index from kasa_b200/synth.py genomes (the generator bench.py uses on the device) and
identifies synthetic reads with
    k = 12..7  (the BASELINE configuration; sparse 7-mer space at this size) and
    k = 11..6  against 50 x 2 Mbp (a 6-letter space smaller than the index: every k-mer hits, groups of several entries and
                taxa, taxa above k_low by chance — what k = 12..7 looks like against C2's 1.67e9 entries).
(k_low < 6 is deliberately not used: there the reference's range lookup (Trie.hpp:162-198) no longer follows the set semantics
of SURVEY.md §3.1 — a probe with k = 10..5 on these inputs gives different values in every cell — and parity is only claimed
for k_low >= 6, which includes every BASELINE configuration.)
Stored per case: sha256 of the index file's records, the printed profile rows, "Number of k-mers", and every (read, taxon,
k-mer score) the reference printed (float32 as parsed from its JSON).  The inputs are NOT stored: the test regenerates them
from the seeds.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kasa_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
DST = os.path.join(ROOT, "tests", "golden", "scale")

# name -> (genome seed, read seed, taxa, genome length, reads, k_high, k_low)
CASES = {
    "c1_k12_7": (31, 32, 50, 400_000, 6000, 12, 7),
    "sat_k11_6": (34, 35, 50, 2_000_000, 2500, 11, 6),
}


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        print(p.stdout)
        raise SystemExit("reference failed: " + " ".join(cmd))
    return p.stdout


def main():
    os.makedirs(DST, exist_ok=True)
    exe = os.path.join(REF, "kASA_scalar")
    built = {}
    for name, (gseed, rseed, n_taxa, glen, n_reads, kh, kl) in CASES.items():
        work = tempfile.mkdtemp(prefix="kasa_scale_")
        os.makedirs(os.path.join(work, "tmp"))
        synth.write_reference(os.path.join(work, "ref.fa"), os.path.join(work, "content.txt"), gseed, n_taxa, glen)
        key = (gseed, n_taxa, glen)
        if key in built:
            for suf in ("", "_info.txt", "_trie", "_trie.txt", "_f.txt"):
                shutil.copy(built[key] + suf, os.path.join(work, "idx" + suf))
        else:
            run([exe, "build", "-c", "content.txt", "-d", "idx", "-i", "ref.fa", "-m", "8", "-n", "8", "-t", "tmp/"], work)
            built[key] = os.path.join(work, "idx")
        n_idx = int(open(os.path.join(work, "idx_info.txt")).read().split()[0])
        sha = hashlib.sha256(open(os.path.join(work, "idx"), "rb").read()[:n_idx * 12]).hexdigest()
        reads = synth.reads_ascii(gseed, rseed, 0, n_reads, n_taxa, glen, 150)
        synth.write_fastq(os.path.join(work, "reads.fq"), reads)
        out = run([exe, "identify", "-c", "content.txt", "-d", "idx", "-i", "reads.fq", "-p", "p.csv", "-q", "q.jsonl", "--jsonl", "-r",
                   "-m", "8", "-n", "1", "-t", "tmp/", "-b", "100000", "-v", "--six", "-k", str(kh), str(kl)], work)
        n_kmers = [int(l.split(":")[2].split()[0]) for l in out.splitlines() if "Number of k-mers in input" in l][0]
        rid, tax, score = [], [], []
        for r, line in enumerate(open(os.path.join(work, "q.jsonl"))):
            j = json.loads(line)
            assert j["Read number"] == r
            for h in j["Top hits"] + j["Further hits"]:
                rid.append(r); tax.append(int(h["tax ID"])); score.append(np.float32(h["k-mer Score"]))
        prof = open(os.path.join(work, "p.csv")).read()
        np.savez_compressed(os.path.join(DST, name + ".npz"), params=np.array([gseed, rseed, n_taxa, glen, n_reads, kh, kl], dtype=np.int64),
                            index_entries=np.int64(n_idx), index_sha256=np.array(sha), n_kmers=np.int64(n_kmers),
                            read=np.array(rid, dtype=np.uint32), taxid=np.array(tax, dtype=np.uint32), score=np.array(score, dtype=np.float32),
                            profile=np.array(prof))
        print(name, "index entries", n_idx, "k-mers", n_kmers, "cells", len(rid), "bytes", os.path.getsize(os.path.join(DST, name + ".npz")))
        if key not in built or built[key] != os.path.join(work, "idx"):
            shutil.rmtree(work, ignore_errors=True)
    for path in set(built.values()):
        shutil.rmtree(os.path.dirname(path), ignore_errors=True)


if __name__ == "__main__":
    main()
