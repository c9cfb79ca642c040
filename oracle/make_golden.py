#!/usr/bin/env python
"""Generate tests/golden/* by running the reference itself (oracle/_ref binaries).

TEST INFRASTRUCTURE.  Run in the authoring container after `make -C oracle ref`:
    python oracle/make_golden.py
The reference has no golden vectors of its own, so parity is pinned
on these outputs of the unmodified reference sources: `kASA_scalar -n 1` for the
64-bit cases, `kASA_scalar_fix128 -n 1` for k=25 (SURVEY.md §8c explains both).
Inputs are seeded (kasa_b200/synth.py) and adversarial on purpose: genomes share
regions (non-unique k-mers), contain runs of N, tandem repeats (low complexity:
exercises the fp32 flush order) and reads have substitutions, Ns, both strands
and ragged lengths (including reads too short to yield a k-mer).
"""
from __future__ import annotations

import gzip
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
GOLD = os.path.join(ROOT, "tests", "golden")


def gz_write(path: str, data: bytes) -> None:
    with open(path, "wb") as f, gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as g:
        g.write(data)


def make_reference(d: str, seed: int, n_taxa: int, glen: int, repeats: bool = True) -> None:
    """FASTA with shared regions, N runs and (taxon 0) tandem repeats."""
    seqs = [bytearray(synth.genome_ascii(seed, t, glen)) for t in range(n_taxa)]
    q = glen // 8
    # regions shared by 2, 3 and 4 taxa
    seqs[1][q:3 * q] = seqs[0][q:3 * q]
    seqs[2][2 * q:3 * q] = seqs[0][2 * q:3 * q]
    if n_taxa > 3:
        seqs[3][2 * q + q // 2:3 * q] = seqs[0][2 * q + q // 2:3 * q]
    if repeats:
        unit = bytes(seqs[0][10:31])                      # 21-mer tandem repeat (period 7 codons)
        seqs[0][4 * q:5 * q] = (unit * (q // len(unit) + 1))[:q]
        seqs[n_taxa - 1][4 * q:4 * q + q // 2] = seqs[0][4 * q:4 * q + q // 2]
        seqs[0][5 * q:5 * q + 300] = b"ACG" * 100        # codon-periodic repeat
        seqs[0][5 * q + 300:5 * q + 500] = b"A" * 200    # homopolymer
    for t in range(n_taxa):
        for r in range(3):
            h = int(synth.splitmix64(np.array([seed * 1000 + t * 10 + r], dtype=np.uint64))[0])
            p = h % (glen - 10)
            run = 1 + (h >> 40) % 6
            seqs[t][p:p + run] = b"N" * run
    with open(os.path.join(d, "ref.fa"), "wb") as f:
        for t in range(n_taxa):
            f.write(b">" + synth.accession(t).encode() + b" synthetic taxon %d\n" % t)
            s = bytes(seqs[t])
            for a in range(0, len(s), 80):
                f.write(s[a:a + 80] + b"\n")
    with open(os.path.join(d, "content.txt"), "w") as f:
        for t in range(n_taxa):
            f.write("Synthetic taxon %d\t%d\t%d\t%s\n" % (t, 1000 + t, 1000 + t, synth.accession(t)))
    return [bytes(s) for s in seqs]


def make_reads(seqs, seed: int, n: int, glen: int, ragged: bool = True):
    """list of read byte strings: random-position reads with errors, repeat-region reads, ragged lengths."""
    n_taxa = len(seqs)
    rng = np.random.default_rng(seed)
    out = []
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    q = glen // 8
    for i in range(n):
        t = int(rng.integers(n_taxa))
        L = 150
        if ragged and i % 7 == 3:
            L = int(rng.integers(20, 260))
        if i % 5 == 0:       # from the shared / repeated part of taxon 0
            p = int(rng.integers(q, 5 * q + 500 - L)) if 5 * q + 500 - L > q else 0
            t = 0
        else:
            p = int(rng.integers(0, glen - L))
        s = bytearray(seqs[t][p:p + L])
        for j in range(len(s)):
            u = rng.random()
            if u < 0.01:
                s[j] = b"ACGT"[int(rng.integers(4))]
            elif u < 0.012:
                s[j] = ord("N")
        s = bytes(s)
        if rng.random() < 0.5:
            s = s.translate(comp)[::-1]
        if i % 37 == 0:
            s = s.lower()
        out.append(s)
    out.append(b"ACGTACGTACGTACGTACGTAC")          # 22 bases: no k-mer at kLow=7 (Read.hpp:50-52)
    out.append(b"ACGTACGTACGTACGTACGTACG")         # 23 bases: exactly 3 k-mers
    out.append(b"N" * 60)
    out.append(b"A" * 150)
    return out


def write_fastq(path: str, reads, prefix: str = "r", suffix: str = "") -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(reads):
            f.write(b"@%s%d%s\n" % (prefix.encode(), i, suffix.encode()) + s + b"\n+\n" + b"I" * len(s) + b"\n")


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        print(p.stdout)
        raise SystemExit("reference failed: " + " ".join(cmd))
    return p.stdout


def store_index(work: str, dst: str, name: str = "idx") -> None:
    info = open(os.path.join(work, name + "_info.txt")).read().split()
    n = int(info[0])
    rec = 20 if len(info) > 1 and info[1] == "128" else 12
    gz_write(os.path.join(dst, name + ".gz"), open(os.path.join(work, name), "rb").read()[:n * rec])
    nt = int(open(os.path.join(work, name + "_trie.txt")).read().split()[0])
    gz_write(os.path.join(dst, name + "_trie.gz"), open(os.path.join(work, name + "_trie"), "rb").read()[:nt * 12])
    for suf in ("_info.txt", "_trie.txt", "_f.txt"):
        shutil.copy(os.path.join(work, name + suf), os.path.join(dst, name + suf))


def case(name: str, *, seed: int, n_taxa: int, glen: int, n_reads: int, build_flags, identify_runs, binary: str,
         paired: bool = False):
    dst = os.path.join(GOLD, name)
    shutil.rmtree(dst, ignore_errors=True)
    os.makedirs(dst)
    work = tempfile.mkdtemp(prefix="kasa_gold_")
    os.makedirs(os.path.join(work, "tmp"))
    seqs = make_reference(work, seed, n_taxa, glen)
    exe = os.path.join(REF, binary)
    cmds = []
    b = [exe, "build", "-c", "content.txt", "-d", "idx", "-i", "ref.fa", "-m", "4", "-n", "1", "-t", "tmp/"] + build_flags
    run(b, work)
    cmds.append(" ".join([binary] + b[1:]))
    store_index(work, dst)
    gz_write(os.path.join(dst, "ref.fa.gz"), open(os.path.join(work, "ref.fa"), "rb").read())
    shutil.copy(os.path.join(work, "content.txt"), os.path.join(dst, "content.txt"))
    if n_reads:
        reads = make_reads(seqs, seed + 1, n_reads, glen)
        if paired:
            comp = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")
            mates = make_reads(seqs, seed + 2, n_reads, glen)
            write_fastq(os.path.join(work, "reads_1.fq"), reads, suffix="/1")
            write_fastq(os.path.join(work, "reads_2.fq"), [m.translate(comp)[::-1] for m in mates], suffix="/2")
            for f in ("reads_1.fq", "reads_2.fq"):
                gz_write(os.path.join(dst, f + ".gz"), open(os.path.join(work, f), "rb").read())
        else:
            write_fastq(os.path.join(work, "reads.fq"), reads)
            gz_write(os.path.join(dst, "reads.fq.gz"), open(os.path.join(work, "reads.fq"), "rb").read())
    for tag, flags in identify_runs:
        inp = ["-1", "reads_1.fq", "-2", "reads_2.fq"] if paired else ["-i", "reads.fq"]
        c = [exe, "identify", "-c", "content.txt", "-d", "idx"] + inp + ["-p", tag + ".csv", "-q", tag + ".jsonl",
             "--jsonl", "-r", "-m", "4", "-n", "1", "-t", "tmp/", "-b", "100000", "-v"] + flags
        out = run(c, work)
        cmds.append(" ".join([binary] + c[1:]))
        shutil.copy(os.path.join(work, tag + ".csv"), os.path.join(dst, tag + ".csv"))
        gz_write(os.path.join(dst, tag + ".jsonl.gz"), open(os.path.join(work, tag + ".jsonl"), "rb").read())
        with open(os.path.join(dst, tag + ".stdout.txt"), "w") as f:
            f.write("\n".join(l for l in out.splitlines() if "Number of" in l) + "\n")
    with open(os.path.join(dst, "cmd.txt"), "w") as f:
        f.write("\n".join(cmds) + "\n")
    shutil.rmtree(work, ignore_errors=True)
    print("golden case", name, "done")


def main():
    os.makedirs(GOLD, exist_ok=True)
    case("k12", seed=11, n_taxa=6, glen=8000, n_reads=500, build_flags=[], binary="kASA_scalar",
         identify_runs=[("six_12_7", ["--six"]), ("three_11_8", ["-k", "11", "8"]), ("six_12_12", ["--six", "-k", "12", "12"])])
    case("k12_build6", seed=12, n_taxa=3, glen=3000, n_reads=0, build_flags=["--six"], binary="kASA_scalar",
         identify_runs=[])
    case("k25", seed=13, n_taxa=4, glen=4000, n_reads=200, build_flags=["--kH", "25"], binary="kASA_scalar_fix128",
         identify_runs=[("six_25_7", ["--six", "-k", "25", "7"])], paired=True)


if __name__ == "__main__":
    main()
