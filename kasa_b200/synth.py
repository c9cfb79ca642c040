"""Seeded synthetic references and reads (SURVEY.md §8(d) configurations C1..C5).

Everything is a pure function of (seed, taxon, position) through splitmix64, so
the CUDA generator in csrc/synth.cu (used by bench.py to create the 20 GB-index
workload on device) and this numpy version produce identical data.  Base codes
follow kASA's own 3-bit mapping (kASA.hpp:70-75): A=0 C=1 T=2 G=3.
"""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
CODE2CHAR = np.frombuffer(b"ACTG", dtype=np.uint8)
COMP_CHAR = {65: 84, 67: 71, 84: 65, 71: 67}

# stream tags mixed into the key so the different random draws are independent
TAG_GENOME, TAG_TAXON, TAG_OFFSET, TAG_STRAND, TAG_SUBST, TAG_SUBST_TO, TAG_N = (
    np.uint64(x) for x in (0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x07))


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64 finaliser; identical to csrc/common.cuh::splitmix64."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _key(seed: int, tag: np.uint64, a, b) -> np.ndarray:
    """hash(seed, tag, a, b) -> u64; a < 2^24 (taxon / misc), b < 2^40 (position / read)."""
    with np.errstate(over="ignore"):
        s = splitmix64(np.uint64(seed) * np.uint64(0x100) + tag)
        return splitmix64(s ^ ((np.asarray(a, dtype=np.uint64) << np.uint64(40)) | np.asarray(b, dtype=np.uint64)))


def genome_codes(seed: int, taxon: int, start: int, length: int) -> np.ndarray:
    """2-bit codes (kASA order A0 C1 T2 G3) of genome `taxon`, positions [start, start+length)."""
    pos = np.arange(start, start + length, dtype=np.uint64)
    return (_key(seed, TAG_GENOME, taxon, pos) >> np.uint64(62)).astype(np.uint8)


def genome_ascii(seed: int, taxon: int, length: int) -> bytes:
    return CODE2CHAR[genome_codes(seed, taxon, 0, length)].tobytes()


def read_params(seed: int, read_ids: np.ndarray, n_taxa: int, genome_len: int, read_len: int):
    rid = np.asarray(read_ids, dtype=np.uint64)
    taxon = (_key(seed, TAG_TAXON, 0, rid) % np.uint64(n_taxa)).astype(np.int64)
    offset = (_key(seed, TAG_OFFSET, 0, rid) % np.uint64(genome_len - read_len + 1)).astype(np.int64)
    rc = (_key(seed, TAG_STRAND, 0, rid) & np.uint64(1)).astype(bool)
    return taxon, offset, rc


def reads_ascii(genome_seed: int, read_seed: int, first_read: int, n_reads: int, n_taxa: int,
                genome_len: int, read_len: int = 150, subst_ppm: int = 10000, n_ppm: int = 1000) -> np.ndarray:
    """[n_reads, read_len] uint8 ASCII reads: taxon/offset uniform, 50 % reverse-complemented,
    `subst_ppm` substitutions and `n_ppm` 'N' per million bases (C1: 1 % / 0.1 %)."""
    rid = np.arange(first_read, first_read + n_reads, dtype=np.uint64)
    taxon, offset, rc = read_params(read_seed, rid, n_taxa, genome_len, read_len)
    j = np.arange(read_len, dtype=np.int64)[None, :]
    # position in the genome of read base j: forward reads walk up, rc reads walk down
    gpos = np.where(rc[:, None], offset[:, None] + (read_len - 1 - j), offset[:, None] + j).astype(np.uint64)
    code = (_key(genome_seed, TAG_GENOME, taxon[:, None].astype(np.uint64), gpos) >> np.uint64(62)).astype(np.uint8)
    code = np.where(rc[:, None], code ^ 2, code)          # complement: A0<->T2, C1<->G3
    bkey = rid[:, None] * np.uint64(65536) + j.astype(np.uint64)
    hs = _key(read_seed, TAG_SUBST, 0, bkey) % np.uint64(1000000)
    sub = hs < np.uint64(subst_ppm)
    to = (_key(read_seed, TAG_SUBST_TO, 0, bkey) % np.uint64(3)).astype(np.uint8) + 1
    code = np.where(sub, (code + to) & 3, code).astype(np.uint8)
    out = CODE2CHAR[code]
    isn = (_key(read_seed, TAG_N, 0, bkey) % np.uint64(1000000)) < np.uint64(n_ppm)
    out = np.where(isn, np.uint8(ord("N")), out)
    return np.ascontiguousarray(out, dtype=np.uint8)


def accession(i: int) -> str:
    return "SYN%04d.1" % i


def write_reference(path_fa: str, path_content: str, seed: int, n_taxa: int, genome_len: int,
                    n_runs: int = 0, cols: int = 80, shared: list[tuple[int, int, int, int]] | None = None) -> None:
    """FASTA (80 cols, one sequence per taxon) + kASA content file (README.md:192-203), sorted by taxid.
    `shared` = [(src_taxon, src_pos, dst_taxon, length)] copies regions between genomes so that k-mers are
    shared by several taxa (non-unique matches); `n_runs` > 0 sprinkles runs of 'N' into each genome."""
    seqs = []
    for t in range(n_taxa):
        seqs.append(bytearray(genome_ascii(seed, t, genome_len)))
    for (st, sp, dt, ln) in (shared or []):
        seqs[dt][sp:sp + ln] = seqs[st][sp:sp + ln]
    for t in range(n_taxa):
        for r in range(n_runs):
            h = int(_key(seed, TAG_N, t, np.array([r], dtype=np.uint64))[0])
            p = h % max(1, genome_len - 8)
            run = 1 + (h >> 40) % 5
            seqs[t][p:p + run] = b"N" * run
    with open(path_fa, "wb") as f:
        for t in range(n_taxa):
            f.write(b">" + accession(t).encode() + b" synthetic taxon %d\n" % t)
            s = bytes(seqs[t])
            for a in range(0, len(s), cols):
                f.write(s[a:a + cols] + b"\n")
    with open(path_content, "w") as f:
        for t in range(n_taxa):
            f.write("Synthetic taxon %d\t%d\t%d\t%s\n" % (t, 1000 + t, 1000 + t, accession(t)))


def write_fastq(path: str, reads: np.ndarray, first_id: int = 0, prefix: str = "r") -> None:
    n, L = reads.shape
    qual = b"I" * L
    with open(path, "wb") as f:
        for i in range(n):
            f.write(b"@%s%d\n" % (prefix.encode(), first_id + i) + reads[i].tobytes() + b"\n+\n" + qual + b"\n")
