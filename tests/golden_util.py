"""Loaders for tests/golden/* (outputs of the reference itself, see oracle/make_golden.py)."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REC64 = np.dtype([("lo", "<u8"), ("tax", "<u4")])
REC128 = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("tax", "<u4")])
TRIE = np.dtype([("count", "<u8"), ("prefix", "<u4")])


def gz(case: str, name: str) -> bytes:
    return gzip.open(os.path.join(GOLD, case, name), "rb").read()


def text(case: str, name: str) -> str:
    return open(os.path.join(GOLD, case, name)).read()


def index(case: str) -> np.ndarray:
    info = text(case, "idx_info.txt").split()
    dt = REC128 if len(info) > 1 and info[1] == "128" else REC64
    a = np.frombuffer(gz(case, "idx.gz"), dtype=dt)
    assert len(a) == int(info[0])
    return a


def trie(case: str) -> np.ndarray:
    a = np.frombuffer(gz(case, "idx_trie.gz"), dtype=TRIE)
    assert len(a) == int(text(case, "idx_trie.txt").split()[0])
    return a


def content(case: str):
    names, taxids, accs = ["non_unique"], [0], [None]
    for line in text(case, "content.txt").splitlines():
        if line:
            p = line.split("\t")
            names.append(p[0]); taxids.append(int(p[1])); accs.append(p[3])
    return names, np.array(taxids, dtype=np.uint32), accs


def acc2tax(case: str) -> dict:
    _, taxids, accs = content(case)
    return {a: int(t) for a, t in zip(accs[1:], taxids[1:])}


def fasta(case: str):
    """[(sequence bytes, taxid)] in file order."""
    _, taxids, accs = content(case)
    acc2tax = {a: int(t) for a, t in zip(accs[1:], taxids[1:])}
    out = []
    for line in gz(case, "ref.fa.gz").split(b"\n"):
        if line.startswith(b">"):
            out.append([b"", acc2tax[line[1:].split(b" ")[0].decode()]])
        elif line:
            out[-1][0] += line
    return [(s, t) for s, t in out]


def fastq(case: str, name: str = "reads.fq.gz"):
    lines = gz(case, name).split(b"\n")
    return [(lines[i][1:], lines[i + 1]) for i in range(0, len(lines) - 1, 4)]


def frequency(case: str) -> np.ndarray:
    rows = [l.split("\t")[1:] for l in text(case, "idx_f.txt").splitlines() if l]
    return np.array(rows, dtype=np.uint64)


def profile(case: str, tag: str):
    """-> (header list, {taxid: row of floats as printed strings})."""
    lines = text(case, tag + ".csv").splitlines()
    head = lines[0].split(",")
    rows = {}
    for l in lines[1:]:
        p = l.split(",")
        rows[int(p[0])] = p[2:]
    return head, rows


def per_read(case: str, tag: str):
    """-> list over reads of dict(taxid -> (kmer score f32, relative score f64, error f32)), plus names/lengths."""
    out = []
    for line in gz(case, tag + ".jsonl.gz").decode().splitlines():
        j = json.loads(line)
        hits = {}
        for h in j["Top hits"] + j["Further hits"]:
            hits[int(h["tax ID"])] = (np.float32(h["k-mer Score"]), float(h["Relative Score"]), np.float32(h["Error"]))
        out.append({"n": j["Read number"], "name": j["Specifier from input file"], "len": j["Length"], "hits": hits,
                    "top": [int(h["tax ID"]) for h in j["Top hits"]]})
    return out


def n_kmers(case: str, tag: str) -> int:
    for l in text(case, tag + ".stdout.txt").splitlines():
        if "Number of k-mers in input" in l:
            return int(l.split(":")[2].split()[0])
    raise KeyError
