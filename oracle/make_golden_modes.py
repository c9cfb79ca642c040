#!/usr/bin/env python
"""Golden outputs of the reference's index-maintenance modes (tests/golden/k12/modes_sha256.txt).

TEST INFRASTRUCTURE.  Run in the authoring container after `make -C oracle ref`:
    python oracle/make_golden_modes.py
The reference binary (`kASA_scalar`) runs on the committed k12 golden inputs:
    trie          -d idx                                   -> idx_trie, idx_trie.txt          (main.cpp:1422-1458)
    getFrequency  -c content.txt -d idx                    -> idx_f.txt                       (main.cpp:1336-1362)
    delete        -c content.txt -d idx -o del -l del.dmp  -> del, del_info.txt, del_trie, del_trie.txt, del_f.txt
                                                                                              (Update.hpp:28-94)
    build         twice, on the first and the second half of ref.fa's records -> a, b
    merge         -c1 content.txt -c2 content.txt -co m_content.txt --firstIndex a --secondIndex b -o m
                                                           -> m, m_info.txt, m_trie, m_trie.txt, m_f.txt, m_content.txt
                                                                                              (Read.hpp:3180-3243)
    update        -c content.txt -d a -i b.fa -o u         -> u, u_info.txt, u_trie, u_trie.txt, u_f.txt (Update.hpp:99-)
    half          -c content.txt -d idx -o idxh            -> idxh (6-byte records), idxh_info.txt, idxh_trie, idxh_trie.txt
                                                                                              (Shrink.hpp:78-143)
    redundancy    -c content.txt -d idx -v                 -> stdout: histogram of taxa per k-mer (Shrink.hpp:33-70)
Stored: one line per produced file, `<case>/<file> <bytes that count> <sha256>` (index and trie files are hashed over their
records, without the STXXL block padding), the deleted taxids and how ref.fa was cut.  tests/test_host_cli.py runs
kasa_b200_cli with the same commands and compares the same hashes.
"""
from __future__ import annotations

import gzip
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "kASA_scalar")
GOLD = os.path.join(ROOT, "tests", "golden", "k12")
DELETED = [1003, 1005]          # taxids of content.txt that `delete` removes


def run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        print(p.stdout)
        raise SystemExit("reference failed: " + " ".join(cmd))
    return p.stdout


def split_fasta(text):
    """records of a FASTA text; first half / second half by record count"""
    recs = [">" + r for r in text.split(">") if r]
    h = len(recs) // 2
    return "".join(recs[:h]), "".join(recs[h:])


def record_hash(work, name, rec_bytes):
    raw = open(os.path.join(work, name), "rb").read()
    info = os.path.join(work, name + "_info.txt")
    if os.path.exists(info):
        n = int(open(info).read().split()[0])
    else:
        # `merge` writes no _info.txt (Read.hpp:3180-3243): the records end where the zero padding of the last STXXL block begins
        n = len(raw) // rec_bytes
        while n and raw[(n - 1) * rec_bytes:n * rec_bytes] == bytes(rec_bytes):
            n -= 1
    data = raw[:n * rec_bytes]
    return n * rec_bytes, hashlib.sha256(data).hexdigest()


def trie_hash(work, name):
    n = int(open(os.path.join(work, name + "_trie.txt")).read().split()[0])
    data = open(os.path.join(work, name + "_trie"), "rb").read()[:n * 12]
    return n * 12, hashlib.sha256(data).hexdigest()


def text_hash(work, name):
    data = open(os.path.join(work, name), "rb").read()
    return len(data), hashlib.sha256(data).hexdigest()


def main():
    work = tempfile.mkdtemp(prefix="kasa_modes_")
    os.makedirs(os.path.join(work, "tmp"))
    for f in ("content.txt", "idx_info.txt", "idx_f.txt", "idx_trie.txt"):
        shutil.copy(os.path.join(GOLD, f), work)
    for f in ("idx", "idx_trie", "ref.fa"):
        open(os.path.join(work, f), "wb").write(gzip.open(os.path.join(GOLD, f + ".gz")).read())
    out = []
    # ---- trie / getFrequency on copies, so that the originals stay for comparison
    for suf in ("", "_info.txt"):
        shutil.copy(os.path.join(work, "idx" + suf), os.path.join(work, "t" + suf))
    run([REF, "trie", "-d", "t", "-m", "4", "-t", "tmp/"], work)
    out.append(("trie/t_trie",) + trie_hash(work, "t"))
    for suf in ("", "_info.txt"):
        shutil.copy(os.path.join(work, "idx" + suf), os.path.join(work, "f" + suf))
    run([REF, "getFrequency", "-c", "content.txt", "-d", "f", "-m", "4", "-t", "tmp/"], work)
    out.append(("getFrequency/f_f.txt",) + text_hash(work, "f_f.txt"))
    # ---- delete
    open(os.path.join(work, "del.dmp"), "w").write("".join("%d\t|\n" % t for t in DELETED))
    run([REF, "delete", "-c", "content.txt", "-d", "idx", "-o", "del", "-l", "del.dmp", "-m", "4", "-t", "tmp/"], work)
    out.append(("delete/del",) + record_hash(work, "del", 12))
    out.append(("delete/del_trie",) + trie_hash(work, "del"))
    out.append(("delete/del_f.txt",) + text_hash(work, "del_f.txt"))
    # ---- merge of two builds over the halves of ref.fa
    a, b = split_fasta(open(os.path.join(work, "ref.fa")).read())
    open(os.path.join(work, "a.fa"), "w").write(a)
    open(os.path.join(work, "b.fa"), "w").write(b)
    for name in ("a", "b"):
        run([REF, "build", "-c", "content.txt", "-d", name, "-i", name + ".fa", "-m", "4", "-n", "1", "-t", "tmp/"], work)
        out.append(("merge/" + name,) + record_hash(work, name, 12))
    # (with a single `-c` the reference stops with "Content file not found": MergeTwoIndices wants the merged content file, Read.hpp:3184)
    run([REF, "merge", "-c1", "content.txt", "-c2", "content.txt", "-co", "m_content.txt", "--firstIndex", "a", "--secondIndex", "b", "-o", "m",
         "-m", "4", "-t", "tmp/"], work)
    out.append(("merge/m_content.txt",) + text_hash(work, "m_content.txt"))
    out.append(("merge/m",) + record_hash(work, "m", 12))
    out.append(("merge/m_trie",) + trie_hash(work, "m"))
    out.append(("merge/m_f.txt",) + text_hash(work, "m_f.txt"))
    # ---- update: the first half's index + the second half's sequences, content file given
    run([REF, "update", "-c", "content.txt", "-d", "a", "-i", "b.fa", "-o", "u", "-m", "4", "-n", "1", "-t", "tmp/"], work)
    out.append(("update/u",) + record_hash(work, "u", 12))
    out.append(("update/u_trie",) + trie_hash(work, "u"))
    out.append(("update/u_f.txt",) + text_hash(work, "u_f.txt"))
    # ---- identify_multiple: the two halves of reads.fq in one folder; the names of the output files are what is recorded (their
    # contents equal those of `identify` per file, which tests/golden/k12/six_12_7.* pin)
    fq = gzip.open(os.path.join(GOLD, "reads.fq.gz"), "rt").read().splitlines(keepends=True)
    half = (len(fq) // 8) * 4
    os.makedirs(os.path.join(work, "in"))
    open(os.path.join(work, "in", "first.part.fq"), "w").write("".join(fq[:half]))
    open(os.path.join(work, "in", "second.fq"), "w").write("".join(fq[half:]))
    os.makedirs(os.path.join(work, "res"))
    run([REF, "identify_multiple", "-c", "content.txt", "-d", "idx", "-i", "in/", "-q", "res/q_", "-p", "res/p_", "--jsonl", "-r", "-m", "8", "-n", "2",
         "-t", "tmp/", "--six"], work)
    names = sorted(os.listdir(os.path.join(work, "res")))
    for nme in names:
        out.append(("identify_multiple/" + nme,) + text_hash(work, os.path.join("res", nme)))
    for part in ("first.part.fq", "second.fq"):
        stem = part.rsplit(".", 1)[0]
        run([REF, "identify", "-c", "content.txt", "-d", "idx", "-i", "in/" + part, "-q", "res/single_" + stem + ".jsonl", "--jsonl", "-r", "-m", "8",
             "-n", "1", "-t", "tmp/", "--six"], work)
        same = open(os.path.join(work, "res", "single_" + stem + ".jsonl")).read() == open(os.path.join(work, "res", "q_" + stem + ".jsonl")).read()
        print("identify_multiple == identify for", part, ":", same)
    # ---- half: the 6-byte index + its prefix table; identify with it prints what identify with the full index prints
    run([REF, "half", "-c", "content.txt", "-d", "idx", "-o", "idxh", "-m", "4", "-t", "tmp/"], work)
    out.append(("half/idxh",) + record_hash(work, "idxh", 6))
    out.append(("half/idxh_trie",) + trie_hash(work, "idxh"))
    # the same on the index cut after a prefix run of four records: the reference closes the last run one short (Shrink.hpp:131-136)
    CUT = 45491
    open(os.path.join(work, "cut"), "wb").write(open(os.path.join(work, "idx"), "rb").read()[:CUT * 12])
    open(os.path.join(work, "cut_info.txt"), "w").write(str(CUT))
    shutil.copy(os.path.join(work, "idx_f.txt"), os.path.join(work, "cut_f.txt"))
    run([REF, "half", "-c", "content.txt", "-d", "cut", "-o", "cuth", "-m", "4", "-t", "tmp/"], work)
    out.append(("half/cuth_of_the_first_%d_records" % CUT,) + record_hash(work, "cuth", 6))
    out.append(("half/cuth_trie",) + trie_hash(work, "cuth"))
    open(os.path.join(work, "reads.fq"), "w").write("".join(fq))
    res = {}
    for name in ("idx", "idxh"):
        run([REF, "identify", "-c", "content.txt", "-d", name, "-i", "reads.fq", "-q", name + ".jsonl", "-p", name + ".csv", "--jsonl", "-r", "-m", "4",
             "-n", "1", "-t", "tmp/", "--six"], work)
        res[name] = (open(os.path.join(work, name + ".jsonl")).read(), open(os.path.join(work, name + ".csv")).read())
    print("identify with the halved index == identify with the full index:", res["idx"] == res["idxh"])
    assert res["idx"] == res["idxh"]
    out.append(("half/identify_equals_full_index", 1, hashlib.sha256(res["idxh"][0].encode()).hexdigest()))
    # ---- redundancy: the histogram of taxa per k-mer and the advice the binary prints
    red = run([REF, "redundancy", "-c", "content.txt", "-d", "idx", "-m", "4", "-t", "tmp/", "-v"], work)
    keep = [l for l in red.splitlines() if not l.startswith("OUT: kASA version") and "STXXL" not in l and "redundancy -c" not in l and not l.startswith("OUT: Time")]
    open(os.path.join(work, "redundancy.txt"), "w").write("\n".join(keep) + "\n")
    os.makedirs(os.path.join(GOLD, "modes"), exist_ok=True)
    for name in ("del_f.txt", "m_f.txt", "del.dmp", "redundancy.txt"):          # the small text files are kept as they are
        shutil.copy(os.path.join(work, name), os.path.join(GOLD, "modes", name))
    with open(os.path.join(GOLD, "modes_sha256.txt"), "w") as f:
        f.write("# written by oracle/make_golden_modes.py from kASA_scalar; deleted taxids: %s; ref.fa cut after half of its records\n"
                % ",".join(str(t) for t in DELETED))
        for name, nbytes, sha in out:
            f.write("%s %d %s\n" % (name, nbytes, sha))
            print(name, nbytes, sha)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
