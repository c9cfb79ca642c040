"""Pin the CPU restatement (oracle/kasa_oracle.c) against outputs of the reference itself.

tests/golden/* were produced by oracle/_ref/kASA_scalar (and kASA_scalar_fix128 for k=25) —
the unmodified reference sources — via oracle/make_golden.py.  Comparison rules follow
SURVEY.md §8(c): index/trie/frequency byte-exact, unique counts exact, non-unique counts at
print precision, per-read k-mer scores bit-exact fp32.
"""
import numpy as np
import pytest

import golden_util as G
from oracle import oracle as O


def _idx_bytes(a):
    return np.ascontiguousarray(a).tobytes()


@pytest.mark.parametrize("case,K,six", [("k12", 12, False), ("k12_build6", 12, True), ("k25", 25, False)])
def test_build_matches_reference(case, K, six):
    idx, trie = O.build_index(G.gz(case, "ref.fa.gz"), G.acc2tax(case), K, six=six)
    ref = G.index(case)
    assert len(idx) == len(ref)
    assert _idx_bytes(idx) == _idx_bytes(ref)
    assert _idx_bytes(trie) == _idx_bytes(G.trie(case))
    _, taxids, _ = G.content(case)
    assert np.array_equal(O.frequency(idx, K, taxids), G.frequency(case))


def _reads(case, paired):
    if paired:
        a, b = G.fastq(case, "reads_1.fq.gz"), G.fastq(case, "reads_2.fq.gz")
        return [[x[1], y[1]] for x, y in zip(a, b)], [x[0] + b" " + y[0] + b" " for x, y in zip(a, b)]
    a = G.fastq(case)
    return [[x[1]] for x in a], [x[0] + b" " for x in a]


def _check_identify(case, tag, K, k_high, k_low, six, paired=False, one=False):
    _, taxids, _ = G.content(case)
    idx = G.index(case)
    reads, names = _reads(case, paired)
    res = O.identify_reads(reads, idx, K, k_high, k_low, six, taxids, one=one)
    assert res.n_kmers == G.n_kmers(case, tag)
    head, rows = G.profile(case, tag)
    numK = k_high - k_low + 1
    for t in range(1, len(taxids)):
        row = rows.get(int(taxids[t]))
        if row is None:       # printed only if non-unique count at kLow > 0 (Compare.hpp:3587)
            assert res.count_all[numK - 1, t] == 0
            continue
        uniq = [int(x) for x in row[0:numK]]
        assert uniq == [int(x) for x in res.count_unique[:, t]]
        nonu = [float(x) for x in row[2 * numK:3 * numK]]
        np.testing.assert_allclose(res.count_all[:, t], nonu, rtol=2e-5)   # 6 printed digits
    # per-read: same taxa with score > 0, k-mer score bit-exact (fp32 in the reference's flush order)
    pr = G.per_read(case, tag)
    assert len(pr) == len(reads)
    freq = G.frequency(case)
    tax2idx = {int(t): i for i, t in enumerate(taxids)}
    frames = 1 if one else (6 if six else 3)
    n_cells = 0
    for r, rec in enumerate(pr):
        assert rec["name"].encode() == names[r]
        length = sum(len(m) + 1 for m in reads[r])       # +1: getChunk counts the newline
        assert rec["len"] == length
        mine = {int(taxids[t]): res.scores[r, t] for t in range(1, len(taxids)) if res.scores[r, t] > 0}
        assert set(mine) == set(rec["hits"])
        best = O.best_score(length, k_high, k_low, frames)
        for tax, (score, rel, err) in rec["hits"].items():
            n_cells += 1
            assert mine[tax] == score, (r, tax, mine[tax], score)
            assert O.relative_score(mine[tax], int(freq[tax2idx[tax]][K - k_high]), length, K) == pytest.approx(rel, rel=1e-12)
            assert np.float32((best - mine[tax]) / best) == err
    assert n_cells > len(reads) // 2


def test_identify_k12_six():
    _check_identify("k12", "six_12_7", 12, 12, 7, True)


def test_identify_k12_three_frames_k11_8():
    _check_identify("k12", "three_11_8", 12, 11, 8, False)


def test_identify_k12_single_level():
    _check_identify("k12", "six_12_12", 12, 12, 12, True)


def test_build_one_frame_matches_reference():
    import hashlib
    idx, trie = O.build_index(G.gz("k12", "ref.fa.gz"), G.acc2tax("k12"), 12, one=True)       # build --one (Read.hpp:2143-2200)
    n, sha, nt, tsha = G.text("k12", "one_build_sha256.txt").split()
    assert len(idx) == int(n) and hashlib.sha256(idx.tobytes()).hexdigest() == sha
    assert len(trie) == int(nt) and hashlib.sha256(trie.tobytes()).hexdigest() == tsha


def test_identify_k12_one_frame():
    _check_identify("k12", "one_12_7", 12, 12, 7, False, one=True)      # --one (Read.hpp:223-260)


def test_identify_k25_paired_six():
    _check_identify("k25", "six_25_7", 25, 25, 7, True, paired=True)


def test_codon_table_properties():
    t = O.default_codon_table()
    code = {"A": 0, "C": 1, "T": 2, "G": 3, "X": 4, "Z": 5}
    def aa(c):
        return chr(t[code[c[0]] << 6 | code[c[1]] << 3 | code[c[2]]])
    assert aa("ATG") == "M" and aa("TGG") == "W" and aa("TAA") == "[" and aa("TAG") == "[" and aa("TGA") == "]"
    assert aa("AAX") == "^" and aa("XXX") == "^" and aa("AXZ") == "_" and aa("ZAA") == "_"
    assert aa("CTG") == "L" and aa("AGA") == "R" and aa("GGG") == "G"


def test_kmer_count_edge_cases():
    L = O.lib()
    assert L.ko_kmer_count(37, 12) == 0      # exactly 3K+1 yields nothing (Read.hpp:50-52)
    assert L.ko_kmer_count(38, 12) == 3
    assert L.ko_kmer_count(165, 12) == 130
    assert L.ko_kmer_count(0, 12) == 0


def test_half_index_restatement_equals_the_reference_files():
    """oracle.half_index (Shrink::putHalfInTrie, Shrink.hpp:78-143) against the files `kASA half` wrote from the k12 golden index
    (tests/golden/k12/modes_sha256.txt, oracle/make_golden_modes.py): 6-byte records and prefix table byte for byte."""
    import hashlib
    want = {}
    for line in G.text("k12", "modes_sha256.txt").splitlines():
        if line.startswith("half/"):
            name, nbytes, sha = line.split()
            want[name] = (int(nbytes), sha)
    _, taxids, _ = G.content("k12")
    rec, trie = O.half_index(G.index("k12"), taxids)
    assert (rec.nbytes, hashlib.sha256(rec.tobytes()).hexdigest()) == want["half/idxh"]
    assert (trie.nbytes, hashlib.sha256(trie.tobytes()).hexdigest()) == want["half/idxh_trie"]
    assert want["half/identify_equals_full_index"][0] == 1      # the reference printed the same JSONL with either index
    # the index cut after a prefix run of four records: the reference closes the last run one short (Shrink.hpp:131-136)
    cut = [int(k.split("_")[-2]) for k in want if k.startswith("half/cuth_of_the_first_")][0]
    rec, trie = O.half_index(G.index("k12")[:cut], taxids)
    assert (rec.nbytes, hashlib.sha256(rec.tobytes()).hexdigest()) == want["half/cuth_of_the_first_%d_records" % cut]
    assert (trie.nbytes, hashlib.sha256(trie.tobytes()).hexdigest()) == want["half/cuth_trie"]
    assert int(trie["count"].sum()) == len(rec) - 1
