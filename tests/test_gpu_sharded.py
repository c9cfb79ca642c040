"""GPU parity of the prefix-sharded identify (SURVEY.md §8e mode 2; include/kasa_b200.h kasa_shard_*): n contexts,
each holding one prefix range of the index and its own reads, must reproduce the unsharded results — per-read
scores bit for bit (KASA_SCORE_ORDERED is a batch- and shard-independent order), unique counts exactly, non-unique counts to 1e-12 —
which are themselves pinned on the reference binary in test_gpu_parity.py.  The contexts live in one process on
one GPU here (peers reached through raw device pointers); the cross-process cudaIpc route is exercised by
`bench.py --sharding prefix` under torchrun."""
import numpy as np
import pytest

import golden_util as G
from kasa_b200 import Context, KasaError, Params
from kasa_b200 import capi
from kasa_b200.dist import LocalShardGroup, shard_bounds

pytestmark = pytest.mark.gpu

LONG_READ_RTOL = 5e-6      # 10 kbp reads: ~10^4 float32 adds per cell, against the reference's own flush order


def _reads(case, paired):
    if not paired:
        seqs = [s for _, s in G.fastq(case)]
        return seqs, None
    a, b = G.fastq(case, "reads_1.fq.gz"), G.fastq(case, "reads_2.fq.gz")
    return [(x[1], y[1]) for x, y in zip(a, b)], True


def _flatten(reads, paired):
    if not paired:
        return list(reads), None
    seqs, first = [], [0]
    for x, y in reads:
        seqs += [x, y]
        first.append(len(seqs))
    return seqs, np.array(first, dtype=np.uint32)


def _run(case, K, k_high, k_low, six, n_shards, paired=False):
    _, taxids, _ = G.content(case)
    idx = G.index(case)
    reads, paired = _reads(case, paired)
    kw = dict(highest_k=K, k_high=k_high, k_low=k_low, frames=6 if six else 3, score_mode=capi.SCORE_ORDERED)
    # unsharded
    ref = Context(Params(**kw))
    ref.load_index(idx, taxids)
    seqs, first = _flatten(reads, paired)
    nk, want_scores, _ = ref.identify_batch(seqs, read_first_seq=first)
    want_ca, want_cu, want_total = ref.identify_finish()
    ref.close()
    # sharded: index cut on prefix boundaries, reads dealt to the ranks
    eb, pb = capi.shard_plan(idx, k_low, n_shards)
    assert eb[0] == 0 and eb[-1] == len(idx) and np.all(np.diff(eb.astype(np.int64)) >= 0)
    ctxs = []
    for r in range(n_shards):
        c = Context(Params(**kw))
        c.load_index(idx[int(eb[r]):int(eb[r + 1])], taxids)
        ctxs.append(c)
    grp = LocalShardGroup(ctxs, pb)
    batches, firsts, bounds = [], [], []
    for r in range(n_shards):
        lo, hi = shard_bounds(len(reads), n_shards, r)
        s, f = _flatten(reads[lo:hi], paired)
        batches.append(s)
        firsts.append(f)
        bounds.append((lo, hi))
    for rnd in range(2):                                     # twice: buffers and phases are reusable
        out = grp.identify_batches(batches, firsts if paired else None)
        got = np.concatenate([o[1] for o in out], axis=0)
        assert sum(o[0] for o in out) == nk
        assert np.array_equal(got, want_scores), "per-read scores differ from the unsharded run"
    ca = np.zeros_like(want_ca)
    cu = np.zeros_like(want_cu)
    total = 0
    for c in ctxs:
        a, u, t = c.identify_finish()
        ca += a
        cu += u
        total += t
        c.close()
    assert total == 2 * want_total
    assert np.array_equal(cu, 2 * want_cu)
    # counts are sums of 1 / n in fixed point; the unsharded run may take level k_low + 1 in 2^-24 units (dense scoring kernel of
    # presorted batches), the sharded one in 2^-32: both far inside the 1e-6 the oracle comparison asks for
    np.testing.assert_allclose(ca, 2 * want_ca, rtol=1e-7)


def test_sharded_k12_three_shards():
    _run("k12", 12, 12, 7, True, 3)


def test_sharded_k12_low_k_boundaries_respect_short_prefixes():
    # k_low = 3 < 6 letters: shard boundaries must fall on 3-letter prefixes
    _run("k12", 12, 12, 3, False, 4)


def test_sharded_k25_paired_two_shards():
    _run("k25", 25, 25, 7, True, 2, paired=True)


def test_sharded_single_shard_equals_plain_identify():
    _run("k12", 12, 11, 8, False, 1)


def test_sharded_protocol_errors():
    _, taxids, _ = G.content("k12")
    idx = G.index("k12")
    c = Context(Params(score_mode=capi.SCORE_EXACT))
    c.load_index(idx, taxids)
    with pytest.raises(KasaError):                           # exact mode replays one global order
        c.shard_init(0, 2, np.array([0, 1 << 29, 1 << 30], dtype=np.uint32))
    c.close()
    c = Context(Params())
    c.load_index(idx, taxids)
    with pytest.raises(KasaError):                           # not spanning [0, 2^30]
        c.shard_init(0, 2, np.array([0, 5, 9], dtype=np.uint32))
    mem = c.shard_init(0, 1, np.array([0, 1 << 30], dtype=np.uint32))
    with pytest.raises(KasaError):                           # sort before connect
        c.shard_sort([b"ACGT" * 30])
    c.shard_connect([mem])
    with pytest.raises(KasaError):                           # score before lookup
        c.shard_score([capi.CShardRuns()], 1)
    c.close()


# ----------------------------------------------------------------------------------------- sharded build (C4)
@pytest.mark.parametrize("case,K,six,n_shards", [("k12", 12, False, 3), ("k12_build6", 12, True, 2), ("k25", 25, False, 4)])
def test_sharded_build_concatenates_to_the_reference_files(case, K, six, n_shards):
    """Every rank pushes part of the reference strings; the shards written in rank order must be byte for byte the index
    and `_trie` the reference binary wrote, and the per-shard frequency tables must add up to its `_f.txt`."""
    from kasa_b200.dist import local_sharded_build_finish
    from oracle import oracle as O
    strings = O.fasta_build_strings(G.gz(case, "ref.fa.gz"), G.acc2tax(case), K, six)
    _, taxids, _ = G.content(case)
    ctxs = [Context(Params(highest_k=K, k_high=K, k_low=7, max_batch_kmers=1 << 20)) for _ in range(n_shards)]
    for i, (s, t) in enumerate(strings):                     # deal the strings (so that every taxon is on several ranks)
        ctxs[i % n_shards].build_push([s], [t])
    for c in ctxs:                                           # duplicates across ranks must be removed by the owner
        c.build_push([strings[0][0]], [strings[0][1]])
    pb, shards = local_sharded_build_finish(ctxs, 7)
    rec = np.concatenate([r for r, _ in shards])
    trie = np.concatenate([t for _, t in shards])
    assert rec.tobytes() == G.index(case).tobytes()
    assert trie.tobytes() == G.trie(case).tobytes()
    freq = sum(c.build_frequency(taxids) for c in ctxs)
    assert np.array_equal(freq, G.frequency(case))
    sizes = [len(r) for r, _ in shards]
    assert max(sizes) < 2.0 * len(rec) / n_shards + 4096     # about equal shards
    # the shards are directly usable as a prefix-sharded index
    for c in ctxs:
        c.index_from_build(taxids)
    assert sum(c.index_size for c in ctxs) == len(rec)
    for c in ctxs:
        c.close()


# ----------------------------------------------------------------------------------------- long reads (C5 shape)
def test_long_reads_sharded_and_plain_match_the_oracle():
    """BASELINE config C5 in miniature: 10 kbp reads with 10 % errors (substitutions and indels) against a prefix-sharded
    index.  19 960 k-mers per read."""
    from kasa_b200 import synth
    from oracle import oracle as O
    n_taxa, glen, n_reads, rlen = 6, 40_000, 12, 10_000
    rng = np.random.default_rng(77)
    fasta = b"".join(b">" + synth.accession(t).encode() + b"\n" + synth.genome_ascii(11, t, glen) + b"\n" for t in range(n_taxa))
    acc2tax = {synth.accession(t): 1000 + t for t in range(n_taxa)}
    taxids = np.array([0] + [1000 + t for t in range(n_taxa)], dtype=np.uint32)
    index, _ = O.build_index(fasta, acc2tax, 12, six=False)
    reads = []
    for i in range(n_reads):
        g = synth.genome_ascii(11, int(rng.integers(n_taxa)), glen)
        o = int(rng.integers(0, glen - rlen))
        out = bytearray()
        for ch in g[o:o + rlen]:
            u = rng.random()
            if u < 0.05:
                out.append(b"ACGT"[int(rng.integers(4))])          # substitution
            elif u < 0.075:
                continue                                            # deletion
            elif u < 0.10:
                out.append(ch); out.append(b"ACGT"[int(rng.integers(4))])   # insertion
            else:
                out.append(ch)
        reads.append(bytes(out))
    ora = O.identify_reads([[r] for r in reads], index, 12, 12, 7, True, taxids)
    kw = dict(frames=6, max_batch_kmers=1 << 19, max_batch_seqs=64, max_batch_bases=1 << 18)
    plain = Context(Params(**kw))
    plain.load_index(index, taxids)
    nk, scores, _ = plain.identify_batch(reads)
    ca, cu, _ = plain.identify_finish()
    plain.close()
    assert nk == ora.n_kmers
    assert np.array_equal(cu, ora.count_unique)
    np.testing.assert_allclose(ca, ora.count_all, rtol=1e-6, atol=1e-9)
    # ~20 000 k-mers per read: ordered rank slice by rank slice (score.cu), bit-identical to the CPU restatement of that order
    mine = O.identify_reads_ordered([[r] for r in reads], index, 12, 12, 7, True, taxids)
    assert np.array_equal(scores, mine.scores)
    np.testing.assert_allclose(scores, ora.scores, rtol=LONG_READ_RTOL)
    eb, pb = capi.shard_plan(index, 7, 2)
    ctxs = []
    for r in range(2):
        c = Context(Params(**kw))
        c.load_index(index[int(eb[r]):int(eb[r + 1])], taxids)
        ctxs.append(c)
    out = LocalShardGroup(ctxs, pb).identify_batches([reads[:5], reads[5:]])
    assert np.array_equal(np.concatenate([o[1] for o in out]), scores)
    for c in ctxs:
        c.close()
