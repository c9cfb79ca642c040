"""CPU-side checks: the C ABI library loads and exports what include/kasa_b200.h declares, fails loudly without a
GPU (no CPU fallback), and the host-side helpers behave.  No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

import kasa_b200
from kasa_b200 import capi, synth
from kasa_b200.dist import shard_bounds, shard_first_read

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "kasa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kasa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = kasa_b200.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libkasa_b200.so does not export %s" % n
    assert set(capi.EXPORTS) <= set(names)
    assert lib.kasa_abi_version() == 2


def test_header_is_plain_c_and_ctypes_structs_match_it(tmp_path):
    """Compile include/kasa_b200.h as C (gcc) and compare struct sizes/offsets with the ctypes mirror."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "kasa_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\\n", offsetof(kasa_params, max_batch_hits), offsetof(kasa_batch_result, hit_taxon), offsetof(kasa_batch_result, hit_count),
         offsetof(kasa_batch_result, hit_capacity), offsetof(kasa_batch_result, n_hits), offsetof(kasa_batch_result, hits_location), sizeof(kasa_batch_result));
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(kasa_params), sizeof(kasa_read_batch), sizeof(kasa_batch_result),
         sizeof(kasa_seq_batch), offsetof(kasa_params, max_batch_kmers), offsetof(kasa_params, codon_table),
         offsetof(kasa_read_batch, read_first_seq), offsetof(kasa_batch_result, gpu_ms),
         sizeof(kasa_shard_mem), offsetof(kasa_shard_mem, handle), sizeof(kasa_shard_runs), offsetof(kasa_shard_runs, begin));
  return 0; }''')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [capi.CParams.max_batch_hits.offset, capi.CBatchResult.hit_taxon.offset, capi.CBatchResult.hit_count.offset,
            capi.CBatchResult.hit_capacity.offset, capi.CBatchResult.n_hits.offset, capi.CBatchResult.hits_location.offset,
            ctypes.sizeof(capi.CBatchResult),
            ctypes.sizeof(capi.CParams), ctypes.sizeof(capi.CReadBatch), ctypes.sizeof(capi.CBatchResult),
            ctypes.sizeof(capi.CSeqBatch), capi.CParams.max_batch_kmers.offset, capi.CParams.codon_table.offset,
            capi.CReadBatch.read_first_seq.offset, capi.CBatchResult.gpu_ms.offset,
            ctypes.sizeof(capi.CShardMem), capi.CShardMem.handle.offset, ctypes.sizeof(capi.CShardRuns), capi.CShardRuns.begin.offset]
    assert got == want
    assert capi.REC64.itemsize == 12 and capi.REC128.itemsize == 20 and capi.TRIE.itemsize == 12


def test_no_cpu_fallback():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(kasa_b200.KasaError) as e:
        kasa_b200.Context(kasa_b200.Params())
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_parameter_validation_happens_before_device_use():
    lib = kasa_b200.load_library()
    h = ctypes.c_void_p()
    bad = capi.CParams(0, 13, 12, 7, 3, 1, 1 << 20, 1 << 10, 1 << 20, None)     # highest_k must be 12 or 25
    assert lib.kasa_ctx_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert b"highest_k" in lib.kasa_last_error(None)
    bad = capi.CParams(0, 12, 12, 13, 3, 1, 1 << 20, 1 << 10, 1 << 20, None)    # k_low > k_high
    assert lib.kasa_ctx_create(ctypes.byref(bad), ctypes.byref(h)) == -1
    assert lib.kasa_identify_finish(None, None, None, None) == -1                # null context


def test_synth_is_deterministic_and_reverse_strand_is_the_complement():
    a = synth.reads_ascii(3, 4, 10, 64, 5, 4000, 150, subst_ppm=0, n_ppm=0)
    b = synth.reads_ascii(3, 4, 10, 64, 5, 4000, 150, subst_ppm=0, n_ppm=0)
    assert np.array_equal(a, b)
    taxon, offset, rc = synth.read_params(4, np.arange(10, 74), 5, 4000, 150)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for i in range(64):
        g = synth.genome_ascii(3, int(taxon[i]), 4000)[int(offset[i]):int(offset[i]) + 150]
        want = g.translate(comp)[::-1] if rc[i] else g
        assert bytes(a[i]) == want


def test_read_sharding_covers_everything_once():
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            lo, hi = shard_bounds(1001, world, rank)
            seen += list(range(lo, hi))
        assert seen == list(range(1001))
        firsts = sorted(shard_first_read(s, world, r, 10) for s in range(3) for r in range(world))
        assert firsts == [10 * i for i in range(3 * world)]


def _prefix30(idx, K):
    if K == 12:
        return (idx["lo"] >> np.uint64(30)).astype(np.uint32)
    return (idx["hi"] >> np.uint64(5 * K - 30 - 64)).astype(np.uint32)


@pytest.mark.parametrize("case,K", [("k12", 12), ("k25", 25)])
@pytest.mark.parametrize("k_low", [1, 3, 6, 7])
@pytest.mark.parametrize("n_shards", [1, 2, 5, 16])
def test_shard_plan_cuts_on_prefix_groups(case, K, k_low, n_shards):
    """kasa_shard_plan is host-only: boundaries never split a group of min(6, k_low) leading letters
    (so every level-k taxa set of SURVEY.md §3.1 stays on one shard) and shards are about equal."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as G
    idx = G.index(case)
    eb, pb = capi.shard_plan(idx, k_low, n_shards)
    p30 = _prefix30(idx, K)
    align = min(6, k_low)
    group = p30 >> np.uint32(5 * (6 - align))
    assert eb[0] == 0 and eb[-1] == len(idx) and pb[0] == 0 and pb[-1] == 1 << 30
    assert np.all(np.diff(eb.astype(np.int64)) >= 0) and np.all(np.diff(pb.astype(np.int64)) >= 0)
    for r in range(1, n_shards):
        b = int(eb[r])
        assert pb[r] % (1 << (5 * (6 - align))) == 0
        assert np.all(p30[:b] < pb[r]) and np.all(p30[b:] >= pb[r])
        if 0 < b < len(idx):
            assert group[b - 1] != group[b]
        # the cut is the start of the group holding the equal-size candidate
        cand = len(idx) * r // n_shards
        assert b <= cand and (b == 0 or group[cand] == group[b])


def test_shard_plan_rejects_bad_arguments():
    idx = np.zeros(4, dtype=capi.REC64)
    with pytest.raises(kasa_b200.KasaError):
        capi.shard_plan(idx, 7, 0)
    with pytest.raises(kasa_b200.KasaError):
        capi.shard_plan(idx, 7, 17)
    eb, pb = capi.shard_plan(np.zeros(0, dtype=capi.REC64), 7, 3)           # empty index: every shard empty
    assert list(eb) == [0, 0, 0, 0] and pb[0] == 0 and pb[-1] == 1 << 30


def test_shard_plan_hist_balances_on_aligned_buckets():
    rng = np.random.default_rng(5)
    hist = rng.integers(0, 1000, size=1 << capi.SHARD_HIST_BITS).astype(np.uint64)
    hist[:5000] = 0
    for k_low, n in [(7, 8), (2, 3), (1, 4), (3, 16)]:
        pb = capi.shard_plan_hist(hist, k_low, n)
        align_bits = min(15, 5 * min(6, k_low))
        assert pb[0] == 0 and pb[-1] == 1 << 30 and np.all(np.diff(pb.astype(np.int64)) >= 0)
        assert np.all(pb[:-1] % (1 << (30 - align_bits)) == 0)
        cum = np.concatenate([[0], np.cumsum(hist)])
        total = int(cum[-1])
        step = 1 << (15 - align_bits)
        for r in range(1, n):
            b = int(pb[r]) >> 15
            want = total * r // n
            assert int(cum[b]) <= want < int(cum[min(b + step, len(hist))]) or b == len(hist)
