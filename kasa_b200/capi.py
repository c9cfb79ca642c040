"""ctypes binding of include/kasa_b200.h plus a small host-side wrapper.

The wrapper mirrors the reference's call sites for the path:
  Context.load_index        <- Compare::index.loadIndex / loadContentAndFrequencyFiles (Compare.hpp:96-337)
  Context.identify_batch    <- convertAndSort + sortInputAndCheckInvalidkMers_sta + compareWithDatabase per batch
  Context.identify_finish   <- count reduction (Compare.hpp:3445-3454)
  Context.build_push/finish <- dnaTokMers/addToInt, IntToExtPart/mergeTemporaries/SaveToStxxlVec, GetFrequencyK
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MEM_HOST, MEM_DEVICE = 0, 1
SEQ_ASCII, SEQ_PACKED2 = 0, 1
SCORE_NONE, SCORE_ORDERED, SCORE_EXACT = 0, 1, 2
SCORE_FAST = SCORE_ORDERED      # name used by ABI version 1

REC64 = np.dtype([("lo", "<u8"), ("tax", "<u4")])                     # packedBigPair, packedPairs.hpp:107-111
REC128 = np.dtype([("lo", "<u8"), ("hi", "<u8"), ("tax", "<u4")])     # packedLargePair, packedPairs.hpp:132-136
TRIE = np.dtype([("count", "<u8"), ("prefix", "<u4")])                # packedBigPairTrie, packedPairs.hpp:157-166

EXPORTS = [
    "kasa_abi_version", "kasa_ctx_create", "kasa_ctx_destroy", "kasa_last_error", "kasa_ctx_set_stream",
    "kasa_ctx_synchronize", "kasa_ctx_launch_count", "kasa_ctx_stage_times", "kasa_index_load", "kasa_index_from_build", "kasa_index_size",
    "kasa_identify_batch", "kasa_identify_submit", "kasa_identify_wait", "kasa_identify_finish", "kasa_build_push", "kasa_build_finish", "kasa_build_fetch",
    "kasa_shard_plan", "kasa_index_from_build_shard", "kasa_shard_init", "kasa_shard_connect", "kasa_shard_sort", "kasa_shard_lookup",
    "kasa_shard_score", "kasa_build_shard_local", "kasa_shard_plan_hist", "kasa_build_shard_runs", "kasa_build_shard_pull", "kasa_build_frequency", "kasa_dbg_extract", "kasa_dbg_sort_pairs", "kasa_synth_build_reference", "kasa_synth_build_reference_range", "kasa_synth_reads", "kasa_host_alloc", "kasa_host_free", "kasa_builtin_codon_table", "kasa_build_push_index", "kasa_index_load_half",
]


class KasaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("kasa_b200 error %d: %s" % (code, msg))
        self.code = code


class CParams(C.Structure):
    _fields_ = [("device", C.c_int32), ("highest_k", C.c_int32), ("k_high", C.c_int32), ("k_low", C.c_int32),
                ("frames", C.c_int32), ("score_mode", C.c_int32), ("max_batch_kmers", C.c_uint64),
                ("max_batch_seqs", C.c_uint32), ("max_batch_bases", C.c_uint64), ("codon_table", C.c_void_p),
                ("max_batch_hits", C.c_uint64)]


class CReadBatch(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_seqs", C.c_uint32), ("format", C.c_int32), ("location", C.c_int32),
                ("bases", C.c_void_p), ("nmask", C.c_void_p), ("seq_start", C.c_void_p), ("seq_len", C.c_void_p),
                ("read_first_seq", C.c_void_p)]


class CBatchResult(C.Structure):
    _fields_ = [("n_kmers", C.c_uint64), ("scores", C.c_void_p), ("scores_location", C.c_int32), ("gpu_ms", C.c_float),
                ("hit_taxon", C.c_void_p), ("hit_score", C.c_void_p), ("hit_begin", C.c_void_p), ("hit_count", C.c_void_p),
                ("hit_capacity", C.c_uint64), ("n_hits", C.c_uint64), ("hits_location", C.c_int32)]


class CSeqBatch(C.Structure):
    _fields_ = [("n_seqs", C.c_uint32), ("location", C.c_int32), ("bases", C.c_void_p), ("seq_start", C.c_void_p),
                ("seq_len", C.c_void_p), ("taxid", C.c_void_p)]


MAX_SHARDS, SHARD_NBUF, SHARD_HIST_BITS = 16, 9, 15


class CShardMem(C.Structure):
    _fields_ = [("pid", C.c_uint64), ("device", C.c_int32), ("reserved", C.c_int32), ("ptr", C.c_uint64 * SHARD_NBUF),
                ("handle", (C.c_ubyte * 64) * SHARD_NBUF)]


class CShardRuns(C.Structure):
    _fields_ = [("n_kmers", C.c_uint64), ("cur", C.c_int32), ("reserved", C.c_int32), ("begin", C.c_uint64 * (MAX_SHARDS + 1))]


def library_path() -> str:
    # KASA_B200_LIB: another build of the same library (kernel variants compared by the tools; never a fallback)
    return os.environ.get("KASA_B200_LIB") or os.path.join(_HERE, "libkasa_b200.so")


_LIB = None


def load_library():
    """dlopen libkasa_b200.so; raises if it has not been built (python __graft_entry__.py / make -C kasa_b200/csrc)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise KasaError(-2, "libkasa_b200.so is not built (run `make -C kasa_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(path)
    L.kasa_last_error.restype = C.c_char_p
    L.kasa_last_error.argtypes = [C.c_void_p]
    L.kasa_ctx_create.argtypes = [C.POINTER(CParams), C.POINTER(C.c_void_p)]
    L.kasa_ctx_destroy.argtypes = [C.c_void_p]
    L.kasa_ctx_destroy.restype = None
    L.kasa_ctx_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.kasa_ctx_synchronize.argtypes = [C.c_void_p]
    L.kasa_ctx_launch_count.argtypes = [C.c_void_p]
    L.kasa_ctx_launch_count.restype = C.c_uint64
    L.kasa_ctx_stage_times.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64), C.c_int]
    L.kasa_index_size.argtypes = [C.c_void_p]
    L.kasa_index_size.restype = C.c_uint64
    L.kasa_index_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32]
    L.kasa_index_from_build.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.kasa_identify_batch.argtypes = [C.c_void_p, C.POINTER(CReadBatch), C.POINTER(CBatchResult)]
    L.kasa_identify_submit.argtypes = [C.c_void_p, C.POINTER(CReadBatch), C.POINTER(CBatchResult)]
    L.kasa_identify_wait.argtypes = [C.c_void_p]
    L.kasa_identify_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.kasa_shard_plan.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kasa_index_from_build_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.kasa_shard_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(CShardMem)]
    L.kasa_shard_connect.argtypes = [C.c_void_p, C.POINTER(CShardMem)]
    L.kasa_shard_sort.argtypes = [C.c_void_p, C.POINTER(CReadBatch), C.POINTER(CShardRuns)]
    L.kasa_shard_lookup.argtypes = [C.c_void_p, C.POINTER(CShardRuns)]
    L.kasa_shard_score.argtypes = [C.c_void_p, C.POINTER(CShardRuns), C.POINTER(CBatchResult)]
    L.kasa_build_shard_local.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
    L.kasa_shard_plan_hist.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.kasa_build_shard_runs.argtypes = [C.c_void_p, C.POINTER(CShardRuns)]
    L.kasa_build_shard_pull.argtypes = [C.c_void_p, C.POINTER(CShardRuns)]
    L.kasa_build_push.argtypes = [C.c_void_p, C.POINTER(CSeqBatch)]
    L.kasa_build_finish.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.kasa_build_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.kasa_build_frequency.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.kasa_dbg_extract.argtypes = [C.c_void_p, C.POINTER(CReadBatch), C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.POINTER(C.c_uint64)]
    L.kasa_dbg_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    L.kasa_synth_build_reference.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64]
    L.kasa_synth_build_reference_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
    L.kasa_synth_reads.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64,
                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
    _LIB = L
    return L


def shard_plan(records: np.ndarray, k_low: int, n_shards: int):
    """Host-only: cut a sorted index into n_shards prefix ranges (kasa_shard_plan) -> (entry_begin, prefix_begin)."""
    L = load_library()
    records = np.ascontiguousarray(records)
    bits = 128 if records.dtype.itemsize == 20 else 64
    eb = np.zeros(n_shards + 1, dtype=np.uint64)
    pb = np.zeros(n_shards + 1, dtype=np.uint32)
    rc = L.kasa_shard_plan(_ptr(records) if len(records) else None, len(records), bits, k_low, n_shards, _ptr(eb), _ptr(pb))
    if rc != 0:
        raise KasaError(rc, L.kasa_last_error(None).decode())
    return eb, pb


def shard_plan_hist(hist: np.ndarray, k_low: int, n_shards: int) -> np.ndarray:
    """Host-only: prefix_begin[n_shards+1] from the global leading-15-bit histogram (kasa_shard_plan_hist)."""
    L = load_library()
    hist = np.ascontiguousarray(hist, dtype=np.uint64)
    assert len(hist) == 1 << SHARD_HIST_BITS
    pb = np.zeros(n_shards + 1, dtype=np.uint32)
    rc = L.kasa_shard_plan_hist(_ptr(hist), k_low, n_shards, _ptr(pb))
    if rc != 0:
        raise KasaError(rc, L.kasa_last_error(None).decode())
    return pb


@dataclass
class Params:
    device: int = 0
    highest_k: int = 12
    k_high: int = 12
    k_low: int = 7
    frames: int = 3
    score_mode: int = SCORE_ORDERED
    max_batch_kmers: int = 1 << 22
    max_batch_seqs: int = 1 << 16
    max_batch_bases: int = 1 << 24
    codon_table: np.ndarray | None = None
    max_batch_hits: int = 0


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


class Context:
    """One GPU context (one per process / rank)."""

    def __init__(self, params: Params):
        self.lib = load_library()
        self.params = params
        self._table = None
        cp = CParams(params.device, params.highest_k, params.k_high, params.k_low, params.frames, params.score_mode,
                     params.max_batch_kmers, params.max_batch_seqs, params.max_batch_bases, None, params.max_batch_hits)
        if params.codon_table is not None:
            self._table = np.ascontiguousarray(params.codon_table, dtype=np.uint8)
            cp.codon_table = self._table.ctypes.data
        h = C.c_void_p()
        rc = self.lib.kasa_ctx_create(C.byref(cp), C.byref(h))
        if rc != 0:
            raise KasaError(rc, self.lib.kasa_last_error(None).decode())
        self.h = h
        self.n_taxa = 0
        self.num_k = params.k_high - params.k_low + 1
        self.wide = params.highest_k > 12

    def close(self):
        if getattr(self, "h", None):
            self.lib.kasa_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise KasaError(rc, self.lib.kasa_last_error(self.h).decode())

    def set_stream(self, cuda_stream: int):
        self._check(self.lib.kasa_ctx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.kasa_ctx_synchronize(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.kasa_ctx_launch_count(self.h))

    STAGES = ("input", "extract", "sort_hist", "sort_passes", "lookup", "per_read", "output")

    def stage_times(self, reset: bool = False):
        """-> (dict stage -> ms summed over batches, batches, radix passes, k-mers)."""
        ms = np.zeros(len(self.STAGES), dtype=np.float64)
        b, p, k = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.kasa_ctx_stage_times(self.h, _ptr(ms), C.byref(b), C.byref(p), C.byref(k), int(reset)))
        return dict(zip(self.STAGES, ms.tolist())), int(b.value), int(p.value), int(k.value)

    @property
    def index_size(self) -> int:
        return int(self.lib.kasa_index_size(self.h))

    # ---------------------------------------------------------------- index
    def load_index(self, records: np.ndarray, taxid_of_idx: np.ndarray):
        records = np.ascontiguousarray(records)
        bits = 128 if records.dtype.itemsize == 20 else 64
        tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
        self._check(self.lib.kasa_index_load(self.h, _ptr(records), len(records), bits, _ptr(tax), len(tax)))
        self.n_taxa = len(tax)

    def load_index_half(self, records6: np.ndarray, trie: np.ndarray, n_taxa: int):
        """halved index (`kASA half`): 6-byte records {u32 lower letters, u16 content index} + the records of <index>_trie (kasa_index_load_half)."""
        records6 = np.ascontiguousarray(records6)
        trie = np.ascontiguousarray(trie)
        assert records6.dtype.itemsize == 6 and trie.dtype.itemsize == 12
        self._check(self.lib.kasa_index_load_half(self.h, _ptr(records6), C.c_uint64(len(records6)), _ptr(trie), C.c_uint64(len(trie)), C.c_uint32(n_taxa)))
        self.n_taxa = n_taxa

    def index_from_build(self, taxid_of_idx: np.ndarray):
        tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
        self._check(self.lib.kasa_index_from_build(self.h, _ptr(tax), len(tax)))
        self.n_taxa = len(tax)

    # ---------------------------------------------------------------- identify
    @staticmethod
    def pack_reads(seqs: list[bytes]):
        """Concatenate raw sequences -> (ascii u8 array, seq_start u64, seq_len u32)."""
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        starts = np.zeros(len(seqs), dtype=np.uint64)
        if len(seqs) > 1:
            starts[1:] = np.cumsum(lens[:-1], dtype=np.uint64)
        data = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, np.uint8)
        return data, starts, lens

    def _read_batch(self, seqs, read_first_seq):
        data, starts, lens = self.pack_reads(seqs) if isinstance(seqs, list) else seqs
        n_seqs = len(lens)
        first = None
        n_reads = n_seqs
        if read_first_seq is not None:
            first = np.ascontiguousarray(read_first_seq, dtype=np.uint32)
            n_reads = len(first) - 1
        rb = CReadBatch(n_reads, n_seqs, SEQ_ASCII, MEM_HOST, data.ctypes.data if len(data) else None, None,
                        starts.ctypes.data if n_seqs else None, lens.ctypes.data if n_seqs else None,
                        first.ctypes.data if first is not None else None)
        return rb, (data, starts, lens, first), n_reads

    def identify_batch(self, seqs, read_first_seq=None):
        """seqs: list of raw sequences (bytes) or (ascii, seq_start, seq_len) arrays.  Returns (n_kmers, scores, gpu_ms)."""
        rb, keep, n_reads = self._read_batch(seqs, read_first_seq)
        want = self.params.score_mode != SCORE_NONE
        scores = np.zeros((n_reads, self.n_taxa), dtype=np.float32) if want else None
        res = CBatchResult(0, scores.ctypes.data if want and scores.size else None, MEM_HOST, 0.0)
        self._check(self.lib.kasa_identify_batch(self.h, C.byref(rb), C.byref(res)))
        return int(res.n_kmers), scores, float(res.gpu_ms)

    def identify_batch_hits(self, seqs, read_first_seq=None):
        """Sparse results (Params.max_batch_hits > 0): -> (n_kmers, hit_begin u64[reads], hit_count u32[reads], hit_taxon u32[n_hits],
        hit_score f32[n_hits]); the hits of read r are [hit_begin[r], hit_begin[r] + hit_count[r])."""
        rb, keep, n_reads = self._read_batch(seqs, read_first_seq)
        cap = self.params.max_batch_hits
        tax = np.zeros(cap, dtype=np.uint32)
        sc = np.zeros(cap, dtype=np.float32)
        beg = np.zeros(max(n_reads, 1), dtype=np.uint64)
        cnt = np.zeros(max(n_reads, 1), dtype=np.uint32)
        res = CBatchResult(0, None, MEM_HOST, 0.0, tax.ctypes.data, sc.ctypes.data, beg.ctypes.data, cnt.ctypes.data, cap, 0, MEM_HOST)
        self._check(self.lib.kasa_identify_batch(self.h, C.byref(rb), C.byref(res)))
        n = int(res.n_hits)
        return int(res.n_kmers), beg[:n_reads], cnt[:n_reads], tax[:n], sc[:n]

    def identify_batch_raw(self, rb: CReadBatch, res: CBatchResult):
        self._check(self.lib.kasa_identify_batch(self.h, C.byref(rb), C.byref(res)))

    # pipelined pair (kasa_identify_submit / kasa_identify_wait): at most three batches in flight; the caller keeps
    # rb's buffers and res (+ its score buffer) alive until the matching wait
    def identify_submit_raw(self, rb: CReadBatch, res: CBatchResult):
        self._check(self.lib.kasa_identify_submit(self.h, C.byref(rb), C.byref(res)))

    def identify_wait(self):
        self._check(self.lib.kasa_identify_wait(self.h))

    def identify_stream(self, batches, read_first_seqs=None):
        """Pipelined identify over an iterable of batches (each as for identify_batch); yields (n_kmers, scores, gpu_ms)
        per batch in order."""
        want = self.params.score_mode != SCORE_NONE
        inflight = []
        for i, seqs in enumerate(batches):
            rb, keep, n_reads = self._read_batch(seqs, read_first_seqs[i] if read_first_seqs is not None else None)
            scores = np.zeros((n_reads, self.n_taxa), dtype=np.float32) if want else None
            res = CBatchResult(0, scores.ctypes.data if want and scores.size else None, MEM_HOST, 0.0)
            self.identify_submit_raw(rb, res)
            inflight.append((res, scores, keep))
            if len(inflight) == 3:
                self.identify_wait()
                r, sc, _ = inflight.pop(0)
                yield int(r.n_kmers), sc, float(r.gpu_ms)
        while inflight:
            self.identify_wait()
            r, sc, _ = inflight.pop(0)
            yield int(r.n_kmers), sc, float(r.gpu_ms)

    def identify_finish(self):
        ca = np.zeros((self.num_k, self.n_taxa), dtype=np.float64)
        cu = np.zeros((self.num_k, self.n_taxa), dtype=np.uint64)
        n = C.c_uint64(0)
        self._check(self.lib.kasa_identify_finish(self.h, _ptr(ca), _ptr(cu), C.byref(n)))
        return ca, cu, int(n.value)

    # ---------------------------------------------------------------- prefix-sharded index (kasa_shard_*)
    def index_from_build_shard(self, taxid_of_idx: np.ndarray, rank: int, n_shards: int):
        """Keep shard `rank` of the index just built; -> (entry_begin u64[n_shards+1], prefix_begin u32[n_shards+1])."""
        tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
        eb = np.zeros(n_shards + 1, dtype=np.uint64)
        pb = np.zeros(n_shards + 1, dtype=np.uint32)
        self._check(self.lib.kasa_index_from_build_shard(self.h, _ptr(tax), len(tax), rank, n_shards, _ptr(eb), _ptr(pb)))
        self.n_taxa = len(tax)
        return eb, pb

    def shard_init(self, rank: int, n_shards: int, prefix_begin: np.ndarray) -> CShardMem:
        pb = np.ascontiguousarray(prefix_begin, dtype=np.uint32)
        assert len(pb) == n_shards + 1
        mem = CShardMem()
        self._check(self.lib.kasa_shard_init(self.h, rank, n_shards, _ptr(pb), C.byref(mem)))
        return mem

    def shard_connect(self, mems: list):
        arr = (CShardMem * len(mems))(*mems)
        self._check(self.lib.kasa_shard_connect(self.h, arr))

    def shard_sort_raw(self, rb: CReadBatch) -> CShardRuns:
        runs = CShardRuns()
        self._check(self.lib.kasa_shard_sort(self.h, C.byref(rb), C.byref(runs)))
        return runs

    def shard_sort(self, seqs, read_first_seq=None):
        rb, keep, n_reads = self._read_batch(seqs, read_first_seq)
        return self.shard_sort_raw(rb), n_reads

    def shard_lookup(self, runs: list):
        arr = (CShardRuns * len(runs))(*runs)
        self._check(self.lib.kasa_shard_lookup(self.h, arr))

    def shard_score_raw(self, runs: list, res: CBatchResult):
        arr = (CShardRuns * len(runs))(*runs)
        self._check(self.lib.kasa_shard_score(self.h, arr, C.byref(res)))

    def shard_score(self, runs: list, n_reads: int):
        want = self.params.score_mode != SCORE_NONE
        scores = np.zeros((n_reads, self.n_taxa), dtype=np.float32) if want else None
        res = CBatchResult(0, scores.ctypes.data if want and scores.size else None, MEM_HOST, 0.0)
        self.shard_score_raw(runs, res)
        return int(res.n_kmers), scores, float(res.gpu_ms)

    def build_shard_local(self):
        """Sort + dedup the local brick -> (hist u64[2^15] of leading-15-bit buckets, n_local)."""
        hist = np.zeros(1 << SHARD_HIST_BITS, dtype=np.uint64)
        n = C.c_uint64(0)
        self._check(self.lib.kasa_build_shard_local(self.h, _ptr(hist), C.byref(n)))
        return hist, int(n.value)

    def build_shard_runs(self) -> CShardRuns:
        runs = CShardRuns()
        self._check(self.lib.kasa_build_shard_runs(self.h, C.byref(runs)))
        return runs

    def build_shard_pull(self, runs: list):
        arr = (CShardRuns * len(runs))(*runs)
        self._check(self.lib.kasa_build_shard_pull(self.h, arr))

    # ---------------------------------------------------------------- build
    def build_push(self, strings: list[bytes], taxids):
        data, starts, lens = self.pack_reads(strings)
        tax = np.ascontiguousarray(taxids, dtype=np.uint32)
        sb = CSeqBatch(len(lens), MEM_HOST, data.ctypes.data if len(data) else None, starts.ctypes.data, lens.ctypes.data,
                       tax.ctypes.data)
        self._check(self.lib.kasa_build_push(self.h, C.byref(sb)))

    def build_push_index(self, records: np.ndarray):
        """records of an existing index (file layout) join the build: update / merge (kasa_build_push_index)."""
        rec = np.ascontiguousarray(records)
        self._check(self.lib.kasa_build_push_index(self.h, _ptr(rec) if len(rec) else None, C.c_uint64(len(rec)), 128 if self.wide else 64))

    def build_finish(self, fetch: bool = True):
        n, nt = C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.kasa_build_finish(self.h, C.byref(n), C.byref(nt)))
        if not fetch:
            return int(n.value), int(nt.value)
        rec = np.zeros(n.value, dtype=REC128 if self.wide else REC64)
        trie = np.zeros(nt.value, dtype=TRIE)
        self._check(self.lib.kasa_build_fetch(self.h, _ptr(rec) if n.value else None, _ptr(trie) if nt.value else None))
        return rec, trie

    def build_frequency(self, taxid_of_idx: np.ndarray) -> np.ndarray:
        tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
        f = np.zeros((len(tax), self.params.highest_k), dtype=np.uint64)
        self._check(self.lib.kasa_build_frequency(self.h, _ptr(tax), len(tax), _ptr(f)))
        return f

    # ---------------------------------------------------------------- stage-level (tests)
    def dbg_extract(self, seqs, read_first_seq=None):
        rb, keep, n_reads = self._read_batch(seqs, read_first_seq)
        cap = self.params.max_batch_kmers
        lo = np.zeros(cap, dtype=np.uint64)
        hi = np.zeros(cap if self.wide else 1, dtype=np.uint64)
        rid = np.zeros(cap, dtype=np.uint32)
        n = C.c_uint64(0)
        self._check(self.lib.kasa_dbg_extract(self.h, C.byref(rb), _ptr(lo), _ptr(hi), _ptr(rid), C.byref(n)))
        n = int(n.value)
        return lo[:n], (hi[:n] if self.wide else np.zeros(n, np.uint64)), rid[:n]

    def dbg_sort_pairs(self, keys: np.ndarray, vals: np.ndarray, bit_lo: int = 0, bit_hi: int = 64):
        keys = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._check(self.lib.kasa_dbg_sort_pairs(self.h, _ptr(keys), _ptr(vals), len(keys), bit_lo, bit_hi))
        return keys, vals

    def synth_build_reference(self, seed: int, n_taxa: int, genome_len: int, first_taxon: int = 0):
        self._check(self.lib.kasa_synth_build_reference_range(self.h, seed, first_taxon, n_taxa, genome_len))

    def synth_reads(self, genome_seed, read_seed, first_read, n_reads, n_taxa, genome_len, read_len=150,
                    subst_ppm=10000, n_ppm=1000, out=None, out_location=MEM_HOST):
        if out is None:
            out = np.zeros((n_reads, read_len), dtype=np.uint8)
        self._check(self.lib.kasa_synth_reads(self.h, genome_seed, read_seed, first_read, n_reads, n_taxa, genome_len,
                                              read_len, subst_ppm, n_ppm, _ptr(out), out_location))
        return out
