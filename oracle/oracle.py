"""ctypes front-end of oracle/kasa_oracle.c — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (the product, kasa_b200/, never does).

Host-side read/sequence preparation is restated here in Python (it is string
handling), citing the reference; the arithmetic lives in kasa_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def build_port() -> str:
    """Compile kasa_oracle.c -> libkasa_oracle.so (gcc only; always possible)."""
    so = os.path.join(_HERE, "libkasa_oracle.so")
    src = os.path.join(_HERE, "kasa_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", so, src, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build_port())
        L.ko_default_codon_table.argtypes = [u8p]
        L.ko_kmer_count.restype = C.c_uint64
        L.ko_kmer_count.argtypes = [C.c_uint64, C.c_int]
        L.ko_extract.restype = C.c_uint64
        L.ko_extract.argtypes = [u8p, C.c_char_p, C.c_uint64, C.c_int, u64p]
        L.ko_extract_build.restype = C.c_uint64
        L.ko_extract_build.argtypes = [u8p, C.c_char_p, C.c_uint64, C.c_int, u64p]
        L.ko_extract_build_step.restype = C.c_uint64
        L.ko_extract_build_step.argtypes = [u8p, C.c_char_p, C.c_uint64, C.c_int, C.c_int, u64p]
        L.ko_sort_dedup.restype = C.c_uint64
        L.ko_sort_dedup.argtypes = [u64p, u32p, C.c_uint64, C.c_int, u8p]
        L.ko_trie.restype = C.c_uint64
        L.ko_trie.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, u8p]
        L.ko_frequency.restype = C.c_int
        L.ko_frequency.argtypes = [u8p, C.c_uint64, C.c_int, C.c_int, u32p, C.c_uint32, u64p]
        L.ko_identify.restype = C.c_int
        L.ko_identify.argtypes = [u64p, u32p, C.c_uint64, u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                  u32p, C.c_uint32, f64p, u64p, f32p, u64p]
        L.ko_identify_ordered.restype = C.c_int
        L.ko_identify_ordered.argtypes = [u64p, u32p, C.c_uint64, u8p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          u32p, C.c_uint32, f64p, u64p, f32p]
        L.ko_best_score.restype = C.c_float
        L.ko_best_score.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.ko_relative_score.restype = C.c_double
        L.ko_relative_score.argtypes = [C.c_float, C.c_uint64, C.c_uint32, C.c_int]
        _LIB = L
    return _LIB


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def default_codon_table() -> np.ndarray:
    t = np.zeros(512, dtype=np.uint8)
    lib().ko_default_codon_table(_p(t, u8p))
    return t


# --------------------------------------------------------------------------- host prep
_REVC = {ord("A"): "T", ord("C"): "G", ord("G"): "C", ord("T"): "A", ord("a"): "T", ord("c"): "G", ord("g"): "C",
         ord("t"): "A"}


def clean(seq: bytes) -> bytes:
    """Read.hpp:657-675 / 2504-2520: every char not in ACGTacgt -> 'Z'."""
    if b" " in seq or b"\t" in seq:
        raise ValueError("Spaces or tabs inside read, please check your input.")
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    ok = np.isin(a, np.frombuffer(b"ACGTacgt", dtype=np.uint8))
    a[~ok] = ord("Z")
    return a.tobytes()


_RC_TAB = np.full(256, ord("?"), dtype=np.uint8)
for _c in range(256):
    _RC_TAB[_c] = b"TGACXZ??"[(_c >> 1) & 7]


def revcomp(seq: bytes) -> bytes:
    """kASA.hpp:214-221 with _aRevComp (kASA.hpp:54)."""
    return _RC_TAB[np.frombuffer(seq, dtype=np.uint8)][::-1].tobytes()


def prepare_read(seq: bytes, K: int, k_low: int, six: bool, one: bool = False) -> list[bytes]:
    """Read.hpp:699-760 processInput + :612-654: clean, pad with 'X' until len+marker >= 3K (--one: until (len+marker)/3 >= K),
    append the marker 'X'*3*(K-kLow) (Read.hpp:1068-1078).  Returns [forward] or [forward, rc]
    in the reference's emission order per read (forward first: Read.hpp:278-281)."""
    marker = b"X" * (3 * (K - k_low))
    s = clean(seq)
    if len(s) > 0:
        if one:
            while (len(s) + len(marker)) // 3 < K:
                s += b"X"
        else:
            while len(s) + len(marker) < 3 * K:
                s += b"X"
    out = [s + marker]
    if six and not one:
        out.append(revcomp(s) + marker)
    return out


def extract_one_frame(seq_prepared: bytes, K: int, table: np.ndarray | None = None) -> np.ndarray:
    """--one (Read.hpp:223-260 convert_dnaTokMerOneFrame, count Read.hpp:43-48): the k-mers of the first reading frame only, i.e.
    every third k-mer of the all-frames walk, len/3 - K + 1 of them when len/3 > K + 1."""
    t = len(seq_prepared) // 3
    n = t - K + 1 if t > K + 1 else 0
    return np.ascontiguousarray(extract(seq_prepared, K, table)[0::3][:n])


def extract(seq_prepared: bytes, K: int, table: np.ndarray | None = None, build: bool = False, one: bool = False) -> np.ndarray:
    """[n,2] u64 {low, high} k-mers of one prepared string (identify or build rules; build with one = True: `build --one`)."""
    table = default_codon_table() if table is None else table
    n = max(0, len(seq_prepared) - 3 * K + 1)
    out = np.zeros((max(n, 1), 2), dtype=np.uint64)
    if build:
        w = lib().ko_extract_build_step(_p(table, u8p), seq_prepared, len(seq_prepared), K, 3 if one else 1, _p(out, u64p))
    else:
        w = lib().ko_extract(_p(table, u8p), seq_prepared, len(seq_prepared), K, _p(out, u64p))
    return out[:w]


# --------------------------------------------------------------------------- build
def rec_dtype(bits: int) -> np.dtype:
    """packedPairs.hpp:107-111 (12 B) / :132-136 (20 B), packed little-endian."""
    if bits == 128:
        return np.dtype([("lo", "<u8"), ("hi", "<u8"), ("tax", "<u4")])
    return np.dtype([("lo", "<u8"), ("tax", "<u4")])


TRIE_DTYPE = np.dtype([("count", "<u8"), ("prefix", "<u4")])


def file_chunks(data: bytes):
    """Utilities.hpp:448-550 FileReader::getChunk: the file is consumed through a 2048-byte buffer and a
    chunk ends at a newline OR at the buffer end, so a line straddling a 2048-byte file offset arrives as
    two chunks.  Yields (text, complete_line); after the data, yields (b"", None) forever (eof)."""
    pos, n = 0, len(data)
    while True:
        if pos >= n:
            yield b"", None
            continue
        bufend = (pos // 2048 + 1) * 2048
        nl = data.find(b"\n", pos, min(bufend, n))
        if nl >= 0:
            yield data[pos:nl], True
            pos = nl + 1
        elif bufend <= n:
            yield data[pos:bufend], False
            pos = bufend
        else:                       # last partial buffer: the reader plants a '\n' sentinel behind it (:478-481)
            yield data[pos:n], True
            pos = n


def _acc_of(header: bytes) -> str:
    """Read.hpp:2346-2353: first blank-separated token, split on '|', first part containing a '.'."""
    first = header.split(b" ")[0]
    for part in first.split(b"|"):
        if b"." in part:
            return part.decode()
    return ""


def fasta_build_strings(data: bytes, acc2tax: dict, K: int, six: bool) -> list[tuple[bytes, int]]:
    """Read.hpp:2306-2578 readFasta, restated over the chunk stream: returns the (string, taxid) pairs handed
    to dnaTokMers in order.  Pieces are 20 chunks long with a 3K-1 overhang; the forward marker
    'X'*3*(K-1) (Read.hpp:2323-2333) closes an entry; the reverse strand of a piece gets the marker when it is
    the first piece of an entry or a piece closed by the next '>' (bAddFalseRCMarker, Read.hpp:2443-2445, 2529-2532),
    which makes the six-frame index depend on where 2048-byte buffer edges cut the lines."""
    marker = b"X" * (3 * (K - 1))
    it = file_chunks(data)
    out: list[tuple[bytes, int]] = []
    idx = next_idx = 0
    found = False
    eof = False
    # header of the first entry
    while not eof:
        chunk, complete = next(it)
        if complete is None:
            eof = True
            break
        if chunk == b"":
            continue
        if chunk[:1] != b">":
            raise ValueError("No > found in input.")
        hdr = chunk[1:]
        acc = _acc_of(hdr)
        if acc in acc2tax:
            idx, found = acc2tax[acc], True
        elif hdr.decode() in acc2tax:
            idx, found = acc2tax[hdr.decode()], True
        else:
            found = False
        break
    add_marker, next_flag, add_rc_marker = False, False, True
    overhang, dna = b"", b""
    while (not eof) or overhang != b"":
        if next_flag:
            add_marker, add_rc_marker = False, True
        create_overhang = True
        i = 0
        while i < 20:
            chunk, complete = next(it)
            if complete is None:
                eof = True
            if chunk == b"":
                if not eof:
                    break
                create_overhang = False
                add_marker = True
                break
            if chunk[:1] == b">":
                hdr = chunk[1:]
                while not complete:
                    more, complete = next(it)
                    if complete is None:
                        eof = True
                        break
                    hdr += more
                acc = _acc_of(hdr)
                if acc in acc2tax:
                    next_idx, found, next_flag = acc2tax[acc], True, True
                elif hdr.decode() in acc2tax:
                    next_idx, found, next_flag = acc2tax[hdr.decode()], True, True
                else:
                    found = False
                create_overhang = False
                add_marker = True
                add_rc_marker = True
                break
            if found:
                dna += chunk
                i += 1
            # lines of an unknown accession are skipped without counting (Read.hpp:2463-2465)
        dna = overhang + dna
        if len(dna) < 3 * K + 2 and create_overhang:
            overhang, dna = dna, b""
            continue
        dna = clean(dna)
        rc = revcomp(dna) if six else b""
        if add_rc_marker and rc:
            rc += marker
            add_rc_marker = False
        if add_marker and dna:
            dna += marker
            add_marker = False
        out.append((dna, idx))
        if six:
            out.append((rc, idx))
        if next_flag:
            idx, next_flag = next_idx, False
        overhang = dna[len(dna) + 1 - 3 * K:] if create_overhang else b""
        dna = b""
    return out


def build_index_from_strings(strings: list[tuple[bytes, int]], K: int, table: np.ndarray | None = None, one: bool = False):
    """(string, taxid) pairs -> (index records, trie records): Read.hpp:1991-2139 + Build.hpp:305-477 + Trie.hpp:366-394."""
    bits = 128 if K > 12 else 64
    table = default_codon_table() if table is None else table
    ks, ts = [], []
    for st, tax in strings:
        k = extract(st, K, table, build=True, one=one)
        ks.append(k)
        ts.append(np.full(len(k), tax, dtype=np.uint32))
    kk = np.ascontiguousarray(np.concatenate(ks)) if ks else np.zeros((0, 2), np.uint64)
    tt = np.ascontiguousarray(np.concatenate(ts)) if ts else np.zeros(0, np.uint32)
    out = np.zeros(max(1, len(kk)), dtype=rec_dtype(bits))
    n = lib().ko_sort_dedup(_p(kk, u64p), _p(tt, u32p), len(kk), bits, out.ctypes.data_as(u8p))
    index = out[:n].copy()
    tr = np.zeros(max(1, n), dtype=TRIE_DTYPE)
    nt = lib().ko_trie(index.ctypes.data_as(u8p), n, bits, K, tr.ctypes.data_as(u8p))
    return index, tr[:nt].copy()


def build_index(fasta: bytes, acc2tax: dict, K: int, six: bool = False, table: np.ndarray | None = None, one: bool = False):
    """Whole `kASA build` restated: FASTA bytes + accession->taxid -> (index, trie)."""
    return build_index_from_strings(fasta_build_strings(fasta, acc2tax, K, six and not one), K, table, one)


def frequency(index: np.ndarray, K: int, taxid_of_idx: np.ndarray) -> np.ndarray:
    bits = 128 if index.dtype.itemsize == 20 else 64
    T = len(taxid_of_idx)
    f = np.zeros((T, K), dtype=np.uint64)
    tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
    rc = lib().ko_frequency(index.ctypes.data_as(u8p), len(index), bits, K, _p(tax, u32p), T, _p(f, u64p))
    if rc:
        raise KeyError("index taxid missing from content")
    return f


# --------------------------------------------------------------------------- identify
@dataclass
class IdentifyResult:
    n_kmers: int
    count_all: np.ndarray      # [numK, T] f64
    count_unique: np.ndarray   # [numK, T] u64
    scores: np.ndarray         # [nreads, T] f32


def identify_kmers(kmers: np.ndarray, read_id: np.ndarray, n_reads: int, index: np.ndarray, K: int, k_high: int,
                   k_low: int, taxid_of_idx: np.ndarray, want_scores: bool = True) -> IdentifyResult:
    bits = 128 if index.dtype.itemsize == 20 else 64
    T = len(taxid_of_idx)
    numK = k_high - k_low + 1
    ca = np.zeros((numK, T), dtype=np.float64)
    cu = np.zeros((numK, T), dtype=np.uint64)
    sc = np.zeros((n_reads, T), dtype=np.float32) if want_scores else None
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    read_id = np.ascontiguousarray(read_id, dtype=np.uint32)
    tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
    index = np.ascontiguousarray(index)
    rc = lib().ko_identify(_p(kmers, u64p), _p(read_id, u32p), len(read_id), index.ctypes.data_as(u8p), len(index),
                           bits, K, k_high, k_low, _p(tax, u32p), T, _p(ca, f64p), _p(cu, u64p),
                           _p(sc, f32p) if want_scores else None, None)
    if rc:
        raise KeyError("index taxid missing from content")
    return IdentifyResult(len(read_id), ca, cu, sc)


def identify_kmers_ordered(kmers: np.ndarray, read_id: np.ndarray, n_reads: int, index: np.ndarray, K: int, k_high: int,
                           k_low: int, taxid_of_idx: np.ndarray, sort_bits: int = 32) -> IdentifyResult:
    """The CUDA library's KASA_SCORE_ORDERED summation order on the CPU (ko_identify_ordered): the bit-exact check of the GPU path."""
    bits = 128 if index.dtype.itemsize == 20 else 64
    T = len(taxid_of_idx)
    numK = k_high - k_low + 1
    ca = np.zeros((numK, T), dtype=np.float64)
    cu = np.zeros((numK, T), dtype=np.uint64)
    sc = np.zeros((n_reads, T), dtype=np.float32)
    kmers = np.ascontiguousarray(kmers, dtype=np.uint64)
    read_id = np.ascontiguousarray(read_id, dtype=np.uint32)
    tax = np.ascontiguousarray(taxid_of_idx, dtype=np.uint32)
    index = np.ascontiguousarray(index)
    rc = lib().ko_identify_ordered(_p(kmers, u64p), _p(read_id, u32p), len(read_id), index.ctypes.data_as(u8p), len(index),
                                   bits, K, k_high, k_low, sort_bits, _p(tax, u32p), T, _p(ca, f64p), _p(cu, u64p), _p(sc, f32p))
    if rc:
        raise KeyError("index taxid missing from content")
    return IdentifyResult(len(read_id), ca, cu, sc)


def presort_limit(K: int) -> int:
    """kasa_b200/csrc/extract.cuh presort_limit(): most k-mers of a read that the extraction still emits in ascending order."""
    return 12288 if K > 12 else 24576


def presorted_emission(kmers: np.ndarray, read_id: np.ndarray, presort_max: int = 24576):
    """The CUDA library's emission order (extract.cu "presorted emission"): the k-mers of a read of at most `presort_max` k-mers in
    ascending order (equal k-mers keep the reference's emission order), longer reads untouched.  read_id is non-decreasing."""
    n = len(read_id)
    if n == 0:
        return kmers, read_id
    pos = np.arange(n, dtype=np.uint64)
    long_read = np.bincount(read_id)[read_id] > presort_max
    hi = np.where(long_read, np.uint64(0), kmers[:, 1])
    lo = np.where(long_read, pos, kmers[:, 0])
    order = np.lexsort((pos, lo, hi, read_id))
    return np.ascontiguousarray(kmers[order]), np.ascontiguousarray(read_id[order])


def identify_reads_ordered(reads: list[list[bytes]], index: np.ndarray, K: int, k_high: int, k_low: int, six: bool,
                           taxid_of_idx: np.ndarray, table: np.ndarray | None = None, sort_bits: int = 32, one: bool = False) -> IdentifyResult:
    """KASA_SCORE_ORDERED: the library emits the k-mers of ordinary reads presorted and sorts the batch stably on the leading
    `sort_bits` key bits, so a read's scores are summed in ascending k-mer order (reads of more than presort_limit(K) k-mers: ascending
    on the leading sort_bits, then emission order)."""
    k, ids = presorted_emission(*read_kmers(reads, K, k_low, six, table, one), presort_limit(K))
    return identify_kmers_ordered(k, ids, len(reads), index, K, k_high, k_low, taxid_of_idx, sort_bits)


def read_kmers(reads: list[list[bytes]], K: int, k_low: int, six: bool, table: np.ndarray | None = None, one: bool = False):
    """reads[i] = list of mate sequences sharing read id i (paired-end mates pool their k-mers,
    Read.hpp:963-970).  Returns (kmers [n,2] u64, read_id [n] u32) in the reference's emission order."""
    table = default_codon_table() if table is None else table
    ks, ids = [], []
    for rid, mates in enumerate(reads):
        for m in mates:
            for st in prepare_read(m, K, k_low, six, one):
                k = extract_one_frame(st, K, table) if one else extract(st, K, table)
                ks.append(k)
                ids.append(np.full(len(k), rid, dtype=np.uint32))
    if not ks:
        return np.zeros((0, 2), np.uint64), np.zeros(0, np.uint32)
    return np.ascontiguousarray(np.concatenate(ks)), np.ascontiguousarray(np.concatenate(ids))


def identify_reads(reads: list[list[bytes]], index: np.ndarray, K: int, k_high: int, k_low: int, six: bool,
                   taxid_of_idx: np.ndarray, table: np.ndarray | None = None, one: bool = False) -> IdentifyResult:
    k, ids = read_kmers(reads, K, k_low, six, table, one)
    return identify_kmers(k, ids, len(reads), index, K, k_high, k_low, taxid_of_idx)


def best_score(read_len: int, k_high: int, k_low: int, frames: int) -> np.float32:
    return np.float32(lib().ko_best_score(read_len, k_high, k_low, frames))


def relative_score(kmer_score, freq_highest: int, read_len: int, K: int) -> float:
    return lib().ko_relative_score(float(kmer_score), int(freq_highest), read_len, K)


HALF_DTYPE = np.dtype([("lo", "<u4"), ("idx", "<u2")])


def half_index(index: np.ndarray, taxid_of_idx: np.ndarray):
    """`kASA half` (Shrink::putHalfInTrie, Shrink.hpp:78-143) of a 64-bit index: the records keep their lower six letters and the
    position of their taxid in the content table; records whose lower six letters are all '^' (30) are dropped; the prefix table
    counts the kept records per upper six letters.  The reference closes the last run with `count - 1` unless the count is 1
    (Shrink.hpp:131-136) — restated as it is.  Returns (records, prefix table)."""
    lo30 = index["lo"] & np.uint64((1 << 30) - 1)
    keep = lo30 != np.uint64(1039104990)                       # "^^^^^^"
    pos = {int(t): i for i, t in enumerate(taxid_of_idx)}
    rec = np.zeros(int(keep.sum()), dtype=HALF_DTYPE)
    rec["lo"] = lo30[keep].astype(np.uint32)
    rec["idx"] = np.array([pos[int(t)] for t in index["tax"][keep]], dtype=np.uint16)
    prefix = (index["lo"][keep] >> np.uint64(30)).astype(np.uint32)
    # the reference starts from the first record's prefix whether that record is kept or not (Shrink.hpp:93); runs by change of prefix
    starts = np.flatnonzero(np.concatenate(([True], prefix[1:] != prefix[:-1])))
    counts = np.diff(np.concatenate((starts, [len(prefix)]))).astype(np.uint64)
    if len(counts) and counts[-1] != 1:
        counts[-1] -= np.uint64(1)
    trie = np.zeros(len(starts), dtype=TRIE_DTYPE)
    trie["count"] = counts
    trie["prefix"] = prefix[starts]
    return rec, trie


# --------------------------------------------------------------------------- file helpers (formats of SURVEY.md §3.4)
def load_index_file(path: str):
    info = open(path + "_info.txt").read().split()
    n = int(info[0])
    bits = 128 if len(info) > 1 and info[1] == "128" else 64
    return np.fromfile(path, dtype=rec_dtype(bits), count=n)


def load_trie_file(path: str):
    n = int(open(path + "_trie.txt").read().split()[0])
    return np.fromfile(path + "_trie", dtype=TRIE_DTYPE, count=n)


def load_content(path: str):
    """Compare.hpp:111-180: dense idx 0 = non_unique, idx i = i-th content line."""
    names, taxids = ["non_unique"], [0]
    for line in open(path):
        line = line.rstrip("\n")
        if not line:
            continue
        p = line.split("\t")
        names.append(p[0])
        taxids.append(int(p[1]))
    return names, np.array(taxids, dtype=np.uint32)
