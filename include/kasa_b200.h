/*
 * kasa_b200.h — C ABI of libkasa_b200.so: kASA's `identify` hot path and the `build`
 * index-sort path as hand-written sm_100a CUDA kernels.
 *
 * kASA (the reference) has no plugin / FFI interface; it is a monolithic C++ CLI.
 * The entry points below are what a kASA maintainer would bind in place of the
 * reference's own call sites (cited per function; paths relative to
 * /root/reference/source).  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success
 * and a negative kasa_status otherwise; no exception crosses the boundary;
 * kasa_last_error(ctx) returns the message of the last failure on that context
 * (the reference throws runtime_error and main prints "ERROR: <what>", main.cpp:1717-1720).
 * The caller owns all host buffers, the library owns all device memory.  Calls on
 * one context are not re-entrant (the reference's driver thread is the only caller
 * too: Compare.hpp:3100-3440).  The library never falls back to a CPU path: without
 * a CUDA device every compute entry point fails with KASA_ERR_CUDA.
 */
#ifndef KASA_B200_H
#define KASA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KASA_B200_ABI_VERSION 2

typedef enum {
    KASA_OK = 0,
    KASA_ERR_ARG = -1,        /* invalid argument / unsupported parameter combination */
    KASA_ERR_CUDA = -2,       /* CUDA runtime error (no device, launch failure, ...) */
    KASA_ERR_NOMEM = -3,      /* device or host allocation failed / workspace too small */
    KASA_ERR_STATE = -4,      /* call order violated (e.g. identify before index load) */
    KASA_ERR_DATA = -5,       /* inconsistent input (taxid missing from content, unsorted index, ...) */
    KASA_ERR_INTERNAL = -6    /* kernel watchdog tripped (would have hung) */
} kasa_status;

typedef struct kasa_ctx kasa_ctx;

/* Where a buffer handed to the library lives. */
#define KASA_MEM_HOST   0
#define KASA_MEM_DEVICE 1

/* Encoding of sequence data handed to the library. */
#define KASA_SEQ_ASCII   0    /* one char per base, as read from FASTA/FASTQ                          */
#define KASA_SEQ_PACKED2 1    /* 2 bits per base (A0 C1 T2 G3, kASA.hpp:70-75), 16 bases per uint32,
                                 LSB first, plus a 1-bit-per-base non-ACGT plane (32 bases per uint32) */

/* Per-read scoring mode (Compare.hpp:516-532 scoreMatchNonAVX is the spec: a float32 running sum per (read, taxon), so the
 * value depends on the order of the additions). */
#define KASA_SCORE_NONE    0  /* profile only: counts, no per-read scores (kASA without -q)            */
#define KASA_SCORE_ORDERED 1  /* float32 adds in (k-mer ascending, k ascending) order per read: deterministic, independent of
                                 the batch composition and of index sharding.  Bit-identical to `kASA -n 1` for every read whose
                                 k-mers fall into different 6-letter prefix ranges (all reads without low-complexity stretches);
                                 otherwise within the spread the reference itself shows when the same read is scored in a
                                 different batch (<= 2e-6 relative on the adversarial goldens, tests/test_gpu_parity.py)   */
#define KASA_SCORE_FAST KASA_SCORE_ORDERED   /* name used by ABI version 1 */
#define KASA_SCORE_EXACT   2  /* float32 adds in the reference's own flush order (Compare.hpp:917-955,
                                 1032-1041) — bit-exact with `kASA -n 1` when the batch is the same; audits      */

typedef struct {
    int32_t device;           /* CUDA device ordinal                                                   */
    int32_t highest_k;        /* index width: 12 (64-bit k-mers) or 25 (128-bit); kASA _iHighestK      */
    int32_t k_high, k_low;    /* -k <high> <low> (main.cpp:372-385); k_low >= 1, k_high <= highest_k   */
    int32_t frames;           /* 3 = forward frames (default), 6 = --six, 1 = --one: first frame only — identify
                                 Read.hpp:223-260; build: every third window of a pushed string, Read.hpp:2143-2200 */
    int32_t score_mode;       /* KASA_SCORE_*                                                          */
    uint64_t max_batch_kmers; /* workspace sizing: largest sum of kMerCount per batch                  */
    uint32_t max_batch_seqs;  /* largest number of sequences (reads/mates or build pieces) per batch   */
    uint64_t max_batch_bases; /* largest number of bases per batch                                     */
    const uint8_t *codon_table; /* 512-byte table indexed c0<<6|c1<<3|c2 (kASA.hpp:621-667; `-a` file,
                                 kASA.hpp:579-615) or NULL for the built-in standard code               */
    uint64_t max_batch_hits;  /* sparse per-read results (kasa_batch_result.hit_*): most (read, taxon) pairs with a
                                 score per batch; 0 = dense score matrix only                           */
} kasa_params;

/* --------------------------------------------------------------------------------------------------- */
/* context                                                                                              */

/* Replaces: Compare<...> kASAObj construction (main.cpp:1023/1070) resp. Read<...> (main.cpp:664).     */
int kasa_ctx_create(const kasa_params *params, kasa_ctx **out);
/* The built-in table kasa_params.codon_table == NULL stands for (kASA.hpp:621-667), for hosts that patch single codons from a
 * `-a <gc.prt> <id>` file the way kASA::setCodonTable does (kASA.hpp:579-615) before handing the table over.              */
int kasa_builtin_codon_table(uint8_t *table512);
void kasa_ctx_destroy(kasa_ctx *ctx);
const char *kasa_last_error(const kasa_ctx *ctx);   /* never NULL; ctx may be NULL (creation errors)   */
int kasa_abi_version(void);

/* Run all kernels of this context on an existing CUDA stream (cudaStream_t passed as void*), e.g.
 * torch's current stream so that torch.cuda.Event timing sees them.  Default: a private stream.       */
int kasa_ctx_set_stream(kasa_ctx *ctx, void *cuda_stream);
int kasa_ctx_synchronize(kasa_ctx *ctx);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t kasa_ctx_launch_count(const kasa_ctx *ctx);

/* Device time per identify stage, summed over the batches since the last reset (CUDA events on the
 * context's stream).  ms[KASA_N_STAGES]; *batches = number of batches summed; *sort_passes = radix passes
 * summed (so ms[KASA_STAGE_SORT_PASSES] / *sort_passes is the average duration of the dominant kernel).   */
#define KASA_STAGE_INPUT        0   /* H2D copies, 2-bit packing, k-mer counts and offsets                */
#define KASA_STAGE_EXTRACT      1   /* translation + k-mer packing                                        */
#define KASA_STAGE_SORT_HIST    2   /* digit histograms                                                   */
#define KASA_STAGE_SORT_PASSES  3   /* all radix passes                                                   */
#define KASA_STAGE_LOOKUP       4   /* partition + index lookup (match records)                           */
#define KASA_STAGE_PER_READ     5   /* per-read scores and per-taxon counts                               */
#define KASA_STAGE_OUTPUT       6   /* D2H of the score matrix                                            */
#define KASA_N_STAGES           7
int kasa_ctx_stage_times(kasa_ctx *ctx, double *ms, uint64_t *batches, uint64_t *sort_passes, uint64_t *kmers, int reset);

/* --------------------------------------------------------------------------------------------------- */
/* index                                                                                                */

/* Replaces: Compare::index.loadIndex / loadTrie / loadContentAndFrequencyFiles (Compare.hpp:96-337,
 * main.cpp:1050-1060).  `entries` = the index file bytes: n records of packedBigPair {u64 kmer, u32 taxid}
 * (12 B, packedPairs.hpp:107-111) when bits == 64, or packedLargePair {u64 low, u64 high, u32 taxid}
 * (20 B, :132-136) when bits == 128; must be sorted by (kmer, taxid) as `kASA build` writes it.
 * taxid_of_idx[0] is the "non_unique" pseudo taxon (0); taxid_of_idx[i] the taxid of content line i
 * (Compare.hpp:124-147).  n_taxa (incl. index 0) <= 65535 — the distinct-taxa sets of the lookup are lists, not
 * bit masks, so any content file the 16-bit dense index can address is accepted (the reference's sBitArray,
 * BitArray.hpp:18-147, is unbounded; its 6-byte half index has the same 16-bit limit, packedPairs.hpp:100-105).
 * The prefix-table file (`_trie`) is NOT an input: the lookup kernel searches the HBM-resident index directly
 * (DESIGN.md §4).  The library copies what it needs; the caller may free its buffers after return.      */
int kasa_index_load(kasa_ctx *ctx, const void *entries, uint64_t n, int bits,
                    const uint32_t *taxid_of_idx, uint32_t n_taxa);

/* Replaces: the loading of a halved index (`kASA half` / shrink strategy TrieHalf, Shrink::putHalfInTrie Shrink.hpp:78-143; read by
 * identify through index_t_p, Compare.hpp:185-222).  entries = n records {u32 lower six letters, u16 index into the content table}
 * (6 bytes each), trie = the n_trie records {u64 count, u32 upper six letters} of <index>_trie.  The keys are put together again on
 * the device (prefix << 30 | lower letters) and the index then serves identify like a full one: the reference's own results with
 * the halved and the full index are the same bytes (tests/golden/k12/modes_sha256.txt).  highest_k must be 12.                     */
int kasa_index_load_half(kasa_ctx *ctx, const void *entries, uint64_t n, const void *trie, uint64_t n_trie, uint32_t n_taxa);

/* Adopt the index just built on this context by kasa_build_finish (stays in HBM, no host round trip).  */
int kasa_index_from_build(kasa_ctx *ctx, const uint32_t *taxid_of_idx, uint32_t n_taxa);

uint64_t kasa_index_size(const kasa_ctx *ctx);

/* --------------------------------------------------------------------------------------------------- */
/* identify                                                                                             */

typedef struct {
    uint32_t n_reads;         /* read ids 0..n_reads-1 are batch local (Read.hpp:1107)                  */
    uint32_t n_seqs;          /* sequences; paired-end mates are consecutive and share a read id        */
    int32_t  format;          /* KASA_SEQ_ASCII | KASA_SEQ_PACKED2                                      */
    int32_t  location;        /* KASA_MEM_HOST | KASA_MEM_DEVICE (applies to all pointers below)        */
    const void *bases;        /* ASCII: chars, raw as in the file (cleaning Read.hpp:657-675 happens on
                                 device). PACKED2: uint32 words                                         */
    const uint32_t *nmask;    /* PACKED2 only: non-ACGT plane                                           */
    const uint64_t *seq_start;/* [n_seqs] first base of each sequence (PACKED2: multiple of 32)         */
    const uint32_t *seq_len;  /* [n_seqs] bases                                                         */
    const uint32_t *read_first_seq; /* [n_reads+1] first sequence of each read, or NULL = one seq/read  */
} kasa_read_batch;

typedef struct {
    uint64_t n_kmers;         /* out: sum of kMerCount = this batch's part of iNumberOfkMersInInput
                                 (Compare.hpp:3109-3122)                                                */
    float *scores;            /* in: caller buffer [n_reads][n_taxa] (row = Non_contiguousArray row,
                                 Utilities.hpp:592-636) or NULL; filled unless score_mode == NONE       */
    int32_t scores_location;  /* KASA_MEM_HOST | KASA_MEM_DEVICE                                        */
    float gpu_ms;             /* out: device time of this batch's kernels (CUDA events)                 */
    /* Sparse results (needs kasa_params.max_batch_hits > 0; the way to score against thousands of taxa, where the
     * dense matrix does not fit): the (taxon index, score) pairs of read r are hit_taxon / hit_score
     * [hit_begin[r] .. hit_begin[r] + hit_count[r]).  Caller buffers, all in `hits_location`; hit_taxon / hit_score
     * hold hit_capacity entries, hit_begin / hit_count n_reads entries.  All NULL = not wanted.          */
    uint32_t *hit_taxon;
    float *hit_score;
    uint64_t *hit_begin;
    uint32_t *hit_count;
    uint64_t hit_capacity;
    uint64_t n_hits;          /* out                                                                    */
    int32_t hits_location;
} kasa_batch_result;

/* Replaces, per batch: Read::convertAndSort (Read.hpp:763-827) + Compare::sortInputAndCheckInvalidkMers_sta
 * (Compare.hpp:1074-1260) + Compare::compareWithDatabase (Compare.hpp:679-1069) + the score matrix
 * (Compare.hpp:3192).  Counts are accumulated inside ctx across batches.                               */
int kasa_identify_batch(kasa_ctx *ctx, const kasa_read_batch *in, kasa_batch_result *out);

/* The same work, pipelined (what the reference gets from its reader / worker threads, Compare.hpp:2983-3440):
 * kasa_identify_submit() queues the batch's host->device copies on a copy stream and its kernels on the compute stream and
 * returns without waiting for the device when the batch is host memory (its k-mer total is computed from seq_len on the
 * host); kasa_identify_wait() completes the OLDEST submitted batch (its scores are copied back on a second copy stream,
 * overlapping later batches' kernels and input copies) and fills the `out` given at submit time.  At most three batches are
 * in flight; `in` itself may go away when submit returns, but the host buffers it points to, `out` and out->scores must
 * stay valid until the matching wait.  Typical loop: submit(0); submit(1); for i >= 2 { submit(i); wait(i-2); }  then
 * wait the rest.  kasa_identify_batch == submit + wait.                                                              */
int kasa_identify_submit(kasa_ctx *ctx, const kasa_read_batch *in, kasa_batch_result *out);
int kasa_identify_wait(kasa_ctx *ctx);

/* Page-locked host memory for the buffers of a batch.  kasa_identify_submit() is only asynchronous when the host buffers it
 * copies from and into (bases, seq_start, seq_len, read_first_seq, scores, hit_*) are page-locked: from pageable memory the CUDA
 * copies are staged synchronously and every submit waits for its own batch (results are the same, the overlap is lost).  The
 * reference has no counterpart (its batches live in std::vector / STXXL memory, Compare.hpp:2983-3110); a caller that keeps
 * its own buffers may cudaHostRegister them instead.                                                                    */
int kasa_host_alloc(void **ptr, uint64_t bytes);
int kasa_host_free(void *ptr);

/* Replaces: the per-thread count reduction (Compare.hpp:3445-3454).  count_all / count_unique are
 * [numK][n_taxa], level 0 = k_high (kASA _aOfK, kASA.hpp:299-302).  Resets the accumulators
 * (Compare.hpp:3748-3752).                                                                             */
int kasa_identify_finish(kasa_ctx *ctx, double *count_all, uint64_t *count_unique, uint64_t *n_kmers_total);

/* --------------------------------------------------------------------------------------------------- */
/* identify against a PREFIX-SHARDED index (SURVEY.md §8e mode 2: one context per GPU, every context holds the
 * index entries of one k-mer prefix range AND its own read batches).
 *
 * Replaces: the reference's "index does not fit memory" route — Trie::GetIndexRange per query (Trie.hpp:494-520)
 * into an STXXL-paged index file (Compare.hpp:3130-3160): here the range owner is another GPU and the pages are
 * reached over NVLink.  There is no staging all-to-all: after a context has sorted its batch, the owner of a prefix
 * range runs the lookup kernel directly on the peer's sorted run (peer loads) and stores the 32-byte match records
 * straight into the peer's record array (peer stores); the origin then scores its own reads as usual.  Host-side
 * protocol per batch (the host exchanges the two small structs, e.g. with torch.distributed / MPI):
 *     every rank:  kasa_shard_sort(ctx, in, &runs[rank])        -> all-gather `runs`
 *     every rank:  kasa_shard_lookup(ctx, runs)                  -> barrier
 *     every rank:  kasa_shard_score(ctx, &out)                   -> barrier before the next kasa_shard_sort
 * Shard boundaries never split a group of k-mers sharing min(6, k_low) leading letters, so every level-k taxa set
 * of SURVEY.md §3.1 lives on one shard and the results equal the unsharded ones bit for bit (KASA_SCORE_ORDERED and NONE;
 * KASA_SCORE_EXACT replays one global flush order and is rejected).                                            */

#define KASA_MAX_SHARDS 16
#define KASA_SHARD_NBUF 9          /* sort ping-pong w0[2], w1[2], v[2], the match-record array, the record stage, the overflow list */

typedef struct {
    uint64_t pid;                  /* owner process: peers in the same process use `ptr`, others open `handle`  */
    int32_t  device;
    int32_t  reserved;
    uint64_t ptr[KASA_SHARD_NBUF]; /* device addresses in the owner process (0 = buffer not used)               */
    unsigned char handle[KASA_SHARD_NBUF][64];   /* cudaIpcMemHandle_t                                           */
} kasa_shard_mem;

typedef struct {
    uint64_t n_kmers;              /* k-mers of the origin's batch                                              */
    int32_t  cur;                  /* which half of the ping-pong holds the sorted run                          */
    int32_t  reserved;             /* set by kasa_shard_sort, read by the owners in kasa_shard_lookup: bit 0 = the origin scores this batch
                                      straight from two-mask match records; pass the struct on unchanged                        */
    uint64_t begin[KASA_MAX_SHARDS + 1];   /* sorted positions: [begin[r], begin[r+1]) belongs to shard r       */
} kasa_shard_runs;

/* Host-only: cut a sorted index (file records as for kasa_index_load) into n_shards ranges of about equal size on
 * boundaries of min(6, k_low)-letter prefixes.  entry_begin[n_shards+1] = record offsets; prefix_begin[n_shards+1] =
 * the first 30-bit prefix (top six letters, the `_trie` key of Trie.hpp:371-386) of each shard, prefix_begin[0] = 0,
 * prefix_begin[n_shards] = 2^30.                                                                               */
int kasa_shard_plan(const void *entries, uint64_t n, int bits, int k_low, int n_shards,
                    uint64_t *entry_begin, uint32_t *prefix_begin);

/* The same plan for the index just built on this context (kasa_build_finish); keeps only shard `rank` of it in
 * HBM as the context's index (like kasa_index_from_build).                                                     */
int kasa_index_from_build_shard(kasa_ctx *ctx, const uint32_t *taxid_of_idx, uint32_t n_taxa, int rank, int n_shards,
                                uint64_t *entry_begin, uint32_t *prefix_begin);

/* Declare this context shard `rank` of n_shards (its index = the records [entry_begin[rank], entry_begin[rank+1]),
 * loaded with kasa_index_load) and export its exchange buffers.                                                */
int kasa_shard_init(kasa_ctx *ctx, int rank, int n_shards, const uint32_t *prefix_begin, kasa_shard_mem *mine);
/* Map the peers' buffers (all[n_shards], all[rank] = mine).                                                    */
int kasa_shard_connect(kasa_ctx *ctx, const kasa_shard_mem *all);
int kasa_shard_sort(kasa_ctx *ctx, const kasa_read_batch *in, kasa_shard_runs *mine);
int kasa_shard_lookup(kasa_ctx *ctx, const kasa_shard_runs *all);
int kasa_shard_score(kasa_ctx *ctx, const kasa_shard_runs *all, kasa_batch_result *out);

/* --- build of a prefix-sharded index (BASELINE config C4): every rank pushes its part of the reference sequences
 * with kasa_build_push; the owners of the prefix ranges then pull "their" records from every peer's sorted brick
 * (contiguous peer copies over NVLink), and sort / dedup / run-length encode them locally.  The shards, written one
 * after the other in rank order, are byte for byte the file `kASA build` writes (index and `_trie`); the frequency
 * tables add up.  Replaces Build::mergeTemporaries (Build.hpp:360-477): the k-way merge of sorted bricks becomes
 * "route by prefix, then sort".  Protocol:
 *     kasa_build_shard_local(ctx, hist, &n)      -> all-reduce(sum) hist over the ranks
 *     kasa_shard_plan_hist(hist, ...)            (host only, same result on every rank)
 *     kasa_shard_init / all-gather / kasa_shard_connect
 *     kasa_build_shard_runs(ctx, &runs[rank])    -> all-gather runs
 *     kasa_build_shard_pull(ctx, runs)           -> barrier
 *     kasa_build_finish / _fetch / _frequency    as for a single GPU, now for this rank's shard                  */
#define KASA_SHARD_HIST_BITS 15
/* Sorts and dedups the local brick; hist[1 << KASA_SHARD_HIST_BITS] = local records per leading-15-bit bucket.   */
int kasa_build_shard_local(kasa_ctx *ctx, uint64_t *hist, uint64_t *n_local);
/* Host only: prefix_begin[n_shards+1] with about equal record counts per shard, on min(6, k_low)-letter boundaries. */
int kasa_shard_plan_hist(const uint64_t *hist, int k_low, int n_shards, uint32_t *prefix_begin);
int kasa_build_shard_runs(kasa_ctx *ctx, kasa_shard_runs *mine);
int kasa_build_shard_pull(kasa_ctx *ctx, const kasa_shard_runs *all);

/* --------------------------------------------------------------------------------------------------- */
/* build                                                                                                */

typedef struct {
    uint32_t n_seqs;          /* prepared strings: cleaned ('Z'), marker 'X'*3*(K-1) already appended where
                                 the reference appends it (Read.hpp:2323-2333, 2529-2538)               */
    int32_t  location;
    const char *bases;        /* ASCII incl. 'X' / 'Z'                                                  */
    const uint64_t *seq_start;/* [n_seqs]                                                               */
    const uint32_t *seq_len;  /* [n_seqs]                                                               */
    const uint32_t *taxid;    /* [n_seqs]                                                               */
} kasa_seq_batch;

/* Replaces: Read::dnaTokMers + Build::addToInt (Read.hpp:1991-2139, Build.hpp:116-130).
 * The records collect in a device brick of max_batch_kmers records.  A full brick is sorted and deduplicated in place; if it
 * is still more than half full of distinct records it leaves for host memory (the reference's spill to a temporary file,
 * Build.hpp:328-352) and a new brick starts, so the size of an index is bounded by host memory, not by max_batch_kmers or HBM.
 * One push may not yield more than max_batch_kmers windows.                                                              */
int kasa_build_push(kasa_ctx *ctx, const kasa_seq_batch *in);

/* Replaces: the reading of an existing index in `update` and `merge` (Build.hpp:153-300 `updateIndex`, Update.hpp,
 * Read.hpp:3180-3243): the records of an index file (as for kasa_index_load) join the build like pushed sequences do, so
 *   update = kasa_build_push_index(old index) + kasa_build_push(new sequences) + kasa_build_finish,
 *   merge  = kasa_build_push_index(first) + kasa_build_push_index(second) + kasa_build_finish
 * (the sort + dedup of kasa_build_finish is the merge; bricks spill to host memory as for any build).                 */
int kasa_build_push_index(kasa_ctx *ctx, const void *entries, uint64_t n, int bits);

/* Replaces: Build::IntToExtPart + mergeTemporaries (Build.hpp:305-477), Trie::SaveToStxxlVec
 * (Trie.hpp:366-394).  Sorts by (kmer, taxid), removes duplicates, run-length encodes the 6-letter
 * prefixes.  Outputs the counts; fetch the data with kasa_build_fetch.  When bricks were spilled they are merged here, one
 * range of six-letter prefixes at a time through the device brick; the merged index then stays in host memory
 * (kasa_build_fetch / _frequency / kasa_index_from_build work as usual, kasa_index_from_build needs it to fit the device). */
int kasa_build_finish(kasa_ctx *ctx, uint64_t *n_entries, uint64_t *n_trie);

/* Copy out the built index in file layout: entries = packedBigPair / packedLargePair records,
 * trie = packedBigPairTrie {u64 count, u32 prefix30} records (packedPairs.hpp:157-166).  Either may be
 * NULL.  Replaces the STXXL export (Build.hpp:457-470).                                                 */
int kasa_build_fetch(kasa_ctx *ctx, void *entries, void *trie);

/* Replaces: kASA::GetFrequencyK (kASA.hpp:449-575).  freq is [n_taxa][highest_k]; column j counts the
 * entries of that taxon whose letter at position highest_k - j is not '^' (30).                        */
int kasa_build_frequency(kasa_ctx *ctx, const uint32_t *taxid_of_idx, uint32_t n_taxa, uint64_t *freq);

/* --------------------------------------------------------------------------------------------------- */
/* stage-level entry points (parity tests drive each kernel against the oracle through these)           */

/* k-mers of a read batch in emission order (Read.hpp:84-220): kmers_lo/hi [n_kmers], ids = read id.
 * Buffers are host memory sized max_batch_kmers.                                                       */
int kasa_dbg_extract(kasa_ctx *ctx, const kasa_read_batch *in, uint64_t *kmers_lo, uint64_t *kmers_hi,
                     uint32_t *read_id, uint64_t *n_kmers);

/* LSD radix sort of (key, value) pairs on bits [bit_lo, bit_hi) of the key; host buffers, in place.    */
int kasa_dbg_sort_pairs(kasa_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t n, int bit_lo, int bit_hi);

/* Synthetic workload generators (device side of kasa_b200/synth.py; used by bench.py to create the
 * C2 workload in HBM).  kasa_synth_build_reference pushes `n_taxa` genomes of `genome_len` bases
 * (taxid 1000+t) through the build path.  kasa_synth_reads writes n_reads x read_len ASCII bases.       */
int kasa_synth_build_reference(kasa_ctx *ctx, uint64_t seed, uint32_t n_taxa, uint64_t genome_len);
/* the same for taxa first_taxon .. first_taxon + n_taxa - 1 only (ranks of a sharded build push disjoint taxa)  */
int kasa_synth_build_reference_range(kasa_ctx *ctx, uint64_t seed, uint32_t first_taxon, uint32_t n_taxa, uint64_t genome_len);
int kasa_synth_reads(kasa_ctx *ctx, uint64_t genome_seed, uint64_t read_seed, uint64_t first_read,
                     uint32_t n_reads, uint32_t n_taxa, uint64_t genome_len, uint32_t read_len,
                     uint32_t subst_ppm, uint32_t n_ppm, char *out, int out_location);

#ifdef __cplusplus
}
#endif
#endif /* KASA_B200_H */
