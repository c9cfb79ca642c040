// score.cuh — per-read scores and per-taxon counts from the match records (lookup.cuh).
//   Compare.hpp:516-532 scoreMatchNonAVX: score[read][taxon] += float(w(k) * (1.f / |taxa_k|)) once per hit instance and level,
//   vCount_all += hits / |taxa_k|, vCount_unique += hits when |taxa_k| == 1; flush order Compare.hpp:917-955, 1032-1041.
//
// The reference's per-read score is a float32 running sum, so its value depends on the order of the additions
// (SURVEY.md §3.1 "Floating-point order").
// PRESORTED BATCHES (the common case: extract.cu emits every read of up to presort_limit() k-mers in ascending k-mer order, so its
// match records already lie in summation order) of content files with <= 64 taxa are scored by ONE kernel straight from the
// records: launch_score_sorted -> score_dense_kernel (warp per read, lane per taxon, uniform control flow; two-mask records).
// Everything else — longer reads, more than 64 taxa, sparse results, KASA_SCORE_EXACT — goes through two stages:
//   A  order_ops_kernel   warp per read: the read's match records are ordered by their rank in the batch's sorted k-mer
//                         order (KASA_SCORE_ORDERED: k-mer ascending, k ascending — the reference's order whenever its
//                         level-by-level flushes follow the k-mer order) or by the reference's flush positions
//                         (KASA_SCORE_EXACT), and written out as a compact stream of 4-byte words in that order;
//   B  accumulate_kernel  thread per read (<= 64 taxa) or warp per read with hashed cells (any number of taxa): the
//                         stream is replayed with sequential float32 adds, the counts are exact integers / 2^-32 fixed point.
#pragma once
#include "common.cuh"
#include "lookup.cuh"

namespace kasa {

// ---- op stream of a read: a sequence of groups  [header word][extra header words x NX][n pair words]
// ordered stream (KASA_SCORE_ORDERED), one group per k-mer instance with a hit:
//   header  bit 31 = 1 | n (10 bits, 21..30) = |taxa_klow| = number of pair words | nibbles 1..5 (bits 1..20) = |taxa_j| of levels j = 1..5
//   extra   NX = ceil((num_k - 6) / 8) words with the nibbles of the levels j = 6 .. (eight per word); constant per context
//   pair    bit 31 = 0 | levels h (5 bits, 16..20) | taxon (16): the taxon matches the levels j = 0 .. h-1 of the instance
//   The group adds, pair by pair and for j = 0 .. h-1 in this order, float(w_j * (1.f / n_j)) to score[read][taxon].
// exact stream (KASA_SCORE_EXACT), one group per (instance, level) in the reference's flush order:
//   header  bit 31 = 1 | n (10 bits) = |taxa_j| = number of pair words | level j (bits 0..4);  NX = 0;  pair = taxon
// escapes: n == kOpNExt and nibble == 15 mean "the value (u16) is the next one of the read's ext stream".
constexpr uint32_t kOpHeader = 1u << 31;
constexpr uint32_t kOpNExt = 1023u;
constexpr uint32_t kOpNibExt = 15u;
__host__ __device__ __forceinline__ int ops_extra_words(int num_k) { return num_k > 6 ? (num_k - 6 + 7) / 8 : 0; }
// mask stream (KASA_SCORE_ORDERED with <= 64 taxa and <= 6 levels — the default k = 12..7 against a small content file): 16-byte groups
//   [header | |taxa_klow| << 24 | nup << 17][taxa mask lo][taxa mask hi][up pair 0]  + blocks of four further up pairs when nup > 1,
//   an up pair = taxon | levels << 16 for a taxon that matches more than the k_low level (score.cu accumulate_mask_kernel)
inline bool ops_mask_stream(const LevelParams &lp) { return lp.n_taxa <= kMaskTaxa && lp.num_k <= 6; }

struct OpsWorkspace {
    uint32_t *ops;                    // op streams, one contiguous run per read (start multiple of 4)
    uint64_t ops_cap;
    uint16_t *ext;                    // escaped level sizes
    uint64_t ext_cap;
    unsigned long long *cursors;      // [0] ops used, [1] ext used (reset per chunk of reads)
    uint64_t *ops_start;              // [reads]
    uint32_t *ops_count;              // [reads]
    uint64_t *ext_start;              // [reads]
};

struct OverflowPeers {
    const uint32_t *base[16];         // overflow array of every shard (own shard: the context's own array)
};

struct HitsOut {                      // sparse per-read results (any number of taxa); all device pointers, may be null
    uint32_t *taxon;
    float *score;
    uint64_t *begin;                  // [reads]
    uint32_t *count;                  // [reads]
    unsigned long long *cursor;       // [1]
    uint64_t cap;
};

// Reads [r0, r1).  trig == nullptr: KASA_SCORE_ORDERED; else trig[instance][num_k] = flush position of (instance, level) (exact.cuh).
// max_inst = largest number of k-mer instances of one read in this batch; n_kmers = k-mers in the batch (rank range).
// err_flag |= 32: op / ext workspace full; |= 64: a read is too long for KASA_SCORE_EXACT.
// counts = true (content files of <= 64 taxa, see scores_counted_by_order_ops): the per-taxon counts of these reads are accumulated here;
// emit = false (only with counts): no op streams are written (profile-only runs need no kernel B).
void launch_order_ops(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t r0, uint32_t r1, uint64_t n_kmers, uint32_t max_inst,
                      LevelParams lp, OverflowPeers peers, OpsWorkspace ws, const uint32_t *trig, bool counts, bool emit, unsigned long long *cnt_fx,
                      unsigned long long *cnt_unique, uint32_t *err_flag, uint64_t *launches);
// which of the two kernels takes the counts: A for small taxonomies scored densely, B (hashed cells) otherwise
inline bool scores_counted_by_order_ops(const LevelParams &lp, bool sparse_hits) { return lp.n_taxa <= kMaskTaxa && !sparse_hits; }

// Reads [r0, r1).  scores: dense [reads][n_taxa] or nullptr; hits: sparse output or all-null.  cnt_fx / cnt_unique: [num_k][n_taxa], level 0 = k_low.
// err_flag |= 128: sparse output full; |= 256: a read hits more distinct taxa than the hashed cells hold.
void launch_accumulate(cudaStream_t s, OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, bool exact_stream, float *scores, HitsOut hits,
                       unsigned long long *cnt_fx, unsigned long long *cnt_unique, uint32_t *err_flag, uint64_t *launches);

// Presorted batches (every read has <= presort_limit() k-mers, extract.cuh) of a content file with <= 64 taxa: scores (dense
// [reads][n_taxa] or nullptr = profile only) and counts straight from the match records, which lie in summation order already;
// replaces launch_order_ops + launch_accumulate.  lanes_hint (KASA_TUNE_SCORE_LANES): 0 / 64 = score_dense_kernel (the default; any
// number of levels), 32 = score_blocks_kernel, 100 = score_items_kernel, 1 / 2 / 4 = score_sorted_kernel (these three: <= 6 levels,
// mask records; kept as the measured alternatives of DESIGN.md §4).  err_flag |= 512: a record is not in the format the kernel reads.
void launch_score_sorted(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp, OverflowPeers peers,
                         float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique, int lanes_hint, uint32_t *err_flag, uint64_t *launches);
// which record format that kernel reads: the dense organisation (lanes_hint 0 / 64) wants two-mask records (lookup.cuh), the others mask records
inline bool score_sorted_two_masks(int lanes_hint, const LevelParams &lp) { return lanes_hint == 0 || lanes_hint == 64 || lp.num_k > 6; }

// largest number of instances of any read -> *out (device u32)
void launch_max_instances(cudaStream_t s, const uint64_t *read_koff, uint32_t n_reads, uint32_t *out, uint64_t *launches);

}  // namespace kasa
