// lookup.cuh — sorted read k-mers against the sorted (k-mer, taxon) index for every k in
// [k_low, k_high] at once, and per-read / per-taxon accumulation.
//   Compare::sortInputAndCheckInvalidkMers_sta range assignment (Compare.hpp:1074-1260, Trie.hpp:494-520)
//   Compare::compareWithDatabase (Compare.hpp:679-1069), scoreMatchNonAVX (Compare.hpp:516-532),
//   markTaxIDs / sBitArray distinct-taxa sets (Compare.hpp:396-441, BitArray.hpp:18-147)
// restated as the set semantics of SURVEY.md §3.1 (what the state machine computes), not transcribed.
#pragma once
#include "common.cuh"

namespace kasa {

// Index in HBM, structure of arrays.  The top bit of an entry's most significant key word (bit 63 of key_hi for
// 128-bit k-mers, of key_lo otherwise; k-mers use 125 / 60 bits) is the entry's TAXON-FIRST flag: set when no earlier
// entry of the same k_low-letter group has the same taxon.  The flagged entries of a group are exactly its distinct
// taxa — the reference's sBitArray (BitArray.hpp:18-147) precomputed once at index load for any number of taxa.
struct IndexDev {
    const uint64_t *key_lo;
    const uint64_t *key_hi;   // nullptr for 64-bit k-mers
    const uint16_t *tax;      // dense taxon index (1-based: position in the content file)
    uint64_t n;
    const uint64_t *gmask;    // content files of <= 64 taxa: bit t-1 set for every taxon t of the entry's k_low-letter group (same word for
                              // all entries of a group) — markTaxIDs of level k_low (Compare.hpp:396-441) done once at index load; else nullptr
};
constexpr uint64_t kKeyFlag = 1ull << 63;

struct LevelParams {
    int K;         // highest_k (letters per k-mer)
    int k_high, k_low;
    int num_k;     // k_high - k_low + 1 (<= 25)
    int n_taxa;    // incl. index 0 ("non_unique"), <= 65535
};

constexpr int kLookupTile = 1024;      // most sorted queries per CTA (see lookup_tile_queries)
constexpr int kLookupBlock = 256;

// ---- match record: 4 x u64 = one 32-byte sector per k-mer instance, written at rec[instance] (emission order)
//   w0   bits  0..31  rank: position of the instance in the batch's sorted k-mer order (fixes the fp32 summation order)
//        bits 32..47  count = |taxa_klow| = number of (taxon, levels) pairs
//        bits 48..52  hmax  = largest number of levels any taxon matches
//        bit  53      the pairs live in the overflow list (count > kRecInline)
//        bits 54..57  shard that owns the overflow list (prefix-sharded index)
//        bit  58      mask format (content files of <= 64 taxa)
//   list format, count <= kRecInline:  w1..w3 hold the pairs, 21 bits each, three per word: taxon (16) | levels (5);
//                         levels = number of levels from k_low upwards at which the taxon is in taxa_k(prefix_k(q))
//   list format, otherwise:            w1 = first entry of the list in the owner's overflow array (u32 per pair: taxon | levels << 16)
//   mask format:          w1 = bit t-1 set for every taxon t of taxa_klow (any set size), w2..w3 = up to kRecUpPairs pairs for the taxa
//                         that match more than one level (the others match exactly the k_low level)
//   two-mask format (bit 59; only on request, for the dense scoring kernel of presorted batches): w1 as above, w2 = the same kind
//                         of mask for level k_low + 1, w3 = up to three pairs for the taxa that match three or more levels.  Against
//                         a saturated index taxa reach k_low + 1 by chance in every fourth k-mer; as a mask they cost the scorer one
//                         more predicated add per lane instead of a chain per pair.
constexpr int kRecWords = 4;
constexpr int kRecInline = 9;
constexpr int kPairBits = 21;
constexpr uint32_t kPairMask = (1u << kPairBits) - 1u;
constexpr uint64_t kRecOverflowBit = 1ull << 53;
constexpr uint64_t kRecMaskBit = 1ull << 58;
constexpr uint64_t kRecMask2Bit = 1ull << 59;
constexpr int kRecUpPairs = 6;
constexpr int kMaskTaxa = 65;          // n_taxa (incl. index 0) up to which records use the mask format

__host__ __device__ __forceinline__ uint32_t rec_rank(uint64_t w0) { return (uint32_t)w0; }
__host__ __device__ __forceinline__ uint32_t rec_count(uint64_t w0) { return (uint32_t)(w0 >> 32) & 0xffffu; }
__host__ __device__ __forceinline__ uint32_t rec_hmax(uint64_t w0) { return (uint32_t)(w0 >> 48) & 31u; }
__host__ __device__ __forceinline__ bool rec_overflow(uint64_t w0) { return (w0 & kRecOverflowBit) != 0; }
__host__ __device__ __forceinline__ uint32_t rec_owner(uint64_t w0) { return (uint32_t)(w0 >> 54) & 15u; }
__host__ __device__ __forceinline__ bool rec_is_mask(uint64_t w0) { return (w0 & kRecMaskBit) != 0; }
__host__ __device__ __forceinline__ bool rec_is_mask2(uint64_t w0) { return (w0 & kRecMask2Bit) != 0; }

// append-only side list for instances whose taxa set does not fit a record (reset per batch)
struct OverflowBuf {
    uint32_t *data;
    unsigned long long *cursor;
    uint64_t cap;
};

// per-tile index spans: part[2t], part[2t+1] = [first entry sharing the sorted prefix of the tile's first query,
// one past the last entry sharing the sorted prefix of its last query).  The queries only need to be ordered
// on their bits >= sorted_from_bit: every query searches the staged span on its own, so the order inside a
// group of equal leading bits does not matter — this is what lets identify sort 32 of the 60 (125) key bits.
uint32_t lookup_tile_queries(uint64_t n_queries, uint64_t n_index);
void launch_partition(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, uint64_t n, IndexDev idx,
                      LevelParams lp, uint64_t *part, uint32_t tq, int sorted_from_bit, uint64_t *launches);

// one match record per k-mer instance, written at rec[inst[i]]; rank = rank_base + i.
// inst == nullptr: record i is written to rec[i] (sorted position; the prefix-sharded path stages records this way).
// err_flag |= 8 when the overflow list is full.
void launch_lookup(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n, uint64_t rank_base,
                   IndexDev idx, LevelParams lp, const uint64_t *part, uint64_t *rec, uint32_t tq, OverflowBuf ovf, uint32_t owner,
                   uint32_t *err_flag, uint64_t *launches, bool two_masks = false);

// rec[inst[i]] = staged[i] for i < n (staged may be peer memory)
void launch_scatter_records(cudaStream_t s, const uint64_t *staged, const uint32_t *inst, uint64_t n, uint64_t *rec, uint64_t *launches);

// sets the taxon-first flags of a freshly loaded index (keys must be unflagged); k_low-letter groups
void launch_index_flags(cudaStream_t s, uint64_t *key_lo, uint64_t *key_hi, const uint16_t *tax, uint64_t n, int K, int k_low,
                        uint64_t *launches);

// group masks of a freshly loaded index (IndexDev::gmask), in three steps:
//   launch_group_heads   out[e] = 1 if entry e starts a k_low-letter group, else 0            (out holds n + 1 words)
//   exclusive_scan_u64   in place: out[e] = groups started before e, out[n] = number of groups (scan.cuh)
//   launch_group_masks   gm[groups] (zeroed by the caller) collects the taxon bits per group; inout[e] (the scan) becomes the mask of e's group
void launch_group_heads(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int k_low, uint64_t *out, uint64_t *launches);
void launch_group_masks(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint16_t *tax, uint64_t n, int K, int k_low,
                        uint64_t *gm, uint64_t *inout, uint64_t *launches);

// read_koff[r] = koff[read_first_seq[r]] (or koff[r])
void launch_read_offsets(cudaStream_t s, const uint64_t *koff, const uint32_t *read_first_seq, uint32_t n_reads,
                         uint64_t *read_koff, uint64_t *launches);

// total[ik][t] += fx[j][t] * 2^-32 (double) and unique likewise, flipping the level order to ik = 0 <-> k_high
void launch_fold_counts(cudaStream_t s, unsigned long long *cnt_fx, unsigned long long *cnt_unique, LevelParams lp,
                        double *total_all, unsigned long long *total_unique, uint64_t *launches);

}  // namespace kasa
