// lookup.cuh — sorted read k-mers against the sorted (k-mer, taxon) index for every k in
// [k_low, k_high] at once, and per-read / per-taxon accumulation.
//   Compare::sortInputAndCheckInvalidkMers_sta range assignment (Compare.hpp:1074-1260, Trie.hpp:494-520)
//   Compare::compareWithDatabase (Compare.hpp:679-1069), scoreMatchNonAVX (Compare.hpp:516-532)
// restated as the set semantics of SURVEY.md §3.1 (what the state machine computes), not transcribed.
#pragma once
#include "common.cuh"

namespace kasa {

struct IndexDev {
    const uint64_t *key_lo;
    const uint64_t *key_hi;   // nullptr for 64-bit k-mers
    const uint16_t *tax;      // dense taxon index (1-based: position in the content file)
    uint64_t n;
};

struct LevelParams {
    int K;         // highest_k (letters per k-mer)
    int k_high, k_low;
    int num_k;     // k_high - k_low + 1
    int n_taxa;    // incl. index 0 ("non_unique")
    int hb;        // bits per level field in a match record (3, 4 or 5)
};

constexpr int kLookupTile = 1024;      // most sorted queries per CTA (see lookup_tile_queries)
constexpr int kLookupBlock = 256;
// match record = 4 x u64 = one 32-byte sector for 3- and 4-bit level fields, 8 x u64 = two sectors for 5-bit fields
// (16..31 levels, e.g. k = 25..7: 12 fields per word)
constexpr int kRecWords = 4;
template <int HB>
constexpr int kRW = HB == 5 ? 8 : 4;
__host__ __device__ inline int record_words(int hb) { return hb == 5 ? 8 : 4; }

// bits needed to store a level count 0..num_k, and how many taxa one record can describe
inline int record_level_bits(int num_k) { return num_k <= 7 ? 3 : (num_k <= 15 ? 4 : 5); }
// 3-bit records keep their fourth word for the level sizes (see lookup.cu); the lookup kernel's taxa set is a 64-bit mask
inline int record_taxa_capacity(int hb) { return hb == 3 ? 63 : 64; }

// per-tile index spans: part[2t], part[2t+1] = [first entry sharing the sorted prefix of the tile's first query,
// one past the last entry sharing the sorted prefix of its last query).  The queries only need to be ordered
// on their bits >= sorted_from_bit: every query searches the staged span on its own, so the order inside a
// group of equal leading bits does not matter — this is what lets identify sort 32 of the 60 (125) key bits.
uint32_t lookup_tile_queries(uint64_t n_queries, uint64_t n_index);
void launch_partition(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, uint64_t n, IndexDev idx,
                      LevelParams lp, uint64_t *part, uint32_t tq, int sorted_from_bit, uint64_t *launches);

// one 32-byte match record per k-mer instance, written at rec[inst[i]] (instance = emission order):
// field t-1 = number of levels (from k_low upwards) at which taxon t is in taxa_k(prefix_k(q))
void launch_lookup(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n,
                   IndexDev idx, LevelParams lp, const uint64_t *part, uint64_t *rec, uint32_t tq, uint64_t *launches);

// inst == nullptr: record i is written to rec[i] (sorted position; the prefix-sharded path stages records this way)
// rec[inst[i]] = staged[i] for i < n (staged may be peer memory)
void launch_scatter_records(cudaStream_t s, const uint64_t *staged, const uint32_t *inst, uint64_t n, uint64_t *rec, int rec_words,
                            uint64_t *launches);

// per read: fp32 scores in canonical order (instance order, k ascending) and the per-taxon per-level
// counts: cnt_fx [num_k][n_taxa] fixed point 2^-32 (sum of 1/|taxa|), cnt_unique [num_k][n_taxa].
// Level j = 0 is k_low.  scores may be nullptr (profile only).
void launch_per_read(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                     float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique, uint64_t *launches);

// 3-bit records (per_read.cu): warp per read, lane per taxon
void launch_per_read_taxon(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                           float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique);

// 4- and 5-bit records (per_read.cu): the same organisation with the level sizes counted from the fields
void launch_per_read_wide(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                          float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique);

// read_koff[r] = koff[read_first_seq[r]] (or koff[r])
void launch_read_offsets(cudaStream_t s, const uint64_t *koff, const uint32_t *read_first_seq, uint32_t n_reads,
                         uint64_t *read_koff, uint64_t *launches);

// total[ik][t] += fx[j][t] * 2^-32 (double) and unique likewise, flipping the level order to ik = 0 <-> k_high
void launch_fold_counts(cudaStream_t s, unsigned long long *cnt_fx, unsigned long long *cnt_unique, LevelParams lp,
                        double *total_all, unsigned long long *total_unique, uint64_t *launches);

}  // namespace kasa
