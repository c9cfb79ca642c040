// shard.cuh — prefix-sharded index (SURVEY.md §8e mode 2): where sorted k-mers change owner.
//   The owner of a k-mer is the shard whose 30-bit prefix range holds the k-mer's first six letters — the key of
//   the reference's prefix table (Trie.hpp:371-386, Trie::GetIndexRange :494-520).
#pragma once
#include "common.cuh"

namespace kasa {

// begin[r] = first position of the sorted keys whose prefix30 >= prefix_begin[r], r = 0..n_shards (begin[n_shards] = n).
// The keys only need to be ordered on their leading 30 bits.
void launch_run_bounds(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K,
                       const uint32_t *prefix_begin, int n_shards, uint64_t *begin, uint64_t *launches);

// plan for a sorted index in HBM: shard r starts at the first entry of the aligned prefix group holding entry
// floor(n * r / n_shards); align_letters = min(6, k_low).  entry_begin / prefix_begin have n_shards + 1 slots.
void launch_shard_plan(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int align_letters,
                       int n_shards, uint64_t *entry_begin, uint32_t *prefix_begin, uint64_t *launches);

// hist[b] = number of sorted keys whose leading `bits` bits equal b, b < 2^bits (bits <= 30)
void launch_prefix_hist(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int bits, uint64_t *hist,
                        uint64_t *launches);

}  // namespace kasa
