// build.cuh — index construction after the sort: duplicate removal, prefix run lengths, frequencies,
// file-layout (de)serialisation.
//   Build::IntToExtPart std::unique (Build.hpp:320-326), mergeTemporaries dedup (Build.hpp:423-428)
//   Trie::SaveToStxxlVec (Trie.hpp:366-394), kASA::GetFrequencyK (kASA.hpp:449-575)
//   record layouts packedPairs.hpp:100-166
#pragma once
#include "common.cuh"

namespace kasa {

// flags[i] = 1 if record i is valid (not the ~0 sentinel) and differs from record i-1; flags has n+1 slots
void launch_unique_flags(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint64_t *flags, uint64_t *launches);
// after exclusive scan of flags: out[pos[i]] = rec[i] where pos[i+1] != pos[i]
void launch_compact(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                    const uint64_t *pos, uint64_t *out_lo, uint64_t *out_hi, uint32_t *out_val, uint64_t *launches);

// prefix30 run starts: flags[i] = 1 where the first 6 letters change (n+1 slots)
void launch_prefix_flags(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, uint64_t *flags,
                         uint64_t *launches);
// after scan: trie record j = {count u64, prefix u32} packed 12 B (packedPairs.hpp:157-166)
void launch_trie_records(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, const uint64_t *pos,
                         uint64_t n_runs, uint64_t *run_start_tmp /* n_runs+1 */, uint32_t *trie_words, uint64_t *launches);

// taxid (u32) -> dense index (u16) through the sorted table; err_flag set if a taxid is unknown
void launch_map_taxids(cudaStream_t s, const uint32_t *taxid, uint64_t n, const uint32_t *sorted_taxid, const uint16_t *sorted_idx,
                       uint32_t n_sorted, uint16_t *dense, uint32_t *err_flag, uint64_t *launches);

// freq[t][j] += entries of taxon t whose letter at position K-j is not '^'
void launch_frequency(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint16_t *dense, uint64_t n, int K,
                      uint32_t n_taxa, unsigned long long *freq, uint64_t *launches);

// structure-of-arrays <-> file records (12 B / 20 B, as 3 / 5 u32 words)
void launch_pack_records(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint32_t *words, uint64_t *launches);
void launch_unpack_records(cudaStream_t s, const uint32_t *words, uint64_t n, bool wide, uint64_t *key_lo, uint64_t *key_hi,
                           uint32_t *val, uint64_t *launches);
// err_flag set if the records are not sorted by (kmer, taxid)
void launch_check_sorted(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint32_t *err_flag, uint64_t *launches);

// halved index (6-byte records + prefix table, Shrink.hpp:78-143) -> full keys: counts and prefixes out of the table's records, then
// (after the exclusive scan of starts[0..n_trie]) key = prefix << 30 | lower six letters; err_flag 2: taxon index or letters out of range
void launch_half_trie_split(cudaStream_t s, const uint32_t *trie_words, uint64_t n_trie, uint64_t *starts, uint32_t *prefix, uint64_t *launches);
void launch_half_expand(cudaStream_t s, const uint16_t *rec, uint64_t n, const uint64_t *starts, const uint32_t *prefix, uint64_t n_trie,
                        uint32_t n_taxa, uint64_t *key_lo, uint16_t *tax, uint32_t *err_flag, uint64_t *launches);

}  // namespace kasa
