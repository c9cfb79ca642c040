// extract.cuh — codon translation + packed k-mer construction.
//   identify: Read::convert_dnaTokMer (Read.hpp:84-220), kASA::dnaToAminoacid (kASA.hpp:69-87),
//             kASA::aminoacidTokMer (kASA.hpp:333-379), reverseComplement (kASA.hpp:214-221),
//             read preparation (Read.hpp:612-695) done analytically on device.
//   build:    Read::dnaTokMers (Read.hpp:1991-2139).
#pragma once
#include "common.cuh"

namespace kasa {

// 2-bit packed sequences on device (KASA_SEQ_PACKED2 of include/kasa_b200.h)
struct PackedSeqs {
    const uint32_t *bases2;     // 16 bases per word, LSB first, codes A0 C1 T2 G3
    const uint32_t *nmask;      // 32 bases per word, bit = non-ACGT ('Z')
    const uint64_t *seq_start;  // [n_seqs] first base (multiple of 32)
    const uint32_t *seq_len;    // [n_seqs]
    uint32_t n_seqs;
};

struct KmerParams {
    int K;          // highest_k: 12 or 25
    int k_low;      // marker = 3*(K - k_low) 'X' (Read.hpp:1068-1078)
    int strands;    // 1 (forward frames) or 2 (--six)
    int one;        // 1: --one, only the first reading frame (Read.hpp:223-260): k-mers at every third base of the forward strand
};

// k-mers of one strand of a prepared string of `len` letters (read + marker): Read.hpp:36-57 calculatekMerCount
__host__ __device__ __forceinline__ uint64_t kmer_count(uint64_t len, int K, int one)
{
    if (one) { const uint64_t t = len / 3; return t > (uint64_t)(K + 1) ? t - K + 1 : 0; }
    return len > (uint64_t)(3 * K + 1) ? len - 3 * K + 1 : 0;
}

// 5-bit letter codes of a 512-entry codon table -> d_table (512 bytes of device memory owned by the context)
void upload_codon_table(const uint8_t table512[512], uint8_t *d_table, cudaStream_t s);

// packed start of each sequence = exclusive scan of roundup32(len); writes pstart[n+1] inputs for the scan
void launch_pack_sizes(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, uint64_t *pstart, uint64_t *launches);
// ASCII -> 2-bit + mask (cleaning of Read.hpp:657-675: anything but ACGTacgt is 'Z')
void launch_pack_ascii(cudaStream_t s, const char *ascii, const uint64_t *ascii_start, const uint32_t *seq_len,
                       uint32_t n_seqs, const uint64_t *pstart, uint32_t *bases2, uint32_t *nmask, uint64_t *launches);

// per-sequence k-mer counts (Read.hpp:36-57) into koff[0..n_seqs); scan afterwards
void launch_kmer_counts(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, KmerParams p, uint64_t *koff, uint64_t *launches);
// Reads of at most kPresortMax k-mers are emitted with their k-mers in ascending order (the order in which the reference scores
// them: extract.cu "presorted emission"); longer reads in the reference's emission order (forward strand, then reverse strand, per
// sequence).  A read = the sequences [read_first[r], read_first[r+1]) (read_first == nullptr: one sequence per read);
// max_inst = most k-mers of any read of the batch (sizes the shared memory).  key_hi may be nullptr for K <= 12; inst[i] = i.
constexpr int kPresortMax = 1024;                 // reads ordered by one warp in its shared memory
// longer reads are ordered by a whole CTA: as many k-mers as 192 KB of shared memory hold (12 kbp reads with --six; 6 kbp for K = 25);
// still longer ones keep the reference's emission order and are ordered by rank in the scoring stage
__host__ __device__ constexpr int presort_limit(bool wide) { return wide ? 12288 : 24576; }
void launch_extract_identify(cudaStream_t s, PackedSeqs seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                             uint64_t *key_hi, uint32_t *inst, const uint32_t *read_first, uint32_t n_reads, uint32_t max_inst, uint64_t *launches);

// the same straight from raw ASCII sequences (KASA_SEQ_ASCII): cleaning, strands and markers on the fly, no packed copy
void launch_extract_identify_ascii(cudaStream_t s, const char *ascii, const uint64_t *ascii_start, const uint32_t *seq_len,
                                   uint32_t n_seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                                   uint64_t *key_hi, uint32_t *inst, const uint32_t *read_first, uint32_t n_reads, uint32_t max_inst,
                                   uint64_t *launches);

// build: one slot per window of each prepared ASCII string; windows holding '_' (31) or giving k-mer 0
// get the sentinel key ~0 (dropped after the sort): Read.hpp:2016-2024, 2048-2059, 2097, 2122
// one = 1: `build --one` (Read.hpp:2143-2200): only the windows at every third base of a string
void launch_build_counts(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, int K, int one, uint64_t *woff, uint64_t *launches);
void launch_extract_build(cudaStream_t s, const char *ascii, const uint64_t *seq_start, const uint32_t *seq_len,
                          const uint32_t *taxid, uint32_t n_seqs, int K, int one, const uint8_t *codon, const uint64_t *woff, uint64_t n_windows,
                          uint64_t *key_lo, uint64_t *key_hi, uint32_t *val, uint64_t *launches);

}  // namespace kasa
