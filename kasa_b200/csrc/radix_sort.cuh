// radix_sort.cuh — on-device LSD radix sort of k-mer records (replaces the reference's
// comparison sorts: Compare.hpp:1077/1130 parallelQuicksort of the read k-mers, Build.hpp:309
// sort of the (kmer, taxid) brick).
//
// Records are structure-of-arrays: w0 (u64, low k-mer word), optional w1 (u64, high word of a
// 128-bit k-mer) and v (u32: instance id for identify, taxid for build).  One pass moves the
// whole record by one 8-bit digit taken from any of the three arrays ("single-pass" radix
// pass: every tile reads its keys once, ranks them in shared memory, learns its global
// offsets by decoupled look-back on a per-(tile,digit) status word, and writes digit-sorted
// runs).  Global digit histograms of all passes are produced by one up-front read of the keys.
//
// Roofline: HBM bound.  Algorithmic bytes per record per pass = 2 x record size (12 B or 20 B),
// plus 8 B (w0) [+8 B w1] [+4 B v] once for the histograms.
#pragma once
#include "common.cuh"

namespace kasa {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxPasses = 24;

// digit source array of a pass
enum : int { SRC_W0 = 0, SRC_W1 = 1, SRC_V = 2 };

struct SortPass {
    int src;         // SRC_*
    int shift;       // bit offset inside the source word
    uint32_t mask;   // (1 << bits) - 1, bits <= 8
};

struct SortPlan {
    int n_passes;
    SortPass pass[kMaxPasses];
};

struct SortBuffers {         // ping-pong record arrays
    uint64_t *w0[2];
    uint64_t *w1[2];         // nullptr for 64-bit k-mers
    uint32_t *v[2];
};

// tile shape, measured on B200 (1.74 G 12-byte records, ms per 4 passes): 512 x 8 x 2 CTAs/SM 64.6 | 384 x 12 x 2 63.8 |
// 256 x 16 x 3 59.2 | 256 x 18 x 3 56.8 | 256 x 24 x 2 64.4 | 512 x 16 x 1 72.2: items per thread buy instruction-level
// parallelism and amortise the per-tile look-back, three resident CTAs overlap the phases of different tiles
constexpr int kSortBlock = 256;
constexpr int kSortItems = 18;
constexpr int kSortCtas = 3;             // resident CTAs per SM (80 registers per thread)
constexpr int kSortTile = kSortBlock * kSortItems;

// workspace: [hist: kMaxPasses*256 u64][ticket+err: 64 B][status: tiles*256 u64]
inline size_t sort_workspace_bytes(uint64_t n)
{
    uint64_t tiles = div_up<uint64_t>(n ? n : 1, kSortTile);
    return (size_t)kMaxPasses * kRadix * 8 + 256 + tiles * kRadix * 8;
}

// Sorts n records on the digits of `plan` (least significant pass first).  Returns the index
// (0/1) of the buffer set holding the sorted records.  Adds the number of kernel launches to
// *launches.  err_flag (device u32) is set to 1 if the look-back watchdog trips.
int radix_sort(cudaStream_t stream, const SortBuffers &buf, int start_buf, uint64_t n, const SortPlan &plan,
               void *workspace, uint32_t *err_flag, uint64_t *launches, cudaEvent_t after_hist = nullptr);

// plan for the identify sort: all significant bits of a K-letter k-mer (5 bits per letter)
SortPlan plan_for_kmer_bits(int key_bits);
// LSD passes over key bits [bit_lo, bit_hi) (bit 64+ = second word)
SortPlan plan_for_bit_range(int bit_lo, int bit_hi);

}  // namespace kasa
