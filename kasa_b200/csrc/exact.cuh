// exact.cuh — KASA_SCORE_EXACT: per-read fp32 scores accumulated in the reference's own flush order
// (Compare.hpp:917-955 flush of a level when the next matching group appears, :1032-1041 flush at the end of a
// trie range, k_low first), so that they are bit-identical to `kASA -n 1` for the same batch.
//
// A contribution (instance i, level j) is applied by the reference when the group holding i at level j is
// flushed, i.e. at sorted position T_j(i) = the first position p > i where either a new matching group of level j
// starts or a new valid trie range starts (SURVEY.md §3.1 "Floating-point order").  Sorting a read's
// contributions by (T_j(i), j) and adding them sequentially reproduces the reference's rounding: this file computes
// the T_j(i), score.cu orders and replays the contributions (the same two kernels as KASA_SCORE_ORDERED, with the
// flush position as the sort key instead of the sorted rank).
#pragma once
#include "common.cuh"
#include "lookup.cuh"

namespace kasa {

struct ExactWorkspace {
    uint32_t *flags;     // [n] per sorted position: bit j = a matching group of level j starts here, bit 31 = a valid range starts
    uint32_t *first;     // [tiles][num_k] first flagged position inside the tile (or ~0)
    uint32_t *next;      // [tiles][num_k] first flagged position in any later tile (or ~0)
    uint32_t *trig;      // [n][num_k] T_j by instance
};

constexpr int kExactTile = 1024;
inline size_t exact_tiles(uint64_t n) { return (size_t)div_up<uint64_t>(n ? n : 1, kExactTile); }

// fills ws.trig from the fully sorted queries, their instance ids and the match records written by the lookup kernel
void launch_exact_triggers(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n,
                           IndexDev idx, LevelParams lp, const uint64_t *rec, ExactWorkspace ws, uint64_t *launches);

}  // namespace kasa
