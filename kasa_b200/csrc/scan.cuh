// scan.cuh — device-wide exclusive prefix sum of u64 (three small kernels; used for sequence
// offsets, k-mer offsets and stream compaction in the build path — none of them is a hot loop).
#pragma once
#include "common.cuh"

namespace kasa {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanBlock * kScanItems;

// In-place exclusive scan of data[0..n); data[n] receives the total (data must hold n+1 values).
// `tmp` must hold scan_tmp_elems(n) u64.
inline size_t scan_tmp_elems(uint64_t n) { return (size_t)div_up<uint64_t>(n ? n : 1, kScanTile) + 1; }
void exclusive_scan_u64(cudaStream_t stream, uint64_t *data, uint64_t n, uint64_t *tmp, uint64_t *launches);

}  // namespace kasa
