// scan.cu — see scan.cuh.
#include "scan.cuh"

namespace kasa {
namespace {

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t x, uint64_t *s_warp, uint64_t *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint64_t v = lane < kScanBlock / 32 ? s_warp[lane] : 0, vi = v;
#pragma unroll
        for (int o = 1; o < kScanBlock / 32; o <<= 1) {
            uint64_t t = __shfl_up_sync(kFull, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < kScanBlock / 32) s_warp[lane] = vi - v;
        if (lane == kScanBlock / 32 - 1) s_warp[kScanBlock / 32] = vi;
    }
    __syncthreads();
    *total = s_warp[kScanBlock / 32];
    return incl - x + s_warp[w];
}

__global__ void __launch_bounds__(kScanBlock) scan_tile_sums(const uint64_t *__restrict__ data, uint64_t n, uint64_t *__restrict__ sums)
{
    __shared__ uint64_t s_warp[kScanBlock / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) acc += data[base + i];
    uint64_t total;
    block_exclusive_scan(acc, s_warp, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single block: exclusive scan of the tile sums, total behind the last one
__global__ void __launch_bounds__(kScanBlock) scan_sums(uint64_t *sums, uint64_t n_tiles)
{
    __shared__ uint64_t s_warp[kScanBlock / 32 + 1];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t b = 0; b < n_tiles; b += kScanBlock) {
        uint64_t i = b + threadIdx.x;
        uint64_t x = i < n_tiles ? sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_scan(x, s_warp, &total);
        uint64_t carry = s_carry;
        if (i < n_tiles) sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n_tiles] = s_carry;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply(uint64_t *__restrict__ data, uint64_t n, const uint64_t *__restrict__ sums, uint64_t n_tiles)
{
    __shared__ uint64_t s_warp[kScanBlock / 32 + 1];
    uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint64_t x[kScanItems], acc = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        x[i] = (base + i < n) ? data[base + i] : 0;
        acc += x[i];
    }
    uint64_t total;
    uint64_t run = block_exclusive_scan(acc, s_warp, &total) + sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) data[base + i] = run;
        run += x[i];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = sums[n_tiles];
}

}  // namespace

void exclusive_scan_u64(cudaStream_t stream, uint64_t *data, uint64_t n, uint64_t *tmp, uint64_t *launches)
{
    if (n == 0) {
        KASA_CUDA_TRY(cudaMemsetAsync(data, 0, 8, stream));
        return;
    }
    uint64_t tiles = div_up<uint64_t>(n, kScanTile);
    scan_tile_sums<<<(unsigned)tiles, kScanBlock, 0, stream>>>(data, n, tmp);
    scan_sums<<<1, kScanBlock, 0, stream>>>(tmp, tiles);
    scan_apply<<<(unsigned)tiles, kScanBlock, 0, stream>>>(data, n, tmp, tiles);
    *launches += 3;
    KASA_CUDA_TRY(cudaGetLastError());
}

}  // namespace kasa
