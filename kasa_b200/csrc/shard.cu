// shard.cu — see shard.cuh.
#include "shard.cuh"

namespace kasa {
namespace {

// first six letters of a K-letter k-mer as a 30-bit number
__device__ __forceinline__ uint32_t prefix30(const uint64_t *lo, const uint64_t *hi, uint64_t i, int K)
{
    const int sh = 5 * K - 30;
    return hi ? (uint32_t)(hi[i] >> (sh - 64)) : (uint32_t)(lo[i] >> sh);
}

__device__ uint64_t first_with_prefix_ge(const uint64_t *lo, const uint64_t *hi, uint64_t n, int K, uint32_t p)
{
    uint64_t a = 0, b = n;
    while (a < b) {
        const uint64_t m = a + ((b - a) >> 1);
        if (prefix30(lo, hi, m, K) < p) a = m + 1; else b = m;
    }
    return a;
}

__global__ void run_bounds_kernel(const uint64_t *lo, const uint64_t *hi, uint64_t n, int K, const uint32_t *prefix_begin,
                                  int n_shards, uint64_t *begin)
{
    const int r = threadIdx.x;
    if (r > n_shards) return;
    begin[r] = r == n_shards ? n : first_with_prefix_ge(lo, hi, n, K, prefix_begin[r]);
}

__global__ void shard_plan_kernel(const uint64_t *lo, const uint64_t *hi, uint64_t n, int K, int align_letters, int n_shards,
                                  uint64_t *entry_begin, uint32_t *prefix_begin)
{
    const int r = threadIdx.x;
    if (r > n_shards) return;
    if (r == 0) { entry_begin[0] = 0; prefix_begin[0] = 0; return; }
    if (r == n_shards || n == 0) { entry_begin[r] = n; prefix_begin[r] = 1u << 30; return; }
    const uint64_t cand = (uint64_t)((unsigned __int128)n * (unsigned)r / (unsigned)n_shards);
    const uint32_t keep = ~((1u << (5 * (6 - align_letters))) - 1u);
    const uint32_t p = prefix30(lo, hi, cand, K) & keep;
    prefix_begin[r] = p;
    entry_begin[r] = first_with_prefix_ge(lo, hi, n, K, p);
}

__global__ void prefix_hist_kernel(const uint64_t *lo, const uint64_t *hi, uint64_t n, int K, int bits, uint64_t *hist)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= (1u << bits)) return;
    const int sh = 30 - bits;
    const uint64_t p0 = first_with_prefix_ge(lo, hi, n, K, b << sh);
    const uint64_t p1 = b + 1 == (1u << bits) ? n : first_with_prefix_ge(lo, hi, n, K, (b + 1) << sh);
    hist[b] = p1 - p0;
}

}  // namespace

void launch_prefix_hist(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int bits, uint64_t *hist,
                        uint64_t *launches)
{
    const unsigned bins = 1u << bits;
    prefix_hist_kernel<<<div_up(bins, 256u), 256, 0, s>>>(key_lo, key_hi, n, K, bits, hist);
    ++*launches;
}

void launch_run_bounds(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K,
                       const uint32_t *prefix_begin, int n_shards, uint64_t *begin, uint64_t *launches)
{
    run_bounds_kernel<<<1, 32, 0, s>>>(key_lo, key_hi, n, K, prefix_begin, n_shards, begin);
    ++*launches;
}

void launch_shard_plan(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int align_letters,
                       int n_shards, uint64_t *entry_begin, uint32_t *prefix_begin, uint64_t *launches)
{
    shard_plan_kernel<<<1, 32, 0, s>>>(key_lo, key_hi, n, K, align_letters, n_shards, entry_begin, prefix_begin);
    ++*launches;
}

}  // namespace kasa
