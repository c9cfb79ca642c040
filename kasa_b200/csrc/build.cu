// build.cu — see build.cuh.
#include "build.cuh"

namespace kasa {
namespace {

__global__ void __launch_bounds__(256) unique_flags_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                           const uint32_t *__restrict__ val, uint64_t n, uint64_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t a = lo[i], b = hi ? hi[i] : 0;
        bool sentinel = hi ? (b == ~0ull) : (a == ~0ull);          // dropped windows (extract.cu)
        bool differs = (i == 0) || a != lo[i - 1] || (hi && b != hi[i - 1]) || val[i] != val[i - 1];
        flags[i] = (!sentinel && differs) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) compact_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                      const uint32_t *__restrict__ val, uint64_t n, const uint64_t *__restrict__ pos,
                                                      uint64_t *__restrict__ out_lo, uint64_t *__restrict__ out_hi,
                                                      uint32_t *__restrict__ out_val)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t p = pos[i];
        if (pos[i + 1] != p) {
            out_lo[p] = lo[i];
            if (hi) out_hi[p] = hi[i];
            out_val[p] = val[i];
        }
    }
}

__device__ __forceinline__ uint32_t prefix30(const uint64_t *lo, const uint64_t *hi, uint64_t i, int K)
{
    // Trie.hpp:371-386: first 6 letters = kmer >> 5*(K-6)
    int sh = 5 * (K - 6);
    if (!hi) return (uint32_t)(lo[i] >> sh);
    return sh >= 64 ? (uint32_t)(hi[i] >> (sh - 64)) : (uint32_t)((hi[i] << (64 - sh)) | (lo[i] >> sh));
}

__global__ void __launch_bounds__(256) prefix_flags_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi, uint64_t n,
                                                           int K, uint64_t *__restrict__ flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flags[i] = (i == 0 || prefix30(lo, hi, i, K) != prefix30(lo, hi, i - 1, K)) ? 1 : 0;
}

__global__ void __launch_bounds__(256) run_starts_kernel(uint64_t n, const uint64_t *__restrict__ pos, uint64_t n_runs,
                                                         uint64_t *__restrict__ starts)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t p = pos[i];
        if (pos[i + 1] != p) starts[p] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) starts[n_runs] = n;
}

__global__ void __launch_bounds__(256) trie_records_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi, int K,
                                                           const uint64_t *__restrict__ starts, uint64_t n_runs,
                                                           uint32_t *__restrict__ words)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += stride) {
        uint64_t st = starts[j], cnt = starts[j + 1] - st;
        words[3 * j] = (uint32_t)cnt;                       // packedBigPairTrie: {u64 count, u32 prefix}
        words[3 * j + 1] = (uint32_t)(cnt >> 32);
        words[3 * j + 2] = prefix30(lo, hi, st, K);
    }
}

// ---- halved index (Shrink::putHalfInTrie, Shrink.hpp:78-143): records {u32 lower six letters, u16 taxon index}, the upper six letters
// in the prefix table {u64 count, u32 prefix}
__global__ void __launch_bounds__(256) half_trie_split_kernel(const uint32_t *__restrict__ trie_words, uint64_t n_trie, uint64_t *__restrict__ starts,
                                                              uint32_t *__restrict__ prefix)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_trie; j += stride) {
        starts[j] = (uint64_t)trie_words[3 * j] | ((uint64_t)trie_words[3 * j + 1] << 32);
        prefix[j] = trie_words[3 * j + 2];
    }
}

// after the exclusive scan of starts: record i belongs to the last prefix whose start is <= i (records behind the table's total — the
// reference counts the last run one short, Shrink.hpp:131-136 — belong to the last prefix)
__global__ void __launch_bounds__(256) half_expand_kernel(const uint16_t *__restrict__ rec, uint64_t n, const uint64_t *__restrict__ starts,
                                                          const uint32_t *__restrict__ prefix, uint64_t n_trie, uint32_t n_taxa,
                                                          uint64_t *__restrict__ key_lo, uint16_t *__restrict__ tax, uint32_t *err_flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t a = 0, b = n_trie;                        // first run whose start is > i
        while (a < b) {
            const uint64_t m = (a + b) >> 1;
            if (starts[m] <= i) a = m + 1; else b = m;
        }
        const uint64_t j = a ? a - 1 : 0;
        const uint32_t lo = (uint32_t)rec[3 * i] | ((uint32_t)rec[3 * i + 1] << 16);
        const uint16_t t = rec[3 * i + 2];
        if (t >= n_taxa || (lo >> 30)) atomicExch(err_flag, 2u);      // the taxon index is not one of the content table
        key_lo[i] = ((uint64_t)prefix[j] << 30) | lo;
        tax[i] = t;
    }
}

__global__ void __launch_bounds__(256) map_taxids_kernel(const uint32_t *__restrict__ taxid, uint64_t n,
                                                         const uint32_t *__restrict__ sorted_taxid,
                                                         const uint16_t *__restrict__ sorted_idx, uint32_t m,
                                                         uint16_t *__restrict__ dense, uint32_t *err_flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t t = taxid[i];
        uint32_t a = 0, b = m;
        while (a < b) {
            uint32_t mid = (a + b) >> 1;
            if (__ldg(sorted_taxid + mid) < t) a = mid + 1; else b = mid;
        }
        if (a < m && __ldg(sorted_taxid + a) == t) dense[i] = __ldg(sorted_idx + a);
        else { dense[i] = 0; atomicExch(err_flag, 2u); }
    }
}

__global__ void __launch_bounds__(256) frequency_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                        const uint16_t *__restrict__ dense, uint64_t n, int K, uint32_t n_taxa,
                                                        unsigned long long *__restrict__ freq, bool use_smem)
{
    extern __shared__ uint32_t s_f[];     // [n_taxa][K] when use_smem
    const uint32_t cells = n_taxa * K;
    if (use_smem) {
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x) s_f[i] = 0;
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t a = lo[i], b = hi ? hi[i] : 0;
        uint32_t t = dense[i];
        for (int j = 0; j < K; ++j) {
            int sh = 5 * j;      // kASA.hpp:520-525: column j <-> letter position K - j
            uint32_t letter = sh >= 64 ? (uint32_t)(b >> (sh - 64)) & 31u
                                       : (sh + 5 <= 64 ? (uint32_t)(a >> sh) & 31u : (uint32_t)((b << (64 - sh)) | (a >> sh)) & 31u);
            if (letter != kLetterCaret) {
                if (use_smem) atomicAdd(&s_f[t * K + j], 1u);
                else atomicAdd(freq + (size_t)t * K + j, 1ull);
            }
        }
    }
    if (use_smem) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cells; i += blockDim.x)
            if (s_f[i]) atomicAdd(freq + i, (unsigned long long)s_f[i]);
    }
}

__global__ void __launch_bounds__(256) pack_records_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                           const uint32_t *__restrict__ val, uint64_t n, uint32_t *__restrict__ words)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t a = lo[i];
        if (hi) {
            uint64_t b = hi[i];
            uint32_t *w = words + 5 * i;     // packedLargePair {u64 low, u64 high, u32 taxid} (uint128_t.hpp:65-75)
            w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = (uint32_t)b; w[3] = (uint32_t)(b >> 32); w[4] = val[i];
        } else {
            uint32_t *w = words + 3 * i;     // packedBigPair {u64 kmer, u32 taxid}
            w[0] = (uint32_t)a; w[1] = (uint32_t)(a >> 32); w[2] = val[i];
        }
    }
}

__global__ void __launch_bounds__(256) unpack_records_kernel(const uint32_t *__restrict__ words, uint64_t n, bool wide,
                                                             uint64_t *__restrict__ lo, uint64_t *__restrict__ hi,
                                                             uint32_t *__restrict__ val)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (wide) {
            const uint32_t *w = words + 5 * i;
            lo[i] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
            hi[i] = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
            val[i] = w[4];
        } else {
            const uint32_t *w = words + 3 * i;
            lo[i] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
            val[i] = w[2];
        }
    }
}

__global__ void __launch_bounds__(256) check_sorted_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                           const uint32_t *__restrict__ val, uint64_t n, uint32_t *err_flag)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; i < n; i += stride) {
        bool bad;
        if (hi && hi[i] != hi[i - 1]) bad = hi[i] < hi[i - 1];
        else if (lo[i] != lo[i - 1]) bad = lo[i] < lo[i - 1];
        else bad = val && val[i] < val[i - 1];
        if (bad) atomicExch(err_flag, 3u);
    }
}

inline unsigned grid_for(uint64_t n)
{
    uint64_t want = div_up<uint64_t>(n ? n : 1, 256);
    return (unsigned)(want < (uint64_t)kSMs * 16 ? want : (uint64_t)kSMs * 16);
}

}  // namespace

void launch_unique_flags(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint64_t *flags, uint64_t *launches)
{
    if (!n) return;
    unique_flags_kernel<<<grid_for(n), 256, 0, s>>>(key_lo, key_hi, val, n, flags);
    ++*launches;
}

void launch_compact(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                    const uint64_t *pos, uint64_t *out_lo, uint64_t *out_hi, uint32_t *out_val, uint64_t *launches)
{
    if (!n) return;
    compact_kernel<<<grid_for(n), 256, 0, s>>>(key_lo, key_hi, val, n, pos, out_lo, out_hi, out_val);
    ++*launches;
}

void launch_prefix_flags(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, uint64_t *flags,
                         uint64_t *launches)
{
    if (!n) return;
    prefix_flags_kernel<<<grid_for(n), 256, 0, s>>>(key_lo, key_hi, n, K, flags);
    ++*launches;
}

void launch_trie_records(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, const uint64_t *pos,
                         uint64_t n_runs, uint64_t *run_start_tmp, uint32_t *trie_words, uint64_t *launches)
{
    if (!n) return;
    run_starts_kernel<<<grid_for(n), 256, 0, s>>>(n, pos, n_runs, run_start_tmp);
    trie_records_kernel<<<grid_for(n_runs), 256, 0, s>>>(key_lo, key_hi, K, run_start_tmp, n_runs, trie_words);
    *launches += 2;
}

void launch_map_taxids(cudaStream_t s, const uint32_t *taxid, uint64_t n, const uint32_t *sorted_taxid, const uint16_t *sorted_idx,
                       uint32_t n_sorted, uint16_t *dense, uint32_t *err_flag, uint64_t *launches)
{
    if (!n) return;
    map_taxids_kernel<<<grid_for(n), 256, 0, s>>>(taxid, n, sorted_taxid, sorted_idx, n_sorted, dense, err_flag);
    ++*launches;
}

void launch_frequency(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint16_t *dense, uint64_t n, int K,
                      uint32_t n_taxa, unsigned long long *freq, uint64_t *launches)
{
    if (!n) return;
    const bool use_smem = (uint64_t)n_taxa * K * 4 <= 40 * 1024;
    size_t sm = use_smem ? (size_t)n_taxa * K * 4 : 0;
    unsigned grid = grid_for(n);
    if (grid > (unsigned)kSMs * 4) grid = kSMs * 4;
    frequency_kernel<<<grid, 256, sm, s>>>(key_lo, key_hi, dense, n, K, n_taxa, freq, use_smem);
    ++*launches;
}

void launch_pack_records(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint32_t *words, uint64_t *launches)
{
    if (!n) return;
    pack_records_kernel<<<grid_for(n), 256, 0, s>>>(key_lo, key_hi, val, n, words);
    ++*launches;
}

void launch_unpack_records(cudaStream_t s, const uint32_t *words, uint64_t n, bool wide, uint64_t *key_lo, uint64_t *key_hi,
                           uint32_t *val, uint64_t *launches)
{
    if (!n) return;
    unpack_records_kernel<<<grid_for(n), 256, 0, s>>>(words, n, wide, key_lo, key_hi, val);
    ++*launches;
}

void launch_check_sorted(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint32_t *val, uint64_t n,
                         uint32_t *err_flag, uint64_t *launches)
{
    if (n < 2) return;
    check_sorted_kernel<<<grid_for(n), 256, 0, s>>>(key_lo, key_hi, val, n, err_flag);
    ++*launches;
}

void launch_half_trie_split(cudaStream_t s, const uint32_t *trie_words, uint64_t n_trie, uint64_t *starts, uint32_t *prefix, uint64_t *launches)
{
    if (!n_trie) return;
    half_trie_split_kernel<<<grid_for(n_trie), 256, 0, s>>>(trie_words, n_trie, starts, prefix);
    ++*launches;
}

void launch_half_expand(cudaStream_t s, const uint16_t *rec, uint64_t n, const uint64_t *starts, const uint32_t *prefix, uint64_t n_trie,
                        uint32_t n_taxa, uint64_t *key_lo, uint16_t *tax, uint32_t *err_flag, uint64_t *launches)
{
    if (!n) return;
    half_expand_kernel<<<grid_for(n), 256, 0, s>>>(rec, n, starts, prefix, n_trie, n_taxa, key_lo, tax, err_flag);
    ++*launches;
}

}  // namespace kasa
