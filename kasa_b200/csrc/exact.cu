// exact.cu — see exact.cuh.
#include "exact.cuh"

namespace kasa {
namespace {

struct Key { uint64_t lo, hi; };

template <bool W>
__device__ __forceinline__ bool key_lt(const Key &a, const Key &b)
{
    if (W && a.hi != b.hi) return a.hi < b.hi;
    return a.lo < b.lo;
}

template <bool W>
__device__ __forceinline__ int common_letters(const Key &a, const Key &b, int K)
{
    int lead;
    if (W) {
        uint64_t xh = a.hi ^ b.hi;
        if (xh) lead = __clzll((long long)xh) - (128 - 5 * K);
        else {
            uint64_t xl = a.lo ^ b.lo;
            if (!xl) return K;
            lead = (5 * K - 64) + __clzll((long long)xl);
        }
    } else {
        uint64_t x = a.lo ^ b.lo;
        if (!x) return K;
        lead = __clzll((long long)x) - (64 - 5 * K);
    }
    return lead / 5;
}

// index key without its taxon-first flag (lookup.cuh: IndexDev)
template <bool W>
__device__ __forceinline__ Key ikey(const IndexDev &idx, int64_t i)
{
    return W ? Key{idx.key_lo[i], idx.key_hi[i] & ~kKeyFlag} : Key{idx.key_lo[i] & ~kKeyFlag, 0ull};
}

// One thread per sorted position.  A level-j group "starts" at i when query i matches the index at level j
// (j < levels matched by i) and differs from query i-1 within its first k_low + j letters; a range starts
// when the first min(6, k_low) letters differ and occur in the index (Trie.hpp:494-520 returns a range).
template <bool W>
__global__ void __launch_bounds__(256) exact_flags_kernel(const uint64_t *__restrict__ q_lo, const uint64_t *__restrict__ q_hi,
                                                          const uint32_t *__restrict__ inst, uint64_t n, IndexDev idx,
                                                          LevelParams lp, const uint64_t *__restrict__ rec, uint32_t *__restrict__ flags)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Key q{q_lo[i], W ? q_hi[i] : 0ull};
    int c = -1;                                       // letters shared with the previous query
    if (i) {
        const Key p{q_lo[i - 1], W ? q_hi[i - 1] : 0ull};
        c = common_letters<W>(p, q, lp.K);
    }
    // levels matched by this instance (record head, lookup.cuh)
    const uint32_t hmax = rec_hmax(rec[(uint64_t)inst[i] * kRecWords]);
    // is the range prefix present?  closest index neighbours of q
    int64_t lo = 0, hi = (int64_t)idx.n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        Key e = ikey<W>(idx, mid);
        if (key_lt<W>(e, q)) lo = mid + 1; else hi = mid;
    }
    int best = -1;
    if (lo < (int64_t)idx.n) best = max(best, common_letters<W>(ikey<W>(idx, lo), q, lp.K));
    if (lo > 0) best = max(best, common_letters<W>(ikey<W>(idx, lo - 1), q, lp.K));
    const int Lr = lp.k_low < 6 ? lp.k_low : 6;
    uint32_t f = 0;
    const int j0 = max(0, c - lp.k_low + 1);
    for (int j = j0; j < (int)hmax; ++j) f |= 1u << j;
    if (best >= Lr && c < Lr) f |= 1u << 31;
    flags[i] = f;
}

__global__ void __launch_bounds__(128) exact_tile_first_kernel(const uint32_t *__restrict__ flags, uint64_t n, int num_k, uint64_t tiles,
                                                               uint32_t *__restrict__ first)
{
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= tiles * num_k) return;
    const uint64_t tile = x / num_k;
    const int j = (int)(x % num_k);
    const uint32_t want = (1u << j) | (1u << 31);
    const uint64_t b = tile * kExactTile, e = min(n, b + kExactTile);
    uint32_t res = 0xffffffffu;
    for (uint64_t p = b; p < e; ++p)
        if (flags[p] & want) { res = (uint32_t)p; break; }
    first[x] = res;
}

__global__ void exact_suffix_kernel(const uint32_t *__restrict__ first, uint64_t tiles, int num_k, uint32_t *__restrict__ next)
{
    const int j = threadIdx.x;
    if (j >= num_k) return;
    uint32_t run = 0xffffffffu;
    for (int64_t t = (int64_t)tiles - 1; t >= 0; --t) {
        next[t * num_k + j] = run;
        const uint32_t f = first[t * num_k + j];
        if (f != 0xffffffffu) run = f;
    }
}

__global__ void __launch_bounds__(128) exact_assign_kernel(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ inst, uint64_t n,
                                                           int num_k, uint64_t tiles, const uint32_t *__restrict__ next,
                                                           uint32_t *__restrict__ trig)
{
    const uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= tiles * num_k) return;
    const uint64_t tile = x / num_k;
    const int j = (int)(x % num_k);
    const uint32_t want = (1u << j) | (1u << 31);
    const uint64_t b = tile * kExactTile, e = min(n, b + kExactTile);
    uint32_t nxt = next[x];                           // ~0 = flushed at the very end
    for (uint64_t p = e; p-- > b;) {
        trig[(uint64_t)inst[p] * num_k + j] = nxt;
        if (flags[p] & want) nxt = (uint32_t)p;
    }
}

}  // namespace

void launch_exact_triggers(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n, IndexDev idx,
                           LevelParams lp, const uint64_t *rec, ExactWorkspace ws, uint64_t *launches)
{
    if (!n) return;
    const uint64_t tiles = exact_tiles(n);
    if (idx.key_hi) exact_flags_kernel<true><<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(q_lo, q_hi, inst, n, idx, lp, rec, ws.flags);
    else exact_flags_kernel<false><<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(q_lo, q_hi, inst, n, idx, lp, rec, ws.flags);
    const unsigned g = (unsigned)div_up<uint64_t>(tiles * lp.num_k, 128);
    exact_tile_first_kernel<<<g, 128, 0, s>>>(ws.flags, n, lp.num_k, tiles, ws.first);
    exact_suffix_kernel<<<1, 32, 0, s>>>(ws.first, tiles, lp.num_k, ws.next);
    exact_assign_kernel<<<g, 128, 0, s>>>(ws.flags, inst, n, lp.num_k, tiles, ws.next, ws.trig);
    *launches += 4;
    KASA_CUDA_TRY(cudaGetLastError());
}

}  // namespace kasa
