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

__device__ __forceinline__ uint32_t rec_field(const uint64_t *R, int slot, int hb)
{
    const int fpw = 64 / hb;
    return (uint32_t)(R[slot / fpw] >> ((slot % fpw) * hb)) & ((1u << hb) - 1u);
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
    // levels matched = largest field of the record
    const int RWn = record_words(lp.hb);
    const uint64_t *R = rec + (uint64_t)inst[i] * RWn;
    uint64_t Rw[8];
    for (int x = 0; x < RWn; ++x) Rw[x] = R[x];
    uint32_t hmax = 0;
    for (int s = 0; s < lp.n_taxa - 1; ++s) hmax = max(hmax, rec_field(Rw, s, lp.hb));
    // is the range prefix present?  closest index neighbours of q
    int64_t lo = 0, hi = (int64_t)idx.n;
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        Key e{idx.key_lo[mid], W ? idx.key_hi[mid] : 0ull};
        if (key_lt<W>(e, q)) lo = mid + 1; else hi = mid;
    }
    int best = -1;
    if (lo < (int64_t)idx.n) best = max(best, common_letters<W>(Key{idx.key_lo[lo], W ? idx.key_hi[lo] : 0ull}, q, lp.K));
    if (lo > 0) best = max(best, common_letters<W>(Key{idx.key_lo[lo - 1], W ? idx.key_hi[lo - 1] : 0ull}, q, lp.K));
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

__global__ void max_instances_kernel(const uint64_t *__restrict__ read_koff, uint32_t n_reads, uint32_t *out)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) atomicMax(out, (uint32_t)(read_koff[r + 1] - read_koff[r]));
}

// One CTA per read.  keys = (T << 29) | (level << 24) | (|taxa| << 16) | instance, sorted ascending in shared memory;
// then thread t replays the contributions of taxon t in that order.
constexpr int kExactBlock = 128;

__global__ void __launch_bounds__(kExactBlock) exact_scores_kernel(const uint64_t *__restrict__ rec, const uint32_t *__restrict__ trig,
                                                                   const uint64_t *__restrict__ read_koff, uint32_t n_reads,
                                                                   LevelParams lp, uint32_t max_inst, uint32_t cap_pairs,
                                                                   float *__restrict__ scores)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_rec = reinterpret_cast<uint64_t *>(smem_raw);              // [max_inst][4]
    const int RWn = record_words(lp.hb);
    uint64_t *s_key = s_rec + (size_t)max_inst * RWn;                      // [cap_pairs]
    __shared__ uint32_t s_n;
    __shared__ float s_w[32];
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    if (tid < 32) { int k = lp.k_low + tid; s_w[tid] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const uint64_t s0 = read_koff[r];
        const uint32_t ni = (uint32_t)(read_koff[r + 1] - s0);
        if (tid == 0) s_n = 0;
        for (uint32_t x = tid; x < ni * RWn; x += kExactBlock) s_rec[x] = rec[s0 * RWn + x];
        __syncthreads();
        // (instance, level) pairs with at least one taxon
        for (uint32_t x = tid; x < ni * (uint32_t)NK; x += kExactBlock) {
            const uint32_t s = x / NK;
            const int j = (int)(x % NK);
            uint32_t nj = 0;
            for (int sl = 0; sl < T - 1; ++sl) nj += rec_field(s_rec + (size_t)s * RWn, sl, lp.hb) > (uint32_t)j;
            if (nj) {
                const uint32_t o = atomicAdd(&s_n, 1u);
                s_key[o] = ((uint64_t)trig[(s0 + s) * NK + j] << 29) | ((uint64_t)j << 24) | ((uint64_t)nj << 16) | s;
            }
        }
        __syncthreads();
        const uint32_t np = s_n;
        uint32_t p2 = 1;
        while (p2 < np) p2 <<= 1;
        for (uint32_t x = np + tid; x < p2; x += kExactBlock) s_key[x] = ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= p2; k <<= 1)                   // bitonic sort, ascending
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t x = tid; x < p2; x += kExactBlock) {
                    const uint32_t y = x ^ j;
                    if (y > x) {
                        const uint64_t a = s_key[x], b = s_key[y];
                        const bool up = (x & k) == 0;
                        if ((a > b) == up) { s_key[x] = b; s_key[y] = a; }
                    }
                }
                __syncthreads();
            }
        // replay: Compare.hpp:924, 528-530  score = weight * (1.f / numOfEntries), added once per hit instance
        for (int t = 1 + tid; t < T; t += kExactBlock) {
            float acc = 0.f;
            for (uint32_t x = 0; x < np; ++x) {
                const uint64_t key = s_key[x];
                const uint32_t s = (uint32_t)key & 0xffffu, nj = (uint32_t)(key >> 16) & 0xffu;
                const int j = (int)((key >> 24) & 31u);
                if (rec_field(s_rec + (size_t)s * RWn, t - 1, lp.hb) > (uint32_t)j) acc = __fadd_rn(acc, __fmul_rn(s_w[j], __fdiv_rn(1.f, (float)nj)));   // Compare.hpp:924, 530: product rounded, then added
            }
            scores[(uint64_t)r * T + t] = acc;
        }
        if (tid == 0) scores[(uint64_t)r * T] = 0.f;
        __syncthreads();
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

void launch_exact_max_instances(cudaStream_t s, const uint64_t *read_koff, uint32_t n_reads, uint32_t *max_inst, uint64_t *launches)
{
    KASA_CUDA_TRY(cudaMemsetAsync(max_inst, 0, 4, s));
    if (!n_reads) return;
    max_instances_kernel<<<div_up(n_reads, 256u), 256, 0, s>>>(read_koff, n_reads, max_inst);
    ++*launches;
}

bool launch_exact_scores(cudaStream_t s, const uint64_t *rec, const uint32_t *trig, const uint64_t *read_koff, uint32_t n_reads,
                         uint32_t max_inst, LevelParams lp, float *scores, uint64_t *launches)
{
    if (!n_reads) return true;
    uint64_t pairs = (uint64_t)max_inst * lp.num_k, cap = 1;
    while (cap < pairs) cap <<= 1;
    const size_t sm = (size_t)max_inst * record_words(lp.hb) * 8 + cap * 8;
    if (max_inst > 65535 || sm > 200 * 1024) return false;
    KASA_CUDA_TRY(cudaFuncSetAttribute(exact_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    unsigned grid = n_reads < (unsigned)kSMs * 8 ? n_reads : (unsigned)kSMs * 8;
    exact_scores_kernel<<<grid, kExactBlock, sm, s>>>(rec, trig, read_koff, n_reads, lp, max_inst, (uint32_t)cap, scores);
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
    return true;
}

}  // namespace kasa
