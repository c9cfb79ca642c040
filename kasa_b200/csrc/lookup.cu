// lookup.cu — see lookup.cuh.
//
// Data flow per batch (all arrays in HBM):
//   sorted (q, inst)  --partition-->  per-tile index span  --lookup-->  rec[inst] (32 B, scattered
//   back to emission order)  --per_read-->  scores[read][taxon], counts[level][taxon]
//
// The reference walks sorted queries and the sorted index with a sequential state machine and scores
// "groups" (all query instances sharing a k-prefix) when the next group appears.  Every quantity it
// accumulates is linear in the hit instances, so here each instance independently
// finds the index entries sharing >= k_low leading letters with it (one binary search + a short scan of
// a shared-memory copy of the tile's index span) and records, per taxon, up to which level it matches.
#include <cstdlib>
#include <cstring>

#include "lookup.cuh"

namespace kasa {
namespace {

struct Key {
    uint64_t lo, hi;
};

template <bool W>
__device__ __forceinline__ bool key_lt(const Key &a, const Key &b)
{
    if (W && a.hi != b.hi) return a.hi < b.hi;
    return a.lo < b.lo;
}

// number of equal leading letters of two K-letter k-mers (5 bits per letter, first letter in the MSBs)
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

// letter at 1-based position k
template <bool W>
__device__ __forceinline__ uint32_t letter_at(const Key &q, int k, int K)
{
    int sh = 5 * (K - k);
    if (!W) return (uint32_t)(q.lo >> sh) & 31u;
    if (sh >= 64) return (uint32_t)(q.hi >> (sh - 64)) & 31u;
    if (sh + 5 <= 64) return (uint32_t)(q.lo >> sh) & 31u;
    return (uint32_t)((q.hi << (64 - sh)) | (q.lo >> sh)) & 31u;
}

// smallest / one-past-largest k-mer sharing all bits above `sh` with q
template <bool W>
__device__ __forceinline__ Key prefix_floor(const Key &q, int sh)
{
    Key r = q;
    if (!W) { r.lo = sh >= 64 ? 0 : (q.lo >> sh) << sh; return r; }
    if (sh >= 64) { r.lo = 0; r.hi = (sh - 64) >= 64 ? 0 : (q.hi >> (sh - 64)) << (sh - 64); }
    else r.lo = (q.lo >> sh) << sh;
    return r;
}
template <bool W>
__device__ __forceinline__ Key prefix_ceil(const Key &q, int sh)
{
    Key r = prefix_floor<W>(q, sh);
    if (!W) { r.lo += 1ull << sh; return r; }
    if (sh >= 64) r.hi += 1ull << (sh - 64);
    else {
        uint64_t old = r.lo;
        r.lo += 1ull << sh;
        if (r.lo < old) r.hi += 1;
    }
    return r;
}

struct GmemAcc {
    const uint64_t *klo, *khi;
    const uint16_t *tax;
    template <bool W>
    __device__ __forceinline__ Key key(int64_t i) const { return Key{__ldg(klo + i), W ? __ldg(khi + i) : 0ull}; }
    __device__ __forceinline__ uint32_t taxon(int64_t i) const { return __ldg(tax + i); }
};

struct SmemAcc {
    const uint64_t *klo, *khi;
    const uint16_t *tax;
    int64_t a;     // global index of element 0
    template <bool W>
    __device__ __forceinline__ Key key(int64_t i) const { return Key{klo[i - a], W ? khi[i - a] : 0ull}; }
    __device__ __forceinline__ uint32_t taxon(int64_t i) const { return tax[i - a]; }
};

// first position in [lo, hi) whose key is >= q.  The trip count depends only on hi - lo, which is uniform
// across a CTA in the lookup kernel, so the search does not diverge.
template <bool W, class Acc>
__device__ __forceinline__ int64_t lower_bound(const Acc &acc, int64_t lo, int64_t hi, const Key &q)
{
    int64_t n = hi - lo;
    if (n <= 0) return lo;
    while (n > 1) {
        int64_t half = n >> 1;
        if (key_lt<W>(acc.template key<W>(lo + half - 1), q)) lo += half;
        n -= half;
    }
    return lo + (key_lt<W>(acc.template key<W>(lo), q) ? 1 : 0);
}

// record held in registers: raise the level field of a taxon (1-based dense index) to h; `np` tracks the level
// sizes |taxa_j| (one byte per level j, the fourth word of a 3-bit record): raising a field from cur to h adds
// the taxon to the levels cur .. h-1
template <int HB>
__device__ __forceinline__ void rec_raise(uint64_t (&rw)[kRW<HB>], uint64_t &np, uint32_t taxon, uint32_t h)
{
    constexpr int FPW = 64 / HB;
    constexpr uint64_t HM = (1ull << HB) - 1;
    constexpr uint64_t kOnes = 0x0101010101010101ull;
    const uint32_t sl = taxon - 1, wi = sl / FPW;
    const int sh = (int)(sl % FPW) * HB;
#pragma unroll
    for (int i = 0; i < kRW<HB>; ++i) {
        if (HB == 3 && i == 3) break;                 // 3-bit records keep their fields in three words
        if ((uint32_t)i == wi) {
            const uint32_t cur = (uint32_t)(rw[i] >> sh) & (uint32_t)HM;
            if (h > cur) {
                rw[i] += (uint64_t)(h - cur) << sh;
                if (HB == 3) np += kOnes & (((1ull << (8 * h)) - 1) ^ ((1ull << (8 * cur)) - 1));
            }
        }
    }
}

// all bits >= sh equal, i.e. the first K - sh/5 letters agree
template <bool W>
__device__ __forceinline__ bool shares_prefix(const Key &a, const Key &b, int sh)
{
    if (!W) return ((a.lo ^ b.lo) >> sh) == 0;
    const uint64_t xh = a.hi ^ b.hi;
    if (sh >= 64) return (xh >> (sh - 64)) == 0;
    return xh == 0 && ((a.lo ^ b.lo) >> sh) == 0;
}

// number of levels (from k_low up) at which an entry sharing >= k_low letters with q matches: min(common, kv) - k_low + 1
template <bool W>
__device__ __forceinline__ uint32_t match_levels(const Key &e, const Key &q, int kv, const LevelParams &lp)
{
    int lead;
    if (W) {
        const uint64_t xh = e.hi ^ q.hi, xl = e.lo ^ q.lo;
        lead = xh ? __clzll((long long)xh) - (128 - 5 * lp.K) : (xl ? (5 * lp.K - 64) + __clzll((long long)xl) : 5 * lp.K);
    } else {
        const uint64_t x = e.lo ^ q.lo;
        lead = x ? __clzll((long long)x) - (64 - 5 * lp.K) : 5 * lp.K;
    }
    const int c = (lead * 205) >> 10;                 // lead / 5 for lead <= 125
    return (uint32_t)(min(c, kv) - lp.k_low + 1);
}

// bit i of x (i < 64 / HB) -> bit i * HB: a taxa bit mask becomes "level field = 1" for each of its taxa
template <int HB>
__device__ __forceinline__ uint64_t spread_fields(uint64_t x)
{
    if (HB == 3) {                                    // 21 bits, stride 3
        x &= 0x1fffffull;
        x = (x | (x << 32)) & 0x001f00000000ffffull;
        x = (x | (x << 16)) & 0x001f0000ff0000ffull;
        x = (x | (x << 8)) & 0x100f00f00f00f00full;
        x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
        x = (x | (x << 2)) & 0x1249249249249249ull;
        return x;
    }
    if (HB == 4) {                                    // 16 bits, stride 4
        x &= 0xffffull;
        x = (x | (x << 24)) & 0x000000ff000000ffull;
        x = (x | (x << 12)) & 0x000f000f000f000full;
        x = (x | (x << 6)) & 0x0303030303030303ull;
        x = (x | (x << 3)) & 0x1111111111111111ull;
        return x;
    }
    uint64_t r = 0;                                   // 12 bits, stride 5
#pragma unroll
    for (int i = 0; i < 12; ++i) r |= ((x >> i) & 1ull) << (5 * i);
    return r;
}

// last level (as a k) not killed by a '^' letter at positions k_low..k_high, or k_low - 1 (Compare.hpp:836, 897)
template <bool W>
__device__ __forceinline__ int last_valid_level(const Key &q, const LevelParams &lp)
{
    if (!W) {
        // exact zero-field test on q ^ ('^' in every field): bit 4 of a 5-bit field is set iff the letter is '^'
        constexpr uint64_t M = 0x07bdef7bdef7bdefull, H = 0x0842108421084210ull;      // 01111 / 10000 in 12 fields
        const uint64_t y = q.lo ^ (M << 1);                                             // 11110 = '^'
        uint64_t z = ~(((y & M) + M) | y) & H;
        const int top = 5 * (lp.K - lp.k_low) + 5, bot = 5 * (lp.K - lp.k_high);         // fields k_low .. k_high
        z &= (top >= 64 ? ~0ull : (1ull << top) - 1) & ~((1ull << bot) - 1);
        if (!z) return lp.k_high;
        const int p = 63 - __clzll((long long)z);                                       // = 5 (K - k) + 4
        return lp.K - (((p - 4) * 205) >> 10) - 1;
    }
    int kv = lp.k_low - 1;
    for (int k = lp.k_low; k <= lp.k_high; ++k) {
        if (letter_at<W>(q, k, lp.K) == kLetterCaret) break;
        kv = k;
    }
    return kv;
}

// Oversized span (rare): Set semantics of Compare.hpp:785-1041 for one query instance straight from the index in
// HBM.  Every entry whose first c >= k_low letters equal the query's belongs to taxa_k(prefix_k(q)) for
// k = k_low .. min(c, kv); those entries start at the lower bound of the query's k_low-letter prefix.
template <bool W, int HB>
__device__ __forceinline__ void match_query_gmem(const GmemAcc &acc, int64_t a, int64_t b, const Key &q, int kv,
                                                 const LevelParams &lp, uint64_t (&rw)[kRW<HB>], uint64_t &np)
{
    const int low_sh = 5 * (lp.K - lp.k_low);
    for (int64_t e = lower_bound<W>(acc, a, b, prefix_floor<W>(q, low_sh)); e < b; ++e) {
        const Key ek = acc.template key<W>(e);
        if (!shares_prefix<W>(ek, q, low_sh)) break;
        rec_raise<HB>(rw, np, acc.taxon(e), match_levels<W>(ek, q, kv, lp));
    }
}

template <bool W>
__global__ void __launch_bounds__(256) partition_kernel(const uint64_t *__restrict__ q_lo, const uint64_t *__restrict__ q_hi,
                                                        uint64_t n, IndexDev idx, LevelParams lp, uint64_t *__restrict__ part,
                                                        uint64_t n_tiles, uint32_t tq, int span_shift)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    uint64_t i0 = t * tq, i1 = min(n, i0 + tq) - 1;
    Key first{q_lo[i0], W ? q_hi[i0] : 0ull}, last{q_lo[i1], W ? q_hi[i1] : 0ull};
    GmemAcc acc{idx.key_lo, idx.key_hi, idx.tax};
    // queries are ordered on the bits above span_shift only; the span must hold every entry that shares the
    // k_low-letter prefix with any query of the tile, which the coarser of the two prefixes guarantees
    Key f = prefix_floor<W>(first, span_shift), c = prefix_ceil<W>(last, span_shift);
    int64_t a = lower_bound<W>(acc, 0, (int64_t)idx.n, f);
    int64_t b = lower_bound<W>(acc, a, (int64_t)idx.n, c);
    part[2 * t] = (uint64_t)a;
    part[2 * t + 1] = (uint64_t)b;
}

constexpr int kSpanCap = 2048;   // index entries staged per tile (the tile size is chosen so that the average span is <= half of it)
static_assert(kSpanCap < 4096, "span positions and counts are packed in 12 bits");

// ---- TMA (bulk async copy) + mbarrier plumbing for the span staging
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared-memory stage of one tile's index span; bulk copies need 16-byte aligned addresses and sizes, so the
// copies start at the aligned entry below the span and the arrays are padded
constexpr int kKeyPad = 2, kTaxPad = 16;

template <bool W>
struct SpanStage {
    uint64_t klo[kSpanCap + kKeyPad];
    uint64_t khi[W ? kSpanCap + kKeyPad : 2];
    uint16_t tax[kSpanCap + kTaxPad];
    uint64_t bar;                        // mbarrier of the bulk copies
};

// One CTA per tile of <= 1024 sorted queries.  The tile's index span (contiguous runs of the key and taxon
// arrays) arrives as TMA bulk copies on an mbarrier while the threads fetch their queries.  Per query:
//   1. one binary search for the first entry sharing the query's first k_low letters — the QPT searches of a thread
//      run in lockstep (independent chains of shared-memory loads, 32-bit positions, CTA-uniform trip count);
//   2. a forward walk over those entries, again QPT queries in lockstep: the taxa seen are OR-ed into a 64-bit
//      mask in a register, and the sub-range sharing more than k_low letters (contiguous, usually 0-2 entries)
//      is remembered;
//   3. the mask spreads into "level 1" fields, the remembered entries raise their taxa's fields, the level sizes
//      follow from popcount + raises, and one 32-byte store puts the record into the instance's slot.
// The record lives in registers throughout; shared memory only holds the span.
template <bool W, int HB>
__global__ void __launch_bounds__(kLookupBlock, W ? 3 : 4) lookup_kernel(const uint64_t *__restrict__ q_lo, const uint64_t *__restrict__ q_hi,
                                                                         const uint32_t *__restrict__ inst, uint64_t n, IndexDev idx,
                                                                         LevelParams lp, const uint64_t *__restrict__ part,
                                                                         uint64_t *__restrict__ rec, uint32_t tq)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpanStage<W> &st = *reinterpret_cast<SpanStage<W> *>(smem_raw);
    constexpr int QPT = kLookupTile / kLookupBlock;
    constexpr int FPW = 64 / HB;
    const int tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    const int64_t a = (int64_t)part[2 * tile], b = (int64_t)part[2 * tile + 1];
    const bool staged = b - a <= kSpanCap;
    const int span = staged ? (int)(b - a) : 0;
    const int64_t ak = a & ~(int64_t)1, ac = a & ~(int64_t)7;
    if (tid == 0 && span > 0) {
        mbar_init(&st.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t kb = (uint32_t)(((b - ak) + 1) & ~(int64_t)1) * 8u, cb = (uint32_t)(((b - ac) + 7) & ~(int64_t)7) * 2u;
        mbar_expect_tx(&st.bar, kb * (W ? 2u : 1u) + cb);
        bulk_g2s(st.klo, idx.key_lo + ak, kb, &st.bar);
        if (W) bulk_g2s(st.khi, idx.key_hi + ak, kb, &st.bar);
        bulk_g2s(st.tax, idx.tax + ac, cb, &st.bar);
    }
    // this thread's queries and instance ids: their latency overlaps the span staging
    Key qv[QPT];
    uint32_t iv[QPT];
    bool live[QPT];
#pragma unroll
    for (int r = 0; r < QPT; ++r) {
        const uint32_t off = r * kLookupBlock + tid;
        const uint64_t i = tile * tq + off;
        live[r] = off < tq && i < n;
        qv[r].lo = live[r] ? ldg_stream(q_lo + i) : 0ull;
        qv[r].hi = (W && live[r]) ? ldg_stream(q_hi + i) : 0ull;
        iv[r] = live[r] ? (inst ? ldg_stream(inst + i) : (uint32_t)i) : 0u;      // inst == nullptr: record i goes to slot i
    }
    __syncthreads();                                          // the barrier word is initialised
    const uint64_t *s_klo = st.klo + (a - ak);
    const uint64_t *s_khi = st.khi + (W ? (a - ak) : 0);
    const uint16_t *s_tax = st.tax + (a - ac);
    const int low_sh = 5 * (lp.K - lp.k_low);                 // bits below the first k_low letters
    const int up_sh = lp.k_low < lp.K ? low_sh - 5 : 0;       // ... below the first k_low + 1 letters
    auto skey = [&](int e) { return Key{s_klo[e], W ? s_khi[e] : 0ull}; };

    int pos[QPT], kv[QPT];
    uint64_t mask[QPT];
    uint32_t up[QPT];                                         // entries sharing > k_low letters: first position << 12 | count (both <= kSpanCap < 4096)
    uint32_t act = 0;
#pragma unroll
    for (int r = 0; r < QPT; ++r) {
        pos[r] = 0; mask[r] = 0; up[r] = 0;
        kv[r] = live[r] ? last_valid_level<W>(qv[r], lp) : 0;
        if (live[r] && kv[r] >= lp.k_low) act |= 1u << r;
    }
    if (span > 0) {
        mbar_wait(&st.bar, 0);
        // ---- 1. lockstep binary searches for the first entry >= the query's k_low-letter prefix
        {
            Key fl[QPT];
#pragma unroll
            for (int r = 0; r < QPT; ++r) fl[r] = prefix_floor<W>(qv[r], low_sh);
            int nn = span;
            while (nn > 1) {
                const int half = nn >> 1;
#pragma unroll
                for (int r = 0; r < QPT; ++r)
                    if (key_lt<W>(skey(pos[r] + half - 1), fl[r])) pos[r] += half;
                nn -= half;
            }
#pragma unroll
            for (int r = 0; r < QPT; ++r)
                if (key_lt<W>(skey(pos[r]), fl[r])) ++pos[r];
        }
        // ---- 2. lockstep forward walks over the entries sharing k_low letters
        uint32_t go = act;
        while (go) {
            Key ek[QPT];
            uint32_t tx[QPT];
#pragma unroll
            for (int r = 0; r < QPT; ++r) {                   // independent loads first
                const int e = min(pos[r], span - 1);
                ek[r] = skey(e);
                tx[r] = s_tax[e];
            }
#pragma unroll
            for (int r = 0; r < QPT; ++r) {
                if (!((go >> r) & 1u)) continue;
                if (pos[r] < span && shares_prefix<W>(ek[r], qv[r], low_sh)) {
                    mask[r] |= 1ull << (tx[r] - 1);
                    if (kv[r] > lp.k_low && shares_prefix<W>(ek[r], qv[r], up_sh)) up[r] = (up[r] ? up[r] : (uint32_t)pos[r] << 12) + 1;
                    ++pos[r];
                } else {
                    go &= ~(1u << r);
                }
            }
        }
    }
    // ---- 3. records
    GmemAcc gacc{idx.key_lo, idx.key_hi, idx.tax};
#pragma unroll
    for (int r = 0; r < QPT; ++r) {
        if (r * kLookupBlock >= (int)tq) break;             // CTA-uniform
        if (!live[r]) continue;
        constexpr int RW = kRW<HB>;
        uint64_t rw[RW];
#pragma unroll
        for (int i = 0; i < RW; ++i) rw[i] = 0;
        uint64_t np = (uint32_t)__popcll(mask[r]);
        if (staged) {
#pragma unroll
            for (int i = 0; i < RW; ++i)
                if (i * FPW < 64 && !(HB == 3 && i == 3)) rw[i] = spread_fields<HB>(mask[r] >> (i * FPW));
            for (uint32_t e = up[r] >> 12, end = (up[r] >> 12) + (up[r] & 4095u); e < end; ++e)
                rec_raise<HB>(rw, np, s_tax[e], match_levels<W>(skey((int)e), qv[r], kv[r], lp));
        } else if ((act >> r) & 1u) {
            match_query_gmem<W, HB>(gacc, a, b, qv[r], kv[r], lp, rw, np);
        }
        if (HB == 3) rw[3] = np;
        // one 256-bit store = one full 32-byte sector per record (STG.256, sm_100)
#pragma unroll
        for (int i = 0; i < RW; i += 4)
            asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec + (uint64_t)iv[r] * RW + i), "l"(rw[i]), "l"(rw[i + 1]),
                         "l"(rw[i + 2]), "l"(rw[i + 3])
                         : "memory");
    }
}

// rec[inst[i]] = staged[i]: the origin of a prefix-sharded batch pulls the records a peer computed for its sorted
// run (contiguous, so the NVLink side is a streaming read) and scatters them into emission order in its own HBM.
// Scattered stores INTO peer memory run ~100x slower than this (measured: DESIGN.md §5), hence the pull.
__global__ void __launch_bounds__(256) scatter_records_kernel(const uint64_t *__restrict__ staged, const uint32_t *__restrict__ inst,
                                                              uint64_t n, uint64_t *__restrict__ rec, int rec_words)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = ldg_stream(inst + i);
    for (int o = 0; o < rec_words; o += 4) {
        uint64_t a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(staged + i * rec_words + o));
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec + (uint64_t)j * rec_words + o), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
    }
}

__global__ void read_offsets_kernel(const uint64_t *__restrict__ koff, const uint32_t *__restrict__ first_seq, uint32_t n_reads,
                                    uint64_t *__restrict__ out)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r <= n_reads) out[r] = koff[first_seq ? first_seq[r] : r];
}

__global__ void fold_counts_kernel(unsigned long long *cnt_fx, unsigned long long *cnt_unique, LevelParams lp, double *total_all,
                                   unsigned long long *total_unique)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = lp.num_k * lp.n_taxa;
    if (i >= n) return;
    int j = i / lp.n_taxa, t = i % lp.n_taxa;
    int o = (lp.num_k - 1 - j) * lp.n_taxa + t;        // ik = 0 <-> k_high (kASA.hpp:299-302)
    total_all[o] += (double)cnt_fx[i] * (1.0 / 4294967296.0) + (double)cnt_unique[i];
    total_unique[o] += cnt_unique[i];
    cnt_fx[i] = 0;
    cnt_unique[i] = 0;
}

template <bool W>
void launch_lookup_hb(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n, IndexDev idx,
                      LevelParams lp, const uint64_t *part, uint64_t *rec, uint32_t tq)
{
    const unsigned tiles = (unsigned)div_up<uint64_t>(n, tq);
    const size_t sm = sizeof(SpanStage<W>);
    auto go = [&](auto kern) {
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        kern<<<tiles, kLookupBlock, sm, s>>>(q_lo, q_hi, inst, n, idx, lp, part, rec, tq);
    };
    if (lp.hb == 3) go(lookup_kernel<W, 3>);
    else if (lp.hb == 4) go(lookup_kernel<W, 4>);
    else go(lookup_kernel<W, 5>);
}

}  // namespace

void launch_scatter_records(cudaStream_t s, const uint64_t *staged, const uint32_t *inst, uint64_t n, uint64_t *rec, int rec_words,
                            uint64_t *launches)
{
    if (!n) return;
    scatter_records_kernel<<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(staged, inst, n, rec, rec_words);
    ++*launches;
}

uint32_t lookup_tile_queries(uint64_t n_queries, uint64_t n_index)
{
    // queries per CTA such that the expected index span (tile * index / queries) is about half the staging capacity
    if (n_index == 0 || n_queries == 0) return kLookupTile;
    double tq = 0.5 * kSpanCap * (double)n_queries / (double)n_index;
    if (tq > kLookupTile) tq = kLookupTile;
    if (tq < 32) tq = 32;
    return (uint32_t)tq;
}

void launch_partition(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, uint64_t n, IndexDev idx, LevelParams lp,
                      uint64_t *part, uint32_t tq, int sorted_from_bit, uint64_t *launches)
{
    if (!n) return;
    const uint64_t tiles = div_up<uint64_t>(n, tq);
    const int klow_shift = 5 * (lp.K - lp.k_low);
    const int sh = sorted_from_bit > klow_shift ? sorted_from_bit : klow_shift;
    if (idx.key_hi) partition_kernel<true><<<(unsigned)div_up<uint64_t>(tiles, 256), 256, 0, s>>>(q_lo, q_hi, n, idx, lp, part, tiles, tq, sh);
    else partition_kernel<false><<<(unsigned)div_up<uint64_t>(tiles, 256), 256, 0, s>>>(q_lo, q_hi, n, idx, lp, part, tiles, tq, sh);
    ++*launches;
}

void launch_lookup(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n, IndexDev idx,
                   LevelParams lp, const uint64_t *part, uint64_t *rec, uint32_t tq, uint64_t *launches)
{
    if (!n) return;
    if (idx.key_hi) launch_lookup_hb<true>(s, q_lo, q_hi, inst, n, idx, lp, part, rec, tq);
    else launch_lookup_hb<false>(s, q_lo, q_hi, inst, n, idx, lp, part, rec, tq);
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

void launch_per_read(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                     float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique, uint64_t *launches)
{
    if (!n_reads) return;
    if (lp.hb == 3) launch_per_read_taxon(s, rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    else launch_per_read_wide(s, rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

void launch_read_offsets(cudaStream_t s, const uint64_t *koff, const uint32_t *read_first_seq, uint32_t n_reads,
                         uint64_t *read_koff, uint64_t *launches)
{
    read_offsets_kernel<<<div_up(n_reads + 1, 256u), 256, 0, s>>>(koff, read_first_seq, n_reads, read_koff);
    ++*launches;
}

void launch_fold_counts(cudaStream_t s, unsigned long long *cnt_fx, unsigned long long *cnt_unique, LevelParams lp,
                        double *total_all, unsigned long long *total_unique, uint64_t *launches)
{
    int n = lp.num_k * lp.n_taxa;
    fold_counts_kernel<<<div_up(n, 256), 256, 0, s>>>(cnt_fx, cnt_unique, lp, total_all, total_unique);
    ++*launches;
}

}  // namespace kasa
