// lookup.cu — see lookup.cuh.
//
// Data flow per batch (all arrays in HBM):
//   sorted (q, inst)  --partition-->  per-tile index span  --lookup-->  rec[inst] (one 32-byte sector, scattered
//   back to emission order)  --score.cu-->  scores[read][taxon], counts[level][taxon]
//
// The reference walks sorted queries and the sorted index with a sequential state machine and scores
// "groups" (all query instances sharing a k-prefix) when the next group appears.  Every quantity it
// accumulates is linear in the hit instances, so here each instance independently
// finds the index entries sharing >= k_low leading letters with it (one binary search + a short scan of
// a shared-memory copy of the tile's index span) and records, per distinct taxon, up to which level it matches.
// The distinct taxa of a k_low group are its flagged entries (lookup.cuh: IndexDev), so no per-query set
// structure is needed and the number of taxa is only bounded by the 16-bit dense index.
//
// Measured on B200 (tools/mem_probe*.cu): HBM3e serves ~35 G random 32-byte sectors per second, reads and full-sector
// writes alike, partial-sector writes ~20 G/s, whatever the size of the region (no gain from windows beyond L2).
// One full-sector store per instance is therefore the cheapest way back from sorted order to emission order, and it
// bounds this kernel: instances without any hit do not store at all (score.cu recognises their stale slots by the
// batch epoch in the record).
#include <algorithm>
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

// index entries as the kernels see them: key with the flag bit stripped, the flag, the taxon
template <bool W>
__device__ __forceinline__ Key strip(uint64_t lo, uint64_t hi)
{
    return W ? Key{lo, hi & ~kKeyFlag} : Key{lo & ~kKeyFlag, 0ull};
}

struct GmemAcc {
    const uint64_t *klo, *khi;
    const uint16_t *tax;
    template <bool W>
    __device__ __forceinline__ Key key(int64_t i) const { return strip<W>(__ldg(klo + i), W ? __ldg(khi + i) : 0ull); }
    template <bool W>
    __device__ __forceinline__ bool first(int64_t i) const { return ((W ? __ldg(khi + i) : __ldg(klo + i)) & kKeyFlag) != 0; }
    __device__ __forceinline__ uint32_t taxon(int64_t i) const { return __ldg(tax + i); }
};

struct SmemAcc {
    const uint64_t *klo, *khi;
    const uint16_t *tax;
    template <bool W>
    __device__ __forceinline__ Key key(int64_t i) const { return strip<W>(klo[i], W ? khi[i] : 0ull); }
    template <bool W>
    __device__ __forceinline__ bool first(int64_t i) const { return ((W ? khi[i] : klo[i]) & kKeyFlag) != 0; }
    __device__ __forceinline__ uint32_t taxon(int64_t i) const { return tax[i]; }
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

// ---- a record's inline pairs, held in three registers
__device__ __forceinline__ void pair_raise(uint64_t (&pw)[3], uint32_t count, uint32_t taxon, uint32_t h)
{
#pragma unroll
    for (int k = 0; k < kRecInline; ++k) {
        const int sh = kPairBits * (k % 3);
        const uint32_t p = (uint32_t)(pw[k / 3] >> sh) & kPairMask;
        if ((uint32_t)k < count && (p & 0xffffu) == taxon) {
            const uint32_t cur = p >> 16;
            if (h > cur) pw[k / 3] += (uint64_t)(h - cur) << (sh + 16);
        }
    }
}

__device__ __forceinline__ uint64_t rec_head(uint64_t rank, uint32_t count, uint32_t hmax, bool overflow, uint32_t owner)
{
    return (rank & 0xffffffffull) | ((uint64_t)(count & 0xffffu) << 32) | ((uint64_t)(hmax & 31u) << 48) | (overflow ? kRecOverflowBit : 0ull) |
           ((uint64_t)(owner & 15u) << 54);
}

// mask format: raise (or add) the level count of a taxon among the <= kRecUpPairs pairs held in two registers;
// false if a seventh distinct taxon would be needed
__device__ __forceinline__ bool up_pair_raise(uint64_t (&uw)[2], uint32_t taxon, uint32_t h)
{
    bool done = false;
#pragma unroll
    for (int k = 0; k < kRecUpPairs; ++k) {
        const int sh = kPairBits * (k % 3);
        const uint32_t p = (uint32_t)(uw[k / 3] >> sh) & kPairMask;
        if (done) continue;
        if (!p) { uw[k / 3] |= (uint64_t)(taxon | (h << 16)) << sh; done = true; }
        else if ((p & 0xffffu) == taxon) {
            if (h > (p >> 16)) uw[k / 3] += (uint64_t)(h - (p >> 16)) << (sh + 16);
            done = true;
        }
    }
    return done;
}

// two-mask format: raise (or add) the level count of a taxon among the <= 3 pairs of one word; false if a fourth taxon would be needed
__device__ __forceinline__ bool up3_pair_raise(uint64_t &uw, uint32_t taxon, uint32_t h)
{
    bool done = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int sh = kPairBits * k;
        const uint32_t p = (uint32_t)(uw >> sh) & kPairMask;
        if (done) continue;
        if (!p) { uw |= (uint64_t)(taxon | (h << 16)) << sh; done = true; }
        else if ((p & 0xffffu) == taxon) {
            if (h > (p >> 16)) uw += (uint64_t)(h - (p >> 16)) << (sh + 16);
            done = true;
        }
    }
    return done;
}

// General path (oversized span read from HBM, or a taxa set that does not fit the record): Set semantics of
// Compare.hpp:785-1041 for one query instance.  g0 = first entry of the query's k_low group, b = end of the span.
// Every entry whose first c >= k_low letters equal the query's belongs to taxa_k(prefix_k(q)) for k = k_low .. min(c, kv).
template <bool W, class Acc>
__device__ __noinline__ void match_general(const Acc &acc, int64_t g0, int64_t b, const Key &q, int kv, const LevelParams &lp, OverflowBuf ovf,
                                           uint32_t owner, uint32_t *err_flag, uint64_t rank, uint64_t (&out)[4])
{
    const int low_sh = 5 * (lp.K - lp.k_low);
    out[0] = rank & 0xffffffffull; out[1] = out[2] = out[3] = 0;
    uint32_t c = 0;
    int64_t gend = g0;
    for (; gend < b && shares_prefix<W>(acc.template key<W>(gend), q, low_sh); ++gend) c += acc.template first<W>(gend) ? 1u : 0u;
    if (!c) return;
    uint32_t hmax = 1;
    if (c <= (uint32_t)kRecInline) {
        uint64_t pw[3] = {0, 0, 0};
        uint32_t k = 0;
        for (int64_t e = g0; e < gend; ++e)
            if (acc.template first<W>(e)) { pw[k / 3] |= (uint64_t)(acc.taxon(e) | (1u << 16)) << (kPairBits * (k % 3)); ++k; }
        for (int64_t e = g0; e < gend; ++e) {
            const uint32_t h = match_levels<W>(acc.template key<W>(e), q, kv, lp);
            if (h > 1) { pair_raise(pw, c, acc.taxon(e), h); hmax = max(hmax, h); }
        }
        out[0] = rec_head(rank, c, hmax, false, owner); out[1] = pw[0]; out[2] = pw[1]; out[3] = pw[2];
        return;
    }
    if (c > 0xffffu) { atomicOr(err_flag, 16u); return; }      // cannot happen: <= 65535 taxa
    const unsigned long long base = atomicAdd(ovf.cursor, (unsigned long long)c);
    if (base + c > ovf.cap) { atomicOr(err_flag, 8u); return; }
    uint32_t *list = ovf.data + base;
    uint32_t k = 0;
    for (int64_t e = g0; e < gend; ++e)
        if (acc.template first<W>(e)) list[k++] = acc.taxon(e) | (1u << 16);
    for (int64_t e = g0; e < gend; ++e) {
        const uint32_t h = match_levels<W>(acc.template key<W>(e), q, kv, lp);
        if (h <= 1) continue;
        hmax = max(hmax, h);
        const uint32_t t = acc.taxon(e);
        for (uint32_t x = 0; x < c; ++x)
            if ((list[x] & 0xffffu) == t) { if ((list[x] >> 16) < h) list[x] = t | (h << 16); break; }
    }
    out[0] = rec_head(rank, c, hmax, true, owner); out[1] = (uint64_t)base;
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

// asks for [src, src + bytes) to be brought into L2 (16-byte aligned address and size); nothing waits for it
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// shared-memory stage of one tile's index span; bulk copies need 16-byte aligned addresses and sizes, so the
// copies start at the aligned entry below the span and the arrays are padded
constexpr int kKeyPad = 2, kTaxPad = 16;

template <bool W, bool MASK>
struct SpanStage {
    uint64_t klo[kSpanCap + kKeyPad];
    uint64_t khi[W ? kSpanCap + kKeyPad : 2];
    uint16_t tax[kSpanCap + kTaxPad];
    alignas(16) uint64_t gm[MASK ? kSpanCap + kKeyPad : 2];   // mask records: taxa mask of every entry's k_low group (IndexDev::gmask); bulk-copy target
    uint64_t pw[3][MASK ? 1 : kLookupTile];   // list records: pair words of the tile's records, word-major (conflict-free per-thread slots)
    uint64_t bar;                        // mbarrier of the bulk copies
};

// One CTA per tile of <= 1024 sorted queries.  The tile's index span (contiguous runs of the key and taxon
// arrays) arrives as TMA bulk copies on an mbarrier while the threads fetch their queries.  Per query:
//   1. one binary search for the first entry sharing the query's first k_low letters — the QPT searches of a thread
//      run in lockstep (independent chains of shared-memory loads, 32-bit positions, CTA-uniform trip count);
//   2. a forward walk over those entries, again QPT queries in lockstep: every flagged entry (first of its taxon in
//      the group) appends a (taxon, 1 level) pair to the query's record in shared memory, and the sub-range sharing
//      more than k_low letters (contiguous, usually 0-2 entries) is remembered;
//   3. the remembered entries raise their taxa's level counts in registers, and one 32-byte store puts the record
//      into the instance's slot.  Taxa sets of more than nine taxa go to the overflow list (match_general).
// MASK (content files of <= 64 taxa): the taxa set of every k_low group was computed once at index load (IndexDev::gmask) and is
// staged with the span, so step 1 searches for the query itself and step 2 shrinks to: one mask load from a neighbour that shares
// k_low letters, and a walk over the entries sharing MORE than k_low letters (they surround the search position; usually 0-2).
// The walk over the whole k_low group — whose trip count is the largest group among the 128 queries of a warp, dozens of
// entries for k-mers of frequent amino acids against a saturated index — took 2/3 of this kernel's instructions.
template <bool W, bool MASK, bool M2>
__global__ void __launch_bounds__(kLookupBlock, W ? 3 : 4) lookup_kernel(const uint64_t *__restrict__ q_lo, const uint64_t *__restrict__ q_hi,
                                                                         const uint32_t *__restrict__ inst, uint64_t n, uint64_t rank_base,
                                                                         IndexDev idx, LevelParams lp, const uint64_t *__restrict__ part,
                                                                         uint64_t *__restrict__ rec, uint32_t tq, OverflowBuf ovf,
                                                                         uint32_t owner, uint32_t *__restrict__ err_flag, uint32_t pf_dist)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpanStage<W, MASK> &st = *reinterpret_cast<SpanStage<W, MASK> *>(smem_raw);
    constexpr int QPT = kLookupTile / kLookupBlock;
    const int tid = threadIdx.x;
    const uint64_t tile = blockIdx.x;
    // the tile's span bounds and its queries are requested together (the queries' first use comes after the span copies are issued)
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
    const int64_t a = (int64_t)part[2 * tile], b = (int64_t)part[2 * tile + 1];
    const bool staged = b - a <= kSpanCap;
    const int span = staged ? (int)(b - a) : 0;
    const int64_t ak = a & ~(int64_t)1, ac = a & ~(int64_t)7;
    // This kernel waits at its start (a CTA is one tile: span bounds -> span copies, queries), so the second warp asks for the tile
    // that will run when the resident CTAs have retired (pf_dist tiles ahead) to be brought into L2: its queries now, its span too.
    if (pf_dist && tid >= 32 && tid < 64 && tile + pf_dist < gridDim.x) {
        const uint64_t t2 = tile + pf_dist, i2 = t2 * tq;
        const int l = tid - 32;
        if (l == 0) {
            const int64_t a2 = (int64_t)part[2 * t2], b2 = (int64_t)part[2 * t2 + 1];
            if (b2 - a2 <= kSpanCap && b2 > a2) {
                const uint32_t kb2 = (uint32_t)(((b2 - (a2 & ~(int64_t)1)) + 1) & ~(int64_t)1) * 8u;
                bulk_prefetch_l2(idx.key_lo + (a2 & ~(int64_t)1), kb2);
                if (W) bulk_prefetch_l2(idx.key_hi + (a2 & ~(int64_t)1), kb2);
                if (MASK) bulk_prefetch_l2(idx.gmask + (a2 & ~(int64_t)1), kb2);
                bulk_prefetch_l2(idx.tax + (a2 & ~(int64_t)7), (uint32_t)(((b2 - (a2 & ~(int64_t)7)) + 7) & ~(int64_t)7) * 2u);
            }
        } else if (l == 1) {
            const uint64_t m2 = min((uint64_t)tq, n - i2);                 // (sizes rounded down to 16 bytes: never past the arrays)
            const uint32_t qb = (uint32_t)(m2 & ~1ull) * 8u, ib = (uint32_t)(m2 & ~3ull) * 4u;
            if (qb) bulk_prefetch_l2(q_lo + (i2 & ~1ull), qb);
            if (W && qb) bulk_prefetch_l2(q_hi + (i2 & ~1ull), qb);
            if (inst && ib) bulk_prefetch_l2(inst + (i2 & ~3ull), ib);
        }
    }
    if (tid == 0 && span > 0) {
        mbar_init(&st.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t kb = (uint32_t)(((b - ak) + 1) & ~(int64_t)1) * 8u, cb = (uint32_t)(((b - ac) + 7) & ~(int64_t)7) * 2u;
        mbar_expect_tx(&st.bar, kb * ((W ? 2u : 1u) + (MASK ? 1u : 0u)) + cb);
        bulk_g2s(st.klo, idx.key_lo + ak, kb, &st.bar);
        if (W) bulk_g2s(st.khi, idx.key_hi + ak, kb, &st.bar);
        if (MASK) bulk_g2s(st.gm, idx.gmask + ak, kb, &st.bar);
        bulk_g2s(st.tax, idx.tax + ac, cb, &st.bar);
    }
    __syncthreads();                                          // the barrier word is initialised
    SmemAcc sacc{st.klo + (a - ak), st.khi + (W ? (a - ak) : 0), st.tax + (a - ac)};
    const int low_sh = 5 * (lp.K - lp.k_low);                 // bits below the first k_low letters
    const int up_sh = lp.k_low < lp.K ? low_sh - 5 : 0;       // ... below the first k_low + 1 letters

    int pos[QPT], kv[QPT];
    uint32_t cnt[QPT];                                        // flagged entries seen = distinct taxa of the k_low group
    uint64_t mask[QPT];                                       // MASK: the taxa of the k_low group
    uint32_t up[QPT];                                         // entries sharing > k_low letters: first position << 12 | count (both <= kSpanCap < 4096)
    uint32_t act = 0;
#pragma unroll
    for (int r = 0; r < QPT; ++r) {
        pos[r] = 0; cnt[r] = 0; up[r] = 0; mask[r] = 0;
        kv[r] = live[r] ? last_valid_level<W>(qv[r], lp) : 0;
        if (live[r] && kv[r] >= lp.k_low) act |= 1u << r;
    }
    int g0[QPT];
    if (span > 0) {
        mbar_wait(&st.bar, 0);
        // ---- 1. lockstep binary searches for the first entry >= the query's k_low-letter prefix
        {
            Key fl[QPT];
#pragma unroll
            for (int r = 0; r < QPT; ++r) fl[r] = MASK ? qv[r] : prefix_floor<W>(qv[r], low_sh);
            int nn = span;
            while (nn > 1) {
                const int half = nn >> 1;
#pragma unroll
                for (int r = 0; r < QPT; ++r)
                    if (key_lt<W>(sacc.template key<W>(pos[r] + half - 1), fl[r])) pos[r] += half;
                nn -= half;
            }
#pragma unroll
            for (int r = 0; r < QPT; ++r)
                if (key_lt<W>(sacc.template key<W>(pos[r]), fl[r])) ++pos[r];
        }
#pragma unroll
        for (int r = 0; r < QPT; ++r) g0[r] = pos[r];
        if (MASK) {
            // ---- 2 (mask records). the group's taxa from a neighbour sharing k_low letters ...
            const uint64_t *gms = st.gm + (a - ak);
            uint32_t goL = 0, goR = 0;
            int lo_p[QPT], hi_p[QPT];
#pragma unroll
            for (int r = 0; r < QPT; ++r) {
                lo_p[r] = hi_p[r] = pos[r];
                if (!((act >> r) & 1u)) continue;
                const int e1 = min(pos[r], span - 1), e0 = max(pos[r] - 1, 0);
                const bool in1 = pos[r] < span && shares_prefix<W>(sacc.template key<W>(e1), qv[r], low_sh);
                const bool in0 = pos[r] > 0 && shares_prefix<W>(sacc.template key<W>(e0), qv[r], low_sh);
                mask[r] = in1 ? gms[e1] : (in0 ? gms[e0] : 0ull);
                if (mask[r] && kv[r] > lp.k_low) { goL |= 1u << r; goR |= 1u << r; }
            }
            // ... and the entries sharing more than k_low letters: a contiguous run around the search position
            while (goL | goR) {
#pragma unroll
                for (int r = 0; r < QPT; ++r) {
                    if ((goL >> r) & 1u) {
                        const int e = lo_p[r] - 1;
                        if (e >= 0 && shares_prefix<W>(sacc.template key<W>(max(e, 0)), qv[r], up_sh)) lo_p[r] = e; else goL &= ~(1u << r);
                    }
                    if ((goR >> r) & 1u) {
                        const int e = hi_p[r];
                        if (e < span && shares_prefix<W>(sacc.template key<W>(min(e, span - 1)), qv[r], up_sh)) hi_p[r] = e + 1; else goR &= ~(1u << r);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < QPT; ++r) up[r] = hi_p[r] > lo_p[r] ? ((uint32_t)lo_p[r] << 12) + (uint32_t)(hi_p[r] - lo_p[r]) : 0u;
        }
        // ---- 2. lockstep forward walks over the entries sharing k_low letters
        uint32_t go = MASK ? 0u : act;
        while (go) {
            uint64_t elo[QPT], ehi[QPT];
            uint32_t tx[QPT];
#pragma unroll
            for (int r = 0; r < QPT; ++r) {                   // independent loads first
                const int e = min(pos[r], span - 1);
                elo[r] = sacc.klo[e];
                ehi[r] = W ? sacc.khi[e] : 0ull;
                tx[r] = sacc.tax[e];
            }
#pragma unroll
            for (int r = 0; r < QPT; ++r) {
                if (!((go >> r) & 1u)) continue;
                const Key ek = strip<W>(elo[r], ehi[r]);
                if (pos[r] < span && shares_prefix<W>(ek, qv[r], low_sh)) {
                    if (MASK) mask[r] |= 1ull << (tx[r] - 1);
                    else if ((W ? ehi[r] : elo[r]) & kKeyFlag) {
                        const uint32_t k = cnt[r]++;
                        if (!MASK && k < (uint32_t)kRecInline) {
                            uint64_t *w = &st.pw[k / 3][MASK ? 0 : r * kLookupBlock + tid];
                            const uint64_t p = (uint64_t)(tx[r] | (1u << 16)) << (kPairBits * (k % 3));
                            *w = (k % 3) ? (*w | p) : p;
                        }
                    }
                    if (kv[r] > lp.k_low && shares_prefix<W>(ek, qv[r], up_sh)) up[r] = (up[r] ? up[r] : (uint32_t)pos[r] << 12) + 1;
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
        const uint64_t rank = rank_base + tile * tq + (uint64_t)(r * kLookupBlock + tid);
        // the record stays in registers on the common paths; only the general path (out of line) goes through an array
        uint64_t o0 = rank & 0xffffffffull, o1 = 0, o2 = 0, o3 = 0;
        bool general = false;
        int64_t gen_g0 = 0, gen_b = 0;
        if (staged && MASK) {
            if (mask[r]) {
                uint64_t uw[2] = {0, 0};
                uint64_t m1 = 0;                                  // M2: the taxa matching at least k_low + 1 letters
                uint32_t hmax = 1;
                bool fits = true;
                for (uint32_t e = up[r] >> 12, end = (up[r] >> 12) + (up[r] & 4095u); e < end; ++e) {
                    const uint32_t h = match_levels<W>(sacc.template key<W>((int64_t)e), qv[r], kv[r], lp);
                    if (M2) { m1 |= 1ull << (sacc.tax[e] - 1); if (h >= 3u) fits = up3_pair_raise(uw[0], sacc.tax[e], h) && fits; }
                    else fits = up_pair_raise(uw, sacc.tax[e], h) && fits;
                    hmax = max(hmax, h);
                }
                if (fits) {
                    o0 = rec_head(rank, (uint32_t)__popcll(mask[r]), hmax, false, owner) | kRecMaskBit | (M2 ? kRecMask2Bit : 0ull);
                    o1 = mask[r]; o2 = M2 ? m1 : uw[0]; o3 = M2 ? uw[0] : uw[1];
                } else {
                    // more than six taxa above k_low: list record from the whole group (its first entry is searched for only now)
                    general = true; gen_b = span;
                    gen_g0 = lower_bound<W>(sacc, 0, (int64_t)span, prefix_floor<W>(qv[r], low_sh));
                }
            }
        } else if (staged) {
            const uint32_t c = cnt[r];
            if (c > (uint32_t)kRecInline) { general = true; gen_g0 = g0[r]; gen_b = span; }
            else if (c) {
                uint64_t pw[3];
#pragma unroll
                for (int w = 0; w < 3; ++w) pw[w] = c > 3u * w ? st.pw[w][MASK ? 0 : r * kLookupBlock + tid] : 0ull;
                uint32_t hmax = 1;
                for (uint32_t e = up[r] >> 12, end = (up[r] >> 12) + (up[r] & 4095u); e < end; ++e) {
                    const uint32_t h = match_levels<W>(sacc.template key<W>((int64_t)e), qv[r], kv[r], lp);
                    pair_raise(pw, c, sacc.tax[e], h);
                    hmax = max(hmax, h);
                }
                o0 = rec_head(rank, c, hmax, false, owner); o1 = pw[0]; o2 = pw[1]; o3 = pw[2];
            }
        } else if ((act >> r) & 1u) {
            uint64_t out[4];
            const int64_t gg = lower_bound<W>(gacc, a, b, prefix_floor<W>(qv[r], low_sh));
            match_general<W>(gacc, gg, b, qv[r], kv[r], lp, ovf, owner, err_flag, rank, out);
            o0 = out[0]; o1 = out[1]; o2 = out[2]; o3 = out[3];
        }
        if (general) {
            uint64_t out[4];
            match_general<W>(sacc, gen_g0, gen_b, qv[r], kv[r], lp, ovf, owner, err_flag, rank, out);
            o0 = out[0]; o1 = out[1]; o2 = out[2]; o3 = out[3];
        }
        // one 256-bit store = one full 32-byte sector per record (STG.256, sm_100)
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec + (uint64_t)iv[r] * kRecWords), "l"(o0), "l"(o1), "l"(o2), "l"(o3) : "memory");
    }
}

// rec[inst[i]] = staged[i]: the origin of a prefix-sharded batch pulls the records a peer computed for its sorted
// run (contiguous, so the NVLink side is a streaming read) and scatters them into emission order in its own HBM.
// Scattered stores INTO peer memory run ~100x slower than this (measured: DESIGN.md §5), hence the pull.
__global__ void __launch_bounds__(256) scatter_records_kernel(const uint64_t *__restrict__ staged, const uint32_t *__restrict__ inst,
                                                              uint64_t n, uint64_t *__restrict__ rec)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = ldg_stream(inst + i);
    uint64_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(staged + i * kRecWords));
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(rec + (uint64_t)j * kRecWords), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// Taxon-first flags (lookup.cuh: IndexDev): entry e is flagged unless an earlier entry of its k_low-letter group has the
// same taxon.  Thread per entry, backward scan to the start of the group (groups are short: 1.3 entries on average for
// the saturated C2 index; the scan is quadratic in the group size, which only shows for k_low <= 3 on large indexes).
// Keys are read with the flag bit stripped, so concurrent flagging of other entries is harmless.
template <bool W>
__global__ void __launch_bounds__(256) index_flags_kernel(uint64_t *__restrict__ key_lo, uint64_t *__restrict__ key_hi,
                                                          const uint16_t *__restrict__ tax, uint64_t n, int low_sh)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    volatile uint64_t *vlo = key_lo, *vhi = key_hi;
    const Key me = strip<W>(vlo[e], W ? vhi[e] : 0ull);
    const uint32_t t = tax[e];
    bool first = true;
    for (uint64_t p = e; p-- > 0;) {
        const Key o = strip<W>(vlo[p], W ? vhi[p] : 0ull);
        if (!shares_prefix<W>(o, me, low_sh)) break;
        if (tax[p] == t) { first = false; break; }
    }
    if (first) {
        if (W) atomicOr(reinterpret_cast<unsigned long long *>(key_hi + e), (unsigned long long)kKeyFlag);
        else atomicOr(reinterpret_cast<unsigned long long *>(key_lo + e), (unsigned long long)kKeyFlag);
    }
}

// ---- taxa mask of every k_low-letter group, stored with each of its entries (IndexDev::gmask; content files of <= 64 taxa).
// Group heads -> exclusive scan = group number -> atomicOr of the entries' taxon bits into one word per group -> every entry
// fetches its group's word.  One pass per step over the index at load time.
template <bool W>
__device__ __forceinline__ bool group_head(const uint64_t *__restrict__ key_lo, const uint64_t *__restrict__ key_hi, uint64_t e, int low_sh)
{
    if (e == 0) return true;
    const Key me = strip<W>(key_lo[e], W ? key_hi[e] : 0ull), pr = strip<W>(key_lo[e - 1], W ? key_hi[e - 1] : 0ull);
    return !shares_prefix<W>(pr, me, low_sh);
}
template <bool W>
__global__ void __launch_bounds__(256) group_heads_kernel(const uint64_t *__restrict__ key_lo, const uint64_t *__restrict__ key_hi, uint64_t n, int low_sh,
                                                          uint64_t *__restrict__ out)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = group_head<W>(key_lo, key_hi, e, low_sh) ? 1ull : 0ull;
}
template <bool W>
__global__ void __launch_bounds__(256) group_or_kernel(const uint64_t *__restrict__ key_lo, const uint64_t *__restrict__ key_hi,
                                                       const uint16_t *__restrict__ tax, uint64_t n, int low_sh, const uint64_t *__restrict__ before,
                                                       unsigned long long *__restrict__ gm)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t g = before[e] + (group_head<W>(key_lo, key_hi, e, low_sh) ? 1ull : 0ull) - 1ull;
    const uint32_t t = tax[e];
    if (t >= 1u && t <= 64u) atomicOr(gm + g, 1ull << (t - 1u));
}
template <bool W>
__global__ void __launch_bounds__(256) group_expand_kernel(const uint64_t *__restrict__ key_lo, const uint64_t *__restrict__ key_hi, uint64_t n, int low_sh,
                                                           const uint64_t *__restrict__ gm, uint64_t *__restrict__ inout)
{
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t g = inout[e] + (group_head<W>(key_lo, key_hi, e, low_sh) ? 1ull : 0ull) - 1ull;
    inout[e] = gm[g];
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

}  // namespace

void launch_scatter_records(cudaStream_t s, const uint64_t *staged, const uint32_t *inst, uint64_t n, uint64_t *rec, uint64_t *launches)
{
    if (!n) return;
    scatter_records_kernel<<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(staged, inst, n, rec);
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

void launch_lookup(cudaStream_t s, const uint64_t *q_lo, const uint64_t *q_hi, const uint32_t *inst, uint64_t n, uint64_t rank_base,
                   IndexDev idx, LevelParams lp, const uint64_t *part, uint64_t *rec, uint32_t tq, OverflowBuf ovf, uint32_t owner,
                   uint32_t *err_flag, uint64_t *launches, bool two_masks)
{
    if (!n) return;
    const unsigned tiles = (unsigned)div_up<uint64_t>(n, tq);
    auto go = [&](auto kern, size_t sm) {
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int per_sm = 4;
        KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kLookupBlock, sm));
        static const int pf_env = getenv("KASA_TUNE_LOOKUP_PREFETCH") ? atoi(getenv("KASA_TUNE_LOOKUP_PREFETCH")) : -1;
        // measured on C2: 62.1 ms with the prefetch one wave (= the resident CTAs) ahead, 59.1 ms without — the kernel sits at 85 % of
        // the random-sector write rate of its record scatter, which a warmer L2 on the load side does not move; off unless asked for
        (void)per_sm;
        const uint32_t pf_dist = pf_env > 0 ? (uint32_t)pf_env : 0u;
        kern<<<tiles, kLookupBlock, sm, s>>>(q_lo, q_hi, inst, n, rank_base, idx, lp, part, rec, tq, ovf, owner, err_flag, pf_dist);
    };
    const bool mask = lp.n_taxa <= kMaskTaxa;
    KASA_REQUIRE(!mask || idx.gmask || idx.n == 0, -4, "the index of a small content file has no group masks (launch_group_masks)");
    KASA_REQUIRE(!two_masks || mask, -6, "two-mask records need a content file of <= 64 taxa");
    if (idx.key_hi) {
        if (mask && two_masks) go(lookup_kernel<true, true, true>, sizeof(SpanStage<true, true>));
        else if (mask) go(lookup_kernel<true, true, false>, sizeof(SpanStage<true, true>));
        else go(lookup_kernel<true, false, false>, sizeof(SpanStage<true, false>));
    } else {
        if (mask && two_masks) go(lookup_kernel<false, true, true>, sizeof(SpanStage<false, true>));
        else if (mask) go(lookup_kernel<false, true, false>, sizeof(SpanStage<false, true>));
        else go(lookup_kernel<false, false, false>, sizeof(SpanStage<false, false>));
    }
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

void launch_index_flags(cudaStream_t s, uint64_t *key_lo, uint64_t *key_hi, const uint16_t *tax, uint64_t n, int K, int k_low,
                        uint64_t *launches)
{
    if (!n) return;
    const int low_sh = 5 * (K - k_low);
    if (key_hi) index_flags_kernel<true><<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(key_lo, key_hi, tax, n, low_sh);
    else index_flags_kernel<false><<<(unsigned)div_up<uint64_t>(n, 256), 256, 0, s>>>(key_lo, key_hi, tax, n, low_sh);
    ++*launches;
}

void launch_group_heads(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, uint64_t n, int K, int k_low, uint64_t *out, uint64_t *launches)
{
    if (!n) return;
    const int low_sh = 5 * (K - k_low);
    const unsigned grid = (unsigned)div_up<uint64_t>(n, 256);
    if (key_hi) group_heads_kernel<true><<<grid, 256, 0, s>>>(key_lo, key_hi, n, low_sh, out);
    else group_heads_kernel<false><<<grid, 256, 0, s>>>(key_lo, key_hi, n, low_sh, out);
    ++*launches;
}

void launch_group_masks(cudaStream_t s, const uint64_t *key_lo, const uint64_t *key_hi, const uint16_t *tax, uint64_t n, int K, int k_low,
                        uint64_t *gm, uint64_t *inout, uint64_t *launches)
{
    if (!n) return;
    const int low_sh = 5 * (K - k_low);
    const unsigned grid = (unsigned)div_up<uint64_t>(n, 256);
    if (key_hi) {
        group_or_kernel<true><<<grid, 256, 0, s>>>(key_lo, key_hi, tax, n, low_sh, inout, reinterpret_cast<unsigned long long *>(gm));
        group_expand_kernel<true><<<grid, 256, 0, s>>>(key_lo, key_hi, n, low_sh, gm, inout);
    } else {
        group_or_kernel<false><<<grid, 256, 0, s>>>(key_lo, key_hi, tax, n, low_sh, inout, reinterpret_cast<unsigned long long *>(gm));
        group_expand_kernel<false><<<grid, 256, 0, s>>>(key_lo, key_hi, n, low_sh, gm, inout);
    }
    *launches += 2;
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
