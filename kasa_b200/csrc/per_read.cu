// per_read.cu — per-read scores and per-taxon counts from the match records (3-bit records: num_k <= 7,
// <= 63 taxa; the default k = 12..7).  Replaces the score matrix updates of scoreMatchNonAVX (Compare.hpp:516-532).
//
// KASA_SCORE_FAST is computed in exact integer arithmetic.  The reference adds, per hit instance and level,
// w(k) * (1.f / |taxa_k|) to score[read][taxon] (Compare.hpp:924, 528-530); mathematically
//     score[r][t] = sum_k w(k) * sum_{instances of r} [t in taxa_k] / |taxa_k| ,
// and the inner sums are exactly the read's share of the profile counts.  They are accumulated here as integers
// (unique hits) and 2^-32 fixed point (shared hits), which is order independent: deterministic, independent of
// the batch composition, and exact to < 1e-7 — the reference's own fp32 running sum drifts from this value by
// up to ~5e-6 relative depending on its flush order; KASA_SCORE_EXACT reproduces that order.
//
// One warp per read, one lane per taxon (lane l owns taxa l+1 and l+33); records are loaded coalesced, 32 at a
// time, and digested by their own lane (pre-pass), then replayed in a warp-uniform loop:
//   * k_low level (the bulk: stray 7-mer hits, heavy-tailed set sizes): bit-matrix walk — every lane multiplies
//     "its" bit of the record's dense taxa mask with 2^32/|taxa_0| into a 64-bit register (SHF/LOP + IMAD.WIDE
//     per taxon).  No shared-memory read-modify-write, no atomics, no divergence.
//   * levels above k_low: almost always a single taxon per instance and the same one for most of the read — the
//     pre-pass finds the chunk's most frequent such taxon (MATCH + REDUX) and adds the packed unique-hit counters
//     of all its instances at once to the owner lane; other taxa are replayed one by one.  Levels shared by
//     several taxa (rare) go to a small per-warp fixed-point array.
#include "lookup.cuh"

namespace kasa {
namespace {

constexpr int kTxBlock = 128;                 // 4 warps, one read each
constexpr int kTxWarps = kTxBlock / 32;
constexpr uint64_t kL = 0x1249249249249249ull;                     // bit 3f, f = 0..20
constexpr unsigned long long kLevelOnes7 = 0x0000040810204080ull;  // 1 << 7j, j = 1..6  (|taxa_j| pack)

struct __align__(16) Entry {      // per instance, written by its lane in the pre-pass
    uint32_t d_lo, d_hi;          // dense mask of all taxa matched at k_low (bit s <-> taxon s + 1); 0 if |taxa_0| == 1
    uint32_t fx;                  // 2^32 / |taxa_0| (0 if |taxa_0| == 1: that hit is unique and counted via `um`)
    uint32_t meta;                // members to replay one by one: first member index | count << 8 | taxon of member 0 << 16; 0 = none
    unsigned long long um;        // member 0: 1 << 8j for every level j it is the only taxon of
    unsigned long long np;        // |taxa_j|, 7 bits per level
};
static_assert(sizeof(Entry) == 32, "Entry is two 16-byte words");

// unique-level mask (8-bit fields) of a taxon matched at h levels; *shared = it shares a level above k_low
__device__ __forceinline__ unsigned long long unique_levels(unsigned long long np, int h, bool *shared)
{
    unsigned long long um = 0;
    bool sh = false;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const uint32_t nj = (uint32_t)(np >> (7 * j)) & 127u;
        if (j < h && nj == 1) um |= 1ull << (8 * j);
        if (j && j < h && nj >= 2) sh = true;
    }
    *shared = sh;
    return um;
}

// 7 x 8-bit packed counters -> two words of 16-bit counters (levels 0-3, 4-6)
__device__ __forceinline__ void widen_add(unsigned long long x, unsigned long long &a, unsigned long long &b)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    a += (unsigned long long)(lo & 0xffu) | ((unsigned long long)((lo >> 8) & 0xffu) << 16) |
         ((unsigned long long)((lo >> 16) & 0xffu) << 32) | ((unsigned long long)(lo >> 24) << 48);
    b += (unsigned long long)(hi & 0xffu) | ((unsigned long long)((hi >> 8) & 0xffu) << 16) | ((unsigned long long)((hi >> 16) & 0xffu) << 32);
}

}  // namespace

__global__ void __launch_bounds__(kTxBlock) per_read_taxon_kernel(const uint4 *__restrict__ rec, const uint64_t *__restrict__ read_koff,
                                                                  uint32_t n_reads, LevelParams lp, float *__restrict__ scores,
                                                                  unsigned long long *__restrict__ cnt_fx,
                                                                  unsigned long long *__restrict__ cnt_unique)
{
    __shared__ Entry s_ent[kTxWarps][32];
    __shared__ uint64_t s_rec[kTxWarps][32][3];
    __shared__ uint32_t s_un[7 * 64];                          // unique counts of this CTA
    __shared__ unsigned long long s_fr[kTxWarps][7 * 64];      // current read: 2^-32 sums of levels above k_low shared by several taxa
    __shared__ unsigned long long s_frtot[7 * 64];             // ... summed over the CTA's reads
    __shared__ uint32_t s_invfx[128];
    __shared__ float s_w[8];
    const int T = lp.n_taxa, NK = lp.num_k;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 7 * 64; i += kTxBlock) {
        s_un[i] = 0; s_frtot[i] = 0;
        for (int ww = 0; ww < kTxWarps; ++ww) s_fr[ww][i] = 0;
    }
    for (int i = tid; i < 128; i += kTxBlock) s_invfx[i] = i >= 2 ? (uint32_t)(4294967296.0 / i + 0.5) : 0u;
    if (tid < 8) { int k = lp.k_low + tid; s_w[tid] = (float)(k * k) / 625.f; }   // Compare.hpp:392
    __syncthreads();

    const int my0 = lane + 1, my1 = lane + 33;                // taxa owned by this lane
    unsigned long long tot0 = 0, tot1 = 0;                    // k_low non-unique counts over all reads of this warp
    Entry *ent = s_ent[w];

    const uint32_t warps_total = gridDim.x * kTxWarps;
#pragma unroll 1
    for (uint32_t r = blockIdx.x * kTxWarps + w; r < n_reads; r += warps_total) {
        const uint64_t s0 = read_koff[r], s1 = read_koff[r + 1];
        unsigned long long acc0 = 0, acc1 = 0;                // k_low level: 2^-32 sums of my two taxa
        unsigned long long ua0 = 0, ub0 = 0, ua1 = 0, ub1 = 0; // unique hits per level of my two taxa (16-bit fields)
        bool any_shared = false;
#pragma unroll 1
        for (uint64_t c0 = s0; c0 < s1; c0 += 32) {
            // ---- pre-pass: lane = instance
            {
                const uint64_t s = c0 + lane;
                uint64_t Rw[3] = {0, 0, 0}, dall = 0;
                if (s < s1) {
                    const uint4 x0 = __ldg(rec + 2 * s), x1 = __ldg(rec + 2 * s + 1);
                    Rw[0] = (uint64_t)x0.x | ((uint64_t)x0.y << 32); Rw[1] = (uint64_t)x0.z | ((uint64_t)x0.w << 32);
                    Rw[2] = (uint64_t)x1.x | ((uint64_t)x1.y << 32); dall = (uint64_t)x1.z | ((uint64_t)x1.w << 32);
                }
                const uint32_t n0 = __popcll(dall);
                uint32_t nB = 0;
                int top_t = 0, top_h = 0;
                unsigned long long np = n0;
#pragma unroll
                for (int wi = 0; wi < 3; ++wi) {
                    uint64_t m = ((Rw[wi] >> 1) | (Rw[wi] >> 2)) & kL;          // fields >= 2: matched above k_low
                    nB += __popcll(m);
                    while (m) {
                        const int f = ((__ffsll((long long)m) - 1) * 171) >> 9;
                        const int h = (int)((Rw[wi] >> (3 * f)) & 7u);
                        m &= m - 1;
                        np += kLevelOnes7 & ((1ull << (7 * h)) - 1);
                        if (!top_t) { top_t = wi * 21 + f + 1; top_h = h; }
                    }
                }
                if (!nB && n0 == 1) { top_t = __ffsll((long long)dall); top_h = 1; }   // lone k_low hit: unique
                bool sh = false;
                const unsigned long long um = top_t ? unique_levels(np, top_h, &sh) : 0ull;
                // the chunk's most frequent top taxon is counted once per chunk, not once per instance
                const uint32_t grp = __match_any_sync(kFull, top_t);
                const uint32_t key = (top_t && !sh) ? ((uint32_t)__popc(grp) << 8 | (uint32_t)top_t) : 0u;
                const int tm = (int)(__reduce_max_sync(kFull, key) & 255u);
                const bool major = tm && top_t == tm && !sh;
                const uint32_t sum_lo = __reduce_add_sync(kFull, major ? (uint32_t)um : 0u);
                const uint32_t sum_hi = __reduce_add_sync(kFull, major ? (uint32_t)(um >> 32) : 0u);
                if (tm && lane == ((tm - 1) & 31)) {
                    const unsigned long long sum = (unsigned long long)sum_lo | ((unsigned long long)sum_hi << 32);
                    if (tm <= 32) widen_add(sum, ua0, ub0); else widen_add(sum, ua1, ub1);
                }
                // what is left to replay one by one: a foreign / sharing top member, and any further members
                const uint32_t first = major ? 1u : 0u;
                const uint32_t count = top_t ? (nB >= 2 ? nB : 1u) : 0u;
                Entry e;
                const bool shared0 = n0 >= 2;
                e.d_lo = shared0 ? (uint32_t)dall : 0u; e.d_hi = shared0 ? (uint32_t)(dall >> 32) : 0u;
                e.fx = s_invfx[n0];
                e.meta = count > first ? (first | (count << 8) | ((uint32_t)top_t << 16) | (sh ? 1u << 31 : 0u) | ((uint32_t)top_h << 24 & 0x7f000000u)) : 0u;
                e.um = um; e.np = np;
                ent[lane] = e;
                if (e.meta) { s_rec[w][lane][0] = Rw[0]; s_rec[w][lane][1] = Rw[1]; s_rec[w][lane][2] = Rw[2]; }
            }
            __syncwarp();
            // ---- main pass: lane = taxon, warp-uniform walk over the instances
            const int cnt = (int)min((uint64_t)32, s1 - c0);
#pragma unroll 1
            for (int i = 0; i < cnt; ++i) {
                const uint4 a = *reinterpret_cast<const uint4 *>(&ent[i]);       // d_lo, d_hi, fx, meta (broadcast)
                acc0 += (unsigned long long)((a.x >> lane) & 1u) * a.z;             // bit * 2^32/|taxa_0|
                acc1 += (unsigned long long)((a.y >> lane) & 1u) * a.z;
                if (!a.w) continue;
                // ---- rare path: members that are not the chunk's majority taxon
                const uint4 b = *reinterpret_cast<const uint4 *>(&ent[i].um);       // um, np
                const unsigned long long np = (unsigned long long)b.z | ((unsigned long long)b.w << 32);
                const int first = (int)(a.w & 255u), count = (int)((a.w >> 8) & 255u);
                int wi = 0;
                uint64_t mwalk = 0;
#pragma unroll 1
                for (int k = 0; k < count; ++k) {
                    int t, h;
                    unsigned long long um;
                    bool sh;
                    if (k == 0) {
                        t = (int)((a.w >> 16) & 255u); h = (int)((a.w >> 24) & 127u); sh = a.w >> 31;
                        um = (unsigned long long)b.x | ((unsigned long long)b.y << 32);
                        if (count > 1) {                                          // position the walk behind the top member
                            mwalk = ((s_rec[w][i][0] >> 1) | (s_rec[w][i][0] >> 2)) & kL;
                            while (!mwalk) { ++wi; mwalk = ((s_rec[w][i][wi] >> 1) | (s_rec[w][i][wi] >> 2)) & kL; }
                            mwalk &= mwalk - 1;
                        }
                    } else {
                        while (!mwalk) { ++wi; mwalk = ((s_rec[w][i][wi] >> 1) | (s_rec[w][i][wi] >> 2)) & kL; }
                        const int f = ((__ffsll((long long)mwalk) - 1) * 171) >> 9;
                        mwalk &= mwalk - 1;
                        t = wi * 21 + f + 1;
                        h = (int)((s_rec[w][i][wi] >> (3 * f)) & 7u);
                        um = unique_levels(np, h, &sh);
                    }
                    if (k < first) continue;                                      // counted with the chunk's majority
                    if (lane == ((t - 1) & 31)) {
                        if (t <= 32) widen_add(um, ua0, ub0); else widen_add(um, ua1, ub1);
                    }
                    if (sh) {                                                     // the warp owns s_fr[w]
                        any_shared = true;
                        if (lane > 0 && lane < h) {
                            const uint32_t nj = (uint32_t)(np >> (7 * lane)) & 127u;
                            if (nj >= 2) s_fr[w][lane * 64 + t] += s_invfx[nj];
                        }
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        // ---- read epilogue: score[r][t] = sum_k w(k) * (unique hits + shared sums) ; counts
        tot0 += acc0;
        tot1 += acc1;
#pragma unroll
        for (int slot = 0; slot < 2; ++slot) {
            const int t = slot ? my1 : my0;
            if (t >= T) continue;
            const unsigned long long ua = slot ? ua1 : ua0, ub = slot ? ub1 : ub0;
            double x = (double)(slot ? acc1 : acc0) * (double)s_w[0];
            if (ua | ub | (unsigned long long)any_shared) {
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const uint32_t u = (uint32_t)((j < 4 ? ua >> (16 * j) : ub >> (16 * (j - 4))) & 0xffffull);
                    unsigned long long fr = 0;
                    if (any_shared && j) { fr = s_fr[w][j * 64 + t]; if (fr) { s_fr[w][j * 64 + t] = 0; atomicAdd(&s_frtot[j * 64 + t], fr); } }
                    if (u) atomicAdd(&s_un[j * 64 + t], u);
                    x += ((double)u * 4294967296.0 + (double)fr) * (double)s_w[j];
                }
            }
            if (scores) scores[(uint64_t)r * T + t] = (float)(x * (1.0 / 4294967296.0));
        }
        if (scores && lane == 0) scores[(uint64_t)r * T] = 0.f;
        __syncwarp();
    }
    if (my0 < T && tot0) atomicAdd(cnt_fx + my0, tot0);        // level 0 (k_low) row of the counts
    if (my1 < T && tot1) atomicAdd(cnt_fx + my1, tot1);
    __syncthreads();
    for (int i = tid; i < 7 * 64; i += kTxBlock) {
        const int j = i / 64, t = i % 64;
        if (j >= NK || t >= T) continue;
        if (s_un[i]) atomicAdd(cnt_unique + j * T + t, (unsigned long long)s_un[i]);
        if (s_frtot[i]) atomicAdd(cnt_fx + j * T + t, s_frtot[i]);
    }
}

void launch_per_read_taxon(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                           float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique)
{
    unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kTxWarps);
    if (grid > (unsigned)kSMs * 16) grid = kSMs * 16;
    per_read_taxon_kernel<<<grid, kTxBlock, 0, s>>>(reinterpret_cast<const uint4 *>(rec), read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
}

}  // namespace kasa
