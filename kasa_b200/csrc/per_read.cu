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
// One warp per read, one lane per taxon (lane l owns taxa l+1 and l+33).  Records are loaded coalesced, 32 at a
// time, by "their" lane, which turns the record's level sizes (|taxa_j|, the record's fourth word, written by
// the lookup kernel) into an 8-entry table indexed by the level count h a taxon can have in this instance:
//     entry[h] = one bit per level j < h with |taxa_j| == 1 (8-bit fields)
// The main pass is then a branch-free, warp-uniform walk over the instances with a non-empty record (rows are stored compacted): every lane extracts its taxon's
// 3-bit field h from the record (LDS + SHF + LOP), fetches entry[h] (one conflict-free LDS.64), adds it to packed
// per-level unique counters, and adds min(h, 1) * 2^32/|taxa_0| (shared k_low hits, the bulk: stray 7-mer hits)
// to a 64-bit accumulator with one IMAD.WIDE.  No shared-memory read-modify-write, no atomics, no divergence.  Instances where a level above k_low is shared by several taxa (a few per cent) are
// replayed afterwards into lane-private fixed-point cells.
#include "lookup.cuh"

namespace kasa {
namespace {

constexpr int kTxBlock = 128;                 // 4 warps, one read each
constexpr int kTxWarps = kTxBlock / 32;
constexpr int kRowBytes = 8 * 8 + 8;          // table row of one instance: 8 entries + pad (rows of a half-warp hit distinct banks)
constexpr int kWidenEvery = 4;                // chunks of 32 instances between widenings of the 8-bit counters (4 * 32 < 256)
constexpr int kFoldEvery = 256;               // widenings between folds of the 16-bit counters (256 * 128 < 65536)

// 7 x 8-bit packed counters -> two words of 16-bit counters (levels 0-3, 4-6)
__device__ __forceinline__ void widen_add(unsigned long long x, unsigned long long &a, unsigned long long &b)
{
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    a += (unsigned long long)(lo & 0xffu) | ((unsigned long long)((lo >> 8) & 0xffu) << 16) |
         ((unsigned long long)((lo >> 16) & 0xffu) << 32) | ((unsigned long long)(lo >> 24) << 48);
    b += (unsigned long long)(hi & 0xffu) | ((unsigned long long)((hi >> 8) & 0xffu) << 16) | ((unsigned long long)((hi >> 16) & 0xffu) << 32);
}

template <int SLOTS>
__global__ void __launch_bounds__(kTxBlock) per_read_taxon_kernel(const uint4 *__restrict__ rec, const uint64_t *__restrict__ read_koff,
                                                                  uint32_t n_reads, LevelParams lp, float *__restrict__ scores,
                                                                  unsigned long long *__restrict__ cnt_fx,
                                                                  unsigned long long *__restrict__ cnt_unique)
{
    __shared__ __align__(16) unsigned char s_tab[kTxWarps][32 * kRowBytes];
    __shared__ __align__(16) uint32_t s_fx[kTxWarps][32];      // 2^32 / |taxa_0| of the chunk's instances (0 if |taxa_0| < 2)
    __shared__ uint64_t s_rec[kTxWarps][4][32];                // field words of the chunk, word-major (row 3: level sizes)
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

    // taxa owned by this lane and where their 3-bit fields live (field f = taxon - 1: word f / 21, bit 3 * (f % 21))
    const int my[2] = {lane + 1, lane + 33};
    const bool own1 = SLOTS > 1 && my[1] < T;
    const uint64_t *fw0 = &s_rec[w][lane / 21][0];
    const int sh0 = 3 * (lane % 21);
    const uint64_t *fw1 = &s_rec[w][own1 ? (lane + 32) / 21 : 0][0];
    const int sh1 = own1 ? 3 * ((lane + 32) % 21) : 0;
    const unsigned char *tab = s_tab[w];
    unsigned long long tot[2] = {0, 0};                       // k_low shared counts over all reads of this warp

    const uint32_t warps_total = gridDim.x * kTxWarps;
#pragma unroll 1
    for (uint32_t r = blockIdx.x * kTxWarps + w; r < n_reads; r += warps_total) {
        const uint64_t s0 = read_koff[r], s1 = read_koff[r + 1];
        unsigned long long acc0 = 0, acc1 = 0;                // k_low level: 2^-32 sums of my two taxa
        unsigned long long u8a = 0, u8b = 0;                  // unique hits per level, 8-bit fields
        unsigned long long ua[2] = {0, 0}, ub[2] = {0, 0};    // ... 16-bit fields (levels 0-3, 4-6)
        double x[2] = {0.0, 0.0};                             // sum_j w_j * unique_j (in units of 2^-32), folded so far
        bool any_shared = false;
        uint32_t chunk = 0, widened = 0;

        auto fold16 = [&]() {                                  // 16-bit counters -> score sums and the CTA's unique counts
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                if (!(ua[slot] | ub[slot]) || my[slot] >= T) { ua[slot] = ub[slot] = 0; continue; }
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const uint32_t u = (uint32_t)((j < 4 ? ua[slot] >> (16 * j) : ub[slot] >> (16 * (j - 4))) & 0xffffull);
                    if (u) { atomicAdd(&s_un[j * 64 + my[slot]], u); x[slot] += (double)u * 4294967296.0 * (double)s_w[j]; }
                }
                ua[slot] = ub[slot] = 0;
            }
        };

        // software pipeline over the read's chunks: the record of this lane's instance in the NEXT chunk is requested (one
        // 256-bit load, LDG.256) before the main pass of the current one, the chunk after that is pulled into L2
        uint64_t N0 = 0, N1 = 0, N2 = 0, N3 = 0;
        if (s0 + lane < s1)
            asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(N0), "=l"(N1), "=l"(N2), "=l"(N3) : "l"(rec + 2 * (s0 + lane)));
#pragma unroll 1
        for (uint64_t c0 = s0; c0 < s1; c0 += 32, ++chunk) {
            // ---- pre-pass: lane = instance
            uint32_t rare_mask, ne;
            int n_ne;
            {
                const uint64_t s = c0 + lane;
                if (s + 64 < s1) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + 2 * (s + 64)));
                const uint64_t R0 = N0, R1 = N1, R2 = N2, np = N3;
                N0 = N1 = N2 = N3 = 0;
                if (s + 32 < s1)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(N0), "=l"(N1), "=l"(N2), "=l"(N3) : "l"(rec + 2 * (s + 32)));
                // rows are stored compacted: instances with a non-empty record first, so that the main pass only walks those
                // (a quarter of the records of a typical read are empty); integer sums do not care about the order
                ne = __ballot_sync(kFull, (uint32_t)np != 0u);
                n_ne = __popc(ne);
                const int row_i = ((uint32_t)np != 0u) ? __popc(ne & lanemask_lt()) : n_ne + __popc(~ne & lanemask_lt());
                s_rec[w][0][row_i] = R0; s_rec[w][1][row_i] = R1; s_rec[w][2][row_i] = R2; s_rec[w][3][row_i] = np;
                // bytes of np equal to 1 -> 0x01 (exact zero-byte test on np ^ 0x0101..)
                const uint64_t y = np ^ 0x0101010101010101ull, m7 = 0x7f7f7f7f7f7f7f7full;
                const uint64_t um = (~(((y & m7) + m7) | y | m7)) >> 7;
                const uint32_t n0 = (uint32_t)np & 255u;
                const uint32_t fx = s_invfx[n0 & 127u];                           // 0 for |taxa_0| < 2
                rare_mask = __ballot_sync(kFull, ((np >> 8) & 0x00fefefefefefefeull) != 0);
                s_fx[w][row_i] = fx;
                uint64_t *row = reinterpret_cast<uint64_t *>(s_tab[w] + row_i * kRowBytes);
                row[0] = 0;
#pragma unroll
                for (int h = 1; h < 8; ++h) row[h] = um & ((1ull << (8 * h)) - 1);
            }
            __syncwarp();
            // ---- main pass: lane = taxon, warp-uniform, branch-free (rows past the end of the read are zero)
#pragma unroll 4
            for (int i = 0; i < ((n_ne + 3) & ~3); ++i) {
                const uint32_t fx = s_fx[w][i];
                const uint32_t h0 = (uint32_t)(fw0[i] >> sh0) & 7u;
                u8a += *reinterpret_cast<const uint64_t *>(tab + i * kRowBytes + h0 * 8);
                acc0 += (unsigned long long)min(h0, 1u) * fx;
                if (SLOTS > 1) {
                    const uint32_t h1 = (uint32_t)(fw1[i] >> sh1) & 7u;
                    u8b += *reinterpret_cast<const uint64_t *>(tab + i * kRowBytes + h1 * 8);
                    acc1 += (unsigned long long)min(h1, 1u) * fx;
                }
            }
            // ---- instances with a level above k_low shared by several taxa: lane-private cells, no atomics
            if (rare_mask) any_shared = true;
#pragma unroll 1
            for (uint32_t m = rare_mask; m; m &= m - 1) {
                const int i = __popc(ne & ((1u << (__ffs((int)m) - 1)) - 1u));     // row of that lane's (non-empty) record
                const uint64_t np = s_rec[w][3][i];
                const int h0 = (int)((uint32_t)(fw0[i] >> sh0) & 7u), h1 = (int)((uint32_t)(fw1[i] >> sh1) & 7u);
                uint64_t lv = (np >> 8) & 0x00fefefefefefefeull;                   // a bit in byte j-1 <=> |taxa_j| >= 2
#pragma unroll 1
                while (lv) {                                                       // warp-uniform: np is a broadcast
                    const int j = ((__ffsll((long long)lv) - 1) >> 3) + 1;
                    lv &= ~(0xffull << (8 * (j - 1)));
                    const uint32_t v = s_invfx[(uint32_t)(np >> (8 * j)) & 127u];
                    if (h0 > j) s_fr[w][j * 64 + my[0]] += v;
                    if (own1 && h1 > j) s_fr[w][j * 64 + my[1]] += v;
                }
            }
            if ((chunk % kWidenEvery) == kWidenEvery - 1) {
                widen_add(u8a, ua[0], ub[0]); u8a = 0;
                if (SLOTS > 1) { widen_add(u8b, ua[1], ub[1]); u8b = 0; }
                if ((++widened % kFoldEvery) == 0) fold16();
            }
            __syncwarp();
        }
        // ---- read epilogue: score[r][t] = sum_k w(k) * (unique hits + shared sums) ; counts
        widen_add(u8a, ua[0], ub[0]);
        if (SLOTS > 1) widen_add(u8b, ua[1], ub[1]);
        fold16();
        tot[0] += acc0;
        tot[1] += acc1;
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {
            const int t = my[slot];
            if (t >= T) continue;
            double xs = x[slot] + (double)(slot ? acc1 : acc0) * (double)s_w[0];
            if (any_shared) {
#pragma unroll 1
                for (int j = 1; j < NK; ++j) {
                    const unsigned long long fr = s_fr[w][j * 64 + t];
                    if (fr) { s_fr[w][j * 64 + t] = 0; atomicAdd(&s_frtot[j * 64 + t], fr); xs += (double)fr * (double)s_w[j]; }
                }
            }
            if (scores) scores[(uint64_t)r * T + t] = (float)(xs * (1.0 / 4294967296.0));
        }
        if (scores && lane == 0) scores[(uint64_t)r * T] = 0.f;
        __syncwarp();
    }
    if (my[0] < T && tot[0]) atomicAdd(cnt_fx + my[0], tot[0]);        // level 0 (k_low) row of the counts
    if (own1 && tot[1]) atomicAdd(cnt_fx + my[1], tot[1]);
    __syncthreads();
    for (int i = tid; i < 7 * 64; i += kTxBlock) {
        const int j = i / 64, t = i % 64;
        if (j >= NK || t >= T) continue;
        if (s_un[i]) atomicAdd(cnt_unique + j * T + t, (unsigned long long)s_un[i]);
        if (s_frtot[i]) atomicAdd(cnt_fx + j * T + t, s_frtot[i]);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// The same organisation for 4- and 5-bit level fields (8..31 levels, e.g. k = 25..7 = 19 levels; 64 taxa, records of 4 / 8
// words without a level-size word): warp per read, lane per taxon (lane l owns taxa l+1 and l+33).  Per 32-instance chunk
//   pre-pass (lane = instance): load the record, count the level sizes |taxa_j| from its (few) non-zero fields into byte
//     counters nl[NW words], derive um = one 0x01 byte per level with |taxa_j| == 1 and fx = 2^32 / |taxa_0|, store the
//     rows compacted (non-empty records first);
//   main pass (lane = taxon): h = my field of instance i; unique-hit counters (8 bits per level, NW words) += um & below[h];
//     k_low shared hits += min(h, 1) * fx — branch-free, no atomics;
//   instances with a level above k_low shared by several taxa (rare) add 2^32 / |taxa_j| per such level to the CTA's
//     count cells (shared-memory atomics) and, times w(k), to a per-lane double.
// Every 4 chunks the 8-bit counters are folded: counts into the CTA's cells, counts x w(k) into the score sum.
// Integer / fixed-point sums in a fixed order: deterministic and batch independent, like the 3-bit kernel.
template <int HB, int NW, int SLOTS>
__global__ void __launch_bounds__(kTxBlock) per_read_wide_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff,
                                                                 uint32_t n_reads, LevelParams lp, float *__restrict__ scores,
                                                                 unsigned long long *__restrict__ cnt_fx,
                                                                 unsigned long long *__restrict__ cnt_unique)
{
    constexpr int RW = kRW<HB>, FPW = 64 / HB, RWU = (64 + FPW - 1) / FPW;     // words that can hold one of the 64 fields
    constexpr uint32_t HM = (1u << HB) - 1u;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = lp.n_taxa, NK = lp.num_k;
    uint64_t *s_below = reinterpret_cast<uint64_t *>(smem_raw);                 // [32][NW] bytes j < h -> 0xff
    uint64_t *s_ones = s_below + 32 * NW;                                       // [32][NW] bytes j < h -> 0x01
    unsigned long long *s_frtot = reinterpret_cast<unsigned long long *>(s_ones + 32 * NW);   // [NK][64] shared upper-level sums of the CTA
    uint64_t *s_rec = reinterpret_cast<uint64_t *>(s_frtot + NK * 64);          // [warps][RWU][32]
    uint64_t *s_um = s_rec + kTxWarps * RWU * 32;                               // [warps][NW][32]
    uint64_t *s_nl = s_um + kTxWarps * NW * 32;                                 // [warps][NW][32]
    uint32_t *s_un = reinterpret_cast<uint32_t *>(s_nl + kTxWarps * NW * 32);   // [NK][64] unique counts of the CTA
    uint32_t *s_fx = s_un + NK * 64;                                            // [warps][32]
    uint32_t *s_invfx = s_fx + kTxWarps * 32;                                   // [128]
    float *s_w = reinterpret_cast<float *>(s_invfx + 128);                      // [32]
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < NK * 64; i += kTxBlock) { s_un[i] = 0; s_frtot[i] = 0; }
    for (int i = tid; i < 128; i += kTxBlock) s_invfx[i] = i >= 2 ? (uint32_t)(4294967296.0 / i + 0.5) : 0u;
    for (int i = tid; i < 32 * NW; i += kTxBlock) {
        const int h = i / NW, k = i % NW;
        const int nb = min(8, max(0, h - 8 * k));
        const uint64_t ff = nb >= 8 ? ~0ull : ((1ull << (8 * nb)) - 1);
        s_below[i] = ff;
        s_ones[i] = ff & 0x0101010101010101ull;
    }
    if (tid < 32) { int k = lp.k_low + tid; s_w[tid] = (float)(k * k) / 625.f; }   // Compare.hpp:392
    __syncthreads();

    const int my[2] = {lane + 1, lane + 33};
    const bool own1 = SLOTS > 1 && my[1] < T;
    uint64_t *w_rec = s_rec + w * RWU * 32, *w_um = s_um + w * NW * 32, *w_nl = s_nl + w * NW * 32;
    uint32_t *w_fx = s_fx + w * 32;
    const uint64_t *fw0 = w_rec + (lane / FPW) * 32;
    const int sh0 = HB * (lane % FPW);
    const uint64_t *fw1 = w_rec + (own1 ? (lane + 32) / FPW : 0) * 32;
    const int sh1 = own1 ? HB * ((lane + 32) % FPW) : 0;
    unsigned long long tot[2] = {0, 0};                       // k_low shared counts over all reads of this warp

    const uint32_t warps_total = gridDim.x * kTxWarps;
#pragma unroll 1
    for (uint32_t r = blockIdx.x * kTxWarps + w; r < n_reads; r += warps_total) {
        const uint64_t s0 = read_koff[r], s1 = read_koff[r + 1];
        unsigned long long acc[2] = {0, 0};                   // k_low level: 2^-32 sums of my two taxa
        unsigned long long u8[2][NW];                         // unique hits per level, 8-bit fields
#pragma unroll
        for (int k = 0; k < NW; ++k) { u8[0][k] = 0; u8[1][k] = 0; }
        double x[2] = {0.0, 0.0};                             // sum_j w_j * (unique_j * 2^32 + shared_j), folded so far
        uint32_t chunk = 0;

        auto fold = [&]() {                                    // 8-bit counters -> the CTA's unique counts and the score sums
#pragma unroll
            for (int slot = 0; slot < SLOTS; ++slot) {
                if (my[slot] >= T) continue;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    unsigned long long v = u8[slot][k];
                    u8[slot][k] = 0;
                    while (v) {                                // few levels are non-zero
                        const int byte = (__ffsll((long long)v) - 1) >> 3;
                        const uint32_t u = (uint32_t)(v >> (8 * byte)) & 255u;
                        v &= ~(0xffull << (8 * byte));
                        const int j = 8 * k + byte;
                        atomicAdd(&s_un[j * 64 + my[slot] - 1], u);            // cells are indexed by taxon - 1 (64 taxa)
                        x[slot] += (double)u * 4294967296.0 * (double)s_w[j];
                    }
                }
            }
        };

#pragma unroll 1
        for (uint64_t c0 = s0; c0 < s1; c0 += 32, ++chunk) {
            // ---- pre-pass: lane = instance
            uint32_t rare_mask, ne;
            int n_ne;
            {
                const uint64_t s = c0 + lane;
                if (s + 32 < s1) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec + RW * (s + 32)));   // next chunk
                uint64_t R[RW];
#pragma unroll
                for (int i = 0; i < RW; ++i) R[i] = 0;
                if (s < s1) {
#pragma unroll
                    for (int i = 0; i < RW; i += 4)
                        asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];"
                                     : "=l"(R[i]), "=l"(R[i + 1]), "=l"(R[i + 2]), "=l"(R[i + 3]) : "l"(rec + RW * s + i));
                }
                uint64_t nl[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) nl[k] = 0;
#pragma unroll
                for (int wi = 0; wi < RWU; ++wi) {
                    uint64_t left = R[wi];
                    while (left) {
                        const int bit = __ffsll((long long)left) - 1;
                        const int f = HB == 4 ? bit >> 2 : (bit * 205) >> 10;       // bit / HB
                        const uint32_t h = (uint32_t)(R[wi] >> (f * HB)) & HM;
                        left &= ~((uint64_t)HM << (f * HB));
#pragma unroll
                        for (int k = 0; k < NW; ++k) nl[k] += s_ones[h * NW + k];
                    }
                }
                const bool nonempty = ((uint32_t)nl[0] & 255u) != 0u;
                ne = __ballot_sync(kFull, nonempty);
                n_ne = __popc(ne);
                const int row_i = nonempty ? __popc(ne & lanemask_lt()) : n_ne + __popc(~ne & lanemask_lt());
                bool rare = false;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    // bytes equal to 1 -> 0x01 (exact zero-byte test on nl ^ 0x0101..)
                    const uint64_t y = nl[k] ^ 0x0101010101010101ull, m7 = 0x7f7f7f7f7f7f7f7full;
                    w_um[k * 32 + row_i] = (~(((y & m7) + m7) | y | m7)) >> 7;
                    w_nl[k * 32 + row_i] = nl[k];
                    rare |= ((k == 0 ? nl[k] >> 8 : nl[k]) & 0xfefefefefefefefeull & (k == 0 ? 0x00ffffffffffffffull : ~0ull)) != 0;
                }
                rare_mask = __ballot_sync(kFull, rare);
#pragma unroll
                for (int wi = 0; wi < RWU; ++wi) w_rec[wi * 32 + row_i] = R[wi];
                w_fx[row_i] = s_invfx[(uint32_t)nl[0] & 127u];                    // 0 for |taxa_0| < 2
            }
            __syncwarp();
            // ---- main pass: lane = taxon, warp-uniform, branch-free (rows of empty records are zero)
#pragma unroll 2
            for (int i = 0; i < ((n_ne + 1) & ~1); ++i) {
                const uint32_t fx = w_fx[i];
                uint64_t um[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) um[k] = w_um[k * 32 + i];
                const uint32_t h0 = (uint32_t)(fw0[i] >> sh0) & HM;
#pragma unroll
                for (int k = 0; k < NW; ++k) u8[0][k] += um[k] & s_below[h0 * NW + k];
                acc[0] += (unsigned long long)min(h0, 1u) * fx;
                if (SLOTS > 1) {
                    const uint32_t h1 = own1 ? (uint32_t)(fw1[i] >> sh1) & HM : 0u;
#pragma unroll
                    for (int k = 0; k < NW; ++k) u8[1][k] += um[k] & s_below[h1 * NW + k];
                    acc[1] += (unsigned long long)min(h1, 1u) * fx;
                }
            }
            // ---- instances with a level above k_low shared by several taxa
#pragma unroll 1
            for (uint32_t m = rare_mask; m; m &= m - 1) {
                const int i = __popc(ne & ((1u << (__ffs((int)m) - 1)) - 1u));     // row of that lane's (non-empty) record
                const int h0 = (int)((uint32_t)(fw0[i] >> sh0) & HM), h1 = own1 ? (int)((uint32_t)(fw1[i] >> sh1) & HM) : 0;
#pragma unroll 1
                for (int j = 1; j < NK; ++j) {                                     // warp-uniform: the level sizes are a broadcast
                    const uint32_t nj = (uint32_t)(w_nl[(j >> 3) * 32 + i] >> (8 * (j & 7))) & 255u;
                    if (nj < 2) continue;
                    const uint32_t v = s_invfx[nj & 127u];
                    if (h0 > j) { atomicAdd(&s_frtot[j * 64 + my[0] - 1], (unsigned long long)v); x[0] += (double)v * (double)s_w[j]; }
                    if (own1 && h1 > j) { atomicAdd(&s_frtot[j * 64 + my[1] - 1], (unsigned long long)v); x[1] += (double)v * (double)s_w[j]; }
                }
            }
            if ((chunk % kWidenEvery) == kWidenEvery - 1) fold();
            __syncwarp();
        }
        // ---- read epilogue
        fold();
        tot[0] += acc[0];
        tot[1] += acc[1];
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {
            const int t = my[slot];
            if (t >= T) continue;
            const double xs = x[slot] + (double)acc[slot] * (double)s_w[0];
            if (scores) scores[(uint64_t)r * T + t] = (float)(xs * (1.0 / 4294967296.0));
        }
        if (scores && lane == 0) scores[(uint64_t)r * T] = 0.f;
        __syncwarp();
    }
    if (my[0] < T && tot[0]) atomicAdd(cnt_fx + my[0], tot[0]);        // level 0 (k_low) row of the counts
    if (own1 && tot[1]) atomicAdd(cnt_fx + my[1], tot[1]);
    __syncthreads();
    for (int i = tid; i < NK * 64; i += kTxBlock) {
        const int j = i / 64, t = i % 64 + 1;
        if (t >= T) continue;
        if (s_un[i]) atomicAdd(cnt_unique + j * T + t, (unsigned long long)s_un[i]);
        if (s_frtot[i]) atomicAdd(cnt_fx + j * T + t, s_frtot[i]);
    }
}

template <int HB, int NW>
void launch_wide(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp, float *scores,
                 unsigned long long *cnt_fx, unsigned long long *cnt_unique)
{
    constexpr int FPW = 64 / HB, RWU = (64 + FPW - 1) / FPW;
    const size_t sm = (size_t)2 * 32 * NW * 8 + (size_t)lp.num_k * 64 * 8 + (size_t)kTxWarps * RWU * 32 * 8 + (size_t)2 * kTxWarps * NW * 32 * 8 +
                      (size_t)lp.num_k * 64 * 4 + (size_t)kTxWarps * 32 * 4 + 128 * 4 + 32 * 4;
    unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kTxWarps);
    if (grid > (unsigned)kSMs * 16) grid = kSMs * 16;
    auto go = [&](auto kern) {
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        kern<<<grid, kTxBlock, sm, s>>>(rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    };
    if (lp.n_taxa <= 33) go(per_read_wide_kernel<HB, NW, 1>); else go(per_read_wide_kernel<HB, NW, 2>);
}

}  // namespace

void launch_per_read_taxon(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                           float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique)
{
    unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kTxWarps);
    if (grid > (unsigned)kSMs * 16) grid = kSMs * 16;
    if (lp.n_taxa <= 33)
        per_read_taxon_kernel<1><<<grid, kTxBlock, 0, s>>>(reinterpret_cast<const uint4 *>(rec), read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    else
        per_read_taxon_kernel<2><<<grid, kTxBlock, 0, s>>>(reinterpret_cast<const uint4 *>(rec), read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
}

// 4- and 5-bit records (per_read_wide_kernel): 8..15 levels -> two words of byte counters, 16..24 -> three, 25..31 -> four
void launch_per_read_wide(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp,
                          float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique)
{
    if (lp.hb == 4) launch_wide<4, 2>(s, rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    else if (lp.num_k <= 24) launch_wide<5, 3>(s, rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
    else launch_wide<5, 4>(s, rec, read_koff, n_reads, lp, scores, cnt_fx, cnt_unique);
}

}  // namespace kasa
