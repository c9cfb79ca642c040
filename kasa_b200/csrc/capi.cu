// capi.cu — the C ABI of include/kasa_b200.h: context, HBM layout, kernel orchestration.
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/kasa_b200.h"
#include "build.cuh"
#include "common.cuh"
#include "exact.cuh"
#include "extract.cuh"
#include "lookup.cuh"
#include "radix_sort.cuh"
#include "scan.cuh"
#include "score.cuh"
#include "shard.cuh"
#include "synth.cuh"

using namespace kasa;

constexpr int kSlots = 3;                // batches in flight (kasa_identify_submit / _wait)

struct kasa_ctx {
    kasa_params p{};
    std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    uint64_t launches = 0;
    int K = 12, num_k = 6;
    int sort_bits = 0;                  // leading k-mer bits identify sorts (tuning knob KASA_TUNE_SORT_BITS)
    bool score_sorted = true;           // presorted batches are scored straight from the records (KASA_TUNE_SCORE_SORTED=0: always via op streams)
    int score_lanes = 0;                // threads per read of that kernel (KASA_TUNE_SCORE_LANES; 0 = default)
    bool wide = false;
    std::vector<void *> allocs;

    // record arrays (sort ping-pong; shared by identify batches and the build brick)
    bool ws_ready = false;
    uint64_t cap_kmers = 0;
    SortBuffers sb{};
    char *sort_slab = nullptr;          // the six record arrays of `sb` in one allocation; the op streams of score.cu reuse it after the lookup
    size_t sort_slab_bytes = 0;
    void *sort_ws = nullptr;
    uint32_t *d_err = nullptr;          // [4] device error flags
    uint8_t *d_codon = nullptr;         // [512] 5-bit letter of every codon (built-in table or `-a`)
    // sequence staging
    char *d_ascii = nullptr;            // = slot[0].ascii etc.: the build path stages through slot 0
    uint32_t *d_bases2 = nullptr, *d_nmask = nullptr;
    uint64_t *d_seq_start = nullptr, *d_pstart = nullptr, *d_koff = nullptr, *d_read_koff = nullptr, *d_scan_tmp = nullptr;
    uint32_t *d_seq_len = nullptr, *d_read_first = nullptr, *d_taxid = nullptr;
    // lookup / accumulation
    uint64_t *d_part = nullptr, *d_rec = nullptr;
    OverflowBuf ovf{};                  // taxa sets that do not fit a match record
    OpsWorkspace ops{};                 // op streams (alias the sort slab) + per-read tables
    uint32_t *d_max_inst = nullptr;
    unsigned long long *d_cnt_fx = nullptr, *d_cnt_un = nullptr, *d_tot_un = nullptr;
    double *d_tot_all = nullptr;
    uint64_t n_kmers_total = 0;
    // index (HBM resident, structure of arrays)
    uint64_t idx_n = 0;
    uint64_t *idx_lo = nullptr, *idx_hi = nullptr;
    uint16_t *idx_tax = nullptr;
    uint64_t *idx_gmask = nullptr;      // taxa mask of every entry's k_low group (content files of <= 64 taxa; lookup.cuh IndexDev)
    uint32_t n_taxa = 0;
    bool dense_scores = false;
    uint32_t batch_epoch = 0;
    // build
    int build_buf = 0;
    uint64_t build_fill = 0;
    bool build_done = false;
    uint64_t n_trie = 0;
    uint32_t *d_trie_words = nullptr;
    // build larger than one brick: full bricks (sorted, distinct) wait in host memory and are merged by prefix range when the build
    // finishes (Build::IntToExtPart / mergeTemporaries, Build.hpp:305-477); the merged index then lives in host memory too
    struct HostBrick { std::vector<uint64_t> lo, hi; std::vector<uint32_t> tax; };
    std::vector<HostBrick> spills;
    bool host_result = false;
    uint64_t host_n = 0;
    std::vector<uint32_t> host_words, host_trie;      // file records (3 or 5 words each) / prefix-table records (3 words each)
    ExactWorkspace xws{};
    bool xws_ready = false;
    // pipelined identify (kasa_identify_submit / _wait): two slots of input staging, score buffers and events;
    // host->device copies run on cs_in, the scores' device->host copy on cs_out, kernels on `stream`
    struct Slot {
        char *ascii = nullptr;
        uint64_t *seq_start = nullptr;
        uint32_t *seq_len = nullptr, *read_first = nullptr;
        float *scores = nullptr;
        uint32_t *h_err = nullptr;                      // pinned copy of the device error flags at the end of the batch
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, sev[KASA_N_STAGES + 1] = {};
        cudaEvent_t h2d = nullptr, in_free = nullptr, done = nullptr, out = nullptr;
        uint32_t *hit_taxon = nullptr, *hit_count = nullptr;      // sparse results (max_batch_hits > 0)
        float *hit_score = nullptr;
        uint64_t *hit_begin = nullptr;
        unsigned long long *hit_cursor = nullptr;
        unsigned long long *h_hits = nullptr;           // pinned copy of hit_cursor
        uint32_t n_reads = 0;
        bool in_free_set = false, out_set = false, pending = false;
        kasa_batch_result *res = nullptr;
        uint64_t n = 0;
        int n_passes = 0;
    } slot[kSlots];
    cudaStream_t cs_in = nullptr, cs_out = nullptr;
    uint64_t submitted = 0, completed = 0;
    double stage_ms[KASA_N_STAGES] = {};
    uint64_t stage_batches = 0, stage_passes = 0, stage_kmers = 0;
    // prefix-sharded index (kasa_shard_*): this context is shard `rank` of `n`; peer[o][b] = shard o's exchange buffers
    struct Shard {
        bool on = false, connected = false;
        int rank = 0, n = 1;
        uint32_t *d_prefix_begin = nullptr;
        uint64_t *d_begin = nullptr;
        uint64_t peer[KASA_MAX_SHARDS][KASA_SHARD_NBUF] = {};
        std::vector<void *> opened;                       // cudaIpcOpenMemHandle mappings
        std::vector<kasa_shard_mem> connected_to;         // what `opened` / `peer` were made from
        uint64_t *d_stage = nullptr;                      // records this shard computed for its peers, in run order
        cudaEvent_t ev_score0 = nullptr;
        bool pull_pending = false;                        // kasa_build_shard_pull filled the free brick half with pulled_total records
        uint64_t pulled_total = 0;
        int phase = 0;                                    // 0 idle, 1 sorted, 2 looked up
        uint32_t n_reads = 0, max_inst = 0;
    } sh;
};

namespace {

thread_local std::string g_create_error;

// leading k-mer bits the identify sort orders (4 radix passes); see launch_partition for why that is enough
constexpr int kIdentifySortBits = 32;

template <typename T>
T *dev_alloc(kasa_ctx *c, size_t n)
{
    void *p = nullptr;
    size_t bytes = std::max<size_t>(n * sizeof(T), 256);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw Error{KASA_ERR_NOMEM, "cudaMalloc of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e)};
    }
    c->allocs.push_back(p);
    return static_cast<T *>(p);
}

void dev_free(kasa_ctx *c, void *p)
{
    if (!p) return;
    auto it = std::find(c->allocs.begin(), c->allocs.end(), p);
    if (it != c->allocs.end()) c->allocs.erase(it);
    cudaFree(p);
}

// standard genetic code in kASA's letter alphabet (kASA.hpp:621-667): stop TAA/TAG '[', TGA ']',
// codons holding 'Z' -> '_', otherwise codons holding 'X' -> '^'.
void builtin_codon_table(uint8_t t[512])
{
    static const char *tcag = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const int code_of[4] = {2, 1, 0, 3};   // T C A G -> A0 C1 T2 G3
    std::memset(t, ' ', 512);
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b)
            for (int c = 0; c < 6; ++c) {
                int idx = a << 6 | b << 3 | c;
                if (a == 5 || b == 5 || c == 5) t[idx] = '_';
                else if (a == 4 || b == 4 || c == 4) t[idx] = '^';
            }
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) {
                char aa = tcag[a * 16 + b * 4 + c];
                if (aa == '*') aa = (a == 0 && b == 3 && c == 2) ? ']' : '[';
                t[code_of[a] << 6 | code_of[b] << 3 | code_of[c]] = (uint8_t)aa;
            }
}

LevelParams level_params(const kasa_ctx *c)
{
    return LevelParams{c->K, c->p.k_high, c->p.k_low, c->num_k, (int)c->n_taxa};
}

void ensure_workspace(kasa_ctx *c)
{
    if (c->ws_ready) return;
    const uint64_t cap = c->p.max_batch_kmers;
    const uint32_t ns = c->p.max_batch_seqs;
    const uint64_t nb = c->p.max_batch_bases;
    {   // sort ping-pong as one slab: [w0[0] | w0[1] | w1[0] | w1[1] | v[0] | v[1]]
        const size_t a8 = ((size_t)cap * 8 + 255) & ~(size_t)255, a4 = ((size_t)cap * 4 + 255) & ~(size_t)255;
        c->sort_slab_bytes = 2 * a8 + (c->wide ? 2 * a8 : 0) + 2 * a4;
        c->sort_slab = dev_alloc<char>(c, c->sort_slab_bytes);
        char *q = c->sort_slab;
        for (int i = 0; i < 2; ++i) { c->sb.w0[i] = reinterpret_cast<uint64_t *>(q); q += a8; }
        for (int i = 0; i < 2; ++i) { c->sb.w1[i] = c->wide ? reinterpret_cast<uint64_t *>(q) : nullptr; if (c->wide) q += a8; }
        for (int i = 0; i < 2; ++i) { c->sb.v[i] = reinterpret_cast<uint32_t *>(q); q += a4; }
        // the op streams of the scoring stage (score.cuh) live in the slab once the lookup has consumed the sorted records:
        // 6 ops per k-mer of capacity for 64-bit k-mers, 10 for 128-bit; the (rarely used) level sizes have their own array
        c->ops.ops = reinterpret_cast<uint32_t *>(c->sort_slab);
        c->ops.ops_cap = c->sort_slab_bytes / 4;
        c->ops.ext_cap = std::max<uint64_t>(cap, 1u << 16);
        c->ops.ext = dev_alloc<uint16_t>(c, c->ops.ext_cap);
    }
    c->sort_ws = dev_alloc<char>(c, sort_workspace_bytes(cap));
    c->d_err = dev_alloc<uint32_t>(c, 4);
    KASA_CUDA_TRY(cudaMemset(c->d_err, 0, 16));
    for (auto &sl : c->slot) {
        sl.ascii = dev_alloc<char>(c, nb + 64);
        sl.seq_start = dev_alloc<uint64_t>(c, (size_t)ns + 1);
        sl.seq_len = dev_alloc<uint32_t>(c, (size_t)ns + 1);
        sl.read_first = dev_alloc<uint32_t>(c, (size_t)ns + 1);
    }
    c->d_ascii = c->slot[0].ascii; c->d_seq_start = c->slot[0].seq_start; c->d_seq_len = c->slot[0].seq_len;
    c->d_read_first = c->slot[0].read_first;
    const uint64_t packed_bases = nb + 32ull * ns + 64;
    c->d_bases2 = dev_alloc<uint32_t>(c, packed_bases / 16 + 4);
    c->d_nmask = dev_alloc<uint32_t>(c, packed_bases / 32 + 4);
    c->d_pstart = dev_alloc<uint64_t>(c, (size_t)ns + 1);
    c->d_koff = dev_alloc<uint64_t>(c, (size_t)ns + 1);
    c->d_read_koff = dev_alloc<uint64_t>(c, (size_t)ns + 1);
    c->d_scan_tmp = dev_alloc<uint64_t>(c, scan_tmp_elems(std::max<uint64_t>(cap + 1, ns + 1)) + 1);
    c->d_taxid = dev_alloc<uint32_t>(c, (size_t)ns + 1);
    c->d_part = dev_alloc<uint64_t>(c, 2 * (size_t)div_up<uint64_t>(cap, 32) + 2);
    c->d_rec = dev_alloc<uint64_t>(c, (size_t)(cap + 2) * kRecWords);
    c->ovf.cursor = dev_alloc<unsigned long long>(c, 4);     // the list itself is sized with the taxonomy (set_taxa)
    c->ops.cursors = c->ovf.cursor + 1;
    KASA_CUDA_TRY(cudaMemset(c->ovf.cursor, 0, 32));
    c->ops.ops_start = dev_alloc<uint64_t>(c, (size_t)ns + 1);
    c->ops.ops_count = dev_alloc<uint32_t>(c, (size_t)ns + 1);
    c->ops.ext_start = dev_alloc<uint64_t>(c, (size_t)ns + 1);
    c->d_max_inst = dev_alloc<uint32_t>(c, 1);
    c->ws_ready = true;
}

[[noreturn]] void throw_device_flags(uint32_t f)
{
    if (f == 1) throw Error{KASA_ERR_INTERNAL, "radix sort look-back watchdog tripped"};
    if (f == 2) throw Error{KASA_ERR_DATA, "a taxid of the index is missing from the content table"};
    if (f == 3) throw Error{KASA_ERR_DATA, "index records are not sorted by (kmer, taxid)"};
    if (f & 8) throw Error{KASA_ERR_NOMEM, "the overflow list of taxa sets is full (max_batch_kmers / 2 entries for <= 64 taxa, 4 x max_batch_kmers beyond): use smaller batches relative to max_batch_kmers"};
    if (f & 32) throw Error{KASA_ERR_NOMEM, "the op-stream workspace of the scoring stage is full: use smaller batches relative to max_batch_kmers"};
    if (f & 64) throw Error{KASA_ERR_ARG, "KASA_SCORE_EXACT: a read has too many (k-mer, level) pairs for the shared-memory replay; use KASA_SCORE_ORDERED for long reads"};
    if (f & 128) throw Error{KASA_ERR_NOMEM, "more (read, taxon) hits than hit_capacity / max_batch_hits"};
    if (f & 256) throw Error{KASA_ERR_NOMEM, "a read hits more than 768 distinct taxa (limit of the sparse scoring cells)"};
    if (f & 512) throw Error{KASA_ERR_INTERNAL, "the dense scoring kernel met a match record that is not in two-mask format"};
    throw Error{KASA_ERR_INTERNAL, "device error flag " + std::to_string(f)};
}

void check_device_flags(kasa_ctx *c)
{
    uint32_t h[4];
    KASA_CUDA_TRY(cudaMemcpyAsync(h, c->d_err, 16, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (h[0] | h[1] | h[2] | h[3]) {
        KASA_CUDA_TRY(cudaMemsetAsync(c->d_err, 0, 16, c->stream));
        uint32_t f = h[0] | h[1] | h[2] | h[3];
        throw_device_flags(f);
    }
}

uint64_t read_u64(kasa_ctx *c, const uint64_t *dptr)
{
    uint64_t v = 0;
    KASA_CUDA_TRY(cudaMemcpyAsync(&v, dptr, 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return v;
}

// taxid -> dense index table (sorted by taxid) on device
struct TaxTable {
    uint32_t *d_taxid = nullptr;
    uint16_t *d_idx = nullptr;
    uint32_t n = 0;
};

TaxTable upload_tax_table(kasa_ctx *c, const uint32_t *taxid_of_idx, uint32_t n_taxa)
{
    std::vector<std::pair<uint32_t, uint16_t>> v;
    for (uint32_t i = 1; i < n_taxa; ++i) v.emplace_back(taxid_of_idx[i], (uint16_t)i);   // idx 0 = "non_unique"
    std::sort(v.begin(), v.end());
    std::vector<uint32_t> t(v.size());
    std::vector<uint16_t> x(v.size());
    for (size_t i = 0; i < v.size(); ++i) { t[i] = v[i].first; x[i] = v[i].second; }
    TaxTable tt;
    tt.n = (uint32_t)v.size();
    tt.d_taxid = dev_alloc<uint32_t>(c, tt.n + 1);
    tt.d_idx = dev_alloc<uint16_t>(c, tt.n + 1);
    KASA_CUDA_TRY(cudaMemcpyAsync(tt.d_taxid, t.data(), tt.n * 4, cudaMemcpyHostToDevice, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(tt.d_idx, x.data(), tt.n * 2, cudaMemcpyHostToDevice, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return tt;
}

void set_taxa(kasa_ctx *c, uint32_t n_taxa)
{
    KASA_REQUIRE(n_taxa >= 1 && n_taxa <= 65535, KASA_ERR_ARG, "n_taxa must be in [1, 65535]");
    c->n_taxa = n_taxa;
    const size_t cells = (size_t)c->num_k * n_taxa;
    c->d_cnt_fx = dev_alloc<unsigned long long>(c, cells);
    c->d_cnt_un = dev_alloc<unsigned long long>(c, cells);
    c->d_tot_un = dev_alloc<unsigned long long>(c, cells);
    c->d_tot_all = dev_alloc<double>(c, cells);
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_cnt_fx, 0, cells * 8, c->stream));
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_cnt_un, 0, cells * 8, c->stream));
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_tot_un, 0, cells * 8, c->stream));
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_tot_all, 0, cells * 8, c->stream));
    // overflow list of taxa sets: content files of <= 64 taxa use mask records and need it only for k-mers whose upper levels name more
    // than six taxa; larger taxonomies list every set of more than nine taxa there
    c->ovf.cap = std::max<uint64_t>(n_taxa <= (uint32_t)kMaskTaxa ? c->p.max_batch_kmers / 2 : c->p.max_batch_kmers * 4, 1u << 22);
    c->ovf.data = dev_alloc<uint32_t>(c, c->ovf.cap);
    // dense score matrix [reads][taxa] (the reference's Non_contiguousArray) per pipeline slot, while that is a sane amount of HBM;
    // larger taxonomies are scored through the sparse hit lists (kasa_params.max_batch_hits)
    const size_t dense_bytes = (size_t)c->p.max_batch_seqs * n_taxa * 4;
    c->dense_scores = c->p.score_mode != KASA_SCORE_NONE && dense_bytes * kSlots <= ((size_t)32 << 30);
    for (auto &sl : c->slot) {
        sl.scores = c->dense_scores ? dev_alloc<float>(c, (size_t)c->p.max_batch_seqs * n_taxa + 1) : nullptr;
        if (c->p.max_batch_hits && c->p.score_mode != KASA_SCORE_NONE && !sl.hit_taxon) {
            sl.hit_taxon = dev_alloc<uint32_t>(c, c->p.max_batch_hits);
            sl.hit_score = dev_alloc<float>(c, c->p.max_batch_hits);
            sl.hit_begin = dev_alloc<uint64_t>(c, (size_t)c->p.max_batch_seqs + 1);
            sl.hit_count = dev_alloc<uint32_t>(c, (size_t)c->p.max_batch_seqs + 1);
            sl.hit_cursor = dev_alloc<unsigned long long>(c, 1);
        }
    }
    c->n_kmers_total = 0;
}

void drop_index(kasa_ctx *c)
{
    dev_free(c, c->idx_lo); dev_free(c, c->idx_hi); dev_free(c, c->idx_tax); dev_free(c, c->idx_gmask);
    c->idx_gmask = nullptr;
    dev_free(c, c->ovf.data); c->ovf.data = nullptr;
    for (auto &sl : c->slot) {
        dev_free(c, sl.hit_taxon); dev_free(c, sl.hit_score); dev_free(c, sl.hit_begin); dev_free(c, sl.hit_count); dev_free(c, sl.hit_cursor);
        sl.hit_taxon = sl.hit_count = nullptr; sl.hit_score = nullptr; sl.hit_begin = nullptr; sl.hit_cursor = nullptr;
    }
    dev_free(c, c->d_cnt_fx); dev_free(c, c->d_cnt_un); dev_free(c, c->d_tot_un); dev_free(c, c->d_tot_all);
    for (auto &sl : c->slot) { dev_free(c, sl.scores); sl.scores = nullptr; }
    c->idx_lo = c->idx_hi = nullptr; c->idx_tax = nullptr;
    c->d_cnt_fx = c->d_cnt_un = c->d_tot_un = nullptr; c->d_tot_all = nullptr;
    c->idx_n = 0; c->n_taxa = 0;
}

// stage a batch of sequences: after this, d_seq_start/d_seq_len (and d_ascii for host ASCII input) are valid.
// Returns device pointers for the sequence data.
struct Staged {
    const char *ascii = nullptr;
    const uint32_t *bases2 = nullptr, *nmask = nullptr;
    const uint64_t *seq_start = nullptr;
    const uint32_t *seq_len = nullptr;
};

template <typename T>
const T *stage_array(kasa_ctx *c, const T *src, size_t n, T *dst, int location)
{
    if (location == KASA_MEM_DEVICE) return src;
    KASA_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    return dst;
}

uint64_t host_total_bases(const uint64_t *seq_start, const uint32_t *seq_len, uint32_t n)
{
    uint64_t m = 0;
    for (uint32_t i = 0; i < n; ++i) m = std::max<uint64_t>(m, seq_start[i] + seq_len[i]);
    return m;
}

// sum of kMerCount over host-resident sequences (Read.hpp:36-57 with the marker of Read.hpp:1068-1078): what
// kmer_counts_kernel + scan produce on the device, so that a host batch needs no device round trip before its launches
uint64_t host_total_kmers(const uint32_t *seq_len, uint32_t n, int K, int k_low, int strands, int one, const uint32_t *read_first_seq,
                           uint32_t n_reads, uint32_t *max_per_read)
{
    uint64_t total = 0, best = 0, cur = 0;
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (read_first_seq) { while (r < n_reads && read_first_seq[r + 1] <= i) { ++r; cur = 0; } } else cur = 0;
        const uint64_t len = (uint64_t)seq_len[i] + 3u * (K - k_low);
        const uint64_t k = kmer_count(len, K, one) * strands;
        total += k;
        cur += k;
        best = std::max(best, cur);
    }
    if (max_per_read) *max_per_read = (uint32_t)std::min<uint64_t>(best, 0xffffffffu);
    return total;
}

// sorts the accumulated build records by (kmer, taxid), removes duplicates and sentinels; the result
// is left in buffer set c->build_buf, c->build_fill = number of distinct records.
void build_compact(kasa_ctx *c)
{
    const uint64_t n = c->build_fill;
    if (n == 0) return;
    SortPlan plan{};
    for (int b = 0; b < 32; b += 8) plan.pass[plan.n_passes++] = SortPass{SRC_V, b, 0xffu};
    SortPlan kp = plan_for_kmer_bits(c->wide ? 128 : 64);   // all bits: sentinels (~0) must sort last
    for (int i = 0; i < kp.n_passes; ++i) plan.pass[plan.n_passes++] = kp.pass[i];
    int cur = radix_sort(c->stream, c->sb, c->build_buf, n, plan, c->sort_ws, c->d_err, &c->launches);
    uint64_t *flags = c->d_rec;
    launch_unique_flags(c->stream, c->sb.w0[cur], c->sb.w1[cur], c->sb.v[cur], n, flags, &c->launches);
    exclusive_scan_u64(c->stream, flags, n, c->d_scan_tmp, &c->launches);
    const uint64_t n_unique = read_u64(c, flags + n);
    const int dst = cur ^ 1;
    launch_compact(c->stream, c->sb.w0[cur], c->sb.w1[cur], c->sb.v[cur], n, flags, c->sb.w0[dst], c->sb.w1[dst], c->sb.v[dst],
                   &c->launches);
    KASA_CUDA_TRY(cudaGetLastError());
    check_device_flags(c);
    c->build_buf = dst;
    c->build_fill = n_unique;
}

// Trie.hpp:366-394 for the brick in c->build_buf: run lengths of the first six letters -> c->d_trie_words, c->n_trie.
// Scratch lives in d_rec: [flags/pos: n+1 u64][run starts: n_runs+1 u64][trie words: 3 u32 per run]
void build_trie(kasa_ctx *c)
{
    const uint64_t n = c->build_fill;
    const int b = c->build_buf;
    c->n_trie = 0;
    if (!n) return;
    uint64_t *flags = c->d_rec;
    launch_prefix_flags(c->stream, c->sb.w0[b], c->sb.w1[b], n, c->K, flags, &c->launches);
    exclusive_scan_u64(c->stream, flags, n, c->d_scan_tmp, &c->launches);
    const uint64_t runs = read_u64(c, flags + n);
    uint64_t *starts = flags + n + 1;
    c->d_trie_words = reinterpret_cast<uint32_t *>(starts + runs + 1);
    launch_trie_records(c->stream, c->sb.w0[b], c->sb.w1[b], n, c->K, flags, runs, starts, c->d_trie_words, &c->launches);
    KASA_CUDA_TRY(cudaGetLastError());
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->n_trie = runs;
}

// the compacted brick (sorted, distinct records) leaves for host memory: Build::IntToExtPart's spill (Build.hpp:328-352)
void spill_brick(kasa_ctx *c)
{
    const uint64_t n = c->build_fill;
    if (!n) return;
    const int b = c->build_buf;
    kasa_ctx::HostBrick hb;
    hb.lo.resize(n); hb.tax.resize(n);
    KASA_CUDA_TRY(cudaMemcpyAsync(hb.lo.data(), c->sb.w0[b], n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (c->wide) { hb.hi.resize(n); KASA_CUDA_TRY(cudaMemcpyAsync(hb.hi.data(), c->sb.w1[b], n * 8, cudaMemcpyDeviceToHost, c->stream)); }
    KASA_CUDA_TRY(cudaMemcpyAsync(hb.tax.data(), c->sb.v[b], n * 4, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->spills.push_back(std::move(hb));
    c->build_fill = 0;
    c->build_buf = 0;
}

// first record of brick `hb` whose six-letter prefix is >= p (the bricks are sorted by k-mer)
static uint64_t brick_lower(const kasa_ctx *c, const kasa_ctx::HostBrick &hb, uint64_t p)
{
    const int sh = c->wide ? 5 * c->K - 30 - 64 : 5 * c->K - 30;
    const std::vector<uint64_t> &key = c->wide ? hb.hi : hb.lo;
    return (uint64_t)(std::partition_point(key.begin(), key.end(), [&](uint64_t k) { return (k >> sh) < p; }) - key.begin());
}

// Build::mergeTemporaries (Build.hpp:360-477) without its priority queue: the key space is cut at six-letter prefixes into
// ranges whose records — over all spilled bricks — fit one device brick; every range is copied in, sorted and deduplicated
// by the same kernels as a single brick, run-length encoded, and appended to the host-side result.  The ranges concatenate
// to the file `kASA build` writes (the prefix table is per prefix, so its pieces concatenate too).
void merge_spills(kasa_ctx *c)
{
    const uint64_t cap = c->p.max_batch_kmers;
    const size_t wpr = c->wide ? 5 : 3;
    c->host_words.clear(); c->host_trie.clear(); c->host_n = 0;
    uint64_t total_trie = 0;
    const uint64_t p_end = 1ull << 30;
    std::vector<uint64_t> at(c->spills.size()), to(c->spills.size());
    for (size_t i = 0; i < c->spills.size(); ++i) at[i] = 0;
    uint64_t p = 0;
    while (p < p_end) {
        // largest q > p with at most cap records in [p, q)
        auto count_to = [&](uint64_t q) { uint64_t n = 0; for (size_t i = 0; i < c->spills.size(); ++i) n += brick_lower(c, c->spills[i], q) - at[i]; return n; };
        uint64_t q = p_end;
        if (count_to(q) > cap) {
            uint64_t lo = p + 1, hi = p_end;                 // count_to(lo) may exceed cap: checked below
            while (lo < hi) { const uint64_t mid = lo + (hi - lo + 1) / 2; if (count_to(mid) <= cap) lo = mid; else hi = mid - 1; }
            q = lo;
            KASA_REQUIRE(count_to(q) <= cap, KASA_ERR_NOMEM, "one six-letter prefix holds more records than max_batch_kmers: the bricks cannot be merged");
        }
        uint64_t n = 0;
        for (size_t i = 0; i < c->spills.size(); ++i) {
            const kasa_ctx::HostBrick &hb = c->spills[i];
            to[i] = brick_lower(c, hb, q);
            const uint64_t m = to[i] - at[i];
            if (m) {
                KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.w0[0] + n, hb.lo.data() + at[i], m * 8, cudaMemcpyHostToDevice, c->stream));
                if (c->wide) KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.w1[0] + n, hb.hi.data() + at[i], m * 8, cudaMemcpyHostToDevice, c->stream));
                KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.v[0] + n, hb.tax.data() + at[i], m * 4, cudaMemcpyHostToDevice, c->stream));
            }
            n += m;
            at[i] = to[i];
        }
        p = q;
        if (!n) continue;
        c->build_buf = 0; c->build_fill = n;
        build_compact(c);                                     // sort + dedup of the range (duplicates across bricks go here)
        build_trie(c);
        const uint64_t m = c->build_fill;
        const int b = c->build_buf;
        // file records through the free half of the sort ping-pong, in chunks
        uint32_t *stage = reinterpret_cast<uint32_t *>(c->sb.w0[b ^ 1]);
        const uint64_t chunk = std::max<uint64_t>(1, cap * 8 / (wpr * 4));
        c->host_words.resize((size_t)(c->host_n + m) * wpr);
        for (uint64_t o = 0; o < m; o += chunk) {
            const uint64_t k = std::min(chunk, m - o);
            launch_pack_records(c->stream, c->sb.w0[b] + o, c->wide ? c->sb.w1[b] + o : nullptr, c->sb.v[b] + o, k, stage, &c->launches);
            KASA_CUDA_TRY(cudaMemcpyAsync(c->host_words.data() + (size_t)(c->host_n + o) * wpr, stage, k * wpr * 4, cudaMemcpyDeviceToHost, c->stream));
            KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        c->host_trie.resize((size_t)(total_trie + c->n_trie) * 3);
        if (c->n_trie) {
            KASA_CUDA_TRY(cudaMemcpyAsync(c->host_trie.data() + (size_t)total_trie * 3, c->d_trie_words, c->n_trie * 12, cudaMemcpyDeviceToHost, c->stream));
            KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
        c->host_n += m;
        total_trie += c->n_trie;
    }
    c->spills.clear(); c->spills.shrink_to_fit();
    c->n_trie = total_trie;
    c->build_fill = 0; c->build_buf = 0;
    c->host_result = true;
}

int fail(kasa_ctx *c, const Error &e)
{
    if (c) c->err = e.msg; else g_create_error = e.msg;
    return e.code;
}

#define KASA_REQUIRE_PLAIN(cond) \
    if (!(cond)) return fail(nullptr, Error{KASA_ERR_ARG, "null argument"})
#define KASA_API_BEGIN try {
#define KASA_API_END(ctx)                                                                   \
    }                                                                                       \
    catch (const Error &e) { return fail((ctx), e); }                                       \
    catch (const std::exception &e) { return fail((ctx), Error{KASA_ERR_INTERNAL, e.what()}); } \
    return KASA_OK;

}  // namespace

extern "C" {

int kasa_abi_version(void) { return KASA_B200_ABI_VERSION; }

const char *kasa_last_error(const kasa_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int kasa_ctx_create(const kasa_params *p, kasa_ctx **out)
{
    kasa_ctx *c = nullptr;
    try {
        KASA_REQUIRE(p && out, KASA_ERR_ARG, "null argument");
        KASA_REQUIRE(p->highest_k == 12 || p->highest_k == 25, KASA_ERR_ARG, "highest_k must be 12 or 25");
        KASA_REQUIRE(p->k_low >= 1 && p->k_low <= p->k_high && p->k_high <= p->highest_k, KASA_ERR_ARG,
                     "need 1 <= k_low <= k_high <= highest_k");
        KASA_REQUIRE(p->frames == 1 || p->frames == 3 || p->frames == 6, KASA_ERR_ARG, "frames must be 1 (--one), 3 or 6 (--six)");
        KASA_REQUIRE(p->score_mode >= KASA_SCORE_NONE && p->score_mode <= KASA_SCORE_EXACT, KASA_ERR_ARG, "unknown score_mode");
        KASA_REQUIRE(p->max_batch_kmers > 0 && p->max_batch_kmers < (1ull << 32) - 2 * kSortTile, KASA_ERR_ARG,
                     "max_batch_kmers must be in (0, 2^32)");
        KASA_REQUIRE(p->max_batch_seqs > 0 && p->max_batch_bases > 0, KASA_ERR_ARG, "max_batch_seqs / max_batch_bases must be > 0");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            throw Error{KASA_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                           " (libkasa_b200 has no CPU fallback)"};
        }
        KASA_REQUIRE(p->device >= 0 && p->device < ndev, KASA_ERR_ARG, "device ordinal out of range");
        KASA_CUDA_TRY(cudaSetDevice(p->device));
        c = new kasa_ctx();
        c->p = *p;
        c->p.codon_table = nullptr;
        c->K = p->highest_k;
        c->wide = p->highest_k > 12;
        c->num_k = p->k_high - p->k_low + 1;
        c->sort_bits = kIdentifySortBits;
        if (const char *e = getenv("KASA_TUNE_SORT_BITS")) c->sort_bits = std::min(5 * c->K, std::max(8, atoi(e)));
        if (const char *e = getenv("KASA_TUNE_SCORE_SORTED")) c->score_sorted = atoi(e) != 0;
        if (const char *e = getenv("KASA_TUNE_SCORE_LANES")) c->score_lanes = atoi(e);
        KASA_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
        KASA_CUDA_TRY(cudaStreamCreateWithFlags(&c->cs_in, cudaStreamNonBlocking));
        KASA_CUDA_TRY(cudaStreamCreateWithFlags(&c->cs_out, cudaStreamNonBlocking));
        for (auto &sl : c->slot) {
            KASA_CUDA_TRY(cudaEventCreate(&sl.ev0));
            KASA_CUDA_TRY(cudaEventCreate(&sl.ev1));
            for (auto &e : sl.sev) KASA_CUDA_TRY(cudaEventCreate(&e));
            for (cudaEvent_t *e : {&sl.h2d, &sl.in_free, &sl.done, &sl.out}) KASA_CUDA_TRY(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
            KASA_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_err), 16, cudaHostAllocDefault));
            KASA_CUDA_TRY(cudaHostAlloc(reinterpret_cast<void **>(&sl.h_hits), 8, cudaHostAllocDefault));
        }
        uint8_t table[512];
        if (p->codon_table) std::memcpy(table, p->codon_table, 512); else builtin_codon_table(table);
        c->d_codon = dev_alloc<uint8_t>(c, 512);
        upload_codon_table(table, c->d_codon, c->stream);
        *out = c;
    } catch (const Error &e) {
        if (c) kasa_ctx_destroy(c);
        return fail(nullptr, e);
    }
    return KASA_OK;
}

void kasa_ctx_destroy(kasa_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->p.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void *p : c->sh.opened) cudaIpcCloseMemHandle(p);
    if (c->sh.ev_score0) cudaEventDestroy(c->sh.ev_score0);
    for (void *p : c->allocs) cudaFree(p);
    if (c->cs_in) { cudaStreamSynchronize(c->cs_in); cudaStreamDestroy(c->cs_in); }
    if (c->cs_out) { cudaStreamSynchronize(c->cs_out); cudaStreamDestroy(c->cs_out); }
    for (auto &sl : c->slot) {
        for (cudaEvent_t e : {sl.ev0, sl.ev1, sl.h2d, sl.in_free, sl.done, sl.out}) if (e) cudaEventDestroy(e);
        for (auto e : sl.sev) if (e) cudaEventDestroy(e);
        if (sl.h_err) cudaFreeHost(sl.h_err);
        if (sl.h_hits) cudaFreeHost(sl.h_hits);
    }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int kasa_ctx_set_stream(kasa_ctx *c, void *cuda_stream)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = static_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    KASA_API_END(c)
}

int kasa_ctx_synchronize(kasa_ctx *c)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    KASA_API_END(c)
}

uint64_t kasa_ctx_launch_count(const kasa_ctx *c) { return c ? c->launches : 0; }

int kasa_ctx_stage_times(kasa_ctx *c, double *ms, uint64_t *batches, uint64_t *sort_passes, uint64_t *kmers, int reset)
{
    if (!c) return KASA_ERR_ARG;
    if (ms) for (int i = 0; i < KASA_N_STAGES; ++i) ms[i] = c->stage_ms[i];
    if (batches) *batches = c->stage_batches;
    if (sort_passes) *sort_passes = c->stage_passes;
    if (kmers) *kmers = c->stage_kmers;
    if (reset) {
        for (auto &x : c->stage_ms) x = 0;
        c->stage_batches = c->stage_passes = c->stage_kmers = 0;
    }
    return KASA_OK;
}
uint64_t kasa_index_size(const kasa_ctx *c) { return c ? c->idx_n : 0; }

// Last step of every index load: the taxon-first flags (any number of taxa) and, for content files of <= 64 taxa, the taxa mask of
// every k_low-letter group (markTaxIDs of level k_low, Compare.hpp:396-441, done once instead of once per query).
static void finalize_index(kasa_ctx *c, uint64_t n)
{
    launch_index_flags(c->stream, c->idx_lo, c->idx_hi, c->idx_tax, n, c->K, c->p.k_low, &c->launches);
    if (c->n_taxa > (uint32_t)kMaskTaxa || n == 0) return;
    c->idx_gmask = dev_alloc<uint64_t>(c, n + 16);
    uint64_t *tmp = dev_alloc<uint64_t>(c, scan_tmp_elems(n) + 1);
    launch_group_heads(c->stream, c->idx_lo, c->idx_hi, n, c->K, c->p.k_low, c->idx_gmask, &c->launches);
    exclusive_scan_u64(c->stream, c->idx_gmask, n, tmp, &c->launches);
    const uint64_t groups = read_u64(c, c->idx_gmask + n);
    uint64_t *gm = dev_alloc<uint64_t>(c, groups + 1);
    KASA_CUDA_TRY(cudaMemsetAsync(gm, 0, (groups + 1) * 8, c->stream));
    launch_group_masks(c->stream, c->idx_lo, c->idx_hi, c->idx_tax, n, c->K, c->p.k_low, gm, c->idx_gmask, &c->launches);
    KASA_CUDA_TRY(cudaMemsetAsync(c->idx_gmask + n, 0, 16 * 8, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    dev_free(c, gm); dev_free(c, tmp);
}

int kasa_index_load(kasa_ctx *c, const void *entries, uint64_t n, int bits, const uint32_t *taxid_of_idx, uint32_t n_taxa)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && entries && taxid_of_idx, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(bits == (c->wide ? 128 : 64), KASA_ERR_ARG, "index width does not match highest_k of the context");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    drop_index(c);
    set_taxa(c, n_taxa);
    const size_t words_per = c->wide ? 5 : 3;
    uint32_t *d_words = dev_alloc<uint32_t>(c, n * words_per + 1);
    uint32_t *d_tax32 = dev_alloc<uint32_t>(c, n + 1);
    c->idx_lo = dev_alloc<uint64_t>(c, n + 16);
    c->idx_hi = c->wide ? dev_alloc<uint64_t>(c, n + 16) : nullptr;
    c->idx_tax = dev_alloc<uint16_t>(c, n + 16);
    KASA_CUDA_TRY(cudaMemcpyAsync(d_words, entries, n * words_per * 4, cudaMemcpyHostToDevice, c->stream));
    launch_unpack_records(c->stream, d_words, n, c->wide, c->idx_lo, c->idx_hi, d_tax32, &c->launches);
    launch_check_sorted(c->stream, c->idx_lo, c->idx_hi, d_tax32, n, c->d_err, &c->launches);
    TaxTable tt = upload_tax_table(c, taxid_of_idx, n_taxa);
    launch_map_taxids(c->stream, d_tax32, n, tt.d_taxid, tt.d_idx, tt.n, c->idx_tax, c->d_err, &c->launches);
    finalize_index(c, n);
    KASA_CUDA_TRY(cudaGetLastError());
    check_device_flags(c);
    dev_free(c, d_words); dev_free(c, d_tax32); dev_free(c, tt.d_taxid); dev_free(c, tt.d_idx);
    c->idx_n = n;
    KASA_API_END(c)
}

int kasa_index_load_half(kasa_ctx *c, const void *entries, uint64_t n, const void *trie, uint64_t n_trie, uint32_t n_taxa)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && entries && trie && n && n_trie, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(!c->wide, KASA_ERR_ARG, "halved indices are 64-bit indices (highest_k 12)");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    drop_index(c);
    set_taxa(c, n_taxa);
    uint16_t *d_rec = dev_alloc<uint16_t>(c, n * 3 + 2);
    uint32_t *d_trie = dev_alloc<uint32_t>(c, n_trie * 3 + 1);
    uint64_t *d_starts = dev_alloc<uint64_t>(c, n_trie + 2);
    uint32_t *d_prefix = dev_alloc<uint32_t>(c, n_trie + 1);
    uint64_t *d_tmp = dev_alloc<uint64_t>(c, scan_tmp_elems(n_trie + 1) + 1);
    c->idx_lo = dev_alloc<uint64_t>(c, n + 16);
    c->idx_hi = nullptr;
    c->idx_tax = dev_alloc<uint16_t>(c, n + 16);
    KASA_CUDA_TRY(cudaMemcpyAsync(d_rec, entries, n * 6, cudaMemcpyHostToDevice, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(d_trie, trie, n_trie * 12, cudaMemcpyHostToDevice, c->stream));
    launch_half_trie_split(c->stream, d_trie, n_trie, d_starts, d_prefix, &c->launches);
    exclusive_scan_u64(c->stream, d_starts, n_trie, d_tmp, &c->launches);
    launch_half_expand(c->stream, d_rec, n, d_starts, d_prefix, n_trie, n_taxa, c->idx_lo, c->idx_tax, c->d_err, &c->launches);
    launch_check_sorted(c->stream, c->idx_lo, nullptr, nullptr, n, c->d_err, &c->launches);
    finalize_index(c, n);
    KASA_CUDA_TRY(cudaGetLastError());
    check_device_flags(c);
    dev_free(c, d_rec); dev_free(c, d_trie); dev_free(c, d_starts); dev_free(c, d_prefix); dev_free(c, d_tmp);
    c->idx_n = n;
    KASA_API_END(c)
}

int kasa_index_from_build(kasa_ctx *c, const uint32_t *taxid_of_idx, uint32_t n_taxa)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && taxid_of_idx, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->build_done, KASA_ERR_STATE, "kasa_index_from_build needs a finished build (kasa_build_finish)");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    if (c->host_result) {
        // a build that spilled bricks left its merged index in host memory: loaded like an index file
        const int rc = kasa_index_load(c, c->host_words.data(), c->host_n, c->wide ? 128 : 64, taxid_of_idx, n_taxa);
        if (rc != KASA_OK) throw Error{rc, c->err};
        return KASA_OK;
    }
    drop_index(c);
    set_taxa(c, n_taxa);
    const uint64_t n = c->build_fill;
    const int b = c->build_buf;
    c->idx_lo = dev_alloc<uint64_t>(c, n + 16);
    c->idx_hi = c->wide ? dev_alloc<uint64_t>(c, n + 16) : nullptr;
    c->idx_tax = dev_alloc<uint16_t>(c, n + 16);
    KASA_CUDA_TRY(cudaMemcpyAsync(c->idx_lo, c->sb.w0[b], n * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (c->wide) KASA_CUDA_TRY(cudaMemcpyAsync(c->idx_hi, c->sb.w1[b], n * 8, cudaMemcpyDeviceToDevice, c->stream));
    TaxTable tt = upload_tax_table(c, taxid_of_idx, n_taxa);
    launch_map_taxids(c->stream, c->sb.v[b], n, tt.d_taxid, tt.d_idx, tt.n, c->idx_tax, c->d_err, &c->launches);
    finalize_index(c, n);
    KASA_CUDA_TRY(cudaGetLastError());
    check_device_flags(c);
    dev_free(c, tt.d_taxid); dev_free(c, tt.d_idx);
    c->idx_n = n;
    KASA_API_END(c)
}

// Host -> device copy of one array of slot `sl` on the copy-in stream (device-resident input is used in place)
extern "C++" template <typename T>
const T *stage_in(kasa_ctx *c, const T *src, size_t n, T *dst, int location)
{
    if (location == KASA_MEM_DEVICE) return src;
    KASA_CUDA_TRY(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, c->cs_in));
    return dst;
}

// front half of identify for slot `sl`: stage (copy-in stream) + pack + count + extract (compute stream);
// returns the number of k-mers; *read_first = device copy of in->read_first_seq (or nullptr)
static uint64_t extract_batch(kasa_ctx *c, const kasa_read_batch *in, kasa_ctx::Slot &sl, const uint32_t **read_first, uint32_t *max_inst = nullptr)
{
    KASA_REQUIRE(c->build_fill == 0 || c->build_done, KASA_ERR_STATE,
                 "a build is open on this context (kasa_build_push without kasa_build_finish): its brick shares the record arrays");
    if (c->build_done) { c->build_done = false; c->build_fill = 0; c->n_trie = 0; }   // the finished brick is overwritten from here on
    KASA_REQUIRE(in->n_seqs <= c->p.max_batch_seqs, KASA_ERR_NOMEM, "batch has more sequences than max_batch_seqs");
    KASA_REQUIRE(in->n_reads <= in->n_seqs, KASA_ERR_ARG, "n_reads > n_seqs");
    KASA_REQUIRE(in->format == KASA_SEQ_ASCII || in->format == KASA_SEQ_PACKED2, KASA_ERR_ARG, "unknown sequence format");
    KASA_REQUIRE(in->read_first_seq || in->n_reads == in->n_seqs, KASA_ERR_ARG, "read_first_seq is required when n_reads != n_seqs");
    ensure_workspace(c);
    *read_first = nullptr;
    const uint32_t ns = in->n_seqs;
    if (ns == 0) return 0;
    // the slot's staging arrays were last read by the extraction of the batch two submissions ago
    if (sl.in_free_set) KASA_CUDA_TRY(cudaStreamWaitEvent(c->cs_in, sl.in_free, 0));
    const uint64_t *seq_start = stage_in(c, in->seq_start, ns, sl.seq_start, in->location);
    const uint32_t *seq_len = stage_in(c, in->seq_len, ns, sl.seq_len, in->location);
    if (in->read_first_seq) *read_first = stage_in(c, in->read_first_seq, (size_t)in->n_reads + 1, sl.read_first, in->location);
    const char *ascii = static_cast<const char *>(in->bases);
    if (in->format == KASA_SEQ_ASCII && in->location == KASA_MEM_HOST) {
        uint64_t total = host_total_bases(in->seq_start, in->seq_len, ns);
        KASA_REQUIRE(total <= c->p.max_batch_bases, KASA_ERR_NOMEM, "batch has more bases than max_batch_bases");
        KASA_CUDA_TRY(cudaMemcpyAsync(sl.ascii, in->bases, total, cudaMemcpyHostToDevice, c->cs_in));
        ascii = sl.ascii;
    }
    KASA_CUDA_TRY(cudaEventRecord(sl.h2d, c->cs_in));
    KASA_CUDA_TRY(cudaStreamWaitEvent(c->stream, sl.h2d, 0));
    PackedSeqs ps{};
    ps.n_seqs = ns;
    ps.seq_len = seq_len;
    if (in->format == KASA_SEQ_PACKED2) {
        KASA_REQUIRE(in->nmask, KASA_ERR_ARG, "PACKED2 input needs the non-ACGT plane");
        if (in->location == KASA_MEM_HOST) {              // packed planes are not double-buffered: copied in stream order
            uint64_t total = host_total_bases(in->seq_start, in->seq_len, ns);
            KASA_REQUIRE(total <= c->p.max_batch_bases + 32ull * ns, KASA_ERR_NOMEM, "batch has more bases than max_batch_bases");
            KASA_CUDA_TRY(cudaMemcpyAsync(c->d_bases2, in->bases, div_up<uint64_t>(total, 16) * 4, cudaMemcpyHostToDevice, c->stream));
            KASA_CUDA_TRY(cudaMemcpyAsync(c->d_nmask, in->nmask, div_up<uint64_t>(total, 32) * 4, cudaMemcpyHostToDevice, c->stream));
            ps.bases2 = c->d_bases2; ps.nmask = c->d_nmask;
        } else {
            ps.bases2 = static_cast<const uint32_t *>(in->bases); ps.nmask = in->nmask;
        }
        ps.seq_start = seq_start;
    }
    KmerParams kp{c->K, c->p.k_low, c->p.frames == 6 ? 2 : 1, c->p.frames == 1 ? 1 : 0};
    launch_kmer_counts(c->stream, seq_len, ns, kp, c->d_koff, &c->launches);
    exclusive_scan_u64(c->stream, c->d_koff, ns, c->d_scan_tmp, &c->launches);
    // host batches: the total is known without asking the device, so submit returns without a stream synchronisation
    uint64_t n = 0;
    uint32_t mi = 0;
    if (in->location == KASA_MEM_HOST) {
        n = host_total_kmers(in->seq_len, ns, kp.K, kp.k_low, kp.strands, kp.one, in->read_first_seq, in->n_reads, &mi);
    } else {
        launch_read_offsets(c->stream, c->d_koff, *read_first, in->n_reads, c->d_read_koff, &c->launches);
        launch_max_instances(c->stream, c->d_read_koff, in->n_reads, c->d_max_inst, &c->launches);
        KASA_CUDA_TRY(cudaMemcpyAsync(&mi, c->d_max_inst, 4, cudaMemcpyDeviceToHost, c->stream));
        n = read_u64(c, c->d_koff + ns);
    }
    if (max_inst) *max_inst = mi;
    KASA_REQUIRE(n <= c->p.max_batch_kmers, KASA_ERR_NOMEM, "batch yields more k-mers than max_batch_kmers");
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[1], c->stream));
    if (in->format == KASA_SEQ_ASCII)
        launch_extract_identify_ascii(c->stream, ascii, seq_start, seq_len, ns, kp, c->d_codon, c->d_koff, c->sb.w0[0], c->sb.w1[0], c->sb.v[0], *read_first,
                                      in->n_reads, mi, &c->launches);
    else
        launch_extract_identify(c->stream, ps, kp, c->d_codon, c->d_koff, c->sb.w0[0], c->sb.w1[0], c->sb.v[0], *read_first, in->n_reads, mi, &c->launches);
    KASA_CUDA_TRY(cudaGetLastError());
    return n;
}

// Presorted batch (every read was emitted in ascending k-mer order, extract.cuh) against a small content file: scores and counts come
// straight from the match records (score.cuh launch_score_sorted); decided before the lookup, which then writes the record format
// that kernel reads.
static bool scores_straight_from_records(const kasa_ctx *c, uint32_t max_inst, bool want_hits)
{
    // (prefix-sharded: the owners of the index ranges write the records, so the choice may only depend on what every rank knows when the
    // batch is sorted — sparse results, chosen per batch, then keep the op-stream path)
    return c->p.score_mode != KASA_SCORE_EXACT && !want_hits && c->n_taxa <= (uint32_t)kMaskTaxa && max_inst <= (uint32_t)presort_limit(c->wide) && c->score_sorted &&
           (!c->sh.on || c->p.max_batch_hits == 0);
}

// Scoring stage (score.cuh) after the match records of the batch are in c->d_rec: op streams in the reference's summation
// order, then scores and counts.  `cur` = sort buffer holding the sorted queries (KASA_SCORE_EXACT reads them).
static void score_stage(kasa_ctx *c, kasa_ctx::Slot &sl, uint64_t n, uint32_t n_reads, uint32_t max_inst, int cur, bool want_dense, bool want_hits)
{
    LevelParams lp = level_params(c);
    const bool exact = c->p.score_mode == KASA_SCORE_EXACT;
    const uint32_t *trig = nullptr;
    if (exact) {
        if (!c->xws_ready) {
            const uint64_t cap = c->p.max_batch_kmers;
            c->xws.flags = dev_alloc<uint32_t>(c, cap + 1);
            c->xws.first = dev_alloc<uint32_t>(c, exact_tiles(cap) * c->num_k + 1);
            c->xws.next = dev_alloc<uint32_t>(c, exact_tiles(cap) * c->num_k + 1);
            c->xws.trig = dev_alloc<uint32_t>(c, cap * c->num_k + 1);
            c->xws_ready = true;
        }
        IndexDev idx{c->idx_lo, c->idx_hi, c->idx_tax, c->idx_n, c->idx_gmask};
        launch_exact_triggers(c->stream, c->sb.w0[cur], c->sb.w1[cur], c->sb.v[cur], n, idx, lp, c->d_rec, c->xws, &c->launches);
        trig = c->xws.trig;
    }
    OverflowPeers peers{};
    for (int o = 0; o < 16; ++o) peers.base[o] = c->ovf.data;
    if (c->sh.on) for (int o = 0; o < c->sh.n; ++o) peers.base[o] = o == c->sh.rank ? c->ovf.data : reinterpret_cast<const uint32_t *>(c->sh.peer[o][8]);
    HitsOut hits{};
    if (want_hits) {
        hits = HitsOut{sl.hit_taxon, sl.hit_score, sl.hit_begin, sl.hit_count, sl.hit_cursor, c->p.max_batch_hits};
        KASA_CUDA_TRY(cudaMemsetAsync(sl.hit_cursor, 0, 8, c->stream));
    }
    // Presorted batch (every read was emitted in ascending k-mer order, extract.cuh) against a small content file: the records lie
    // in summation order, scores and counts come straight from them.
    if (scores_straight_from_records(c, max_inst, want_hits)) {
        launch_score_sorted(c->stream, c->d_rec, c->d_read_koff, n_reads, lp, peers, want_dense ? sl.scores : nullptr, c->d_cnt_fx, c->d_cnt_un,
                            c->score_lanes, c->d_err, &c->launches);
        launch_fold_counts(c->stream, c->d_cnt_fx, c->d_cnt_un, lp, c->d_tot_all, c->d_tot_un, &c->launches);
        return;
    }
    // The op streams live in the sort slab (6 words per k-mer of capacity for 64-bit k-mers); the reads are scored in four
    // chunks, so a batch may average ~24 (taxon, read k-mer) pairs per k-mer before the workspace is the limit.
    // KASA_SCORE_EXACT writes one group per (instance, level): up to ~2 num_k words per k-mer, hence more and smaller chunks.
    uint32_t n_chunks = n_reads >= 4096 ? 4 : 1;
    if (exact) n_chunks = std::max<uint32_t>(1, std::min<uint32_t>(n_reads, std::max<uint32_t>(n_chunks, 2u * (uint32_t)c->num_k)));
    for (uint32_t ch = 0; ch < n_chunks; ++ch) {
        const uint32_t r0 = (uint32_t)((uint64_t)n_reads * ch / n_chunks), r1 = (uint32_t)((uint64_t)n_reads * (ch + 1) / n_chunks);
        KASA_CUDA_TRY(cudaMemsetAsync(c->ops.cursors, 0, 16, c->stream));
        const bool counts_in_a = scores_counted_by_order_ops(lp, want_hits), emit = want_dense || want_hits || !counts_in_a;
        launch_order_ops(c->stream, c->d_rec, c->d_read_koff, r0, r1, n, max_inst, lp, peers, c->ops, trig, counts_in_a, emit, c->d_cnt_fx, c->d_cnt_un,
                         c->d_err, &c->launches);
        if (emit) launch_accumulate(c->stream, c->ops, r0, r1, lp, exact, want_dense ? sl.scores : nullptr, hits, c->d_cnt_fx, c->d_cnt_un, c->d_err, &c->launches);
    }
    launch_fold_counts(c->stream, c->d_cnt_fx, c->d_cnt_un, lp, c->d_tot_all, c->d_tot_un, &c->launches);
    if (want_hits) KASA_CUDA_TRY(cudaMemcpyAsync(sl.h_hits, sl.hit_cursor, 8, cudaMemcpyDeviceToHost, c->stream));
}

// sparse results of a finished batch -> the caller's buffers
static void deliver_hits(kasa_ctx *c, kasa_ctx::Slot &sl, kasa_batch_result *out, uint32_t n_reads)
{
    out->n_hits = 0;
    if (!out->hit_taxon || !sl.hit_taxon || !n_reads) return;
    const uint64_t nh = *sl.h_hits;
    KASA_REQUIRE(nh <= out->hit_capacity, KASA_ERR_NOMEM, "hit_capacity is smaller than the number of (read, taxon) hits of the batch");
    const cudaMemcpyKind kind = out->hits_location == KASA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    KASA_CUDA_TRY(cudaMemcpyAsync(out->hit_taxon, sl.hit_taxon, nh * 4, kind, c->cs_out));
    KASA_CUDA_TRY(cudaMemcpyAsync(out->hit_score, sl.hit_score, nh * 4, kind, c->cs_out));
    KASA_CUDA_TRY(cudaMemcpyAsync(out->hit_begin, sl.hit_begin, (size_t)n_reads * 8, kind, c->cs_out));
    KASA_CUDA_TRY(cudaMemcpyAsync(out->hit_count, sl.hit_count, (size_t)n_reads * 4, kind, c->cs_out));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->cs_out));
    out->n_hits = nh;
}

// completes the oldest submitted batch: waits for its kernels and score copy, reports timings and device errors
static void finish_slot(kasa_ctx *c, kasa_ctx::Slot &sl)
{
    sl.pending = false;
    c->completed += 1;
    KASA_CUDA_TRY(cudaEventSynchronize(sl.out_set ? sl.out : sl.done));
    kasa_batch_result *out = sl.res;
    KASA_CUDA_TRY(cudaEventElapsedTime(&out->gpu_ms, sl.ev0, sl.ev1));
    for (int i = 0; i < KASA_N_STAGES; ++i) {
        float ms = 0.f;
        KASA_CUDA_TRY(cudaEventElapsedTime(&ms, sl.sev[i], sl.sev[i + 1]));
        c->stage_ms[i] += ms;
    }
    c->stage_batches += 1;
    c->stage_passes += sl.n_passes;
    c->stage_kmers += sl.n;
    c->n_kmers_total += sl.n;
    const uint32_t f = sl.h_err[0] | sl.h_err[1] | sl.h_err[2] | sl.h_err[3];
    if (f) {
        KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        KASA_CUDA_TRY(cudaMemsetAsync(c->d_err, 0, 16, c->stream));
        throw_device_flags(f);
    }
    deliver_hits(c, sl, out, sl.n_reads);
}

int kasa_identify_submit(kasa_ctx *c, const kasa_read_batch *in, kasa_batch_result *out)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && in && out, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->idx_lo, KASA_ERR_STATE, "kasa_identify_batch before an index was loaded");
    KASA_REQUIRE(c->submitted - c->completed < (uint64_t)kSlots, KASA_ERR_STATE, "three batches are in flight: call kasa_identify_wait first");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    kasa_ctx::Slot &sl = c->slot[c->submitted % kSlots];
    KASA_CUDA_TRY(cudaEventRecord(sl.ev0, c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[0], c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[1], c->stream));      // re-recorded by extract_batch when there is work
    const uint32_t *first = nullptr;
    uint32_t max_inst = 0;
    const bool want_scores = c->p.score_mode != KASA_SCORE_NONE;
    const bool want_hits = want_scores && out->hit_taxon != nullptr;
    // the scores are computed whenever the score mode asks for them (into the slot's matrix); out->scores only decides about the copy
    const bool want_dense = want_scores && (out->scores != nullptr || (!want_hits && c->dense_scores));
    KASA_REQUIRE(!want_dense || c->dense_scores, KASA_ERR_ARG,
                 "the dense score matrix [max_batch_seqs][n_taxa] is too large for this taxonomy: use the sparse results (max_batch_hits, hit_*)");
    KASA_REQUIRE(!want_hits || (sl.hit_taxon && out->hit_score && out->hit_begin && out->hit_count), KASA_ERR_ARG,
                 "sparse results need kasa_params.max_batch_hits > 0 and all four hit_* buffers");
    const uint64_t n = extract_batch(c, in, sl, &first, &max_inst);
    KASA_CUDA_TRY(cudaEventRecord(sl.in_free, c->stream));
    sl.in_free_set = true;
    out->n_kmers = n;
    out->n_hits = 0;
    // identify orders the k-mers on their leading kIdentifySortBits only (lookup.cuh: launch_partition); the exact
    // score mode needs the reference's complete order (its flush points are positions in the sorted array)
    const bool exact = c->p.score_mode == KASA_SCORE_EXACT;
    const int sorted_from = exact ? 0 : std::max(0, 5 * c->K - c->sort_bits);
    SortPlan plan = plan_for_bit_range(sorted_from, 5 * c->K);
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[2], c->stream));
    // the slot's score buffer may still be on its way to the host for the batch two submissions ago
    if (sl.out_set) KASA_CUDA_TRY(cudaStreamWaitEvent(c->stream, sl.out, 0));
    sl.out_set = false;
    if (n) {
        const int cur = radix_sort(c->stream, c->sb, 0, n, plan, c->sort_ws, c->d_err, &c->launches, sl.sev[3]);
        KASA_CUDA_TRY(cudaEventRecord(sl.sev[4], c->stream));
        launch_read_offsets(c->stream, c->d_koff, first, in->n_reads, c->d_read_koff, &c->launches);
        IndexDev idx{c->idx_lo, c->idx_hi, c->idx_tax, c->idx_n, c->idx_gmask};
        LevelParams lp = level_params(c);
        const uint32_t tq = lookup_tile_queries(n, c->idx_n);
        KASA_CUDA_TRY(cudaMemsetAsync(c->ovf.cursor, 0, 8, c->stream));
        launch_partition(c->stream, c->sb.w0[cur], c->sb.w1[cur], n, idx, lp, c->d_part, tq, sorted_from, &c->launches);
        launch_lookup(c->stream, c->sb.w0[cur], c->sb.w1[cur], c->sb.v[cur], n, 0, idx, lp, c->d_part, c->d_rec, tq, c->ovf, 0, c->d_err,
                      &c->launches, scores_straight_from_records(c, max_inst, want_hits) && score_sorted_two_masks(c->score_lanes, lp));
        KASA_CUDA_TRY(cudaEventRecord(sl.sev[5], c->stream));
        score_stage(c, sl, n, in->n_reads, max_inst, cur, want_dense, want_hits);
    } else {
        for (int i = 3; i <= 5; ++i) KASA_CUDA_TRY(cudaEventRecord(sl.sev[i], c->stream));
        if (want_dense && in->n_reads) KASA_CUDA_TRY(cudaMemsetAsync(sl.scores, 0, (size_t)in->n_reads * c->n_taxa * 4, c->stream));
        if (want_hits && in->n_reads) {
            KASA_CUDA_TRY(cudaMemsetAsync(sl.hit_count, 0, (size_t)in->n_reads * 4, c->stream));
            KASA_CUDA_TRY(cudaMemsetAsync(sl.hit_begin, 0, (size_t)in->n_reads * 8, c->stream));
            *sl.h_hits = 0;
        }
    }
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[6], c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.ev1, c->stream));
    const bool copy_scores = want_dense && out->scores && in->n_reads;
    if (copy_scores && out->scores_location == KASA_MEM_DEVICE)
        KASA_CUDA_TRY(cudaMemcpyAsync(out->scores, sl.scores, (size_t)in->n_reads * c->n_taxa * 4, cudaMemcpyDeviceToDevice, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(sl.h_err, c->d_err, 16, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[7], c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.done, c->stream));
    if (copy_scores && out->scores_location != KASA_MEM_DEVICE) {
        // device -> host on its own stream: it overlaps the next batch's kernels
        KASA_CUDA_TRY(cudaStreamWaitEvent(c->cs_out, sl.done, 0));
        KASA_CUDA_TRY(cudaMemcpyAsync(out->scores, sl.scores, (size_t)in->n_reads * c->n_taxa * 4, cudaMemcpyDeviceToHost, c->cs_out));
        KASA_CUDA_TRY(cudaEventRecord(sl.out, c->cs_out));
        sl.out_set = true;
    }
    sl.res = out;
    sl.n = n;
    sl.n_reads = in->n_reads;
    sl.n_passes = n ? plan.n_passes : 0;
    sl.pending = true;
    c->submitted += 1;
    KASA_API_END(c)
}

int kasa_identify_wait(kasa_ctx *c)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_REQUIRE(c->submitted > c->completed, KASA_ERR_STATE, "kasa_identify_wait without a submitted batch");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    finish_slot(c, c->slot[c->completed % kSlots]);
    KASA_API_END(c)
}

int kasa_builtin_codon_table(uint8_t *table512)
{
    if (!table512) return KASA_ERR_ARG;
    builtin_codon_table(table512);
    return KASA_OK;
}

int kasa_host_alloc(void **ptr, uint64_t bytes)
{
    if (!ptr) return KASA_ERR_ARG;
    *ptr = nullptr;
    if (bytes == 0) return KASA_OK;
    return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? KASA_OK : KASA_ERR_NOMEM;
}

int kasa_host_free(void *ptr)
{
    if (!ptr) return KASA_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? KASA_OK : KASA_ERR_CUDA;
}

int kasa_identify_batch(kasa_ctx *c, const kasa_read_batch *in, kasa_batch_result *out)
{
    KASA_REQUIRE_PLAIN(c && in && out);
    while (c->submitted > c->completed) {
        int rc = kasa_identify_wait(c);
        if (rc != KASA_OK) return rc;
    }
    int rc = kasa_identify_submit(c, in, out);
    if (rc != KASA_OK) {                                   // nothing stays in flight after a failed synchronous call
        cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->cs_in); cudaStreamSynchronize(c->cs_out);
        return rc;
    }
    return kasa_identify_wait(c);
}

int kasa_identify_finish(kasa_ctx *c, double *count_all, uint64_t *count_unique, uint64_t *n_kmers_total)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_REQUIRE(c->idx_lo, KASA_ERR_STATE, "kasa_identify_finish before an index was loaded");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    while (c->submitted > c->completed) finish_slot(c, c->slot[c->completed % kSlots]);
    const size_t cells = (size_t)c->num_k * c->n_taxa;
    if (count_all) KASA_CUDA_TRY(cudaMemcpyAsync(count_all, c->d_tot_all, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    if (count_unique) KASA_CUDA_TRY(cudaMemcpyAsync(count_unique, c->d_tot_un, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (n_kmers_total) *n_kmers_total = c->n_kmers_total;
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_tot_all, 0, cells * 8, c->stream));
    KASA_CUDA_TRY(cudaMemsetAsync(c->d_tot_un, 0, cells * 8, c->stream));
    c->n_kmers_total = 0;
    KASA_API_END(c)
}

// ---------------------------------------------------------------------------------------------------------------
// prefix-sharded index (include/kasa_b200.h: kasa_shard_*)

// top six letters of file record i (12-byte / 20-byte records, packedPairs.hpp:107-136)
static uint32_t host_prefix30(const unsigned char *entries, uint64_t i, int bits, int K)
{
    uint64_t lo, hi = 0;
    const unsigned char *p = entries + i * (bits == 128 ? 20 : 12);
    std::memcpy(&lo, p, 8);
    if (bits == 128) std::memcpy(&hi, p + 8, 8);
    const int sh = 5 * K - 30;
    return bits == 128 ? (uint32_t)(hi >> (sh - 64)) : (uint32_t)(lo >> sh);
}

int kasa_shard_plan(const void *entries, uint64_t n, int bits, int k_low, int n_shards, uint64_t *entry_begin, uint32_t *prefix_begin)
{
    KASA_API_BEGIN
    KASA_REQUIRE((entries || n == 0) && entry_begin && prefix_begin, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(bits == 64 || bits == 128, KASA_ERR_ARG, "bits must be 64 or 128");
    KASA_REQUIRE(n_shards >= 1 && n_shards <= KASA_MAX_SHARDS, KASA_ERR_ARG, "n_shards must be in [1, KASA_MAX_SHARDS]");
    KASA_REQUIRE(k_low >= 1, KASA_ERR_ARG, "k_low must be >= 1");
    const int K = bits == 128 ? 25 : 12;
    const int align = std::min(6, k_low);
    const uint32_t keep = ~((1u << (5 * (6 - align))) - 1u);
    const unsigned char *e = static_cast<const unsigned char *>(entries);
    entry_begin[0] = 0; prefix_begin[0] = 0;
    for (int r = 1; r < n_shards; ++r) {
        if (n == 0) { entry_begin[r] = 0; prefix_begin[r] = 1u << 30; continue; }
        const uint64_t cand = (uint64_t)((unsigned __int128)n * (unsigned)r / (unsigned)n_shards);
        const uint32_t p = host_prefix30(e, cand, bits, K) & keep;
        uint64_t a = 0, b = n;
        while (a < b) {
            const uint64_t m = a + ((b - a) >> 1);
            if (host_prefix30(e, m, bits, K) < p) a = m + 1; else b = m;
        }
        entry_begin[r] = a; prefix_begin[r] = p;
    }
    entry_begin[n_shards] = n; prefix_begin[n_shards] = 1u << 30;
    KASA_API_END(nullptr)
}

static void shard_common_init(kasa_ctx *c, int rank, int n_shards, const uint32_t *prefix_begin)
{
    KASA_REQUIRE(n_shards >= 1 && n_shards <= KASA_MAX_SHARDS && rank >= 0 && rank < n_shards, KASA_ERR_ARG,
                 "need 0 <= rank < n_shards <= KASA_MAX_SHARDS");
    KASA_REQUIRE(c->p.score_mode != KASA_SCORE_EXACT, KASA_ERR_ARG,
                 "KASA_SCORE_EXACT replays one global flush order and cannot run on a prefix-sharded index");
    KASA_REQUIRE(c->sort_bits >= 30, KASA_ERR_ARG, "a prefix-sharded index needs the leading 30 k-mer bits sorted");
    for (int r = 0; r < n_shards; ++r)
        KASA_REQUIRE(prefix_begin[r] <= prefix_begin[r + 1], KASA_ERR_ARG, "prefix_begin must be non-decreasing");
    KASA_REQUIRE(prefix_begin[0] == 0 && prefix_begin[n_shards] == (1u << 30), KASA_ERR_ARG, "prefix_begin must span [0, 2^30]");
    ensure_workspace(c);
    if (!c->sh.d_prefix_begin) {
        c->sh.d_prefix_begin = dev_alloc<uint32_t>(c, KASA_MAX_SHARDS + 1);
        c->sh.d_begin = dev_alloc<uint64_t>(c, KASA_MAX_SHARDS + 1);
        KASA_CUDA_TRY(cudaEventCreate(&c->sh.ev_score0));
    }
    // the record stage is only needed to answer peers' identify queries, i.e. when this context holds an index shard
    if (n_shards > 1 && c->idx_lo && !c->sh.d_stage) c->sh.d_stage = dev_alloc<uint64_t>(c, (size_t)(c->p.max_batch_kmers + 2) * kRecWords);
    KASA_CUDA_TRY(cudaMemcpyAsync(c->sh.d_prefix_begin, prefix_begin, (n_shards + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->sh.on = true; c->sh.connected = false; c->sh.rank = rank; c->sh.n = n_shards; c->sh.phase = 0;
}

static void *const *shard_buffers(kasa_ctx *c, void *out[KASA_SHARD_NBUF])
{
    out[0] = c->sb.w0[0]; out[1] = c->sb.w0[1]; out[2] = c->sb.w1[0]; out[3] = c->sb.w1[1];
    out[4] = c->sb.v[0]; out[5] = c->sb.v[1]; out[6] = c->d_rec; out[7] = c->sh.d_stage; out[8] = c->ovf.data;
    return out;
}

int kasa_shard_init(kasa_ctx *c, int rank, int n_shards, const uint32_t *prefix_begin, kasa_shard_mem *mine)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && prefix_begin && mine, KASA_ERR_ARG, "null argument");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    shard_common_init(c, rank, n_shards, prefix_begin);
    std::memset(mine, 0, sizeof(*mine));
    mine->pid = (uint64_t)getpid();
    mine->device = c->p.device;
    void *buf[KASA_SHARD_NBUF];
    shard_buffers(c, buf);
    for (int b = 0; b < KASA_SHARD_NBUF; ++b) {
        mine->ptr[b] = (uint64_t)(uintptr_t)buf[b];
        if (!buf[b]) continue;
        cudaIpcMemHandle_t h;
        KASA_CUDA_TRY(cudaIpcGetMemHandle(&h, buf[b]));
        static_assert(sizeof(h) <= 64, "cudaIpcMemHandle_t does not fit kasa_shard_mem::handle");
        std::memcpy(mine->handle[b], &h, sizeof(h));
    }
    KASA_API_END(c)
}

int kasa_shard_connect(kasa_ctx *c, const kasa_shard_mem *all)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && all, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on, KASA_ERR_STATE, "kasa_shard_connect before kasa_shard_init");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    // mapping a peer's buffers costs ~0.1 s per GB-scale allocation: keep the mappings when the same buffers come again
    if ((int)c->sh.connected_to.size() == c->sh.n && std::memcmp(c->sh.connected_to.data(), all, sizeof(kasa_shard_mem) * c->sh.n) == 0) {
        c->sh.connected = true;
        return KASA_OK;
    }
    for (void *p : c->sh.opened) cudaIpcCloseMemHandle(p);
    c->sh.opened.clear();
    c->sh.connected_to.clear();
    const uint64_t me = (uint64_t)getpid();
    for (int o = 0; o < c->sh.n; ++o) {
        const kasa_shard_mem &m = all[o];
        if (m.pid == me && m.device != c->p.device) {       // same process, other GPU: plain peer access
            int can = 0;
            KASA_CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->p.device, m.device));
            KASA_REQUIRE(can, KASA_ERR_CUDA, "no peer access between device " + std::to_string(c->p.device) + " and " + std::to_string(m.device));
            cudaError_t e = cudaDeviceEnablePeerAccess(m.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) KASA_CUDA_TRY(e);
            cudaGetLastError();
        }
        for (int b = 0; b < KASA_SHARD_NBUF; ++b) {
            if (!m.ptr[b]) { c->sh.peer[o][b] = 0; continue; }
            if (m.pid == me) { c->sh.peer[o][b] = m.ptr[b]; continue; }
            cudaIpcMemHandle_t h;
            std::memcpy(&h, m.handle[b], sizeof(h));
            void *p = nullptr;
            KASA_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            c->sh.opened.push_back(p);
            c->sh.peer[o][b] = (uint64_t)(uintptr_t)p;
        }
    }
    c->sh.connected_to.assign(all, all + c->sh.n);
    c->sh.connected = true;
    KASA_API_END(c)
}

int kasa_index_from_build_shard(kasa_ctx *c, const uint32_t *taxid_of_idx, uint32_t n_taxa, int rank, int n_shards,
                                uint64_t *entry_begin, uint32_t *prefix_begin)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && taxid_of_idx && entry_begin && prefix_begin, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->build_done, KASA_ERR_STATE, "kasa_index_from_build_shard needs a finished build (kasa_build_finish)");
    KASA_REQUIRE(!c->host_result, KASA_ERR_STATE, "the build spilled bricks to host memory: fetch it (kasa_build_fetch) and load the shards with kasa_index_load");
    KASA_REQUIRE(n_shards >= 1 && n_shards <= KASA_MAX_SHARDS && rank >= 0 && rank < n_shards, KASA_ERR_ARG,
                 "need 0 <= rank < n_shards <= KASA_MAX_SHARDS");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    drop_index(c);
    set_taxa(c, n_taxa);
    const uint64_t n_all = c->build_fill;
    const int b = c->build_buf;
    uint64_t *d_eb = dev_alloc<uint64_t>(c, KASA_MAX_SHARDS + 1);
    uint32_t *d_pb = dev_alloc<uint32_t>(c, KASA_MAX_SHARDS + 1);
    launch_shard_plan(c->stream, c->sb.w0[b], c->wide ? c->sb.w1[b] : nullptr, n_all, c->K, std::min(6, c->p.k_low), n_shards, d_eb, d_pb,
                      &c->launches);
    KASA_CUDA_TRY(cudaMemcpyAsync(entry_begin, d_eb, (n_shards + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(prefix_begin, d_pb, (n_shards + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    dev_free(c, d_eb); dev_free(c, d_pb);
    const uint64_t o = entry_begin[rank], n = entry_begin[rank + 1] - o;
    c->idx_lo = dev_alloc<uint64_t>(c, n + 16);
    c->idx_hi = c->wide ? dev_alloc<uint64_t>(c, n + 16) : nullptr;
    c->idx_tax = dev_alloc<uint16_t>(c, n + 16);
    KASA_CUDA_TRY(cudaMemcpyAsync(c->idx_lo, c->sb.w0[b] + o, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    if (c->wide) KASA_CUDA_TRY(cudaMemcpyAsync(c->idx_hi, c->sb.w1[b] + o, n * 8, cudaMemcpyDeviceToDevice, c->stream));
    TaxTable tt = upload_tax_table(c, taxid_of_idx, n_taxa);
    launch_map_taxids(c->stream, c->sb.v[b] + o, n, tt.d_taxid, tt.d_idx, tt.n, c->idx_tax, c->d_err, &c->launches);
    finalize_index(c, n);
    KASA_CUDA_TRY(cudaGetLastError());
    check_device_flags(c);
    dev_free(c, tt.d_taxid); dev_free(c, tt.d_idx);
    c->idx_n = n;
    KASA_API_END(c)
}

int kasa_shard_plan_hist(const uint64_t *hist, int k_low, int n_shards, uint32_t *prefix_begin)
{
    KASA_API_BEGIN
    KASA_REQUIRE(hist && prefix_begin, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(n_shards >= 1 && n_shards <= KASA_MAX_SHARDS && k_low >= 1, KASA_ERR_ARG, "need 1 <= n_shards <= KASA_MAX_SHARDS, k_low >= 1");
    const int bits = KASA_SHARD_HIST_BITS;
    const int align_bits = std::min(bits, 5 * std::min(6, k_low));       // boundaries on min(6, k_low)-letter prefixes
    const uint32_t step = 1u << (bits - align_bits);
    unsigned __int128 total = 0;
    for (uint32_t b = 0; b < (1u << bits); ++b) total += hist[b];
    prefix_begin[0] = 0;
    uint32_t b = 0;
    unsigned __int128 cum = 0;                                            // records in bins < b
    for (int r = 1; r < n_shards; ++r) {
        const unsigned __int128 want = total * (unsigned)r / (unsigned)n_shards;
        // advance whole aligned groups while the group's END stays <= want (cut = start of the group holding the candidate)
        while (b < (1u << bits)) {
            unsigned __int128 g = 0;
            for (uint32_t i = 0; i < step; ++i) g += hist[b + i];
            if (cum + g > want) break;
            cum += g; b += step;
        }
        prefix_begin[r] = b << (30 - bits);
    }
    prefix_begin[n_shards] = 1u << 30;
    KASA_API_END(nullptr)
}

int kasa_build_shard_local(kasa_ctx *c, uint64_t *hist, uint64_t *n_local)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && hist, KASA_ERR_ARG, "null argument");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    KASA_REQUIRE(!c->build_done, KASA_ERR_STATE, "kasa_build_shard_local after kasa_build_finish");
    build_compact(c);
    const size_t bins = (size_t)1 << KASA_SHARD_HIST_BITS;
    uint64_t *d_hist = dev_alloc<uint64_t>(c, bins);
    const int b = c->build_buf;
    launch_prefix_hist(c->stream, c->sb.w0[b], c->wide ? c->sb.w1[b] : nullptr, c->build_fill, c->K, KASA_SHARD_HIST_BITS, d_hist, &c->launches);
    KASA_CUDA_TRY(cudaMemcpyAsync(hist, d_hist, bins * 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    dev_free(c, d_hist);
    if (n_local) *n_local = c->build_fill;
    KASA_API_END(c)
}

int kasa_build_shard_runs(kasa_ctx *c, kasa_shard_runs *mine)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && mine, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on && c->sh.connected, KASA_ERR_STATE, "kasa_build_shard_runs before kasa_shard_init / kasa_shard_connect");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    std::memset(mine, 0, sizeof(*mine));
    const int b = c->build_buf;
    mine->n_kmers = c->build_fill;
    mine->cur = b;
    launch_run_bounds(c->stream, c->sb.w0[b], c->wide ? c->sb.w1[b] : nullptr, c->build_fill, c->K, c->sh.d_prefix_begin, c->sh.n, c->sh.d_begin,
                      &c->launches);
    KASA_CUDA_TRY(cudaMemcpyAsync(mine->begin, c->sh.d_begin, (c->sh.n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    KASA_API_END(c)
}

int kasa_build_shard_pull(kasa_ctx *c, const kasa_shard_runs *all)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && all, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on && c->sh.connected, KASA_ERR_STATE, "kasa_build_shard_pull before kasa_shard_init / kasa_shard_connect");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    const int me = c->sh.rank, dst = c->build_buf ^ 1;
    uint64_t total = 0;
    for (int o = 0; o < c->sh.n; ++o) {
        const kasa_shard_runs &r = all[o];
        KASA_REQUIRE(r.cur == 0 || r.cur == 1, KASA_ERR_ARG, "kasa_shard_runs.cur must be 0 or 1");
        KASA_REQUIRE(r.begin[me] <= r.begin[me + 1] && r.begin[me + 1] <= r.n_kmers, KASA_ERR_ARG, "kasa_shard_runs.begin is not monotone");
        total += r.begin[me + 1] - r.begin[me];
    }
    KASA_REQUIRE(total <= c->p.max_batch_kmers, KASA_ERR_NOMEM, "this shard's records exceed max_batch_kmers");
    uint64_t off = 0;
    for (int i = 0; i < c->sh.n; ++i) {
        const int o = (me + i) % c->sh.n;                  // start with the local run: spreads the NVLink traffic
        const kasa_shard_runs &r = all[o];
        const uint64_t b0 = r.begin[me], m = r.begin[me + 1] - b0;
        if (!m) continue;
        KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.w0[dst] + off, reinterpret_cast<const uint64_t *>(c->sh.peer[o][r.cur]) + b0, m * 8, cudaMemcpyDefault, c->stream));
        if (c->wide)
            KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.w1[dst] + off, reinterpret_cast<const uint64_t *>(c->sh.peer[o][2 + r.cur]) + b0, m * 8, cudaMemcpyDefault, c->stream));
        KASA_CUDA_TRY(cudaMemcpyAsync(c->sb.v[dst] + off, reinterpret_cast<const uint32_t *>(c->sh.peer[o][4 + r.cur]) + b0, m * 4, cudaMemcpyDefault, c->stream));
        off += m;
    }
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->sh.pull_pending = true;                              // adopted by kasa_build_finish, after the host's barrier
    c->sh.pulled_total = total;
    KASA_API_END(c)
}

int kasa_shard_sort(kasa_ctx *c, const kasa_read_batch *in, kasa_shard_runs *mine)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && in && mine, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on && c->sh.connected, KASA_ERR_STATE, "kasa_shard_sort before kasa_shard_init / kasa_shard_connect");
    KASA_REQUIRE(c->idx_lo, KASA_ERR_STATE, "kasa_shard_sort before an index shard was loaded");
    KASA_REQUIRE(c->submitted == c->completed, KASA_ERR_STATE, "pipelined batches are still in flight");
    KASA_REQUIRE(c->sh.phase == 0, KASA_ERR_STATE, "kasa_shard_sort: the previous batch has not been scored");
    KASA_REQUIRE(c->sh.n == 1 || c->sh.d_stage, KASA_ERR_STATE, "load the index shard before kasa_shard_init");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    kasa_ctx::Slot &sl = c->slot[0];
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[0], c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[1], c->stream));
    const uint32_t *first = nullptr;
    uint32_t max_inst = 0;
    const uint64_t n = extract_batch(c, in, sl, &first, &max_inst);
    c->sh.max_inst = max_inst;
    KASA_CUDA_TRY(cudaEventRecord(sl.in_free, c->stream));
    sl.in_free_set = true;
    const int sorted_from = std::max(0, 5 * c->K - c->sort_bits);
    SortPlan plan = plan_for_bit_range(sorted_from, 5 * c->K);
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[2], c->stream));
    std::memset(mine, 0, sizeof(*mine));
    mine->n_kmers = n;
    // what the owners of the index ranges need to know about this batch: bit 0 = its scores come straight from two-mask records
    mine->reserved = scores_straight_from_records(c, max_inst, false) && score_sorted_two_masks(c->score_lanes, level_params(c)) ? 1 : 0;
    if (n) {
        const int cur = radix_sort(c->stream, c->sb, 0, n, plan, c->sort_ws, c->d_err, &c->launches, sl.sev[3]);
        mine->cur = cur;
        launch_read_offsets(c->stream, c->d_koff, first, in->n_reads, c->d_read_koff, &c->launches);
        launch_run_bounds(c->stream, c->sb.w0[cur], c->wide ? c->sb.w1[cur] : nullptr, n, c->K, c->sh.d_prefix_begin, c->sh.n, c->sh.d_begin,
                          &c->launches);
        KASA_CUDA_TRY(cudaMemcpyAsync(mine->begin, c->sh.d_begin, (c->sh.n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    } else {
        KASA_CUDA_TRY(cudaEventRecord(sl.sev[3], c->stream));
    }
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[4], c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));        // the sorted run is complete before any peer is told about it
    sl.n = n;
    sl.n_passes = n ? plan.n_passes : 0;
    c->sh.n_reads = in->n_reads;
    c->sh.phase = 1;
    KASA_API_END(c)
}

int kasa_shard_lookup(kasa_ctx *c, const kasa_shard_runs *all)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && all, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on && c->sh.phase == 1, KASA_ERR_STATE, "kasa_shard_lookup must follow kasa_shard_sort");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    kasa_ctx::Slot &sl = c->slot[0];
    IndexDev idx{c->idx_lo, c->idx_hi, c->idx_tax, c->idx_n, c->idx_gmask};
    LevelParams lp = level_params(c);
    const int sorted_from = std::max(0, 5 * c->K - c->sort_bits);
    const int me = c->sh.rank;
    // serve every origin's run of this shard's prefix range, starting with the local one (spreads the NVLink traffic):
    // queries are read from the origin's sorted buffer; the local run's records go straight to their instance slots,
    // a peer's records are left in run order in the stage (segment o starts where the runs of the origins before o end)
    uint64_t seg[KASA_MAX_SHARDS + 1] = {0};
    KASA_CUDA_TRY(cudaMemsetAsync(c->ovf.cursor, 0, 8, c->stream));
    for (int o = 0; o < c->sh.n; ++o) {
        const kasa_shard_runs &r = all[o];
        KASA_REQUIRE(r.cur == 0 || r.cur == 1, KASA_ERR_ARG, "kasa_shard_runs.cur must be 0 or 1");
        KASA_REQUIRE(r.begin[me] <= r.begin[me + 1] && r.begin[me + 1] <= r.n_kmers, KASA_ERR_ARG, "kasa_shard_runs.begin is not monotone");
        seg[o + 1] = seg[o] + (o == me ? 0 : r.begin[me + 1] - r.begin[me]);
    }
    KASA_REQUIRE(seg[c->sh.n] <= c->p.max_batch_kmers, KASA_ERR_NOMEM,
                 "the peers' runs of this shard's prefix range exceed max_batch_kmers records (unbalanced shards or batches)");
    for (int i = 0; i < c->sh.n; ++i) {
        const int o = (me + i) % c->sh.n;
        const kasa_shard_runs &r = all[o];
        const uint64_t b0 = r.begin[me], m = r.begin[me + 1] - b0;
        if (!m) continue;
        KASA_REQUIRE(m <= c->p.max_batch_kmers, KASA_ERR_NOMEM, "a peer's run exceeds max_batch_kmers");
        const uint64_t *q_lo = reinterpret_cast<const uint64_t *>(c->sh.peer[o][r.cur]) + b0;
        const uint64_t *q_hi = c->wide ? reinterpret_cast<const uint64_t *>(c->sh.peer[o][2 + r.cur]) + b0 : nullptr;
        const uint32_t *inst = o == me ? c->sb.v[r.cur] + b0 : nullptr;
        uint64_t *rec = o == me ? c->d_rec : c->sh.d_stage + seg[o] * kRecWords;
        const uint32_t tq = lookup_tile_queries(m, c->idx_n);
        launch_partition(c->stream, q_lo, q_hi, m, idx, lp, c->d_part, tq, sorted_from, &c->launches);
        // the rank of a k-mer is its position in the ORIGIN's sorted run; overflow lists stay in this shard's array (owner = me)
        launch_lookup(c->stream, q_lo, q_hi, inst, m, b0, idx, lp, c->d_part, rec, tq, c->ovf, (uint32_t)me, c->d_err, &c->launches, (r.reserved & 1) != 0);
    }
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[5], c->stream));
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));        // every record this shard owes is in its origin's HBM
    c->sh.phase = 2;
    KASA_API_END(c)
}

int kasa_shard_score(kasa_ctx *c, const kasa_shard_runs *all, kasa_batch_result *out)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && all && out, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->sh.on && c->sh.phase == 2, KASA_ERR_STATE, "kasa_shard_score must follow kasa_shard_lookup");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    kasa_ctx::Slot &sl = c->slot[0];
    const uint64_t n = sl.n;
    const uint32_t n_reads = c->sh.n_reads;
    const bool want_scores = c->p.score_mode != KASA_SCORE_NONE;
    LevelParams lp = level_params(c);
    KASA_CUDA_TRY(cudaEventRecord(c->sh.ev_score0, c->stream));
    // pull the records the peers computed for this context's sorted run into emission order
    const int me = c->sh.rank;
    const kasa_shard_runs &mine = all[me];
    KASA_REQUIRE(mine.n_kmers == n, KASA_ERR_ARG, "kasa_shard_score: all[rank] is not this context's batch");
    for (int i = 1; i < c->sh.n && n; ++i) {
        const int w = (me + i) % c->sh.n;
        uint64_t seg = 0;                                     // where owner w staged this origin's segment
        for (int o = 0; o < me; ++o) seg += o == w ? 0 : all[o].begin[w + 1] - all[o].begin[w];
        const uint64_t b0 = mine.begin[w], m = mine.begin[w + 1] - b0;
        launch_scatter_records(c->stream, reinterpret_cast<const uint64_t *>(c->sh.peer[w][7]) + seg * kRecWords, c->sb.v[mine.cur] + b0, m,
                               c->d_rec, &c->launches);
    }
    const bool want_hits = want_scores && out->hit_taxon != nullptr;
    const bool want_dense = want_scores && (out->scores != nullptr || (!want_hits && c->dense_scores));
    KASA_REQUIRE(!want_dense || c->dense_scores, KASA_ERR_ARG,
                 "the dense score matrix [max_batch_seqs][n_taxa] is too large for this taxonomy: use the sparse results (max_batch_hits, hit_*)");
    KASA_REQUIRE(!want_hits || (sl.hit_taxon && out->hit_score && out->hit_begin && out->hit_count), KASA_ERR_ARG,
                 "sparse results need kasa_params.max_batch_hits > 0 and all four hit_* buffers");
    (void)lp;
    if (n) {
        score_stage(c, sl, n, n_reads, c->sh.max_inst, mine.cur, want_dense, want_hits);
    } else {
        if (want_dense && n_reads) KASA_CUDA_TRY(cudaMemsetAsync(sl.scores, 0, (size_t)n_reads * c->n_taxa * 4, c->stream));
        if (want_hits && n_reads) {
            KASA_CUDA_TRY(cudaMemsetAsync(sl.hit_count, 0, (size_t)n_reads * 4, c->stream));
            KASA_CUDA_TRY(cudaMemsetAsync(sl.hit_begin, 0, (size_t)n_reads * 8, c->stream));
            *sl.h_hits = 0;
        }
    }
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[6], c->stream));
    if (want_dense && out->scores && n_reads)
        KASA_CUDA_TRY(cudaMemcpyAsync(out->scores, sl.scores, (size_t)n_reads * c->n_taxa * 4,
                                      out->scores_location == KASA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaEventRecord(sl.sev[7], c->stream));
    c->sh.phase = 0;
    check_device_flags(c);
    deliver_hits(c, sl, out, n_reads);
    float total = 0.f;
    for (int i = 0; i < KASA_N_STAGES; ++i) {
        float ms = 0.f;
        if (i == KASA_STAGE_PER_READ) KASA_CUDA_TRY(cudaEventElapsedTime(&ms, c->sh.ev_score0, sl.sev[6]));   // not the host barrier before it
        else KASA_CUDA_TRY(cudaEventElapsedTime(&ms, sl.sev[i], sl.sev[i + 1]));
        c->stage_ms[i] += ms;
        total += ms;
    }
    c->stage_batches += 1;
    c->stage_passes += sl.n_passes;
    c->stage_kmers += n;
    c->n_kmers_total += n;
    out->n_kmers = n;
    out->gpu_ms = total;
    KASA_API_END(c)
}

int kasa_build_push(kasa_ctx *c, const kasa_seq_batch *in)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && in, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(in->n_seqs <= c->p.max_batch_seqs, KASA_ERR_NOMEM, "batch has more sequences than max_batch_seqs");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    if (c->build_done) { c->build_done = false; c->build_fill = 0; c->build_buf = 0; c->host_result = false; c->host_n = 0; c->host_words.clear(); c->host_trie.clear(); }
    const uint32_t ns = in->n_seqs;
    if (ns == 0) return KASA_OK;
    const char *ascii = in->bases;
    if (in->location == KASA_MEM_HOST) {
        uint64_t total = host_total_bases(in->seq_start, in->seq_len, ns);
        KASA_REQUIRE(total <= c->p.max_batch_bases, KASA_ERR_NOMEM, "batch has more bases than max_batch_bases");
        KASA_CUDA_TRY(cudaMemcpyAsync(c->d_ascii, in->bases, total, cudaMemcpyHostToDevice, c->stream));
        ascii = c->d_ascii;
    }
    const uint64_t *seq_start = stage_array(c, in->seq_start, ns, c->d_seq_start, in->location);
    const uint32_t *seq_len = stage_array(c, in->seq_len, ns, c->d_seq_len, in->location);
    const uint32_t *taxid = stage_array(c, in->taxid, ns, c->d_taxid, in->location);
    const int one = c->p.frames == 1 ? 1 : 0;                       // build --one: every third window of a string
    launch_build_counts(c->stream, seq_len, ns, c->K, one, c->d_koff, &c->launches);
    exclusive_scan_u64(c->stream, c->d_koff, ns, c->d_scan_tmp, &c->launches);
    const uint64_t nw = read_u64(c, c->d_koff + ns);
    KASA_REQUIRE(nw <= c->p.max_batch_kmers, KASA_ERR_NOMEM, "one push yields more k-mer windows than max_batch_kmers: push fewer sequences at a time");
    if (c->build_fill + nw > c->p.max_batch_kmers) {                     // the brick is full: Build.hpp:116-122
        build_compact(c);
        // still (more than half) full of distinct records: it leaves for host memory and a new brick starts (Build.hpp:328-352)
        if (c->build_fill + nw > c->p.max_batch_kmers || c->build_fill > c->p.max_batch_kmers / 2) {
            KASA_REQUIRE(!c->sh.on, KASA_ERR_NOMEM, "the brick of a prefix-sharded build is full: use more shards or a larger max_batch_kmers");
            spill_brick(c);
        }
    }
    const int b = c->build_buf;
    const uint64_t o = c->build_fill;
    launch_extract_build(c->stream, ascii, seq_start, seq_len, taxid, ns, c->K, one, c->d_codon, c->d_koff, nw, c->sb.w0[b] + o,
                         c->wide ? c->sb.w1[b] + o : nullptr, c->sb.v[b] + o, &c->launches);
    KASA_CUDA_TRY(cudaGetLastError());
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));    // staging buffers are reused by the next push
    c->build_fill += nw;
    KASA_API_END(c)
}

int kasa_build_push_index(kasa_ctx *c, const void *entries, uint64_t n, int bits)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && (entries || n == 0), KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(bits == (c->wide ? 128 : 64), KASA_ERR_ARG, "index width does not match highest_k of the context");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    if (c->build_done) { c->build_done = false; c->build_fill = 0; c->build_buf = 0; c->host_result = false; c->host_n = 0; c->host_words.clear(); c->host_trie.clear(); }
    const size_t wpr = c->wide ? 5 : 3;
    const uint64_t cap = c->p.max_batch_kmers;
    // the records join the brick in pieces of at most half a brick (staged as file words in the free half of the ping-pong)
    const uint64_t piece = std::max<uint64_t>(1, std::min<uint64_t>(cap / 2, cap * 8 / (wpr * 4)));
    for (uint64_t o = 0; o < n; o += piece) {
        const uint64_t m = std::min(piece, n - o);
        if (c->build_fill + m > cap) {
            build_compact(c);
            if (c->build_fill + m > cap || c->build_fill > cap / 2) {
                KASA_REQUIRE(!c->sh.on, KASA_ERR_NOMEM, "the brick of a prefix-sharded build is full: use more shards or a larger max_batch_kmers");
                spill_brick(c);
            }
        }
        const int b = c->build_buf;
        uint32_t *stage = reinterpret_cast<uint32_t *>(c->sb.w0[b ^ 1]);
        KASA_CUDA_TRY(cudaMemcpyAsync(stage, static_cast<const char *>(entries) + o * wpr * 4, m * wpr * 4, cudaMemcpyHostToDevice, c->stream));
        launch_unpack_records(c->stream, stage, m, c->wide, c->sb.w0[b] + c->build_fill, c->wide ? c->sb.w1[b] + c->build_fill : nullptr,
                              c->sb.v[b] + c->build_fill, &c->launches);
        KASA_CUDA_TRY(cudaGetLastError());
        KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->build_fill += m;
    }
    KASA_API_END(c)
}

int kasa_build_finish(kasa_ctx *c, uint64_t *n_entries, uint64_t *n_trie)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    if (c->sh.pull_pending) {                               // prefix-sharded build: the brick is what this shard pulled from its peers
        c->build_buf ^= 1;
        c->build_fill = c->sh.pulled_total;
        c->sh.pull_pending = false;
    }
    build_compact(c);
    uint64_t n = c->build_fill;
    if (!c->spills.empty()) {
        spill_brick(c);                                      // the last brick joins the others; the merged index lives in host memory
        merge_spills(c);
        n = c->host_n;
    } else {
        build_trie(c);
    }
    c->build_done = true;
    if (n_entries) *n_entries = n;
    if (n_trie) *n_trie = c->n_trie;
    KASA_API_END(c)
}

int kasa_build_fetch(kasa_ctx *c, void *entries, void *trie)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_REQUIRE(c->build_done, KASA_ERR_STATE, "kasa_build_fetch before kasa_build_finish");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    if (c->host_result) {
        if (trie && c->n_trie) std::memcpy(trie, c->host_trie.data(), c->n_trie * 12);
        if (entries && c->host_n) std::memcpy(entries, c->host_words.data(), c->host_words.size() * 4);
        return KASA_OK;
    }
    const uint64_t n = c->build_fill;
    const int b = c->build_buf;
    if (trie && c->n_trie) {
        KASA_CUDA_TRY(cudaMemcpyAsync(trie, c->d_trie_words, c->n_trie * 12, cudaMemcpyDeviceToHost, c->stream));
        KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (entries && n) {
        // file records through the free half of the sort ping-pong (enough for a chunk of n/2 records)
        const size_t wpr = c->wide ? 5 : 3;
        uint32_t *stage = reinterpret_cast<uint32_t *>(c->sb.w0[b ^ 1]);
        const uint64_t chunk = std::max<uint64_t>(1, c->p.max_batch_kmers * 8 / (wpr * 4));
        for (uint64_t o = 0; o < n; o += chunk) {
            const uint64_t m = std::min(chunk, n - o);
            launch_pack_records(c->stream, c->sb.w0[b] + o, c->wide ? c->sb.w1[b] + o : nullptr, c->sb.v[b] + o, m, stage, &c->launches);
            KASA_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(entries) + o * wpr * 4, stage, m * wpr * 4, cudaMemcpyDeviceToHost, c->stream));
            KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    KASA_API_END(c)
}

int kasa_build_frequency(kasa_ctx *c, const uint32_t *taxid_of_idx, uint32_t n_taxa, uint64_t *freq)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && taxid_of_idx && freq, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(c->build_done, KASA_ERR_STATE, "kasa_build_frequency before kasa_build_finish");
    KASA_REQUIRE(n_taxa >= 1 && n_taxa <= 65535, KASA_ERR_ARG, "n_taxa must be in [1, 65535]");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    const uint64_t n = c->build_fill;
    const int b = c->build_buf;
    const size_t cells = (size_t)n_taxa * c->K;
    TaxTable tt = upload_tax_table(c, taxid_of_idx, n_taxa);
    uint16_t *dense = reinterpret_cast<uint16_t *>(c->sb.v[b ^ 1]);     // free half of the ping-pong
    unsigned long long *d_freq = dev_alloc<unsigned long long>(c, cells);
    KASA_CUDA_TRY(cudaMemsetAsync(d_freq, 0, cells * 8, c->stream));
    if (c->host_result) {
        // merged index in host memory: streamed through the device brick by brick
        const size_t wpr = c->wide ? 5 : 3;
        const uint64_t chunk = std::max<uint64_t>(1, c->p.max_batch_kmers * 8 / (wpr * 4));
        uint32_t *stage = reinterpret_cast<uint32_t *>(c->sb.w0[1]);
        for (uint64_t o = 0; o < c->host_n; o += chunk) {
            const uint64_t k = std::min(chunk, c->host_n - o);
            KASA_CUDA_TRY(cudaMemcpyAsync(stage, c->host_words.data() + (size_t)o * wpr, k * wpr * 4, cudaMemcpyHostToDevice, c->stream));
            launch_unpack_records(c->stream, stage, k, c->wide, c->sb.w0[0], c->sb.w1[0], c->sb.v[0], &c->launches);
            launch_map_taxids(c->stream, c->sb.v[0], k, tt.d_taxid, tt.d_idx, tt.n, reinterpret_cast<uint16_t *>(c->sb.v[1]), c->d_err, &c->launches);
            launch_frequency(c->stream, c->sb.w0[0], c->sb.w1[0], reinterpret_cast<uint16_t *>(c->sb.v[1]), k, c->K, n_taxa, d_freq, &c->launches);
            KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    } else {
    launch_map_taxids(c->stream, c->sb.v[b], n, tt.d_taxid, tt.d_idx, tt.n, dense, c->d_err, &c->launches);
    launch_frequency(c->stream, c->sb.w0[b], c->sb.w1[b], dense, n, c->K, n_taxa, d_freq, &c->launches);
    }
    KASA_CUDA_TRY(cudaGetLastError());
    KASA_CUDA_TRY(cudaMemcpyAsync(freq, d_freq, cells * 8, cudaMemcpyDeviceToHost, c->stream));
    check_device_flags(c);
    dev_free(c, d_freq); dev_free(c, tt.d_taxid); dev_free(c, tt.d_idx);
    KASA_API_END(c)
}

int kasa_dbg_extract(kasa_ctx *c, const kasa_read_batch *in, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *read_id,
                     uint64_t *n_kmers)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && in && kmers_lo && n_kmers, KASA_ERR_ARG, "null argument");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    const uint32_t *first_dev = nullptr;
    const uint64_t n = extract_batch(c, in, c->slot[0], &first_dev);
    KASA_CUDA_TRY(cudaEventRecord(c->slot[0].in_free, c->stream));
    c->slot[0].in_free_set = true;
    *n_kmers = n;
    if (n) {
        KASA_CUDA_TRY(cudaMemcpyAsync(kmers_lo, c->sb.w0[0], n * 8, cudaMemcpyDeviceToHost, c->stream));
        if (c->wide && kmers_hi) KASA_CUDA_TRY(cudaMemcpyAsync(kmers_hi, c->sb.w1[0], n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (read_id && in->n_seqs) {
        std::vector<uint64_t> koff((size_t)in->n_seqs + 1);
        KASA_CUDA_TRY(cudaMemcpyAsync(koff.data(), c->d_koff, koff.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        std::vector<uint32_t> first;
        if (in->read_first_seq) {
            first.resize((size_t)in->n_reads + 1);
            KASA_CUDA_TRY(cudaMemcpyAsync(first.data(), in->read_first_seq, first.size() * 4,
                                          in->location == KASA_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost, c->stream));
        }
        KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (uint32_t r = 0; r < in->n_reads; ++r) {
            uint32_t s0 = in->read_first_seq ? first[r] : r, s1 = in->read_first_seq ? first[r + 1] : r + 1;
            for (uint64_t i = koff[s0]; i < koff[s1]; ++i) read_id[i] = r;
        }
    }
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    KASA_API_END(c)
}

int kasa_dbg_sort_pairs(kasa_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t n, int bit_lo, int bit_hi)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && keys && vals, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(n <= c->p.max_batch_kmers, KASA_ERR_NOMEM, "more records than max_batch_kmers");
    KASA_REQUIRE(0 <= bit_lo && bit_lo < bit_hi && bit_hi <= 64, KASA_ERR_ARG, "need 0 <= bit_lo < bit_hi <= 64");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    if (n == 0) return KASA_OK;
    SortBuffers sb = c->sb;
    sb.w1[0] = sb.w1[1] = nullptr;
    KASA_CUDA_TRY(cudaMemcpyAsync(sb.w0[0], keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(sb.v[0], vals, n * 4, cudaMemcpyHostToDevice, c->stream));
    SortPlan plan{};
    for (int b = bit_lo; b < bit_hi; b += 8) {
        int bits = std::min(8, bit_hi - b);
        plan.pass[plan.n_passes++] = SortPass{SRC_W0, b, (1u << bits) - 1u};
    }
    const int cur = radix_sort(c->stream, sb, 0, n, plan, c->sort_ws, c->d_err, &c->launches);
    KASA_CUDA_TRY(cudaMemcpyAsync(keys, sb.w0[cur], n * 8, cudaMemcpyDeviceToHost, c->stream));
    KASA_CUDA_TRY(cudaMemcpyAsync(vals, sb.v[cur], n * 4, cudaMemcpyDeviceToHost, c->stream));
    check_device_flags(c);
    KASA_API_END(c)
}

int kasa_synth_build_reference(kasa_ctx *c, uint64_t seed, uint32_t n_taxa, uint64_t genome_len)
{
    return kasa_synth_build_reference_range(c, seed, 0, n_taxa, genome_len);
}

int kasa_synth_build_reference_range(kasa_ctx *c, uint64_t seed, uint32_t first_taxon, uint32_t n_taxa, uint64_t genome_len)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c, KASA_ERR_ARG, "null context");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    ensure_workspace(c);
    const uint32_t marker = 3 * (c->K - 1);
    KASA_REQUIRE(genome_len + marker <= c->p.max_batch_bases && genome_len < (1ull << 32) - 1024, KASA_ERR_NOMEM,
                 "genome_len exceeds max_batch_bases");
    for (uint32_t t = first_taxon; t < first_taxon + n_taxa; ++t) {
        launch_synth_genome(c->stream, seed, t, genome_len, marker, c->d_ascii, &c->launches);
        uint64_t start = 0;
        uint32_t len = (uint32_t)(genome_len + marker), tax = 1000 + t;
        KASA_CUDA_TRY(cudaMemcpyAsync(c->d_seq_start, &start, 8, cudaMemcpyHostToDevice, c->stream));
        KASA_CUDA_TRY(cudaMemcpyAsync(c->d_seq_len, &len, 4, cudaMemcpyHostToDevice, c->stream));
        KASA_CUDA_TRY(cudaMemcpyAsync(c->d_taxid, &tax, 4, cudaMemcpyHostToDevice, c->stream));
        KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
        kasa_seq_batch sbatch{1, KASA_MEM_DEVICE, c->d_ascii, c->d_seq_start, c->d_seq_len, c->d_taxid};
        int rc = kasa_build_push(c, &sbatch);
        if (rc != KASA_OK) return rc;
    }
    KASA_API_END(c)
}

int kasa_synth_reads(kasa_ctx *c, uint64_t genome_seed, uint64_t read_seed, uint64_t first_read, uint32_t n_reads, uint32_t n_taxa,
                     uint64_t genome_len, uint32_t read_len, uint32_t subst_ppm, uint32_t n_ppm, char *out, int out_location)
{
    KASA_API_BEGIN
    KASA_REQUIRE(c && out, KASA_ERR_ARG, "null argument");
    KASA_REQUIRE(genome_len >= read_len && n_taxa > 0, KASA_ERR_ARG, "genome shorter than a read");
    KASA_CUDA_TRY(cudaSetDevice(c->p.device));
    const uint64_t total = (uint64_t)n_reads * read_len;
    if (out_location == KASA_MEM_DEVICE) {
        launch_synth_reads(c->stream, genome_seed, read_seed, first_read, n_reads, n_taxa, genome_len, read_len, subst_ppm, n_ppm, out,
                           &c->launches);
    } else {
        ensure_workspace(c);
        KASA_REQUIRE(total <= c->p.max_batch_bases, KASA_ERR_NOMEM, "more bases than max_batch_bases");
        launch_synth_reads(c->stream, genome_seed, read_seed, first_read, n_reads, n_taxa, genome_len, read_len, subst_ppm, n_ppm,
                           c->d_ascii, &c->launches);
        KASA_CUDA_TRY(cudaMemcpyAsync(out, c->d_ascii, total, cudaMemcpyDeviceToHost, c->stream));
    }
    KASA_CUDA_TRY(cudaGetLastError());
    KASA_CUDA_TRY(cudaStreamSynchronize(c->stream));
    KASA_API_END(c)
}

}  // extern "C"
